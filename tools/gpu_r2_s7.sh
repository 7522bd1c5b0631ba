#!/bin/bash
# Round 2, GPU session 7: CTC gradient kernel (private bins, no fp64), one-pass fast-FIR filter packing, and the
# first run of the CTA-pair (cta_group::2) conv kernel: parity, then same-box A/B.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s7
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp start
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_properties.py -x -q -m gpu > $O/t_ops.log 2>&1
stamp "ops + properties rc=$?: $(tail -1 $O/t_ops.log)"
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "config1 or every_layer or train_step_parity" > $O/t_model.log 2>&1
stamp "model (single CTA) rc=$?: $(tail -1 $O/t_model.log)"
timeout 120 python tools/ctc_bench.py > $O/ctc_bench.txt 2>&1
stamp "ctc bench rc=$?: $(head -1 $O/ctc_bench.txt) | $(sed -n 5p $O/ctc_bench.txt)"
SPEECHT_B200_PAIR=1 timeout 300 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "config1 or every_layer or train_step_parity or ragged or tiny" > $O/t_pair.log 2>&1
rc=$?
stamp "model parity with CTA pairs rc=$rc: $(tail -1 $O/t_pair.log)"
if [ $rc -eq 0 ]; then
  SPEECHT_B200_PAIR=1 timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -s > $O/t_pair_full.log 2>&1
  stamp "fullsize with CTA pairs rc=$?: $(tail -1 $O/t_pair_full.log)"
  ab() {
    name=$1; cfg=$2; shift 2
    env "$@" timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline --sustained-seconds 1.5 > $O/ab_$name.json 2> $O/ab_$name.err
    stamp "A/B $name rc=$?: $(python - <<P
import json
try:
  d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1])
  r=d['roofline']; L=r['layers_ms_per_step']
  print('ms/step %.3f sustained %.3f value %.0f  L8 %s %s %s L9 %s %s %s' % (d['ms_per_step'], d['sustained']['ms_per_step'], d['value'], L.get('L8.fwd'), L.get('L8.dgrad'), L.get('L8.wgrad'), L.get('L9.fwd'), L.get('L9.dgrad'), L.get('L9.wgrad')))
except Exception as e:
  print('unreadable', e)
P
)"
  }
  ab single 2 A=1
  ab pair 2 SPEECHT_B200_PAIR=1
  ab single_b 2 A=1
  ab pair_b 2 SPEECHT_B200_PAIR=1
  ab cfg3_single 3 A=1
  ab cfg3_pair 3 SPEECHT_B200_PAIR=1
else
  tail -40 $O/t_pair.log >> $S
fi
cat $S
