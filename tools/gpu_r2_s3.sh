#!/bin/bash
# Round 2, GPU session 3: A/B of the 128-wide layer-10 data gradient, CTC kernels timed alone + ncu source view of the
# alpha/beta recursion.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s3
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp start
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "train_step_parity or every_layer or config1" > $O/t_model.log 2>&1
stamp "model parity (L10 n128 default) rc=$?: $(tail -1 $O/t_model.log)"
ab() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-sustained > $O/ab_$name.json 2> $O/ab_$name.err
  stamp "A/B $name rc=$?: $(python - <<P
import json
try:
  d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1])
  r=d['roofline']
  print('ms/step %.3f  value %.0f  L10.dgrad %s L10.fwd %s L10.wgrad %s' % (d['ms_per_step'], d['value'], r['layers_ms_per_step'].get('L10.dgrad'), r['layers_ms_per_step'].get('L10.fwd'), r['layers_ms_per_step'].get('L10.wgrad')))
except Exception as e:
  print('unreadable', e)
P
)"
}
ab n128 A=1
ab n256 SPEECHT_B200_L10_N128=0
ab n128b A=1
ab n256b SPEECHT_B200_L10_N128=0
timeout 300 python tools/ctc_bench.py > $O/ctc_bench.txt 2>&1
stamp "ctc bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ctc_launches.csv python tools/ctc_bench.py 1 > $O/ctc_ncu.log 2>&1
stamp "ctc launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ctc_alpha_beta --launch-skip 3 --launch-count 1 -o $O/ncu_ctc_ab -f python tools/ctc_bench.py 1 > $O/ncu_ctc_ab.log 2>&1
stamp "ncu full ctc rc=$?"
cat $S
