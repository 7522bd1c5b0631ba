#!/bin/bash
# Round 2, GPU session 10: two-level fast-FIR split of layer 8 (SPEECHT_B200_FFA=2): parity, then A/B against one level.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s10
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp start
SPEECHT_B200_FFA=2 timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -s -k "config1 or every_layer or train_step_parity or ragged or tiny or shape_changes" > $O/t_model.log 2>&1
rc=$?
stamp "model parity (two levels) rc=$rc: $(tail -1 $O/t_model.log)"
if [ $rc -ne 0 ]; then tail -60 $O/t_model.log >> $S; fi
SPEECHT_B200_FFA=2 timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -s > $O/t_full.log 2>&1
stamp "fullsize (two levels) rc=$?: $(tail -1 $O/t_full.log) | $(grep 'bf16x3 vs float64' $O/t_full.log)"
ab() {
  name=$1; cfg=$2; shift 2
  env "$@" timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline --sustained-seconds 1.5 > $O/ab_$name.json 2> $O/ab_$name.err
  stamp "A/B $name rc=$?: $(python - <<P
import json
try:
  d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1])
  r=d['roofline']; L=r['layers_ms_per_step']
  print('ms/step %.3f sustained %.3f value %.0f e2e %.0f L8 %s %s %s conv-sum %.3f delta %.1e logits %.2e' % (d['ms_per_step'], d['sustained']['ms_per_step'], d['value'], d['e2e']['value'], L.get('L8.fwd'), L.get('L8.dgrad'), L.get('L8.wgrad'), sum(L.values()), d['ctc_loss_delta']['max_rel'], d['ctc_loss_delta']['logits_max_rel']))
except Exception as e:
  print('unreadable', e)
P
)"
}
ab level1 2 A=1
ab level2 2 SPEECHT_B200_FFA=2
ab level1_b 2 A=1
ab level2_b 2 SPEECHT_B200_FFA=2
ab cfg3_level1 3 A=1
ab cfg3_level2 3 SPEECHT_B200_FFA=2
ab cfg4_level1 4 A=1
ab cfg4_level2 4 SPEECHT_B200_FFA=2
cat $S
