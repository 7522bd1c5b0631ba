#!/bin/bash
# Round 2, GPU session 8: CTA pairs on by default (per-launch policy): whole GPU suite, A/B against PAIR=0, the
# 128-wide-tile experiment for the 250-channel layers, 2 planes.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s8
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp start
timeout 1500 python -m pytest tests -x -q -m gpu > $O/t_all.log 2>&1
stamp "pytest -m gpu rc=$?: $(tail -1 $O/t_all.log)"
SPEECHT_B200_SMALL_N128=1 timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "config1 or every_layer or train_step_parity or ragged or tiny" > $O/t_n128.log 2>&1
stamp "small-layer n128 parity rc=$?: $(tail -1 $O/t_n128.log)"
ab() {
  name=$1; cfg=$2; shift 2
  env "$@" timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline --sustained-seconds 1.5 > $O/ab_$name.json 2> $O/ab_$name.err
  stamp "A/B $name rc=$?: $(python - <<P
import json
try:
  d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1])
  r=d['roofline']; L=r['layers_ms_per_step']
  small=sum(v for k,v in L.items() if k.split('.')[0] in ('L1','L2','L3','L4','L5','L6','L7'))
  print('ms/step %.3f sustained %.3f value %.0f e2e %.0f L8 %s %s %s L9 %s %s %s L1-7 %.3f' % (d['ms_per_step'], d['sustained']['ms_per_step'], d['value'], d['e2e']['value'], L.get('L8.fwd'), L.get('L8.dgrad'), L.get('L8.wgrad'), L.get('L9.fwd'), L.get('L9.dgrad'), L.get('L9.wgrad'), small))
except Exception as e:
  print('unreadable', e)
P
)"
}
ab default 2 A=1
ab nopair 2 SPEECHT_B200_PAIR=0
ab n128 2 SPEECHT_B200_SMALL_N128=1
ab default_b 2 A=1
ab nopair_b 2 SPEECHT_B200_PAIR=0
ab n128_b 2 SPEECHT_B200_SMALL_N128=1
ab cfg3_default 3 A=1
ab cfg3_nopair 3 SPEECHT_B200_PAIR=0
ab cfg4_default 4 A=1
ab cfg4_nopair 4 SPEECHT_B200_PAIR=0
cat $S
