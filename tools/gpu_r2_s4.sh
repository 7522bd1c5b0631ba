#!/bin/bash
# Round 2, GPU session 4: fast-FIR layer 8 in forward, data gradient and filter gradient (default on): parity of the
# whole model suite, same-box A/B against the direct kernels.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s4
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp start
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -s > $O/t_model.log 2>&1
stamp "test_gpu_model (FFA on) rc=$?: $(tail -1 $O/t_model.log)"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -s > $O/t_full.log 2>&1
stamp "test_gpu_fullsize (FFA on) rc=$?: $(tail -1 $O/t_full.log)"
ab() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --sustained-seconds 1.5 > $O/ab_$name.json 2> $O/ab_$name.err
  stamp "A/B $name rc=$?: $(python - <<P
import json
try:
  d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1])
  r=d['roofline']; L=r['layers_ms_per_step']
  print('ms/step %.3f sustained %.3f value %.0f  L8 fwd %s dgrad %s wgrad %s  delta %.1e' % (d['ms_per_step'], d['sustained']['ms_per_step'], d['value'], L.get('L8.fwd'), L.get('L8.dgrad'), L.get('L8.wgrad'), d['ctc_loss_delta']['max_rel']))
except Exception as e:
  print('unreadable', e)
P
)"
}
ab ffa A=1
ab direct SPEECHT_B200_FFA=0
ab ffa2 A=1
ab direct2 SPEECHT_B200_FFA=0
env timeout 300 python bench.py --config 3 --steps 20 --warmup 5 --no-cpu-baseline --no-sustained > $O/cfg3_ffa.json 2> $O/cfg3_ffa.err
stamp "cfg3 ffa rc=$?: $(python -c "import json;d=json.loads(open('$O/cfg3_ffa.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'])" 2>&1 | tail -1)"
env SPEECHT_B200_FFA=0 timeout 300 python bench.py --config 3 --steps 20 --warmup 5 --no-cpu-baseline --no-sustained > $O/cfg3_direct.json 2> $O/cfg3_direct.err
stamp "cfg3 direct rc=$?: $(python -c "import json;d=json.loads(open('$O/cfg3_direct.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'])" 2>&1 | tail -1)"
cat $S
