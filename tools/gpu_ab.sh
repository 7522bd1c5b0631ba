#!/bin/bash
# Same-box A/B of an environment switch: [AB_BASE="<VAR=VALUE ...>"] tools/gpu_ab.sh <tag> "<VAR=VALUE ...>" [configs...]
# (AB_BASE: environment of the base arm, e.g. SPEECHT_B200_LIB=speecht_b200/libspeecht_b200_base.so for an older build;
#  a variant of "X=" sets an empty dummy variable, i.e. the default build)
# Runs the model parity tests under the variant, then bench.py alternately without / with it.  Output: gpurun_out/<tag>/
cd "${GRAFT_REPO_ROOT:-.}"
TAG=$1; VARIANT=$2; shift 2
CFGS=${@:-2 3 4}
O=gpurun_out/$TAG
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp "start variant: $VARIANT"
env $VARIANT timeout 1200 python -m pytest tests -x -q -m gpu > $O/t_variant.log 2>&1
stamp "GPU suite under the variant rc=$?: $(tail -1 $O/t_variant.log)"
line() {
  python - "$1" <<P
import json, sys
try:
  d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  L=d['roofline']['layers_ms_per_step']
  s=d.get('sustained') or {}
  small=sum(v for k,v in L.items() if k.split('.')[0] in ('L1','L2','L3','L4','L5','L6','L7') and 'wgrad' not in k)
  print('ms/step %.3f sustained %.3f value %.0f e2e %.0f | L0 %.3f %.3f | L1-7 fwd+dgrad %.3f wgrad %.3f | L8 %.3f %.3f %.3f | L9 %.3f %.3f %.3f | L10 %.3f %.3f %.3f' % (
    d['ms_per_step'], s.get('ms_per_step',0), d['value'], d['e2e']['value'], L.get('L0.fwd',0), L.get('L0.wgrad',0), small, L.get('L1.wgrad',0),
    L.get('L8.fwd',0), L.get('L8.dgrad',0), L.get('L8.wgrad',0), L.get('L9.fwd',0), L.get('L9.dgrad',0), L.get('L9.wgrad',0),
    L.get('L10.fwd',0), L.get('L10.dgrad',0), L.get('L10.wgrad',0)))
except Exception as e:
  print('unreadable', e)
P
}
for c in $CFGS; do
  for rep in a b; do
    env $AB_BASE timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > $O/cfg${c}_base_$rep.json 2> $O/cfg${c}_base_$rep.err
    stamp "cfg$c base $rep rc=$?: $(line $O/cfg${c}_base_$rep.json)"
    env $VARIANT timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > $O/cfg${c}_var_$rep.json 2> $O/cfg${c}_var_$rep.err
    stamp "cfg$c variant $rep rc=$?: $(line $O/cfg${c}_var_$rep.json)"
  done
done
cat $S
