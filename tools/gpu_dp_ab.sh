#!/bin/bash
# 2-GPU A/B of the data-parallel overlap settings: SMs reserved for the NCCL allreduce x NCCL's CTA budget.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/dp_ab
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
run() {
  name=$1; envs=$2; shift 2
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --no-sustained "$@" > $O/$name.json 2> $O/$name.err
  stamp "$name [$envs] rc=$?: $(python - <<P
import json
try:
  d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1])
  print('ms/step %.3f value %.0f e2e %.0f dp_identical %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d.get('dp_identical')))
except Exception as e:
  print('unreadable', e)
P
)"
}
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-sustained > $O/n1.json 2> $O/n1.err
stamp "N=1 same box: $(python -c "import json;d=json.loads(open('$O/n1.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'])" 2>&1 | tail -1)"
for rep in a b; do
  run r0_$rep "SPEECHT_B200_DP_RESERVE_SMS=0"
  run r8_$rep "SPEECHT_B200_DP_RESERVE_SMS=8"
  run r16_$rep "SPEECHT_B200_DP_RESERVE_SMS=16"
  run r8_c8_$rep "SPEECHT_B200_DP_RESERVE_SMS=8 NCCL_MAX_CTAS=8"
  run r4_c4_$rep "SPEECHT_B200_DP_RESERVE_SMS=4 NCCL_MAX_CTAS=4"
  run r16_c16_$rep "SPEECHT_B200_DP_RESERVE_SMS=16 NCCL_MAX_CTAS=16"
done
run cfg4_r0 "SPEECHT_B200_DP_RESERVE_SMS=0" --config 4
run cfg4_r8 "SPEECHT_B200_DP_RESERVE_SMS=8" --config 4
cat $S
