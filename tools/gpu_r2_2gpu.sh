#!/bin/bash
# Round 2, 2-GPU session (gpurun --gpus 2): NCCL data-parallel parity test, bench at N=2 for configs 2, 4 and 5
# (dp_identical after the timed steps, tensor gather of the decoded rows).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/g2
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp start
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "two_gpu" -s > $O/t_dp.log 2>&1
stamp "2-GPU DP parity test rc=$?: $(tail -1 $O/t_dp.log)"
run() {
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 "$@" > $O/$name.json 2> $O/$name.err
  stamp "$name rc=$?: $(python - <<P
import json
try:
  d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1])
  print('n_gpus %d ms/step %.3f value %.0f e2e %.0f dp_identical %s sustained %s' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('dp_identical'), (d.get('sustained') or {}).get('ms_per_step')))
except Exception as e:
  print('unreadable', e)
P
)"
}
run bench_cfg2_n2 --steps 20 --warmup 5
run bench_cfg4_n2 --config 4 --steps 10 --warmup 3 --no-sustained
run bench_cfg5_n2 --config 5 --steps 5 --warmup 2 --no-sustained
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg2_n1.json 2> $O/bench_cfg2_n1.err
stamp "cfg2 N=1 same box rc=$?: $(python -c "import json;d=json.loads(open('$O/bench_cfg2_n1.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'], d['e2e']['value'])" 2>&1 | tail -1)"
# 2-rank CLI training smoke on a synthetic corpus (termination word over the gloo group, sharded file list)
rm -rf /tmp/synth /tmp/synth_train /tmp/synth_log
python tools/make_synth_data.py /tmp/synth 512 > /dev/null 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
  speecht-cli-b200 train --data-dir /tmp/synth --train-dir /tmp/synth_train --log-dir /tmp/synth_log --run-name dp2 \
  --batch-size 32 --steps-per-checkpoint 50 --max-steps 200 > $O/cli_train_2gpu.log 2>&1
stamp "2-rank CLI train rc=$?: $(grep 'global step' $O/cli_train_2gpu.log | tail -2 | cut -c1-90 | tr '\n' '|')"
timeout 600 python speecht-cli-b200 train --data-dir /tmp/synth --train-dir /tmp/synth_train1 --log-dir /tmp/synth_log --run-name dp1 \
  --batch-size 32 --steps-per-checkpoint 50 --max-steps 200 > $O/cli_train_1gpu.log 2>&1
stamp "1-rank CLI train rc=$?: $(grep 'global step' $O/cli_train_1gpu.log | tail -2 | cut -c1-90 | tr '\n' '|')"
cat $S
