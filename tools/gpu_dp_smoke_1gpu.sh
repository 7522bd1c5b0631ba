#!/bin/bash
# Smoke run of the data-parallel CLI training loop with TWO ranks on ONE GPU (gloo collectives): sharded file list,
# termination agreement, allreduce + identical update, checkpoints.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/dp_smoke; mkdir -p $O
rm -rf /tmp/synth /tmp/synth_train /tmp/synth_log
python tools/make_synth_data.py /tmp/synth 256 > /dev/null 2>&1
SPEECHT_B200_DIST_BACKEND=gloo timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
  speecht-cli-b200 train --data-dir /tmp/synth --train-dir /tmp/synth_train --log-dir /tmp/synth_log --run-name dp2 \
  --batch-size 16 --steps-per-checkpoint 20 --max-steps 60 > $O/cli_train_2rank_1gpu.log 2>&1
echo "rc=$?"; grep -E 'global step|Done|Begin|Error|error' $O/cli_train_2rank_1gpu.log | cut -c1-110 | tail -12
