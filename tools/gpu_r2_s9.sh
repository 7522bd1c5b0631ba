#!/bin/bash
# Round 2, GPU session 9: profiler evidence for the round-2 kernels -- ncu launch list of the default bench command and
# `--set full` captures of the dominant launches (fast-FIR multi-problem forward / data gradient on CTA pairs, the
# multi-problem filter gradient, a 250-channel layer, the CTC recursion).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s9
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp start
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sustained"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_cfg2.csv $B > $O/ncu_list.log 2>&1
stamp "ncu launch list cfg2 rc=$?"
full() {  # name, kernel regex, skip, extra bench args
  name=$1; rx=$2; skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip --launch-count 1 \
    -o $O/ncu_$name -f $B "$@" > $O/ncu_$name.log 2>&1
  stamp "ncu full $name rc=$?"
}
# tc_conv launches: 11 (loss-delta forward) + 21 (warm-up step) + 21 (first timed step), then forward L0..L10, data
# gradients L10, L9, L8, L7..L1; tc_wgrad launches: 11 per step in the order L10, L9, L8, L7..L0
full l8_fwd_ffa_pair tc_conv_kernel 61
full l8_dgrad_ffa_pair tc_conv_kernel 66
full l9_dgrad_pair tc_conv_kernel 65
full l1_fwd tc_conv_kernel 54
full l8_wgrad_ffa tc_wgrad_kernel 24
full ctc_alpha_beta ctc_alpha_beta 3
full l8_fwd_ffa_pair_bf16 tc_conv_kernel 61 --config 3
cat $S
