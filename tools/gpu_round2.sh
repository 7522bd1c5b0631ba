#!/bin/bash
# Round-2 artefacts in one gpurun call: GPU suite, smoke, the bench lines of BASELINE configs 2-5 with their reference
# arms, the 8-seed accuracy sweep, the ncu launch list of the default bench command and `--set full` captures of the
# dominant launches.  Outputs under gpurun_out/final/; tools/collect_round2.py copies the summaries into profiles/.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/final
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp start
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,driver_version --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu -s > $O/t_all.log 2>&1
stamp "pytest -m gpu rc=$?: $(tail -1 $O/t_all.log)"
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
stamp "smoke rc=$?: $(tail -2 $O/smoke.log | tr '\n' ' ')"
line() {
  python - "$1" <<P
import json, sys
try:
  d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  r=d.get('roofline') or {}
  s=d.get('sustained') or {}
  c=d.get('ctc_loss_delta') or {}
  print('ms/step %.3f value %.0f e2e %.0f | frac %.3f (burst %.3f sust %.3f) | sustained %.3f ms | loss delta %.1e labels_equal %s | cpu %s' % (
    d['ms_per_step'], d['value'], d['e2e']['value'], r.get('frac',0), r.get('frac_burst',0), r.get('frac_sustained',0),
    s.get('ms_per_step',0), c.get('max_rel',-1), c.get('greedy_labels_equal'), (d.get('cpu_baseline') or {}).get('value')))
except Exception as e:
  print('unreadable', e)
P
}
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_cfg2_reference.json 2> $O/bench_cfg2_reference.err
stamp "bench cfg2 reference rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err
stamp "bench cfg2 rc=$?: $(line $O/bench_cfg2.json)"
for c in 3 4 5; do
  timeout 900 python bench.py --config $c > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err
  stamp "bench cfg$c rc=$?: $(line $O/bench_cfg$c.json)"
done
timeout 600 python bench.py --config 5 --impl reference > $O/bench_cfg5_reference.json 2> $O/bench_cfg5_reference.err
stamp "bench cfg5 reference rc=$?"
timeout 600 python bench.py --config 5 --precision bf16 --no-cpu-baseline > $O/bench_cfg5_bf16.json 2> $O/bench_cfg5_bf16.err
stamp "bench cfg5 bf16 rc=$?: $(line $O/bench_cfg5_bf16.json)"
timeout 600 python bench.py --config 5 --eval-buckets 8 --no-cpu-baseline > $O/bench_cfg5_buckets8.json 2> $O/bench_cfg5_buckets8.err
stamp "bench cfg5 buckets 8 rc=$?: $(line $O/bench_cfg5_buckets8.json)"
timeout 900 python tools/accuracy_sweep.py 32 1001 8 > $O/accuracy_sweep.txt 2>&1
stamp "accuracy sweep rc=$?: $(tail -1 $O/accuracy_sweep.txt)"
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sustained"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_cfg2.csv $B > $O/ncu_list.log 2>&1
stamp "ncu launch list cfg2 rc=$?"
full() {  # name, kernel regex, skip, extra bench args
  name=$1; rx=$2; skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip --launch-count 1 \
    -o $O/ncu_$name -f $B "$@" > $O/ncu_$name.log 2>&1
  rc=$?
  # gpurun brings back at most 64 MiB: keep the raw metric page (what tools/collect_round2.py reads), drop the report
  ncu -i $O/ncu_$name.ncu-rep --page raw --csv > $O/ncu_$name.raw.csv 2>/dev/null && rm -f $O/ncu_$name.ncu-rep
  stamp "ncu full $name rc=$rc"
}
# tc_conv launches: 11 (loss-delta forward) + 21 (warm-up step) + 21 (first timed step), then forward L0..L10, data
# gradients L10, L9, L8, L7..L1; tc_wgrad launches: 5 per step in the order L10, L9, L8 (nine leaves), L0, L1-7 (merged)
full l8_fwd tc_conv_kernel 61
full l8_dgrad tc_conv_kernel 66
full l9_dgrad tc_conv_kernel 65
full l1_fwd tc_conv_kernel 54
full l8_wgrad tc_wgrad_kernel 12
full l1_7_wgrad tc_wgrad_kernel 14
full ctc_alpha_beta ctc_alpha_beta 3
full pack pack_bwd 2
full l10_dgrad tc_conv_kernel 64
full ffa2_combine ffa2_combine 3
full ffa2_dz_prep ffa2_dz_prep 3
full l8_fwd_cfg3 tc_conv_kernel 61 --config 3
cat $S
