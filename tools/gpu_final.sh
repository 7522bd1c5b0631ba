#!/bin/bash
# Round artefacts in one gpurun call: quick check + same-box A/B of the MMA trimming, GPU suite, smoke, bench lines for
# every precision + the reference arm, the ncu launch list, --set full captures of the dominant launches and the
# per-CTA timelines.  Outputs under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp start
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "not two_gpu" > $O/t_model_default.log 2>&1
rc=$?; stamp "test_gpu_model default rc=$rc: $(tail -1 $O/t_model_default.log)"
if [ $rc -ne 0 ]; then
  export SPEECHT_B200_TRIM=0
  stamp "FALLING BACK to SPEECHT_B200_TRIM=0 for the rest of the session"
fi
ab() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/ab_$name.json 2> $O/ab_$name.err
  stamp "A/B $name rc=$?: $(python - <<P
import json
try:
  d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1])
  r=d['roofline']
  print('ms/step %.3f  value %.0f  e2e %.0f  conv %.3f wgrad %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['kernels']['tc_conv_kernel']['ms_per_step'], r['kernels']['tc_wgrad_kernel']['ms_per_step']))
except Exception as e:
  print('unreadable', e)
P
)"
}
ab trim A=1
ab notrim SPEECHT_B200_TRIM=0
ab trim2 A=1
ab notrim2 SPEECHT_B200_TRIM=0
timeout 1200 python -m pytest tests -x -q -m gpu > $O/t_all.log 2>&1
stamp "pytest -m gpu rc=$?: $(tail -1 $O/t_all.log)"
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
stamp "smoke rc=$?: $(tail -2 $O/smoke.log | tr '\n' ' ')"
timeout 600 python bench.py > $O/bench_bf16x3.json 2> $O/bench_bf16x3.err
stamp "bench default rc=$?"
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
stamp "bench reference rc=$?"
for prec in bf16 bf16x6 fp32; do
  timeout 600 python bench.py --precision $prec --no-cpu-baseline > $O/bench_$prec.json 2> $O/bench_$prec.err
  stamp "bench $prec rc=$?"
done
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_bf16x3_50.json 2> $O/bench_bf16x3_50.err
stamp "bench 50 steps rc=$?"
timeout 300 python tools/conv_timeline.py > $O/timeline_bf16x3.txt 2>&1
stamp "timeline rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
stamp "ncu launch list rc=$?"
full() {  # name, kernel regex, skip, extra bench args
  name=$1; rx=$2; skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip --launch-count 1 \
    -o $O/ncu_$name -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > $O/ncu_$name.log 2>&1
  stamp "ncu full $name rc=$?"
}
# second step: tc_conv launches 21.. (11 forward, 10 data gradients per step), tc_wgrad launches 11.. (L10, L9, L8, ...)
full l8_fwd tc_conv_kernel 29
full l8_wgrad tc_wgrad_kernel 13
full l1_fwd tc_conv_kernel 22
full l7_dgrad tc_conv_kernel 35
full l10_dgrad tc_conv_kernel 32
full l8_fwd_bf16 tc_conv_kernel 29 --precision bf16
cat $S
