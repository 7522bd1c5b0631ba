#!/bin/bash
# One gpurun call: correctness of the current build, A/B benches of the kernel switches, full GPU suite, launch list.
# Everything is written under gpurun_out/ (merged back by gpurun).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
mkdir -p $O
S=$O/summary.txt
: > $S
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/smi.txt 2>&1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }

stamp start
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "not two_gpu" > $O/t_model_default.log 2>&1
rc=$?; stamp "test_gpu_model default rc=$rc: $(tail -1 $O/t_model_default.log)"

bench() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_$name.json 2> $O/bench_$name.err
  stamp "bench $name rc=$?: $(python - <<P
import json
try:
  d=json.loads(open('$O/bench_$name.json').read().strip().splitlines()[-1])
  r=d['roofline']
  print('ms/step %.3f  value %.0f  e2e %.0f  conv %.3f wgrad %.3f  launches %d  ctc %.4f' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['kernels']['tc_conv_kernel']['ms_per_step'], r['kernels']['tc_wgrad_kernel']['ms_per_step'], d['gpu_launches'], d['aux_hbm_kernels']['ctc_loss+grad (a8-a9)']['ms']))
except Exception as e:
  print('unreadable', e)
P
)"
}
bench default A=1
bench base SPEECHT_B200_LIB=$PWD/speecht_b200/libspeecht_b200_base.so
bench default2 A=1
bench base2 SPEECHT_B200_LIB=$PWD/speecht_b200/libspeecht_b200_base.so
bench bf16 SPEECHT_B200_PRECISION=bf16
bench bf16_base SPEECHT_B200_PRECISION=bf16 SPEECHT_B200_LIB=$PWD/speecht_b200/libspeecht_b200_base.so

timeout 1200 python -m pytest tests -x -q -m gpu > $O/t_all_default.log 2>&1
stamp "pytest -m gpu (all) rc=$?: $(tail -1 $O/t_all_default.log)"

timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
stamp "ncu launch list rc=$?"
# --set full captures (second step): layer-1 forward, layer-7 data gradient, layer-10 data gradient
for spec in l1_fwd:22 l7_dgrad:35 l10_dgrad:32; do
  name=${spec%%:*}; skip=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel --launch-skip $skip --launch-count 1 \
    -o $O/ncu_$name -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_$name.log 2>&1
  stamp "ncu full $name rc=$?"
done
cat $S
