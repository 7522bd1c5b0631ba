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
if [ $rc -ne 0 ]; then
  SPEECHT_B200_TMA_STORE=0 timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "not two_gpu" > $O/t_model_nostore.log 2>&1
  stamp "test_gpu_model TMA_STORE=0 rc=$?: $(tail -1 $O/t_model_nostore.log)"
  SPEECHT_B200_PACK_MERGED=0 timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "not two_gpu" > $O/t_model_nomerge.log 2>&1
  stamp "test_gpu_model PACK_MERGED=0 rc=$?: $(tail -1 $O/t_model_nomerge.log)"
  SPEECHT_B200_TMA_STORE=0 SPEECHT_B200_PACK_MERGED=0 timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "not two_gpu" > $O/t_model_neither.log 2>&1
  stamp "test_gpu_model neither rc=$?: $(tail -1 $O/t_model_neither.log)"
fi

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
bench nostore SPEECHT_B200_TMA_STORE=0
bench ctcns2 SPEECHT_B200_CTC_NS=2
bench ctcns4 SPEECHT_B200_CTC_NS=4
bench default2 A=1
bench base2 SPEECHT_B200_LIB=$PWD/speecht_b200/libspeecht_b200_base.so
bench bf16 SPEECHT_B200_PRECISION=bf16

timeout 1200 python -m pytest tests -x -q -m gpu > $O/t_all_default.log 2>&1
stamp "pytest -m gpu (all) rc=$?: $(tail -1 $O/t_all_default.log)"

timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
stamp "ncu launch list rc=$?"
# --set full captures of two small launches whose epilogue is not hidden: layer-1 forward (23rd tc_conv launch) and
# the layer-10 data gradient (33rd), both in the second step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel --launch-skip 22 --launch-count 1 \
  -o $O/ncu_l1_fwd -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_l1_fwd.log 2>&1
stamp "ncu full L1 fwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel --launch-skip 32 --launch-count 1 \
  -o $O/ncu_l10_dgrad -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_l10_dgrad.log 2>&1
stamp "ncu full L10 dgrad rc=$?"
cat $S
