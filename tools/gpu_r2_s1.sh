#!/bin/bash
# Round 2, GPU session 1: does the experimental fast-FIR forward of layer 8 hold on hardware (parity + same-box A/B),
# and where do BASELINE configs 3 / 4 stand with the round-1 kernels.  Outputs under gpurun_out/s1/.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s1
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp start
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
SPEECHT_B200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "fast_fir" > $O/t_ffa.log 2>&1
stamp "fast_fir parity rc=$?: $(tail -1 $O/t_ffa.log)"
ab() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/ab_$name.json 2> $O/ab_$name.err
  stamp "A/B $name rc=$?: $(python - <<P
import json
try:
  d=json.loads(open('$O/ab_$name.json').read().strip().splitlines()[-1])
  r=d['roofline']
  print('ms/step %.3f  value %.0f  e2e %.0f  conv %.3f wgrad %.3f L8.fwd %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['kernels']['tc_conv_kernel']['ms_per_step'], r['kernels']['tc_wgrad_kernel']['ms_per_step'], r['layers_ms_per_step'].get('L8.fwd')))
except Exception as e:
  print('unreadable', e)
P
)"
}
ab base A=1
ab ffa SPEECHT_B200_FFA=1
ab base2 A=1
ab ffa2 SPEECHT_B200_FFA=1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precision bf16 --batch 64 > $O/cfg3.json 2> $O/cfg3.err
stamp "cfg3 (B=64 bf16) rc=$?: $(python -c "import json;d=json.loads(open('$O/cfg3.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'], d['roofline']['frac'])" 2>&1 | tail -1)"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision bf16 --seconds 30 > $O/cfg4.json 2> $O/cfg4.err
stamp "cfg4 (B=32 30s bf16) rc=$?: $(python -c "import json;d=json.loads(open('$O/cfg4.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'], d['roofline']['frac'])" 2>&1 | tail -1)"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --seconds 30 > $O/cfg4_x3.json 2> $O/cfg4_x3.err
stamp "cfg4 shape bf16x3 rc=$?: $(python -c "import json;d=json.loads(open('$O/cfg4_x3.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'], d['roofline']['frac'])" 2>&1 | tail -1)"
timeout 600 python tools/accuracy_sweep.py 32 1001 8 > $O/accuracy_sweep.txt 2>&1
stamp "accuracy sweep rc=$?"
cat $S
