#!/bin/bash
# Round 2, GPU session 2: the whole GPU suite with the new parity tests, smoke, and the bench lines of BASELINE configs
# 2-5 in the new format (ctc_loss_delta, burst/sustained denominators, sustained pass) plus the reference arm.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s2
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp start
timeout 1500 python -m pytest tests -x -q -m gpu -s > $O/t_all.log 2>&1
stamp "pytest -m gpu rc=$?: $(tail -1 $O/t_all.log)"
SPEECHT_B200_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -k "fast_fir" -s > $O/t_ffa.log 2>&1
stamp "fast_fir parity rc=$?: $(grep 'fast-FIR' $O/t_ffa.log | tr '\n' ' ')"
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
timeout 600 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err
stamp "bench cfg2 rc=$?: $(line $O/bench_cfg2.json)"
timeout 600 python bench.py --impl reference > $O/bench_cfg2_reference.json 2> $O/bench_cfg2_reference.err
stamp "bench cfg2 reference rc=$?: $(tail -c 300 $O/bench_cfg2_reference.json | head -c 200)"
for c in 3 4 5; do
  timeout 900 python bench.py --config $c > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err
  stamp "bench cfg$c rc=$?: $(line $O/bench_cfg$c.json)"
done
timeout 600 python bench.py --config 5 --impl reference > $O/bench_cfg5_reference.json 2> $O/bench_cfg5_reference.err
stamp "bench cfg5 reference rc=$?"
timeout 600 python bench.py --config 5 --precision bf16 --no-cpu-baseline > $O/bench_cfg5_bf16.json 2> $O/bench_cfg5_bf16.err
stamp "bench cfg5 bf16 rc=$?: $(line $O/bench_cfg5_bf16.json)"
cat $S
