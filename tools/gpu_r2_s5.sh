#!/bin/bash
# Round 2, GPU session 5: rewritten CTC recursion (zero-free scaled floats, probabilities as emissions, loss-only mode).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s5
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp start
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_properties.py -x -q -m gpu > $O/t_ops.log 2>&1
stamp "ops + properties rc=$?: $(tail -1 $O/t_ops.log)"
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -x -q -m gpu > $O/t_model.log 2>&1
stamp "model + fullsize rc=$?: $(tail -1 $O/t_model.log)"
timeout 300 python tools/ctc_bench.py > $O/ctc_bench.txt 2>&1
stamp "ctc bench rc=$?"
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --sustained-seconds 1.5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err
stamp "cfg2 rc=$?: $(python -c "import json;d=json.loads(open('$O/bench_cfg2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['sustained']['ms_per_step'], d['value'], d['aux_hbm_kernels']['ctc_loss+grad (a8-a9)'])" 2>&1 | tail -1)"
timeout 300 python bench.py --config 4 --steps 20 --warmup 5 --no-cpu-baseline --no-sustained > $O/bench_cfg4.json 2> $O/bench_cfg4.err
stamp "cfg4 rc=$?: $(python -c "import json;d=json.loads(open('$O/bench_cfg4.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'], d['aux_hbm_kernels']['ctc_loss+grad (a8-a9)'])" 2>&1 | tail -1)"
timeout 300 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu-baseline --no-sustained > $O/bench_cfg5.json 2> $O/bench_cfg5.err
stamp "cfg5 rc=$?: $(python -c "import json;d=json.loads(open('$O/bench_cfg5.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'], d['e2e']['value'])" 2>&1 | tail -1)"
cat $S
