#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s12
mkdir -p $O
S=$O/summary.txt
: > $S
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" >> $S; }
stamp start
timeout 1500 python -m pytest tests -x -q -m gpu > $O/t_all.log 2>&1
stamp "pytest -m gpu rc=$?: $(tail -1 $O/t_all.log)"
run() {
  name=$1; shift
  timeout 600 python bench.py "$@" --no-cpu-baseline --no-sustained > $O/$name.json 2> $O/$name.err
  stamp "$name rc=$?: $(python -c "import json;d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]);print('ms %.3f value %.0f e2e %.0f (%.3f ms) h2d %.1f MB' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step']/1e6))" 2>&1 | tail -1)"
}
run cfg5 --config 5 --steps 10 --warmup 3
run cfg5_b8 --config 5 --steps 10 --warmup 3 --eval-buckets 8
run cfg5_b16 --config 5 --steps 10 --warmup 3 --eval-buckets 16
run cfg2 --steps 30 --warmup 5
cat $S
