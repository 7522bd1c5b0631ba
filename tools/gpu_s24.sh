cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s24; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/t_all.log 2>&1; echo "suite rc=$?: $(tail -1 $O/t_all.log)" > $O/summary.txt
for v in 1 0 1 0; do
SPEECHT_B200_EVAL_PIPELINE=$v timeout 600 python bench.py --config 5 --no-cpu-baseline --no-sustained > $O/cfg5_$v.json 2> $O/cfg5_$v.err
python -c "import json;d=json.loads(open('$O/cfg5_$v.json').read().strip().splitlines()[-1]);print('cfg5 pipeline=$v ms %.3f value %.0f e2e %.0f (%.3f ms)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))" >> $O/summary.txt 2>&1
done
cat $O/summary.txt
