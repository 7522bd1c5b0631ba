cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s28; mkdir -p $O
for r in a b; do
SPEECHT_B200_STEP_TRACE=1 timeout 600 python bench.py --config 4 --no-cpu-baseline > $O/cfg4_$r.json 2> $O/cfg4_$r.err
python -c "import json;d=json.loads(open('$O/cfg4_$r.json').read().strip().splitlines()[-1]);print('cfg4 $r ms %.3f value %.0f e2e %.0f (%.3f ms)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))"
grep 'train step trace' $O/cfg4_$r.err | tail -13
done
SPEECHT_B200_STEP_TRACE=1 timeout 600 python bench.py --config 4 --no-cpu-baseline --no-sustained > $O/cfg4_ns.json 2> $O/cfg4_ns.err
python -c "import json;d=json.loads(open('$O/cfg4_ns.json').read().strip().splitlines()[-1]);print('cfg4 no sustained ms %.3f value %.0f e2e %.0f (%.3f ms)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))"
grep 'train step trace' $O/cfg4_ns.err | tail -13
