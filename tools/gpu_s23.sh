cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s23; mkdir -p $O; : > $O/summary.txt
for r in a b c; do
timeout 600 python bench.py --no-cpu-baseline --no-sustained > $O/cfg2_$r.json 2> $O/cfg2_$r.err
python -c "import json;d=json.loads(open('$O/cfg2_$r.json').read().strip().splitlines()[-1]);print('cfg2 $r ms %.3f value %.0f e2e %.0f (%.3f ms)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))" >> $O/summary.txt 2>&1
done
timeout 600 python bench.py --no-cpu-baseline > $O/cfg2_sus.json 2> $O/cfg2_sus.err
python -c "import json;d=json.loads(open('$O/cfg2_sus.json').read().strip().splitlines()[-1]);print('cfg2 with sustained pass ms %.3f value %.0f e2e %.0f (%.3f ms)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))" >> $O/summary.txt 2>&1
nproc >> $O/summary.txt; uptime >> $O/summary.txt
cat $O/summary.txt
