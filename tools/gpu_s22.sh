cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s22; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/t_all.log 2>&1; echo "suite rc=$?: $(tail -1 $O/t_all.log)" > $O/summary.txt
timeout 600 python bench.py --config 5 --no-cpu-baseline > $O/cfg5.json 2> $O/cfg5.err
python -c "import json;d=json.loads(open('$O/cfg5.json').read().strip().splitlines()[-1]);print('cfg5 ms %.3f value %.0f e2e %.0f (%.3f ms)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))" >> $O/summary.txt 2>&1
SPEECHT_B200_EVAL_PIPELINE=0 timeout 600 python bench.py --config 5 --no-cpu-baseline > $O/cfg5_nopipe.json 2> $O/cfg5_nopipe.err
python -c "import json;d=json.loads(open('$O/cfg5_nopipe.json').read().strip().splitlines()[-1]);print('cfg5 no pipeline ms %.3f value %.0f e2e %.0f (%.3f ms)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))" >> $O/summary.txt 2>&1
timeout 600 python bench.py --config 5 --precision bf16 --no-cpu-baseline > $O/cfg5_bf16.json 2> $O/cfg5_bf16.err
python -c "import json;d=json.loads(open('$O/cfg5_bf16.json').read().strip().splitlines()[-1]);print('cfg5 bf16 ms %.3f value %.0f e2e %.0f (%.3f ms)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))" >> $O/summary.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline > $O/cfg2.json 2> $O/cfg2.err
python -c "import json;d=json.loads(open('$O/cfg2.json').read().strip().splitlines()[-1]);print('cfg2 ms %.3f value %.0f e2e %.0f traffic %s'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['traffic']))" >> $O/summary.txt 2>&1
grep -c PASSED $O/t_all.log; tail -5 $O/t_all.log; cat $O/summary.txt
