cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/final; mkdir -p $O
for c in 3 4; do
timeout 900 python bench.py --config $c > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err
python -c "import json;d=json.loads(open('$O/bench_cfg$c.json').read().strip().splitlines()[-1]);print('cfg$c ms %.3f value %.0f e2e %.0f (%.3f ms) sustained %.3f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['sustained']['ms_per_step']))"
done
python - <<'P'
import torch, time
x = torch.empty(64*1024*1024, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device='cuda')
for _ in range(3):
  torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); print('H2D 256 MB pinned: %.1f GB/s' % (x.numel()*4/1e9/(time.perf_counter()-t)))
P
