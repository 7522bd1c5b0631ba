cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/s25; mkdir -p $O
for v in 1 0; do
SPEECHT_B200_STEP_TRACE=1 SPEECHT_B200_EVAL_PIPELINE=$v timeout 600 python bench.py --config 5 --no-cpu-baseline --no-sustained --steps 6 --warmup 3 > $O/cfg5_$v.json 2> $O/cfg5_$v.err
echo "pipeline=$v"; tail -9 $O/cfg5_$v.err
done
