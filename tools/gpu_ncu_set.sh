#!/bin/bash
# `ncu --set full` captures of single launches of the default bench step, exported as raw metric pages (CSV):
#   tools/gpu_ncu_set.sh <tag> name:kernel_regex:skip[:extra bench args] ...
cd "${GRAFT_REPO_ROOT:-.}"
TAG=$1; shift
O=gpurun_out/$TAG
mkdir -p $O
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sustained"
for spec in "$@"; do
  IFS=: read name rx skip extra <<< "$spec"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip --launch-count 1 \
    -o $O/ncu_$name -f $B $extra > $O/ncu_$name.log 2>&1
  rc=$?
  ncu -i $O/ncu_$name.ncu-rep --page raw --csv > $O/ncu_$name.raw.csv 2>/dev/null
  ncu -i $O/ncu_$name.ncu-rep --page details --csv > $O/ncu_$name.details.csv 2>/dev/null
  rm -f $O/ncu_$name.ncu-rep
  echo "ncu $name rc=$rc" >> $O/summary.txt
done
cat $O/summary.txt
