#!/bin/bash
# perf bisect of the ANYmal constraint-contact configuration over old commits (worktrees under exp/)
TAG=${1:-r02_bisect}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
ARGS="--no-cpu-baseline --workload anymal --contact-model constraint --steps 4 --warmup 2"
for sha in ${SHAS:-af7f5a1 c1168b2 504c33c 0a1f8b2 0aa54a8}; do
  echo "== $sha" | tee -a $OUT/bisect.log
  (cd exp/wt_$sha && timeout 600 python bench.py $ARGS 2>> ../../$OUT/bench.err) | tee -a $OUT/bisect.log
done
tail -3 $OUT/bench.err
