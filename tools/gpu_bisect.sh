#!/bin/bash
# perf bisect of the ANYmal constraint-contact configuration over old commits (worktrees under exp/), plus the GPU suite at HEAD
TAG=${1:-r02_bisect}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
ARGS="--no-cpu-baseline --workload anymal --contact-model constraint --steps 4 --warmup 2"
for sha in af7f5a1 0aa54a8 56f875e c822f72 08d2009; do
  echo "== $sha" | tee -a $OUT/bisect.log
  (cd exp/wt_$sha && timeout 600 python bench.py $ARGS 2>> ../../$OUT/bench.err) | tee -a $OUT/bisect.log
done
echo "== HEAD" | tee -a $OUT/bisect.log
timeout 600 python bench.py $ARGS 2>> $OUT/bench.err | tee -a $OUT/bisect.log
echo "== pytest -m gpu" | tee $OUT/pytest_gpu.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/bench.err
