#!/bin/bash
TAG=${1:-r02_head4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="timeout 600 python bench.py --no-cpu-baseline --steps 30 --warmup 3"
for rep in 1 2; do
  for M in 1 2 3; do echo "== mode $M" | tee -a $OUT/modes.log; JB_FAST_BOUNDS_MODE=$M $B 2>> $OUT/bench.err | tee -a $OUT/modes.log; done
  echo "== off" | tee -a $OUT/modes.log; JB_NO_FAST_BOUNDS=1 $B 2>> $OUT/bench.err | tee -a $OUT/modes.log
done
tail -3 $OUT/bench.err
