#!/bin/bash
TAG=${1:-r02_head2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="timeout 600 python bench.py --no-cpu-baseline"
for rep in 1 2; do
  $B --steps 30 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench.log
  JB_NO_FAST_BOUNDS=1 $B --steps 30 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_nofastbounds.log
done
$B --steps 10 --warmup 3 --flagged-fraction 0.1 2>> $OUT/bench.err | tee -a $OUT/bench_flagged.log
tail -3 $OUT/bench.err
