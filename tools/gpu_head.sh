#!/bin/bash
TAG=${1:-r02_head}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="timeout 600 python bench.py --no-cpu-baseline"
for rep in 1 2; do
  $B --steps 30 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench.log
  JB_NO_FAST_BOUNDS=1 $B --steps 30 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_nofastbounds.log
done
$B --steps 10 --warmup 3 --flagged-fraction 0.1 2>> $OUT/bench.err | tee -a $OUT/bench_flagged.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 3 -c 1 -f -o $OUT/prof_step \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_run.log 2>&1
tail -3 $OUT/bench.err
