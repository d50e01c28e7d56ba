#!/bin/bash
# Quick perf iteration: bench (no CPU baseline) + one full ncu capture of the step kernel.
TAG=${1:-q}
WL=${2:-anymal}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline 2> $OUT/bench.err | tee $OUT/bench.log
tail -3 $OUT/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 3 -c 1 -f -o $OUT/prof_step \
    python bench.py --workload $WL --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_run.log 2>&1
ls -la $OUT | tail -4
