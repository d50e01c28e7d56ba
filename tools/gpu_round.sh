#!/bin/bash
# full GPU suite + headline bench + secondary workloads (no ncu): the check after a device-code change
TAG=${1:-r02_round}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu.log
B="timeout 600 python bench.py --no-cpu-baseline"
$B --steps 20 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench.log
$B --steps 20 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench.log
$B --workload anymal --contact-model constraint --steps 4 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_anymal_constraint.log
$B --flagged-fraction 0.1 --steps 5 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_flagged.log
tail -3 $OUT/bench.err
