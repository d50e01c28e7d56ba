#!/bin/bash
TAG=${1:-r02_cons2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="timeout 600 python bench.py --no-cpu-baseline --workload anymal --contact-model constraint --steps 4 --warmup 2"
$B 2>> $OUT/bench.err | tee -a $OUT/bench_default.log
JB_CW_SLOTS=32 $B 2>> $OUT/bench.err | tee -a $OUT/bench_slots32.log
JB_CW_SLOTS=32 timeout 600 python bench.py --no-cpu-baseline --flagged-fraction 0.01 --steps 5 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_flagged_slots32.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 2 -c 1 -f -o $OUT/prof_cons \
    python bench.py --no-cpu-baseline --workload anymal --contact-model constraint --steps 2 --warmup 1 > $OUT/ncu_full_run.log 2>&1
tail -3 $OUT/bench.err
