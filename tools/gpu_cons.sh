#!/bin/bash
# constraint-path investigation: ANYmal with contacts.model = constraint, solver toggles, one ncu capture of the full kernel
TAG=${1:-r02_cons}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="timeout 600 python bench.py --no-cpu-baseline --workload anymal --contact-model constraint --steps 4 --warmup 2"
$B 2>> $OUT/bench.err | tee -a $OUT/bench_default.log
JB_NO_BLOCK_CONS=1 $B 2>> $OUT/bench.err | tee -a $OUT/bench_noblock.log
JB_NO_STRUCTURED_CONS=1 $B --n-env 512 2>> $OUT/bench.err | tee -a $OUT/bench_nostructured512.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 2 -c 1 -f -o $OUT/prof_cons \
    python bench.py --no-cpu-baseline --workload anymal --contact-model constraint --steps 2 --warmup 1 > $OUT/ncu_full_run.log 2>&1
# 1 % of the envs through their joint bounds (spring-damper contacts): what the slow warps spend their time on
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 3 -c 1 -f -o $OUT/prof_flagged \
    python bench.py --no-cpu-baseline --flagged-fraction 0.01 --steps 2 --warmup 2 > $OUT/ncu_flagged_run.log 2>&1
tail -3 $OUT/bench.err
