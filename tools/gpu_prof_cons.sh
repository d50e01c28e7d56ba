#!/bin/bash
# ncu --set full capture of one step of the ANYmal `constraint` contact-model workload (one report per call: 64 MiB cap)
TAG=${1:-r02_prof_cons}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --no-cpu-baseline --workload anymal --contact-model constraint --steps 4 --warmup 2 2>> $OUT/bench.err | tee -a $OUT/bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 2 -c 1 -f -o $OUT/prof_cons \
    python bench.py --no-cpu-baseline --workload anymal --contact-model constraint --steps 2 --warmup 1 > $OUT/ncu_full_run.log 2>&1
tail -3 $OUT/bench.err
ls -la $OUT
