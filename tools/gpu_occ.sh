#!/bin/bash
# step time against the number of resident warps per SM (8 ANYmal envs per warp, 148 SMs): separates a per-warp latency
# bound (flat) from a shared per-SM resource such as instruction fetch (grows with warps per SM)
TAG=${1:-r02_occ}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for N in 1184 2368 3552 4096; do
  timeout 600 python bench.py --no-cpu-baseline --workload anymal --contact-model constraint --n-env $N --steps 5 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_constraint.log
done
for N in 1184 2368 4096; do
  timeout 600 python bench.py --no-cpu-baseline --workload anymal --n-env $N --steps 10 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_spring.log
done
tail -3 $OUT/bench.err
