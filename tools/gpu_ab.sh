#!/bin/bash
# A/B of the two hot-path evaluations on one box: composite-rigid-body form (default) vs ABA sweeps (JB_QUADRUPED_ABA=1)
TAG=${1:-r02_ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for rep in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>> $OUT/bench.err | tee -a $OUT/bench_crba.log
  JB_QUADRUPED_ABA=1 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>> $OUT/bench.err | tee -a $OUT/bench_aba.log
done
echo "== pytest -m gpu" | tee $OUT/pytest_gpu.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee -a $OUT/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 3 -c 1 -f -o $OUT/prof_step \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_run.log 2>&1
tail -3 $OUT/bench.err
