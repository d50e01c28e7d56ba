#!/bin/bash
TAG=${1:-r02_regime2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="timeout 600 python bench.py --no-cpu-baseline"
$B --steps 30 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench.log
JB_NO_FAST_BOUNDS=1 $B --steps 30 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_nofastbounds.log
for F in 0.01 0.1 0.5; do
  $B --steps 10 --warmup 3 --flagged-fraction $F 2>> $OUT/bench.err | tee -a $OUT/bench_flagged.log
done
$B --steps 3 --warmup 3 --action torque --ode-solver euler_explicit --dt-max 1e-4 2>> $OUT/bench.err | tee -a $OUT/bench_torque.log
$B --steps 5 --warmup 3 --action torque --ode-solver runge_kutta_4 --dt-max 2.5e-4 2>> $OUT/bench.err | tee -a $OUT/bench_torque.log
echo "== pytest -m gpu" | tee $OUT/pytest_gpu.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/bench.err
