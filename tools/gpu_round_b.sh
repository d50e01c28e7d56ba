#!/bin/bash
# full GPU suite + the workloads that run the constraint solvers (no ncu)
TAG=${1:-r02_round_b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu.log
B="timeout 600 python bench.py --no-cpu-baseline"
$B --steps 20 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench.log
$B --workload anymal --contact-model constraint --steps 5 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_anymal_constraint.log
$B --workload atlas --contact-model constraint --ode-solver euler_explicit --dt-max 0.005 --steps 5 --warmup 3 2>> $OUT/bench.err | tee $OUT/bench_atlas_reference_settings.log
timeout 300 python tools/bench_pipeline.py --n-env 4096 --steps 10 --warmup 3 2>> $OUT/bench.err | tee $OUT/bench_atlas_pd_pipeline.log
for F in 0.01 0.1 0.5; do $B --steps 10 --warmup 3 --flagged-fraction $F 2>> $OUT/bench.err | tee -a $OUT/bench_flagged.log; done
$B --steps 3 --warmup 3 --action torque --ode-solver euler_explicit --dt-max 1e-4 2>> $OUT/bench.err | tee -a $OUT/bench_torque.log
$B --steps 5 --warmup 3 --action torque --ode-solver runge_kutta_4 --dt-max 2.5e-4 2>> $OUT/bench.err | tee -a $OUT/bench_torque.log
$B --steps 3 --warmup 3 --action torque --ode-solver runge_kutta_dopri 2>> $OUT/bench.err | tee -a $OUT/bench_torque.log
$B --steps 3 --warmup 3 --action torque --contact-model constraint 2>> $OUT/bench.err | tee -a $OUT/bench_torque.log
tail -3 $OUT/bench.err
