#!/bin/bash
TAG=${1:-r02_ab2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="timeout 600 python bench.py --no-cpu-baseline"
for rep in 1 2; do
  $B --steps 30 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_v2.log
  JB_LIBRARY=$PWD/exp/lib_v1.so $B --steps 30 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_v1.log
done
# the regime a random policy produces: envs leaving the hot path (joint bounds -> full body with constraints)
for F in 0.01 0.1 0.5; do
  $B --steps 10 --warmup 3 --flagged-fraction $F 2>> $OUT/bench.err | tee -a $OUT/bench_flagged.log
done
# SURVEY 8(d) config 3 as written: raw torque actions, with the steppers that survive them on the oracle
$B --steps 3 --warmup 3 --action torque --ode-solver euler_explicit --dt-max 1e-4 2>> $OUT/bench.err | tee -a $OUT/bench_torque.log
$B --steps 3 --warmup 3 --action torque --ode-solver runge_kutta_dopri 2>> $OUT/bench.err | tee -a $OUT/bench_torque.log
$B --steps 3 --warmup 3 --action torque --contact-model constraint 2>> $OUT/bench.err | tee -a $OUT/bench_torque.log
# secondary configs at HEAD
$B --workload atlas --steps 5 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_atlas.log
$B --workload anymal --contact-model constraint --steps 5 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_anymal_constraint.log
$B --workload atlas --contact-model constraint --ode-solver euler_explicit --dt-max 0.005 --steps 5 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/bench_atlas_reference_settings.log
tail -5 $OUT/bench.err
