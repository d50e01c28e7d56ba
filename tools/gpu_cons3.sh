#!/bin/bash
# constraint-path check: parity tests that touch the constraint solvers, then the ANYmal / Atlas constraint workloads
TAG=${1:-r02_cons3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q -k "constraint or bounds or contact or foot or hysteresis or pipeline or long_horizon" 2>&1 | tail -5 | tee $OUT/pytest.log
B="timeout 600 python bench.py --no-cpu-baseline --steps 6 --warmup 3"
$B --workload anymal --contact-model constraint 2>> $OUT/bench.err | tee -a $OUT/bench_anymal_constraint.log
JB_NO_UNIFORM_SOLVER=1 $B --workload anymal --contact-model constraint 2>> $OUT/bench.err | tee -a $OUT/bench_anymal_constraint_nouni.log
$B 2>> $OUT/bench.err | tee -a $OUT/bench.log
tail -3 $OUT/bench.err
