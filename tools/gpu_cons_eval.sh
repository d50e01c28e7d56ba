#!/bin/bash
# constraint-path evaluation of a solver change: parity tests, cycle accounting (development build), bench lines
TAG=${1:-r02_cons_eval}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q -k "constraint or bounds or contact or foot or hysteresis or pipeline or long_horizon" 2>&1 | tail -4 | tee $OUT/pytest.log
timeout 600 python tools/prof_clocks.py anymal 4096 3 2>&1 | tee $OUT/clocks_4096.txt
B="timeout 600 python bench.py --no-cpu-baseline --steps 6 --warmup 3"
$B --workload anymal --contact-model constraint 2>> $OUT/bench.err | tee -a $OUT/bench_anymal_constraint.log
$B 2>> $OUT/bench.err | tee -a $OUT/bench.log
tail -3 $OUT/bench.err
