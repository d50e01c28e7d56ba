#!/bin/bash
# hot-path A/B after a device-code change: parity tests that exercise the hot path, then the headline bench three times
TAG=${1:-r02_ab_hot}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q -k "anymal or long_horizon or rhs or bounds or handoff or analytic or energy" 2>&1 | tail -4 | tee $OUT/pytest.log
for i in 1 2 3; do timeout 600 python bench.py --no-cpu-baseline --steps 30 --warmup 5 2>> $OUT/bench.err | tee -a $OUT/bench.log; done
tail -3 $OUT/bench.err
