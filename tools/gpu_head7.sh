#!/bin/bash
TAG=${1:-r02_head7}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="timeout 600 python bench.py --no-cpu-baseline"
echo "== anymal constraint uniform" | tee -a $OUT/modes.log; $B --workload anymal --contact-model constraint --steps 5 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/modes.log
echo "== anymal constraint non-uniform" | tee -a $OUT/modes.log; JB_NO_UNIFORM_SOLVER=1 $B --workload anymal --contact-model constraint --steps 5 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/modes.log
echo "== torque rk4 constraint contacts" | tee -a $OUT/modes.log; $B --steps 3 --warmup 3 --action torque --contact-model constraint 2>> $OUT/bench.err | tee -a $OUT/modes.log
echo "== headline" | tee -a $OUT/modes.log; $B --steps 30 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/modes.log
echo "== pytest -m gpu" | tee $OUT/pytest_gpu.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/bench.err
