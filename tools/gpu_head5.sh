#!/bin/bash
TAG=${1:-r02_head5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
B="timeout 600 python bench.py --no-cpu-baseline"
for rep in 1 2; do
  echo "== on" | tee -a $OUT/modes.log; $B --steps 30 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/modes.log
  echo "== off" | tee -a $OUT/modes.log; JB_NO_FAST_BOUNDS=1 $B --steps 30 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/modes.log
done
echo "== flagged 0.1" | tee -a $OUT/modes.log; $B --steps 10 --warmup 3 --flagged-fraction 0.1 2>> $OUT/bench.err | tee -a $OUT/modes.log
echo "== anymal constraint" | tee -a $OUT/modes.log; $B --workload anymal --contact-model constraint --steps 5 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/modes.log
echo "== atlas" | tee -a $OUT/modes.log; $B --workload atlas --steps 5 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/modes.log
echo "== atlas ref settings" | tee -a $OUT/modes.log; $B --workload atlas --contact-model constraint --ode-solver euler_explicit --dt-max 0.005 --steps 5 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/modes.log
echo "== torque euler" | tee -a $OUT/modes.log; $B --steps 3 --warmup 3 --action torque --ode-solver euler_explicit --dt-max 1e-4 2>> $OUT/bench.err | tee -a $OUT/modes.log
echo "== pytest -m gpu" | tee $OUT/pytest_gpu.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/bench.err
