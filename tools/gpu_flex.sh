#!/bin/bash
# Flexibility joints on the device + a quick look at the headline (no CPU baseline): the last GPU minutes of round 2.
OUT=gpurun_out/flex
mkdir -p $OUT
timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest.log
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> $OUT/bench.err | tee $OUT/bench.json
tail -3 $OUT/bench.err
