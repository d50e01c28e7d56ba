#!/bin/bash
TAG=${1:-r02_clocks}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python tools/prof_clocks.py anymal 4096 3 2>&1 | tee $OUT/clocks_4096.txt
