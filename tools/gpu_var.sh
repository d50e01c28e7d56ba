#!/bin/bash
TAG=${1:-r02_var}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
ARGS="--no-cpu-baseline --workload anymal --contact-model constraint --steps 4 --warmup 2"
for v in head v1 v2; do
  echo "== $v" | tee -a $OUT/var.log
  if [ $v = head ]; then timeout 600 python bench.py $ARGS 2>> $OUT/bench.err | tee -a $OUT/var.log
  else JB_LIBRARY=$PWD/exp/lib_$v.so timeout 600 python bench.py $ARGS 2>> $OUT/bench.err | tee -a $OUT/var.log; fi
done
echo "== v2 flagged 0.01 / 0.1" | tee -a $OUT/var.log
JB_LIBRARY=$PWD/exp/lib_v2.so timeout 600 python bench.py --no-cpu-baseline --flagged-fraction 0.01 --steps 5 --warmup 3 2>> $OUT/bench.err | tee -a $OUT/var.log
tail -3 $OUT/bench.err
