#!/bin/bash
# One-GPU check: parity tests, headline bench, the two small configs, launch list + one full ncu capture.
TAG=${1:-r02_check}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  echo "== pytest -m gpu" | tee $OUT/pytest_gpu.log
  timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 | tee -a $OUT/pytest_gpu.log
fi
echo "== bench" | tee $OUT/bench.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 ${BENCH_FLAGS:---no-cpu-baseline} 2> $OUT/bench.err | tee -a $OUT/bench.log
timeout 300 python bench.py --workload cartpole --n-env 512 --steps 50 --warmup 5 --no-cpu-baseline 2>> $OUT/bench.err | tee $OUT/bench_cartpole512.log
timeout 300 python bench.py --workload double_pendulum --n-env 1 --steps 50 --warmup 5 --no-cpu-baseline 2>> $OUT/bench.err | tee $OUT/bench_double_pendulum1.log
tail -5 $OUT/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_run.log 2>&1
  echo "== ncu full capture of env_step_kernel (launch 0 = start; the steps follow, one kernel each)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 3 -c 1 -f -o $OUT/prof_step \
      python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_run.log 2>&1
fi
ls -la $OUT
