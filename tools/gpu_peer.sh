#!/bin/bash
# Multi-GPU session (gpurun --gpus N): 2-GPU exchange tests, then the bench at 1..N GPUs on the same box.
TAG=${1:-r02_peer}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.csv 2>&1
echo "== peer tests" | tee $OUT/pytest_peer.log
timeout 900 python -m pytest tests/test_gpu_peer.py -x -q -m gpu 2>&1 | tail -30 | tee -a $OUT/pytest_peer.log
echo "== bench 1 GPU" | tee $OUT/bench.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline 2>> $OUT/bench.err | tee -a $OUT/bench.log
for G in 2 4 8; do
  if [ $G -le $N ]; then
    echo "== bench $G GPUs" | tee -a $OUT/bench.log
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611 \
        bench.py --gpus $G --steps 20 --warmup 3 --no-cpu-baseline 2>> $OUT/bench.err | tee -a $OUT/bench.log
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29612 \
        bench.py --gpus $G --steps 20 --warmup 3 --no-cpu-baseline --nccl-gather 2>> $OUT/bench.err | tee -a $OUT/bench_nccl.log
  fi
done
tail -20 $OUT/bench.err
