#!/bin/bash
# Multi-GPU session (gpurun --gpus N): exchange tests on 2 GPUs, then the bench at 1, 2, 4, ... N GPUs on the same box.
TAG=${1:-r02_scale}
N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== peer tests" | tee $OUT/pytest_peer.log
timeout 900 python -m pytest tests/test_gpu_peer.py -x -q -m gpu 2>&1 | tail -10 | tee -a $OUT/pytest_peer.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline 2>> $OUT/bench.err | tee -a $OUT/bench.log
for G in 2 4 8; do
  if [ $G -le $N ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2961$G \
        bench.py --gpus $G --steps 20 --warmup 3 --no-cpu-baseline 2>> $OUT/bench.err | tee -a $OUT/bench.log
  fi
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --nccl-gather 2>> $OUT/bench.err | tee -a $OUT/bench_nccl.log
tail -20 $OUT/bench.err | grep -v "OMP_NUM_THREADS\|^\*\|^$"
