#!/bin/bash
# Round-2 final single-GPU session: parity suite, every bench line quoted in DESIGN.md / README.md, launch list + one full ncu capture.
TAG=${1:-r02_final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu" | tee $OUT/pytest_gpu.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee -a $OUT/pytest_gpu.log
B="timeout 600 python bench.py"
$B --gpus 1 --steps 20 --warmup 3 2> $OUT/bench.err | tee $OUT/bench.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>> $OUT/bench.err | tee $OUT/bench_reference_arm.log
N="--no-cpu-baseline"
JB_QUADRUPED_ABA=1 $B $N --steps 20 --warmup 3 2>> $OUT/bench.err | tee $OUT/bench_aba.log
$B $N --workload cartpole --n-env 512 --steps 50 --warmup 5 2>> $OUT/bench.err | tee $OUT/bench_cartpole512.log
$B $N --workload double_pendulum --n-env 1 --steps 50 --warmup 5 2>> $OUT/bench.err | tee $OUT/bench_double_pendulum1.log
for F in 0.01 0.1 0.5; do $B $N --steps 10 --warmup 3 --flagged-fraction $F 2>> $OUT/bench.err | tee -a $OUT/bench_flagged.log; done
$B $N --steps 3 --warmup 3 --action torque --ode-solver euler_explicit --dt-max 1e-4 2>> $OUT/bench.err | tee -a $OUT/bench_torque.log
$B $N --steps 5 --warmup 3 --action torque --ode-solver runge_kutta_4 --dt-max 2.5e-4 2>> $OUT/bench.err | tee -a $OUT/bench_torque.log
$B $N --steps 3 --warmup 3 --action torque --ode-solver runge_kutta_dopri 2>> $OUT/bench.err | tee -a $OUT/bench_torque.log
$B $N --steps 3 --warmup 3 --action torque --contact-model constraint 2>> $OUT/bench.err | tee -a $OUT/bench_torque.log
$B $N --workload anymal --contact-model constraint --steps 5 --warmup 3 2>> $OUT/bench.err | tee $OUT/bench_anymal_constraint.log
$B $N --workload atlas --steps 5 --warmup 3 2>> $OUT/bench.err | tee $OUT/bench_atlas4096.log
$B $N --workload atlas --contact-model constraint --ode-solver euler_explicit --dt-max 0.005 --steps 5 --warmup 3 2>> $OUT/bench.err | tee $OUT/bench_atlas_reference_settings.log
timeout 300 python tools/bench_pipeline.py --n-env 4096 --steps 10 --warmup 3 2>> $OUT/bench.err | tee $OUT/bench_atlas_pd_pipeline.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_run.log 2>&1
echo "== ncu full capture"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 3 -c 1 -f -o $OUT/prof_step \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_run.log 2>&1
tail -5 $OUT/bench.err
ls -la $OUT
