OUT=gpurun_out/${1:-r01y}; mkdir -p $OUT
A="--workload atlas --contact-model constraint --ode-solver euler_explicit --dt-max 0.005 --steps 2 --warmup 1 --no-cpu-baseline"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 3 -c 1 -f -o $OUT/prof_atlas_bodies python bench.py $A > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
