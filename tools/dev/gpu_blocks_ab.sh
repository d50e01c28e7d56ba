OUT=gpurun_out/r01v; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
A="--workload atlas --contact-model constraint --ode-solver euler_explicit --dt-max 0.005 --steps 5 --warmup 3 --no-cpu-baseline"
timeout 200 python bench.py $A > $OUT/atlas_ref_block.json 2> $OUT/err1.log
JB_NO_BLOCK_CONS=1 timeout 200 python bench.py $A > $OUT/atlas_ref_generic.json 2> $OUT/err2.log
B="--workload atlas --contact-model constraint --n-env 512 --steps 3 --warmup 3 --no-cpu-baseline"
timeout 200 python bench.py $B > $OUT/atlas512_block.json 2> $OUT/err3.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r01v/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
