# A/B of the constraint solvers on Atlas with the reference's own settings (Euler 5 ms, constraint contacts)
OUT=gpurun_out/${1:-r01x}; mkdir -p $OUT
# (the store->load probe of tools/micro/store_load_latency.cu ran here once: profiles/r01_micro_store_load_latency_b200.txt)
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
A="--workload atlas --contact-model constraint --ode-solver euler_explicit --dt-max 0.005 --steps 5 --warmup 3 --no-cpu-baseline"
timeout 200 python bench.py $A > $OUT/atlas_ref_bodies.json 2> $OUT/err1.log
B="--workload atlas --contact-model constraint --n-env 512 --steps 3 --warmup 3 --no-cpu-baseline"
timeout 200 python bench.py $B > $OUT/atlas512_bodies.json 2> $OUT/err3.log
JB_NO_STRUCTURED_CONS=1 timeout 200 python bench.py --workload anymal --contact-model constraint --steps 5 --warmup 3 --no-cpu-baseline > $OUT/anymal_cons_bodies.json 2> $OUT/err4.log
python - $OUT <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
