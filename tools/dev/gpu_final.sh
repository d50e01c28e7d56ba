OUT=gpurun_out/${1:-r01z}; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
timeout 200 python bench.py --workload atlas --contact-model constraint --ode-solver euler_explicit --dt-max 0.005 --steps 10 --warmup 3 > $OUT/bench_atlas_reference_settings.json 2> $OUT/err1.log
timeout 200 python bench.py --workload atlas --contact-model constraint --n-env 512 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_atlas_constraint512.json 2> $OUT/err2.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_anymal4096.json 2> $OUT/err3.log
python - $OUT <<'PY'
import json,glob,sys
for f in sorted(glob.glob(sys.argv[1]+'/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
