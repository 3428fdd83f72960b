set -x
mkdir -p gpurun_out/r3
timeout 600 python -m pytest tests/test_statevec_gpu.py tests/test_patterns_gpu.py -m gpu -x -q > gpurun_out/r3/tests1.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r3/tests1.log
timeout 200 python tools/microbench_batch.py 27 > gpurun_out/r3/mb27.log 2>&1; tail -22 gpurun_out/r3/mb27.log
B="--steps 3 --warmup 3 --no-extra --no-per-command --no-e2e --no-cpu --no-parity"
timeout 200 python bench.py $B > gpurun_out/r3/b_c2_new.json 2> gpurun_out/r3/b_c2_new.err
GXSV_TILE_PREFETCH=0 timeout 200 python bench.py $B > gpurun_out/r3/b_c2_nopf.json 2>/dev/null
GXSV_DEFER_VCZ=0 timeout 200 python bench.py $B > gpurun_out/r3/b_c2_nodefer.json 2>/dev/null
timeout 300 python bench.py $B --workload config3_qaoa29 --steps 2 --warmup 1 > gpurun_out/r3/b_c3_new.json 2>gpurun_out/r3/b_c3_new.err
GXSV_TILE_PREFETCH=0 timeout 300 python bench.py $B --workload config3_qaoa29 --steps 2 --warmup 1 > gpurun_out/r3/b_c3_nopf.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3/b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('ms_per_step'), d.get('roofline'))
    except Exception as e: print(f, 'ERR', e)
PY
