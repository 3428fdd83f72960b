mkdir -p gpurun_out/r3
timeout 500 python -m pytest tests/test_statevec_gpu.py -m gpu -x -q -k "twelve_stage or many_axes" > gpurun_out/r3/tests7.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r3/tests7.log
timeout 500 python -m pytest tests/test_patterns_gpu.py tests/test_full_size_gpu.py -m gpu -x -q > gpurun_out/r3/tests7b.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3/tests7b.log
MB_QUICK=1 timeout 300 python tools/microbench_batch.py 27 2>&1 | grep "n=27" | tee gpurun_out/r3/mb27_io.log
B="--no-extra --no-per-command --no-e2e --no-cpu --no-parity"
for io in 0 1 2; do
  GXSV_TILE_IO=$io timeout 200 python bench.py $B --steps 3 --warmup 3 > gpurun_out/r3/c2_io$io.json 2>/dev/null
  GXSV_TILE_IO=$io timeout 300 python bench.py $B --workload config3_qaoa29 --steps 2 --warmup 1 > gpurun_out/r3/c3_io$io.json 2>/dev/null
  GXSV_TILE_IO=$io timeout 600 python bench.py $B --workload config4_qft32 --max-commands 900 --steps 2 --warmup 1 > gpurun_out/r3/c4_io$io.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3/c?_io*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d.get('value'),1), round(d.get('ms_per_step'),2), round(r['frac'],3), round(r['avg_launch_ms'],4), d.get('projections_per_pass'))
    except Exception as e: print(f, 'ERR', e)
PY
