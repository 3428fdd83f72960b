mkdir -p gpurun_out/r3
timeout 400 python -m pytest tests/test_statevec_gpu.py -m gpu -x -q -k "twelve_stage or many_axes" > gpurun_out/r3/tests3.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r3/tests3.log
for sp in 1 0; do
  echo "== STAGE_PREFETCH=$sp"
  GXSV_STAGE_PREFETCH=$sp MB_QUICK=1 timeout 200 python tools/microbench_batch.py 27 2>&1 | grep "n=27" | tee gpurun_out/r3/mb27_sp$sp.log
done
B="--no-extra --no-per-command --no-e2e --no-cpu --no-parity"
for cfg in "tw0_sp1 0 1" "tw1_sp1 1 1" "tw4_sp1 4 1" "tw8_sp1 8 1" "tw0_sp0 0 0" "tw1_sp0 1 0"; do
  set -- $cfg
  GXSV_TILE_WARPS=$2 GXSV_STAGE_PREFETCH=$3 timeout 200 python bench.py $B --steps 3 --warmup 3 > gpurun_out/r3/c2_$1.json 2>/dev/null
  GXSV_TILE_WARPS=$2 GXSV_STAGE_PREFETCH=$3 timeout 300 python bench.py $B --workload config3_qaoa29 --steps 2 --warmup 1 > gpurun_out/r3/c3_$1.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3/c?_tw*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d.get('value')), round(d.get('ms_per_step'),2), round(r['frac'],3), round(r['avg_launch_ms'],4), d.get('projections_per_pass'))
    except Exception as e: print(f, 'ERR', e)
PY
