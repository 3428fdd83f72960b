mkdir -p gpurun_out/r3
B="--no-extra --no-per-command --no-e2e --no-cpu --no-parity"
run() { # name env...
  name=$1; shift
  env "$@" timeout 200 python bench.py $B --steps 3 --warmup 3 > gpurun_out/r3/c2_$name.json 2>/dev/null
  env "$@" timeout 300 python bench.py $B --workload config3_qaoa29 --steps 2 --warmup 1 > gpurun_out/r3/c3_$name.json 2>/dev/null
  env "$@" timeout 600 python bench.py $B --workload config4_qft32 --max-commands 900 --steps 2 --warmup 1 > gpurun_out/r3/c4_$name.json 2>/dev/null
}
run lc0 GXSV_TILE_LD_CACHED=0
run lc1 GXSV_TILE_LD_CACHED=1
run lc2 GXSV_TILE_LD_CACHED=2
run lc64 GXSV_TILE_LD_CACHED=64
run lc1_pf GXSV_TILE_LD_CACHED=1 GXSV_TILE_PREFETCH=1
run lc0_pf GXSV_TILE_LD_CACHED=0 GXSV_TILE_PREFETCH=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3/c?_lc*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d.get('value'),1), round(d.get('ms_per_step'),2), round(r['frac'],3), round(r['avg_launch_ms'],4), d.get('projections_per_pass'))
    except Exception as e: print(f, 'ERR', e)
PY
