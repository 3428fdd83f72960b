mkdir -p gpurun_out/r3
timeout 300 python -m pytest tests/test_statevec_gpu.py -m gpu -x -q -k "twelve_stage or many_axes" > gpurun_out/r3/tests11.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r3/tests11.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
B="--no-extra --no-per-command --no-e2e --no-cpu --no-parity"
timeout 200 python bench.py $B --steps 3 --warmup 3 > gpurun_out/r3/c2_final.json 2>/dev/null
timeout 300 python bench.py $B --workload config3_qaoa29 --steps 2 --warmup 1 > gpurun_out/r3/c3_final.json 2>/dev/null
timeout 600 python bench.py $B --workload config4_qft32 --max-commands 900 --steps 2 --warmup 1 > gpurun_out/r3/c4_final.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3/c?_final.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d.get('value'),1), round(d.get('ms_per_step'),2), round(r['frac'],3), round(r['avg_launch_ms'],4), d.get('projections_per_pass'))
    except Exception as e: print(f, 'ERR', e)
PY
