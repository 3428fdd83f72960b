mkdir -p gpurun_out/r3
timeout 500 python -m pytest tests/test_patterns_gpu.py tests/test_full_size_gpu.py -m gpu -x -q > gpurun_out/r3/tests4.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r3/tests4.log
B="--no-extra --no-per-command --no-e2e --no-cpu --no-parity"
# config 4 anchor at 33 live qubits (128 GiB): 900-command window
timeout 600 python bench.py $B --workload config4_qft32 --max-commands 900 --steps 2 --warmup 1 > gpurun_out/r3/c4_qft32_w900.json 2> gpurun_out/r3/c4.err; tail -2 gpurun_out/r3/c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/c4_qft32_w900.json').read().strip().splitlines()[-1]); r=d['roofline']
print('config4 qft32 window', round(d['value'],1), round(d['ms_per_step'],1), round(r['frac'],3), round(r['avg_launch_ms'],3), d.get('projections_per_pass'), d['config'].get('live_qubits'))
PY
# ncu: full capture of the shared-tile kernel on config 3, then the same with the L2 prefetch on
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_tile -s 30 -c 2 -f -o gpurun_out/r3/prof_c3_tw4 python bench.py $B --workload config3_qaoa29 --steps 1 --warmup 1 --max-commands 1200 > gpurun_out/r3/ncu_c3_tw4.log 2>&1; tail -2 gpurun_out/r3/ncu_c3_tw4.log
GXSV_TILE_PREFETCH=1 timeout 300 ncu --set full --clock-control none -k regex:k_fused_tile -s 30 -c 2 -f -o gpurun_out/r3/prof_c3_tw4_pf python bench.py $B --workload config3_qaoa29 --steps 1 --warmup 1 --max-commands 1200 > gpurun_out/r3/ncu_c3_tw4_pf.log 2>&1; tail -2 gpurun_out/r3/ncu_c3_tw4_pf.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_tile -s 60 -c 2 -f -o gpurun_out/r3/prof_c2_tw4 python bench.py $B --steps 1 --warmup 1 > gpurun_out/r3/ncu_c2_tw4.log 2>&1; tail -2 gpurun_out/r3/ncu_c2_tw4.log
ls -la gpurun_out/r3/*.ncu-rep
