mkdir -p gpurun_out/r3
timeout 60 python -m pytest tests/test_statevec_gpu.py -m gpu -x -q -k "twelve_stage and 17-" > gpurun_out/r3/tests13.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r3/tests13.log
B="--no-extra --no-per-command --no-e2e --no-cpu --no-parity"
timeout 60 python bench.py $B --steps 3 --warmup 3 > gpurun_out/r3/c2_trim.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r3/c2_trim.json').read().strip().splitlines()[-1]); r=d['roofline']
print('c2', round(d['value'],1), round(r['frac'],3), round(r['avg_launch_ms'],4))"
timeout 100 python bench.py $B --workload config4_qft32 --max-commands 900 --steps 2 --warmup 1 > gpurun_out/r3/c4_trim.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r3/c4_trim.json').read().strip().splitlines()[-1]); r=d['roofline']
print('c4', round(d['value'],1), round(r['frac'],3), round(r['avg_launch_ms'],4))"
