mkdir -p gpurun_out/r3
timeout 1000 python -m pytest tests -m gpu -x -q > gpurun_out/r3/final_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r3/final_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r3/bench_final_n1.json 2> gpurun_out/r3/bench_final_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r3/bench_final_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_final_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'] if d.get('e2e') else None,'frac',d['roofline']['frac'])
for k in ('strict_lazy','random_selector','per_command','config3','config4','config5'):
    v=d.get(k)
    if isinstance(v,dict): print(k, {kk:(vv if not isinstance(vv,dict) else {a:b for a,b in vv.items() if a in ('value','frac','ok')}) for kk,vv in v.items() if kk in ('value','batched','strict_lazy','roofline','parity_check','window','full_pattern','e2e')})
PY
B="--no-extra --no-per-command --no-e2e --no-cpu --no-parity"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3/launches_final.csv python bench.py --steps 1 --warmup 1 $B > gpurun_out/r3/ncu_launches.log 2>&1; tail -1 gpurun_out/r3/ncu_launches.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_tile -s 60 -c 3 -f -o gpurun_out/r3/prof_final_c2 python bench.py $B --steps 1 --warmup 1 > gpurun_out/r3/ncu_final_c2.log 2>&1; tail -1 gpurun_out/r3/ncu_final_c2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_tile -s 30 -c 2 -f -o gpurun_out/r3/prof_final_c3 python bench.py $B --workload config3_qaoa29 --steps 1 --warmup 1 --max-commands 1200 > gpurun_out/r3/ncu_final_c3.log 2>&1; tail -1 gpurun_out/r3/ncu_final_c3.log
