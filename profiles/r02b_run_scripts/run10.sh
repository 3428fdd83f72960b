mkdir -p gpurun_out/r3
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/r3/final_sharded_tests_n2.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r3/final_sharded_tests_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra > gpurun_out/r3/bench_final_n2_noextra.json 2> gpurun_out/r3/bench_final_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_final_n2_noextra.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d.get('e2e',{}).get('value') if d.get('e2e') else None,'frac',d['roofline']['frac'],'parity',d.get('sharded_parity'),'exchange',d.get('exchange'))
PY
