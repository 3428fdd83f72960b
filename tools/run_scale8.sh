# Round-1 scaling evidence on one 8xB200 box: python -m pytest (8-rank golden) + bench.py at N = 8, 4, 2, 1
# + BASELINE config 4 anchor points (33 live qubits on 1 GPU, 36 live qubits on 8 GPUs; first 900 commands).
set -x
python -m pytest tests/test_sharded_gpu.py -m gpu -x -q -k "golden[8] or nonstrict[4]" 2>&1 | tail -3
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 3 --warmup 3 2>gpurun_out/scale_n$n.err | tail -1 > gpurun_out/scale_n$n.json
done
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu 2>gpurun_out/scale_n1.err | tail -1 > gpurun_out/scale_n1.json
python bench.py --workload config4_qft32 --max-commands 900 --steps 2 --warmup 1 --no-cpu --no-per-command 2>gpurun_out/qft32_n1.err | tail -1 > gpurun_out/qft32_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --workload config4_qft35 --max-commands 900 --steps 2 --warmup 1 2>gpurun_out/qft35_n8.err | tail -1 > gpurun_out/qft35_n8.json
tail -c 200 gpurun_out/*.err | grep -v "OMP_NUM\|\*\*\*\*" 
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/scale_n*.json')+glob.glob('gpurun_out/qft3*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), round(d['ms_per_step'],1), d['config']['workload'], d.get('exchange'), round(d['roofline']['kernel_time_share_of_step'],3))
    except Exception as e: print(f,'ERR',e)
PY
