set -x
mkdir -p gpurun_out/r3
B="--no-extra --no-per-command --no-e2e --no-cpu --no-parity"
summ() { python - "$1" <<'PY'
import sys,re,collections
c=collections.Counter(); bits=collections.Counter(); tot=0; st=0; uni=0
for l in open(sys.argv[1]):
    m=re.match(r'gxsv-pass n=(\d+) axes=(\d+) stages=(\d+) phases=(\d+) uniform=(\d+) bits=(.*)',l)
    if not m: continue
    n,a,s,p,u=map(int,m.groups()[:5]); c[(a,s,p)]+=1; tot+=1; st+=s; uni+=u
    F=9-a
    lowrun=min(int(b) for b in m.group(6).split(','))
    bits[(a,min(lowrun,12))]+=1
print(sys.argv[1],'passes',tot,'stages',st,'uniform',uni)
print(' (axes,stages,phases):',sorted(c.items(),key=lambda kv:-kv[1])[:12])
print(' (axes,lowest axis bit capped 12):',sorted(bits.items(),key=lambda kv:-kv[1])[:10])
PY
}
GXSV_BATCH_LOG=1 timeout 200 python bench.py $B --steps 1 --warmup 1 > /dev/null 2> /tmp/shape_c2.err; summ /tmp/shape_c2.err > gpurun_out/r3/shape_c2.txt
GXSV_BATCH_LOG=1 timeout 300 python bench.py $B --workload config3_qaoa29 --steps 1 --warmup 1 --max-commands 3000 > /dev/null 2> /tmp/shape_c3.err; summ /tmp/shape_c3.err > gpurun_out/r3/shape_c3.txt
GXSV_BATCH_LOG=1 timeout 300 python bench.py $B --workload config4_qft29 --steps 1 --warmup 1 --max-commands 3000 > /dev/null 2> /tmp/shape_c4.err; summ /tmp/shape_c4.err > gpurun_out/r3/shape_c4.txt
cat gpurun_out/r3/shape_c*.txt
for ax in 5 6 7; do
  GXSV_BATCH_AXES=$ax timeout 200 python bench.py $B --steps 3 --warmup 3 > gpurun_out/r3/b_c2_ax$ax.json 2>/dev/null
  GXSV_BATCH_AXES=$ax timeout 300 python bench.py $B --workload config3_qaoa29 --steps 2 --warmup 1 > gpurun_out/r3/b_c3_ax$ax.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3/b_c?_ax*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d.get('value')), round(d.get('ms_per_step'),2), round(r['frac'],3), round(r['avg_launch_ms'],4), d.get('projections_per_pass'))
    except Exception as e: print(f, 'ERR', e)
PY
MB_QUICK=1 timeout 200 python tools/microbench_batch.py 27 > gpurun_out/r3/mb27b.log 2>&1; tail -19 gpurun_out/r3/mb27b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_tile -s 30 -c 3 -f -o gpurun_out/r3/prof_c3_tile python bench.py $B --workload config3_qaoa29 --steps 1 --warmup 1 --max-commands 1500 > gpurun_out/r3/ncu_c3.log 2>&1; tail -3 gpurun_out/r3/ncu_c3.log
ls -la gpurun_out/r3/
