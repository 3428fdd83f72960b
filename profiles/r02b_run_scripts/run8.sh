mkdir -p gpurun_out/r3
B="--no-extra --no-per-command --no-e2e --no-cpu --no-parity"
GXSV_BATCH_LOG=2 timeout 600 python bench.py $B --workload config4_qft32 --max-commands 900 --steps 1 --warmup 1 > /dev/null 2> /tmp/c4shape.err
grep gxsv-pass /tmp/c4shape.err > gpurun_out/r3/c4_pass_times_io.txt
python - <<'PY'
import re,collections
rows=[]
for l in open('gpurun_out/r3/c4_pass_times_io.txt'):
    m=re.match(r'gxsv-pass n=(\d+) axes=(\d+) stages=(\d+) phases=(\d+) layouts=(\d+) uniform=(\d+) tw=(\d+) ms=([\d.]+) bits=(.*)',l)
    if m:
        n,a,s,p,nl,u,tw=map(int,m.groups()[:7]); ms=float(m.group(8)); bits=[int(x) for x in m.group(9).split(',')]
        rows.append((n,a,s,p,nl,tw,ms,bits))
rows=rows[len(rows)//2:]
for r in rows: print(r[:6], round(r[6],2), 'frac', round(2*16*2**r[0]/6549.4e9*1e3/r[6],3), r[7])
tot=sum(r[6] for r in rows); fl=sum(2*16*2**r[0]/6549.4e9*1e3 for r in rows)
print('total ms',round(tot,1),'frac',round(fl/tot,3),'passes',len(rows),'stages',sum(r[2] for r in rows))
PY
for bm in 8 10 12; do
  timeout 600 python bench.py $B --workload config4_qft32 --max-commands 900 --steps 2 --warmup 1 --batch $bm > gpurun_out/r3/c4_batch$bm.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3/c4_batch*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d.get('value'),1), round(d.get('ms_per_step'),2), round(r['frac'],3), round(r['avg_launch_ms'],4), d.get('projections_per_pass'))
    except Exception as e: print(f, 'ERR', e)
PY
