mkdir -p gpurun_out/r3
B="--no-extra --no-per-command --no-e2e --no-cpu --no-parity"
GXSV_BATCH_LOG=2 timeout 600 python bench.py $B --workload config4_qft32 --max-commands 900 --steps 1 --warmup 1 > /dev/null 2> /tmp/c4shape.err
grep gxsv-pass /tmp/c4shape.err | tail -n +1 > gpurun_out/r3/c4_pass_times.txt
python - <<'PY'
import re,collections
rows=[]
for l in open('gpurun_out/r3/c4_pass_times.txt'):
    m=re.match(r'gxsv-pass n=(\d+) axes=(\d+) stages=(\d+) phases=(\d+) uniform=(\d+) tw=(\d+) ms=([\d.]+) bits=(.*)',l)
    if m:
        n,a,s,p,u,tw=map(int,m.groups()[:6]); ms=float(m.group(7)); bits=[int(x) for x in m.group(8).split(',')]
        rows.append((n,a,s,p,u,tw,ms,bits))
rows=rows[len(rows)//2:]   # second half = the timed step
agg=collections.defaultdict(lambda:[0,0.0,0.0])
for n,a,s,p,u,tw,ms,bits in rows:
    floor=2*16*2**n/6549.4e9*1e3
    k=(n,a,s,p,min(bits)<2)
    agg[k][0]+=1; agg[k][1]+=ms; agg[k][2]+=floor
print('n axes stages phases lowaxis : count avg_ms frac')
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(k, v[0], round(v[1]/v[0],3), round(v[2]/v[1],3))
tot=sum(r[6] for r in rows); fl=sum(2*16*2**r[0]/6549.4e9*1e3 for r in rows)
print('total ms',round(tot,1),'frac',round(fl/tot,3),'passes',len(rows),'stages',sum(r[2] for r in rows))
PY
MB_QUICK=1 timeout 300 python tools/microbench_batch.py 32 2 2>&1 | grep "n=32" | tee gpurun_out/r3/mb32.log
