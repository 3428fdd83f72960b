mkdir -p gpurun_out/r3
B="--no-extra --no-per-command --no-e2e --no-cpu --no-parity"
GXSV_BATCH_LOG=2 timeout 100 python bench.py $B --steps 1 --warmup 1 > /dev/null 2> /tmp/c2shape.err
grep gxsv-pass /tmp/c2shape.err | tail -n 188 > gpurun_out/r3/c2_pass_times_final.txt
GXSV_BATCH_LOG=2 timeout 200 python bench.py $B --workload config4_qft32 --max-commands 900 --steps 1 --warmup 1 > /dev/null 2> /tmp/c4shape.err
grep gxsv-pass /tmp/c4shape.err | tail -n 22 > gpurun_out/r3/c4_pass_times_final.txt
python - <<'PY'
import re,collections
for f in ('c2','c4'):
    rows=[]
    for l in open(f'gpurun_out/r3/{f}_pass_times_final.txt'):
        m=re.match(r'gxsv-pass n=(\d+) axes=(\d+) stages=(\d+) phases=(\d+) layouts=(\d+) uniform=(\d+) tw=(\d+) ms=([\d.]+) bits=(.*)',l)
        if m:
            n,a,s,p,nl,u,tw=map(int,m.groups()[:7]); rows.append((n,a,s,p,nl,tw,float(m.group(8))))
    agg=collections.defaultdict(lambda:[0,0.0,0.0])
    for n,a,s,p,nl,tw,ms in rows:
        k=(n,a,s,p,nl,tw); agg[k][0]+=1; agg[k][1]+=ms; agg[k][2]+=2*16*2**n/6549.4e9*1e3
    print(f,'(n,axes,stages,phases,layouts,tw): count avg_ms frac')
    for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(' ',k,v[0],round(v[1]/v[0],4),round(v[2]/v[1],3))
    tot=sum(r[6] for r in rows); fl=sum(2*16*2**r[0]/6549.4e9*1e3 for r in rows)
    print('  total ms',round(tot,2),'frac',round(fl/tot,3),'passes',len(rows))
PY
