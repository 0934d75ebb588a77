mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_42_launches.csv python bench.py --steps 5 --warmup 3 --no-jacobian --no-e2e --no-cpu-baseline > gpurun_out/r2_42_bench_under_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_42_launches.csv') if not l.startswith('=='))]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); iu=h.index('Metric Unit')
agg=collections.OrderedDict()
seq=[]
for r in rows[1:]:
    if len(r)<=iv: continue
    v=float(r[iv].replace(',','')); u=r[iu]
    v = v/1e3 if u in ('ns','nsecond') else v*1e3 if u in ('ms','msecond') else v
    k=r[ik].split('(')[0][-48:]
    seq.append((k,v))
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(v[1] for v in agg.values())
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:12]: print('%-50s n=%4d total %9.1f us (%4.1f%%) avg %8.1f us'%(k,n,t,100*t/tot,t/n))
# the last step: kernels after the last-but-one residual kernel
idx=[i for i,(k,v) in enumerate(seq) if 'k_residual_fast_bulk' in k]
if len(idx)>=2:
    step=seq[idx[-2]+1:idx[-1]+1]
    st=sum(v for k,v in step)
    print('one step:', [(k[-24:],round(v,1)) for k,v in step], 'residual share %.3f'%(step[-1][1]/st))
PY
