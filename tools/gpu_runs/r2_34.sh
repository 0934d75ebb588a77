mkdir -p gpurun_out
timeout 300 python tools/strips_only.py 8192x2048 2>&1 | tail -1
BROADCAST_B200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_34_strips_launches.csv python tools/strips_only.py 8192x2048 > gpurun_out/r2_34_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_34_strips_launches.csv') if not l.startswith('=='))]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); iu=h.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    if len(r)<=iv: continue
    v=float(r[iv].replace(',','')); u=r[iu]
    v = v/1e3 if u in ('ns','nsecond') else v*1e3 if u in ('ms','msecond') else v
    k=r[ik].split('(')[0][-50:]
    agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print('%-52s n=%5d total %.2f ms (%.0f%%) avg %.1f us'%(k,n,t/1e3,100*t/tot,t/n))
print('total', tot/1e3,'ms')
PY
