import csv,sys,re,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
stall_cols=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ph=0; P=collections.OrderedDict()
def get(p): return P.setdefault(p,{"inst":0,"samples":0,"fp64":0,"lds":0,"stalls":collections.Counter(),"static":0,"ops":collections.Counter()})
for r in rows[2:]:
    if len(r)<len(hdr): continue
    src=r[ix["Source"]].strip()
    op=re.sub(r"^@!?U?P\d+\s+","",src).split()[0] if src else "?"
    base=op.split(".")[0]
    n=float(r[ix["Instructions Executed"]] or 0); s=float(r[ix["# Samples"]] or 0)
    d=get(ph); d["inst"]+=n; d["samples"]+=s; d["static"]+=1; d["ops"][base]+=n
    if base in("DFMA","DMUL","DADD","DSETP","MUFU"): d["fp64"]+=n
    if base in("LDS","STS"): d["lds"]+=n
    for c in stall_cols: d["stalls"][c]+=float(r[ix[c]] or 0)
    if base=="BAR": ph+=1
T=sum(d["samples"] for d in P.values()); I=sum(d["inst"] for d in P.values())
cells=float(sys.argv[2])
print("phase static inst/cell  share_inst share_time  rel_util fp64/cell lds/cell top stalls")
for p,d in P.items():
    st=", ".join(f"{k[6:]} {100*v/max(d['samples'],1):.0f}%" for k,v in d["stalls"].most_common(4))
    print(f"{p:3d} {d['static']:5d} {d['inst']*32/cells:8.1f} {100*d['inst']/I:6.1f}% {100*d['samples']/T:6.1f}%  {(d['inst']/I)/(d['samples']/T):5.2f}  {d['fp64']*32/cells:7.1f} {d['lds']*32/cells:6.1f}  {st}")
print("total inst/cell",I*32/cells)
for p in (0,1):
    d=P[p]
    print("phase",p,", ".join(f"{k} {v*32/cells:.1f}" for k,v in d["ops"].most_common(22)))
for p in (2,5,7):
    if p in P:
        d=P[p]
        print("phase",p,", ".join(f"{k} {v*32/cells:.1f}" for k,v in d["ops"].most_common(24)))
