"""Aggregate the source page of an ncu report by CUDA line:  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | tail -n +3 > src.csv;  python profiles/stalls_by_line.py src.csv [N lines]"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
for hi,r in enumerate(rows[:5]):
    if "# Samples" in r: break
hdr=rows[hi]
def I(x):
    try: return int(float(x))
    except: return 0
data=rows[hi+1:]
# group sass rows under the preceding cuda line
groups=[]; cur=None
cl=hdr.index("stall_long_sb"); cw=hdr.index("stall_wait"); cb=hdr.index("stall_barrier"); cs=hdr.index("stall_short_sb"); ci=hdr.index("Instructions Executed"); iS=hdr.index("# Samples")
cm=hdr.index("stall_mio"); cg=hdr.index("stall_lg"); csl=hdr.index("stall_sleep") if "stall_sleep" in hdr else None
for r in data:
    if r[0]!='':
        cur={"line":r[0],"src":r[1],"samples":0,"long":0,"wait":0,"bar":0,"short":0,"inst":0,"mio":0,"lg":0,"sass":[]}
        groups.append(cur)
    elif cur is not None and len(r)>cw:
        cur["samples"]+=I(r[iS]); cur["long"]+=I(r[cl]); cur["wait"]+=I(r[cw]); cur["bar"]+=I(r[cb]); cur["short"]+=I(r[cs]); cur["inst"]+=I(r[ci]); cur["mio"]+=I(r[cm]); cur["lg"]+=I(r[cg])
        cur["sass"].append((I(r[iS]), r[3].strip()))
# merge by line no
import collections
by=collections.OrderedDict()
for g in groups:
    k=(g["line"],g["src"])
    if k not in by: by[k]={x:0 for x in ("samples","long","wait","bar","short","inst","mio","lg")}; by[k]["sass"]=[]
    for x in ("samples","long","wait","bar","short","inst","mio","lg"): by[k][x]+=g[x]
    by[k]["sass"]+=g["sass"]
tot=sum(v["samples"] for v in by.values())
print("total samples",tot, "total inst", sum(v["inst"] for v in by.values()))
N=int(sys.argv[2]) if len(sys.argv)>2 else 40
for (ln,src),v in sorted(by.items(),key=lambda kv:-kv[1]["samples"])[:N]:
    print(f"{ln:>5} {v['samples']*100/tot:5.1f}% long={v['long']:>6} short={v['short']:>6} wait={v['wait']:>6} bar={v['bar']:>5} mio={v['mio']:>5} lg={v['lg']:>5} inst={v['inst']:>11}  {src[:95]}")
    if len(sys.argv)>3:
        for s,t in sorted(v["sass"],key=lambda x:-x[0])[:3]: print("            ",s,t[:80])
