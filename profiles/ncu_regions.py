"""Stall samples of an .ncu-rep grouped by code region (regions end at BAR.SYNC / CALL / RET): python profiles/ncu_regions.py <file.ncu-rep>"""
import csv, subprocess, sys
rep=sys.argv[1]
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hi=next(i for i,x in enumerate(rows) if "Source" in x and "# Samples" in x)
h=rows[hi]; data=rows[hi+1:]
si,so,ie=h.index("# Samples"),h.index("Source"),h.index("Instructions Executed")
# stall columns
cols=[(i,c) for i,c in enumerate(h) if c.startswith("stall_")]
reg=[]; cur={"start":0,"samples":0,"inst":0,"n":0,"stalls":{}}
tot=0
for idx,x in enumerate(data):
    if len(x)<=ie or not x[ie].isdigit(): continue
    s=int(x[si]) if x[si].isdigit() else 0
    cur["samples"]+=s; cur["inst"]+=int(x[ie]); cur["n"]+=1; tot+=s
    for i,c in cols:
        try: cur["stalls"][c]=cur["stalls"].get(c,0)+int(x[i])
        except: pass
    if "BAR.SYNC" in x[so] or "RET." in x[so] or "EXIT" in x[so] or "CALL" in x[so]:
        cur["end"]=idx; cur["endop"]=x[so].split()[0 if not x[so].strip().startswith('@') else 1]; reg.append(cur); cur={"start":idx+1,"samples":0,"inst":0,"n":0,"stalls":{}}
reg.append(cur)
for r in reg:
    if r["samples"]<tot*0.005: continue
    top=sorted(r["stalls"].items(), key=lambda kv:-kv[1])[:4]
    print(f"{r['start']:5d}-{r.get('end',0):5d} {r.get('endop','')[:10]:10s} n={r['n']:5d} samples {100*r['samples']/tot:5.1f}% inst {r['inst']:>11d}  " + " ".join(f"{k[6:]}={100*v/max(r['samples'],1):.0f}%" for k,v in top))
