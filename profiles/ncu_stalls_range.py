"""Stall reasons summed over a range of SASS lines of an .ncu-rep: python profiles/ncu_stalls_range.py <file.ncu-rep> <first> <last>"""
import csv, subprocess, sys, collections
rep=sys.argv[1]; lo=int(sys.argv[2]); hi_=int(sys.argv[3])
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hi=next(i for i,x in enumerate(rows) if "Source" in x and "# Samples" in x)
h=rows[hi]; data=rows[hi+1:]
si,so,ie=h.index("# Samples"),h.index("Source"),h.index("Instructions Executed")
cols=[(i,c) for i,c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot=collections.Counter(); n=0; inst=0
for idx,x in enumerate(data[lo:hi_]):
    if len(x)<=ie or not x[ie].isdigit(): continue
    n+=int(x[si]) if x[si].isdigit() else 0; inst+=int(x[ie])
    for i,c in cols:
        try: tot[c]+=int(x[i])
        except: pass
print("samples",n,"inst",inst)
for c,v in tot.most_common(10): print(f"  {c:30s} {100*v/max(n,1):5.1f}%")
