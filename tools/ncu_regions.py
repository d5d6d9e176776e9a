"""Per-region view of an ncu source page (ncu -i x.ncu-rep --page source --csv > src.csv; python tools/ncu_regions.py
src.csv): instruction share, stall-sample share and shared-memory wavefront share of the stretches of SASS between
two barriers of the BP tile kernel (finalize | refill | check phase | variable phase | bookkeeping)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1] if len(sys.argv) > 1 else "/tmp/src_base.csv")))
kern=[];cur=None
for r in rows:
    if r and r[0]=="Kernel Name": cur={"name":r[1],"rows":[]}; kern.append(cur)
    elif r and r[0]=="Address": cur["hdr"]=r
    elif cur is not None and len(r)>5: cur["rows"].append(r)
for k in kern[:2]:
    h=k["hdr"]; iI=h.index("Instructions Executed"); iS=h.index("Source"); iN=h.index("# Samples")
    iW=h.index("L1 Wavefronts Shared"); iWi=h.index("L1 Wavefronts Shared Ideal")
    st=[c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    print(k["name"])
    tot=sum(int(r[iI]) for r in k["rows"]); totS=sum(int(r[iN]) for r in k["rows"])
    # regions delimited by BAR.SYNC
    reg=[]; acc=collections.Counter(); n=0; first=0
    for i,r in enumerate(k["rows"]):
        acc["inst"]+=int(r[iI]); acc["samp"]+=int(r[iN]); acc["wf"]+=int(r[iW]); acc["wfi"]+=int(r[iWi])
        for c in st: acc[c]+=int(r[h.index(c)])
        if "BAR.SYNC" in r[iS] or i==len(k["rows"])-1:
            reg.append((first,i,acc)); acc=collections.Counter(); first=i+1
    for a,b,c in reg:
        if c["inst"]/tot<0.004: continue
        top=sorted(((c[s],s[6:]) for s in st),reverse=True)[:6]
        print("  lines %4d-%4d inst %5.1f%% samples %5.1f%% wf %5.1f%% (excess %4.1f%%) ipc-ish %.2f | %s"%(a,b,100*c["inst"]/tot,100*c["samp"]/totS,100*c["wf"]/max(1,sum(x[2]["wf"] for x in reg)),100*(c["wf"]-c["wfi"])/max(1,c["wfi"]),c["inst"]/max(1,c["samp"]),", ".join("%s %.0f%%"%(n_,100*v/max(1,c["samp"])) for v,n_ in top)))
