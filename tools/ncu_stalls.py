import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print("==", d.get("Kernel Name","?")[:90], d.get("gpu__time_duration.sum"))
    out=[]
    for h in hdr:
        if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try: out.append((float(d[h]),h.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio","")))
            except: pass
    print("  ", [(n,round(v,2)) for v,n in sorted(out,reverse=True)[:9]])
    for k in ("smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","sm__warps_active.avg.pct_of_peak_sustained_active"):
        print("  ",k,d.get(k))
