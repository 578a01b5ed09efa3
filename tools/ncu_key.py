import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; vals=rows[2] if len(rows)>2 else rows[1]
want=["gpu__time_duration.sum","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","launch__registers_per_thread","launch__grid_size","launch__block_size","l1tex__t_sector_hit_rate.pct","smsp__thread_inst_executed_per_inst_executed.ratio","sm__cycles_active.avg","sm__cycles_elapsed.max"]
for w in want:
    for i,h in enumerate(hdr):
        if h==w: print(w, vals[i])
stalls=[(float(vals[i].replace(',','')),h) for i,h in enumerate(hdr) if ("smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"))]
for v,h in sorted(stalls,reverse=True)[:8]: print("  ",h.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio",""),v)
