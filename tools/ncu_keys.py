"""print the handful of ncu --set full metrics the acquisition / tracking notes quote: python tools/ncu_keys.py file.csv"""
import csv, sys
want=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","lts__t_sector_hit_rate.pct","smsp__inst_executed.sum","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","lts__throughput.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
stalls="smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
for f in sys.argv[1:]:
    rows=list(csv.reader(open(f))); hdr=rows[0]; units=rows[1]
    for r in rows[2:]:
        print(f, r[hdr.index("Kernel Name")][:60], r[hdr.index("Grid Size")], r[hdr.index("Block Size")])
        for w in want:
            if w in hdr: print("   %-75s %s %s"%(w, r[hdr.index(w)], units[hdr.index(w)]))
        st=[]
        for i,h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        print("   stalls per issue:", ", ".join("%s %.2f"%(n,v) for v,n in sorted(st,reverse=True)[:8]))
