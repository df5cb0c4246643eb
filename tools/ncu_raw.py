import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','dram__bytes_read.sum','dram__bytes_write.sum','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__shared_mem_per_block_dynamic','sm__cycles_active.avg','smsp__cycles_active.avg']
for r in rows[2:]:
    print("---", r[hdr.index("Kernel Name")][:60] if "Kernel Name" in hdr else "")
    for i, h in enumerate(hdr):
        if h in want or "issue_stalled" in h and "per_issue_active" in h or (len(sys.argv) > 2 and sys.argv[2] in h):
            try:
                if "issue_stalled" in h and float(r[i]) < 0.15: continue
            except ValueError: pass
            print("  %-95s %s %s" % (h, r[i], rows[1][i]))
