"""Key ncu metrics of every kernel in a report: python profiles/ncu_summary.py rep.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct"]
stall = [c for c in h if "smsp__average_warp" in c and "issue_stalled" in c and c.endswith(".ratio") and "not_issued" not in c]
for r in rows[2:]:
    print(r[h.index("Kernel Name")])
    for w in want:
        if w in h:
            print(f"  {w} {r[h.index(w)]} {rows[1][h.index(w)]}")
    for c in stall:
        v = float(r[h.index(c)] or 0)
        if v > 0.15:
            print("   stall", c.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), round(v, 2))
