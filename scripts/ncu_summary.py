"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("kernel:", name)
    for i, h in enumerate(hdr):
        if h in keys:
            print(f"  {h} [{units[i]}] = {r[i]}")
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.15:
                print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]} = {v:.2f}")
