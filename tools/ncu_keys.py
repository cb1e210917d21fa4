#!/usr/bin/env python
"""Print the metrics that explain a kernel's bound from an .ncu-rep (raw page). Usage: ncu_keys.py REP [kernel-substr]"""
import csv, subprocess, sys
rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct", "lts__t_sectors.sum", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "issue_stalled", "sm__inst_executed_pipe_lsu.avg.pct", "l1tex__data_bank_conflicts", "shared_op",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__average_warp_latency", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate"]
for row in rows[2:]:
    d = dict(zip(h, row))
    if sub not in d["Kernel Name"]:
        continue
    print("==", d["Kernel Name"][:90])
    for k, v in d.items():
        if any(s in k for s in KEYS) and "peak_sustained" not in k.replace("pct_of_peak_sustained", ""):
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            if x != 0 and ("stalled" not in k or x > 0.15):
                print(f"  {k:95s} {v}")
