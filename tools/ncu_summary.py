#!/usr/bin/env python
"""Write the judged subset of an .ncu-rep (raw page) as text.  Usage: ncu_summary.py REP OUT.txt "<command captured>" """
import csv
import re
import subprocess
import sys

rep, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, data = rows[0], rows[1], rows[2:]
idx = {n: i for i, n in enumerate(h)}
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on; captured inside: {cmd}\n")
    for d in data:
        name = re.sub(r"\(.*", "", d[idx["Kernel Name"]])[-90:]
        f.write(f"\n== {name}\n")
        for w in WANT:
            if w in idx and d[idx[w]] not in ("", "n/a"):
                f.write(f"  {w:96s} {d[idx[w]]:>16s} {u[idx[w]]}\n")
print(open(out).read()[:1500])
