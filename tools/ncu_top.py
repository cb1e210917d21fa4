"""List the SASS instructions with the most warp-stall samples from an `ncu --page source --csv` dump.
Usage: python tools/ncu_top.py report.ncu-rep <kernel regex> [top N]"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several kernels may match: split on 'Kernel Name' rows, report the first
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
s = start[which]
e = start[which + 1] if which + 1 < len(start) else len(rows)
print(rows[s][1][:150])
hdr = rows[s + 1]
body = [r for r in rows[s + 2:e] if len(r) >= 6]
tot = sum(int(r[2] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
ranked = sorted(enumerate(body), key=lambda t: -int(t[1][2] or 0))[:top]
for i, r in sorted(ranked):
    print(f"{i:5d} {int(r[2]):7d} {100*int(r[2])/max(tot,1):5.1f}%  exec={r[5]:>9}  {r[1].strip()[:110]}")
