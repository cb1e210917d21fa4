#!/usr/bin/env python
"""ncu `--metrics ... --csv` log  ->  one text block per kernel launch.  Usage: ncu_counters.py IN.csv OUT.txt "<command>" [regex]"""
import csv
import re
import sys

src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
pat = re.compile(sys.argv[4]) if len(sys.argv) > 4 else None
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]
ki, mi, ui, vi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
launches = {}
for r in rows[1:]:
    launches.setdefault((int(r[ii]), r[ki]), []).append((r[mi], r[vi], r[ui]))
with open(out, "w") as f:
    f.write(f"# ncu --metrics <list> --clock-control none; command: {cmd}\n")
    for (lid, name), ms in sorted(launches.items()):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("ffcorr::<unnamed>::", "")
        if pat and not pat.search(short):
            continue
        f.write(f"\n== launch {lid}: {short[:110]}\n")
        for m, v, u in ms:
            f.write(f"  {m:72s} {v:>16s} {u}\n")
print(open(out).read()[:600])
