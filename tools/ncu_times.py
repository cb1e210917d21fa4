"""Summarise `ncu --csv --metrics gpu__time_duration.sum[,launch__grid_size]` output: mean us per (kernel, grid)."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
t, g = {}, {}
for r in rows[1:]:
    if r[mi] == "gpu__time_duration.sum":
        t[r[idi]] = (r[ki], float(r[vi].replace(",", "")))
    elif r[mi] == "launch__grid_size":
        g[r[idi]] = r[vi]
acc = collections.OrderedDict()
for i, (k, v) in t.items():
    key = (k.split("(")[0][-60:], g.get(i, "?"))
    acc.setdefault(key, []).append(v)
tot = sum(sum(v) for v in acc.values())
for (k, gs), v in acc.items():
    print(f"{k:62s} grid={gs:>8s} n={len(v):4d} mean={sum(v)/len(v)/1e3:9.2f} us  share={100*sum(v)/tot:5.1f}%")
