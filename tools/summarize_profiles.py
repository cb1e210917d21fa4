"""Turn gpurun_out/<dir> ncu artefacts into the committed summaries under profiles/."""
import collections, csv, json, os, re, subprocess, sys

src = sys.argv[1]
tag = sys.argv[2] if len(sys.argv) > 2 else "r1"
out = "profiles"
os.makedirs(out, exist_ok=True)

# ---- 1. launch list: per-kernel totals and shares -------------------------------------------------
lp = os.path.join(src, "launches.csv")
if not os.path.exists(lp):
    lp = os.path.join(src, "launches_ffcorr.csv")
only_ffcorr = lp.endswith("launches_ffcorr.csv")
rows = [r for r in csv.reader(open(lp)) if len(r) > 10]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
acc = collections.OrderedDict()
n_launch = 0
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    n_launch += 1
    name = re.sub(r"\(.*", "", r[ki])[-70:]
    acc.setdefault(name, []).append(float(r[vi].replace(",", "")) / 1e3)  # us
tot = sum(sum(v) for v in acc.values())
mine = {k: v for k, v in acc.items() if "ffcorr" in k or any(s in k for s in ("lookup_kernel", "volume_gemm", "pyramid_", "operand_prepass", "pwc81"))}
with open(os.path.join(out, f"{tag}_launches_summary.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none" + (" -k <this repo's kernels only>" if only_ffcorr else "") + " python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cudnn-benchmark\n")
    f.write(f"# {n_launch} launches (warm-up step + timed step, + e2e steps), total {tot/1e3:.2f} ms of kernel time (cold-cache, serialised)\n")
    f.write(f"# share of this repo's kernels: {100*sum(sum(v) for v in mine.values())/tot:.2f}%\n")
    f.write(f"{'kernel':72s} {'n':>6s} {'mean_us':>10s} {'total_ms':>10s} {'share%':>8s}\n")
    for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{k:72s} {len(v):6d} {sum(v)/len(v):10.2f} {sum(v)/1e3:10.3f} {100*sum(v)/tot:8.2f}\n")
print(open(os.path.join(out, f"{tag}_launches_summary.txt")).read()[:2500])

# ---- 2. full captures: selected metrics ----------------------------------------------------------
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
traffic = {}
for rep in ("lookup_full", "build_full"):
    p = os.path.join(src, rep + ".ncu-rep")
    if not os.path.exists(p):
        continue
    raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u, data = rr[0], rr[1], rr[2:]
    idx = {n: i for i, n in enumerate(h)}
    with open(os.path.join(out, f"{tag}_{rep}_ncu.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, kernels captured inside `python bench.py --steps 1 --warmup 1`\n")
        for d in data:
            name = re.sub(r"\(.*", "", d[idx["Kernel Name"]])[-70:]
            f.write(f"\n== {name}\n")
            for w in WANT:
                if w in idx:
                    f.write(f"  {w:82s} {d[idx[w]]:>16s} {u[idx[w]]}\n")
            if "lookup" in name:
                def num(key):
                    v = float(d[idx[key]].replace(",", ""))
                    unit = u[idx[key]].lower()
                    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(unit, 1)
                traffic = {"dram_bytes_per_launch": int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum")),
                           "dram_read": int(num("dram__bytes_read.sum")), "dram_write": int(num("dram__bytes_write.sum")),
                           "source": f"profiles/{tag}_{rep}_ncu.txt"}
    print(open(os.path.join(out, f"{tag}_{rep}_ncu.txt")).read()[:1800])
if traffic:
    json.dump(traffic, open(os.path.join(out, "lookup_traffic.json"), "w"), indent=1)
    print(traffic)
for kb in ("kernel_bench_c2.jsonl", "kernel_bench_c4.jsonl", "kernel_bench_c1.jsonl"):
    pth = os.path.join(src, kb)
    if os.path.exists(pth):
        open(os.path.join(out, f"{tag}_{kb}"), "w").write("".join(l for l in open(pth) if l.startswith("{")))
b = os.path.join(src, "bench_r1.json")
if os.path.exists(b):
    open(os.path.join(out, f"{tag}_bench.json"), "w").write(open(b).read())
