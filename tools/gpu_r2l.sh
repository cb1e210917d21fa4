#!/bin/bash
# round 2, GPU call L: fused lookup + convc1 probe and in-model bench
cd /root/repo
OUT=gpurun_out/r2l; mkdir -p $OUT
timeout 120 python tools/fused_convc1_probe.py > $OUT/probe.txt 2> $OUT/probe.err; echo "probe exit=$?"; cat $OUT/probe.txt; tail -5 $OUT/probe.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-stock --no-pwc --no-cpu-baseline --fuse-convc1 > $OUT/bench_fused.json 2> $OUT/bench_fused.err; echo "bench fused exit=$?"
timeout 200 python bench.py --steps 5 --warmup 3 --no-stock --no-pwc --no-cpu-baseline > $OUT/bench_plain.json 2> $OUT/bench_plain.err; echo "bench plain exit=$?"
python - <<'PY'
import json
for n in ("fused","plain"):
    try:
        d=json.load(open(f"gpurun_out/r2l/bench_{n}.json"))
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["launch_ms"], d["roofline"]["frac"], d["roofline"]["kernel"][:40])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 $OUT/bench_fused.err
