#!/bin/bash
# round 2, GPU call L: fused lookup + convc1 in the model (bench, config 2 and 4)
cd /root/repo
OUT=gpurun_out/r2l; mkdir -p $OUT
for tag in fused plain; do
  FLAG=""; [ $tag = fused ] && FLAG="--fuse-convc1"
  timeout 200 python bench.py --steps 5 --warmup 3 --no-stock --no-pwc --no-cpu-baseline $FLAG > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err; echo "bench $tag exit=$?"
done
python - <<'PY'
import json
for n in ("fused","plain"):
    try:
        d=json.load(open(f"gpurun_out/r2l/bench_{n}.json"))
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["launch_ms"], d["roofline"]["frac"], d["roofline"]["kernel"][:40], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 $OUT/bench_fused.err
