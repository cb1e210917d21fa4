#!/bin/bash
# A/B/A/B of the whole model with and without the fused lookup + convc1, 30 timed steps each
cd /root/repo
OUT=gpurun_out/r2l; mkdir -p $OUT
for i in 1 2; do for tag in plain fused; do
  FLAG=""; [ $tag = fused ] && FLAG="--fuse-convc1"
  timeout 200 python bench.py --steps 30 --warmup 5 --no-stock --no-pwc --no-cpu-baseline $FLAG > $OUT/ab_${tag}_$i.json 2> $OUT/ab.err || tail -3 $OUT/ab.err
done; done
python - <<'PY'
import json
for i in (1,2):
    for n in ("plain","fused"):
        d=json.load(open(f"gpurun_out/r2l/ab_{n}_{i}.json"))
        print(n, i, "pairs/s", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "lookup launch ms", d["roofline"]["launch_ms"])
PY
