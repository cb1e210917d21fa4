#!/bin/bash
# round 2, GPU call I: PWC forward with half-height tiles for the small levels -- parity and timing
cd /root/repo
OUT=gpurun_out/r2i; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_pwc_host.py -m gpu -q -k "pwc or PWC or backwarp or empty_batch" > $OUT/pytest_pwc.log 2>&1; echo "pytest pwc exit=$?"; tail -4 $OUT/pytest_pwc.log | cut -c1-300
for c in "16 32 112 256" "16 64 56 128" "16 96 28 64" "16 128 14 32" "16 196 7 16"; do timeout 120 python tools/pwc_level.py $c 2>&1 | tail -1; done | tee $OUT/pwc_levels_back_to_back.txt
timeout 300 python bench.py --config 3 --steps 5 --warmup 3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "c3 exit=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2i/bench_c3.json")); r=d["roofline"]
print(d["value"], d["ms_per_step"], r["ms_per_forward"], r["frac"], [(l["C"], l["ms"], l["frac"]) for l in r["levels"]])
PY
