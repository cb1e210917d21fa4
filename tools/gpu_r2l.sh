#!/bin/bash
cd /root/repo
OUT=gpurun_out/r2l; mkdir -p $OUT
timeout 300 python -m pytest tests/test_pwc_host.py -q -m gpu > $OUT/pytest_pwc_host.log 2>&1; echo "pwc host tests exit=$?"; tail -3 $OUT/pytest_pwc_host.log | cut -c1-300
timeout 300 python bench.py --config 3 --steps 5 --warmup 3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "c3 exit=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2l/bench_c3.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["ms_per_forward"], d["roofline"]["backwarp_ms_per_forward"])
PY
