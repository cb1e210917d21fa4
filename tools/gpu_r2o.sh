#!/bin/bash
# round 2, GPU call O: verification of the committed tree + ncu of the fused kernel as committed
cd /root/repo
OUT=gpurun_out/r2o; mkdir -p $OUT
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 $OUT/smoke.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit=$?"; grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest.log | cut -c1-200 | head
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref exit=$?"; cut -c1-200 $OUT/bench_ref.json
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?"; cut -c1-300 $OUT/bench.json
timeout 300 python bench.py --config 3 --steps 5 --warmup 3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "c3 exit=$?"; cut -c1-200 $OUT/bench_c3.json
CMD="python tools/fused_convc1_probe.py 8x47x156"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:lookup_convc1 -s 2 -c 1 -o $OUT/prof_fused $CMD > $OUT/ncu.log 2>&1; echo "ncu exit=$?"
