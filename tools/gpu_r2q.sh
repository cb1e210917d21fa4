#!/bin/bash
# round 2, GPU call Q: last verification of the committed tree
cd /root/repo
OUT=gpurun_out/r2q; mkdir -p $OUT
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 $OUT/smoke.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit=$?"; grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest.log | cut -c1-200 | head
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?"; cut -c1-200 $OUT/bench.json
