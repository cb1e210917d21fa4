#!/bin/bash
# round 2, GPU call T: context encoder in channels_last (bench default from here on): parity test + both bench arms' line
cd /root/repo; OUT=gpurun_out/r2t; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "context_encoder or e2e_epe" > $OUT/pytest.log 2>&1; echo "tests exit=$?"; grep -E "passed|failed|NHWC vs NCHW|^FAILED" $OUT/pytest.log | cut -c1-250
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?"; cut -c1-200 $OUT/bench.json
timeout 300 python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo "c4 exit=$?"; cut -c1-200 $OUT/bench_c4.json
