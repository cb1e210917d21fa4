#!/bin/bash
# round 2, GPU call T: host plumbing (context encoder NHWC + folded BatchNorm, stacked GRU gates, lerp): parity + bench
cd /root/repo; OUT=gpurun_out/r2t; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_conv.py -q -m gpu -s -k "context_encoder or e2e_epe or host_model or trajectory" > $OUT/pytest.log 2>&1; echo "tests exit=$?"; grep -E "passed|failed|NHWC vs NCHW|^FAILED|e2e_full|fused vs" $OUT/pytest.log | cut -c1-160
timeout 600 python bench.py --steps 10 --warmup 3 --no-stock --no-pwc --no-cpu-baseline > $OUT/bench_gru.json 2> $OUT/bench.err; echo "bench exit=$?"; cut -c1-200 $OUT/bench_gru.json
