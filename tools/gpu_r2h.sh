#!/bin/bash
# round 2, GPU call H: compute-sanitizer memcheck over the kernels added in round 2 (small test shapes)
cd /root/repo
OUT=gpurun_out/r2h; mkdir -p $OUT
SEL="grouped or channels_last_lookup or fp16_storage_build or fp16_storage_refuses or block_scaled or flowformer_ops"
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $OUT/memcheck_corr.log 2>&1; echo "memcheck corr exit=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|=========     at" $OUT/memcheck_corr.log | head -20 | cut -c1-220
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_pwc_host.py -m gpu -q -x -k "backwarp" > $OUT/memcheck_warp.log 2>&1; echo "memcheck backwarp exit=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $OUT/memcheck_warp.log | head -10 | cut -c1-220
