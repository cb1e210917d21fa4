#!/bin/bash
# round 2, GPU call M: tests of the fused lookup + convc1
cd /root/repo
OUT=gpurun_out/r2m; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_fused_conv.py -q -m gpu -s > $OUT/pytest_fused.log 2>&1; echo "fused tests exit=$?"; grep -E "passed|failed|^FAILED|^ERROR|EPE|assert" $OUT/pytest_fused.log | cut -c1-300 | head -30
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "channels_last or golden_lookup" > $OUT/pytest_cl.log 2>&1; echo "channels-last tests exit=$?"; tail -2 $OUT/pytest_cl.log | cut -c1-200
