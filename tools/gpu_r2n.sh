#!/bin/bash
# round 2, GPU call N: fused lookup + convc1: tests, then one ncu --set full capture
cd /root/repo
OUT=gpurun_out/r2n; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_fused_conv.py -q -m gpu -s > $OUT/pytest_fused.log 2>&1; echo "fused tests exit=$?"; grep -E "passed|failed|^FAILED|^ERROR|EPE" $OUT/pytest_fused.log | cut -c1-300 | head
CMD="python tools/fused_convc1_probe.py 8x47x156"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:lookup_convc1 -s 2 -c 1 -o $OUT/prof $CMD > $OUT/ncu.log 2>&1; echo "ncu exit=$?"; tail -2 $OUT/ncu.log | cut -c1-200
ls -la $OUT
