#!/bin/bash
# ncu capture of the hot kernels (one launch each, cold). Usage: bash tools/gpu_prof.sh <tag> [extra kernel_bench args]
set -u
TAG=${1:-prof}; shift || true
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
CMD="python tools/kernel_bench.py --config 2 --iters 1 --warmup 0 $*"
$CMD > "$OUT/plain.log" 2>&1 && \
ncu --set full --import-source on --clock-control none \
    -k regex:"lookup_kernel|volume_gemm|pyramid_fused|pwc81_kernel|operand_prepass" -c 12 \
    -o "$OUT/prof" $CMD > "$OUT/ncu.log" 2>&1
echo "ncu exit=$?"
tail -5 "$OUT/ncu.log"
ls -la "$OUT"
