#!/bin/bash
# Runs the GPU parity suite in isolated processes (a trapped kernel poisons its CUDA context,
# so the riskier groups get their own interpreter), then the kernel micro-benchmarks.
# Usage (under gpurun): bash tools/gpu_check.sh [tag]
set -u
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/nvidia_smi.csv" 2>&1
run() { # name, timeout, cmd...
  local name=$1 to=$2; shift 2
  timeout "$to" "$@" > "$OUT/$name.log" 2>&1
  echo "$name exit=$?" | tee -a "$OUT/summary.txt"
}
run smoke 300 python __graft_entry__.py smoke
run t_lookup_pyr_pwc 600 python -m pytest tests -m gpu -q --timeout 180 -k "pyramid or lookup or pwc"
run t_vol_fp32 300 python -m pytest tests -m gpu -q --timeout 180 -k "volume and fp32"
run t_vol_fp16 300 python -m pytest tests -m gpu -q --timeout 180 -k "volume and fp16"
run t_vol_tf32 300 python -m pytest tests -m gpu -q --timeout 180 -k "volume and tf32"
run t_vol_bf16x3 300 python -m pytest tests -m gpu -q --timeout 180 -k "volume and bf16x3"
run t_rest 600 python -m pytest tests -m gpu -q --timeout 300 -k "not (pyramid or lookup or pwc or (volume and (fp32 or fp16 or tf32 or bf16x3)))"
run kbench 600 python tools/kernel_bench.py --config 2 --pwc --all-precisions
run kbench_c1 300 python tools/kernel_bench.py --config 1
tail -n 3 "$OUT"/t_*.log "$OUT/smoke.log"
cat "$OUT/kbench.log" "$OUT/kbench_c1.log"
