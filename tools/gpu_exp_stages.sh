#!/bin/bash
# lookup streaming kernel: parity + timing for several ring depths (rebuilds lookup.o on the box)
cd /root/repo
python -m pytest tests -m gpu -q -x -k "tiled or lookup or full_size or e2e" 2>&1 | tail -3
python tools/kernel_bench.py --config 2 | grep -E "lookup_tiled"
python tools/kernel_bench.py --config 2 --sigma 0 | grep -E "lookup_tiled"
for st in 2 4 5; do
  ( cd focusflow_official_b200/csrc && rm -f lookup.o && make NVFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr -DFFCORR_STREAM_STAGES=$st" 2>&1 | grep -A1 "lookup_tiled_stream_kernelILi4" | grep Used )
  echo "stages=$st"
  python tools/kernel_bench.py --config 2 | grep -E "lookup_tiled"
  python tools/kernel_bench.py --config 2 --sigma 0 | grep -E "lookup_tiled"
done
