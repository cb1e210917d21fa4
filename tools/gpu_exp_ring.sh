#!/bin/bash
# fused build: operand stages vs store-ring depth (rebuilds volume.o on the box)
cd /root/repo
python tools/kernel_bench.py --config 2 --only volume_tiled,build_fused
for cfg in "2 6" "2 4" "3 4"; do
  set -- $cfg
  ( cd focusflow_official_b200/csrc && rm -f volume.o && make NVFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr -DFFCORR_GEMM_STAGES=$1 -DFFCORR_STORE_BUFS=$2" 2>&1 | grep -E "error" )
  echo "stages=$1 store_bufs=$2"
  python tools/kernel_bench.py --config 2 --only volume_tiled,build_fused
done
python -m pytest tests -m gpu -q -x -k "fused_build" 2>&1 | tail -2
