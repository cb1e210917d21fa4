#!/bin/bash
# quick experiment matrix for the lookup + regression check
set -u
OUT=gpurun_out/${1:-exp}; mkdir -p $OUT
python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -3
echo "== default"; python tools/kernel_bench.py --config 2 2>&1 | tee $OUT/default.log
echo "== QU8"; FFCORR_LOOKUP_QU=8 python tools/kernel_bench.py --config 2 2>&1 | grep lookup
echo "== QU2"; FFCORR_LOOKUP_QU=2 python tools/kernel_bench.py --config 2 2>&1 | grep lookup
echo "== L2 32"; FFCORR_L2_FETCH=32 python tools/kernel_bench.py --config 2 2>&1 | tee $OUT/l2_32.log
echo "== L2 32 QU8"; FFCORR_L2_FETCH=32 FFCORR_LOOKUP_QU=8 python tools/kernel_bench.py --config 2 2>&1 | grep lookup
echo "== L2 128"; FFCORR_L2_FETCH=128 python tools/kernel_bench.py --config 2 2>&1 | grep lookup
