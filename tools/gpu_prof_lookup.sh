#!/bin/bash
# timing (random / smooth / integer flow) + one ncu capture of the streaming tiled lookup
cd /root/repo
TAG=${1:-prof_lk}
OUT=gpurun_out/$TAG; mkdir -p $OUT
python tools/kernel_bench.py --config 2 --only lookup_tiled
python tools/kernel_bench.py --config 2 --only lookup_tiled --smooth
python tools/kernel_bench.py --config 2 --only lookup_tiled --sigma 0
CMD="python tools/kernel_bench.py --config 2 --iters 1 --warmup 0 --only lookup_tiled --smooth"
$CMD > $OUT/plain.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:"lookup_tiled" -c 1 -o $OUT/prof $CMD > $OUT/ncu.log 2>&1
echo "ncu exit=$?"
