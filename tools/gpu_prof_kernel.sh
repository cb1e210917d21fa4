#!/bin/bash
# one ncu --set full capture of one kernel of kernel_bench.  Usage: gpu_prof_kernel.sh <tag> <kernel-regex> <kernel_bench --only value> [extra args]
cd /root/repo
TAG=$1; KRE=$2; ONLY=$3; shift 3
OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python tools/kernel_bench.py --config 2 --iters 1 --warmup 0 --only $ONLY $*"
$CMD > $OUT/plain.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:"$KRE" -c 1 -o $OUT/prof $CMD > $OUT/ncu.log 2>&1
echo "ncu exit=$?"; tail -3 $OUT/plain.log
