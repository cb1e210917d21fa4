#!/bin/bash
# One ncu pass over bench.py (ONE ncu use per gpurun call).  Usage: gpu_profile_bench.sh <mode> [tag]
#   lookup   : --set full capture of one steady-state lookup launch
#   build    : --set full capture of the operand pre-pass + fused GEMM of the timed step
#   launches : launch list (gpu__time_duration.sum) of every kernel of a short run
cd /root/repo
MODE=$1; TAG=${2:-profiles_r1c}
OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cudnn-benchmark"
$CMD > $OUT/bench_plain_$MODE.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/bench_plain_$MODE.log; exit 1; }
case $MODE in
  lookup)   ncu --set full --import-source on --clock-control none -k regex:"lookup_tiled" -s 14 -c 1 -o $OUT/lookup_full $CMD > $OUT/ncu_lookup.log 2>&1 ;;
  build)    ncu --set full --import-source on --clock-control none -k regex:"volume_gemm|operand_prepass" -s 2 -c 2 -o $OUT/build_full $CMD > $OUT/ncu_build.log 2>&1 ;;
  launches) ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1 ;;
esac
echo "ncu $MODE exit=$?"
ls -la $OUT | tail -8
