#!/bin/bash
# round 2, GPU call B: GPU tests incl. training / trajectory goldens; lookup kernel variants (column mask, L2 prefetch hints, stages, warps)
cd /root/repo
OUT=gpurun_out/r2b; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > $OUT/pytest.log 2>&1; echo "pytest exit=$?" | tee -a $OUT/pytest.log
grep -E "passed|failed|FAILED|trajectory|e2e_full" $OUT/pytest.log | head -40
for v in default l2_64 l2_128 l2_256 nomask s4 w4 w1; do
  LIB=""; [ $v != default ] && LIB="--lib build/variants/libffcorr_$v.so"
  for mode in "" "--smooth"; do
    timeout 300 python tools/kernel_bench.py --config 2 --iters 30 $mode --only lookup_tiled,lookup_tiled_nhwc $LIB 2>&1 | sed "s/^/$v $mode /" | tee -a $OUT/variants.txt
  done
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_requests_srcunit_tex_op_read.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
for v in default l2_64 l2_128; do
  LIB=""; [ $v != default ] && LIB="--lib build/variants/libffcorr_$v.so"
  timeout 300 ncu --metrics $M --clock-control none -k regex:"lookup_tiled" --csv --log-file $OUT/ncu_$v.csv python tools/kernel_bench.py --config 2 --iters 1 --warmup 0 --only lookup_tiled,lookup_tiled_nhwc $LIB > $OUT/ncu_$v.log 2>&1
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?"; cat $OUT/bench.json
