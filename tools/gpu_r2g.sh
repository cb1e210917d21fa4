#!/bin/bash
# round 2, GPU call G: the grouped tile order -- parity with the tiled layout, and its effect on build / lookup time
cd /root/repo
OUT=gpurun_out/r2g; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "grouped" > $OUT/pytest_grouped.log 2>&1; echo "pytest grouped exit=$?"; tail -4 $OUT/pytest_grouped.log | cut -c1-300
for mode in "" "--smooth"; do
timeout 200 python tools/kernel_bench.py --config 2 --iters 30 $mode --only build_fused,build_grouped,lookup_tiled,lookup_tiled_nhwc,lookup_grouped,lookup_grouped_nhwc 2>&1 | cut -c1-200 | sed "s/^/$mode /" | tee -a $OUT/kb_c2.jsonl
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 200 ncu --metrics $M --clock-control none -k regex:"lookup_tiled|volume_gemm" --csv --log-file $OUT/ncu_grouped.csv python tools/kernel_bench.py --config 2 --iters 1 --warmup 0 --only build_fused,build_grouped,lookup_tiled_nhwc,lookup_grouped,lookup_grouped_nhwc > $OUT/ncu_grouped.log 2>&1; echo "ncu exit=$?"
