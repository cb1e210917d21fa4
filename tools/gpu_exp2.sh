#!/bin/bash
for ld in 0 1 2 3 4; do echo "== LD=$ld"; FFCORR_LOOKUP_LD=$ld python tools/kernel_bench.py --config 2 2>&1 | grep lookup; done
echo "== LD=1 L2 fetch 32"; FFCORR_LOOKUP_LD=1 FFCORR_L2_FETCH=32 python tools/kernel_bench.py --config 2 2>&1 | grep lookup
mkdir -p gpurun_out/exp2
M=lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
for ld in 0 1; do
FFCORR_LOOKUP_LD=$ld ncu --metrics $M --clock-control none -k regex:lookup_kernel -c 1 python tools/kernel_bench.py --config 2 --iters 1 --warmup 0 2>&1 | grep -E "lts__|dram__|l1tex__|gpu__time|warps_active|issue_active|inst_executed" | sed "s/^/LD=$ld /"
done
FFCORR_L2_FETCH=32 ncu --metrics $M --clock-control none -k regex:lookup_kernel -c 1 python tools/kernel_bench.py --config 2 --iters 1 --warmup 0 2>&1 | grep -E "lts__|dram__|gpu__time" | sed "s/^/L2F32 /"
