#!/bin/bash
# round 2, GPU call D: full GPU suite; fp16-storage kernels after the row-pair lookup; bench config 2 (fp32 / fp16 storage) and config 4
cd /root/repo
OUT=gpurun_out/r2d; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest.log 2>&1; echo "pytest exit=$?" | tee -a $OUT/pytest.log
grep -E "passed|failed|^FAILED|^ERROR|^e2e_full|^trajectory|^FF-PWC" $OUT/pytest.log | cut -c1-220 | head -40
timeout 300 python tools/kernel_bench.py --config 2 --iters 30 --only build_fused,build_fused_f16,lookup_tiled,lookup_tiled_nhwc,lookup_tiled_f16 2>&1 | cut -c1-260 | tee $OUT/kb_c2.jsonl
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
timeout 300 ncu --metrics $M --clock-control none -k regex:"lookup_tiled|volume_gemm|operand_" --csv --log-file $OUT/ncu_f16.csv python tools/kernel_bench.py --config 2 --iters 1 --warmup 0 --only build_fused,build_fused_f16,lookup_tiled_nhwc,lookup_tiled_f16 > $OUT/ncu_f16.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?"; cut -c1-3500 $OUT/bench.json; tail -2 $OUT/bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --storage fp16 --no-cpu-baseline --no-stock --no-pwc > $OUT/bench_f16.json 2> $OUT/bench_f16.err; echo "bench f16 exit=$?"; cut -c1-2500 $OUT/bench_f16.json; tail -2 $OUT/bench_f16.err
timeout 600 python bench.py --config 4 --steps 2 --warmup 1 > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo "bench c4 exit=$?"; cut -c1-1800 $OUT/bench_c4.json; tail -2 $OUT/bench_c4.err
