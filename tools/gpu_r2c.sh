#!/bin/bash
# round 2, GPU call C: full GPU suite; fp16-storage kernels; lookup variants (L2 prefetch hint, stages, warps); bench fp32 / fp16 storage; configs 3/4/5
cd /root/repo
OUT=gpurun_out/r2c; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest.log 2>&1; echo "pytest exit=$?" | tee -a $OUT/pytest.log
grep -E "passed|failed|^FAILED|^ERROR|^e2e_full|^trajectory|^FF-PWC" $OUT/pytest.log | cut -c1-220 | head -60
timeout 300 python tools/kernel_bench.py --config 2 --iters 30 --only build_fused,build_fused_f16,lookup_tiled,lookup_tiled_nhwc,lookup_tiled_f16 2>&1 | cut -c1-260 | tee $OUT/kb_c2.jsonl
for v in l2_64 l2_128 s4 s2 w4; do
  timeout 300 python tools/kernel_bench.py --config 2 --iters 30 --only lookup_tiled,lookup_tiled_nhwc,lookup_tiled_f16 --lib build/variants/libffcorr_$v.so 2>&1 | cut -c1-120 | sed "s/^/$v /" | tee -a $OUT/variants.txt
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
timeout 300 ncu --metrics $M --clock-control none -k regex:"lookup_tiled|volume_gemm" --csv --log-file $OUT/ncu_f16.csv python tools/kernel_bench.py --config 2 --iters 1 --warmup 0 --only build_fused,build_fused_f16,lookup_tiled_nhwc,lookup_tiled_f16 > $OUT/ncu_f16.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?"; cut -c1-3000 $OUT/bench.json
timeout 600 python bench.py --steps 5 --warmup 3 --storage fp16 --no-cpu-baseline --no-stock --no-pwc > $OUT/bench_f16.json 2> $OUT/bench_f16.err; echo "bench f16 exit=$?"; cut -c1-2500 $OUT/bench_f16.json
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "bench c3 exit=$?"; cut -c1-2500 $OUT/bench_c3.json; tail -3 $OUT/bench_c3.err
timeout 600 python bench.py --config 4 --steps 2 --warmup 1 > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo "bench c4 exit=$?"; cut -c1-1500 $OUT/bench_c4.json; tail -3 $OUT/bench_c4.err
timeout 600 python bench.py --config 5 --steps 3 --warmup 2 > $OUT/bench_c5.json 2> $OUT/bench_c5.err; echo "bench c5 exit=$?"; cut -c1-2500 $OUT/bench_c5.json; tail -3 $OUT/bench_c5.err
