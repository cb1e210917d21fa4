#!/bin/bash
# round 2, GPU call A: full GPU test suite, bench (both arms), kernel table, gather micro-benchmark, lookup ncu counters
cd /root/repo
OUT=gpurun_out/r2a; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > $OUT/pytest.log 2>&1; echo "pytest exit=$?" | tee -a $OUT/pytest.log
tail -25 $OUT/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?"; cat $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref exit=$?"; cat $OUT/bench_ref.json
timeout 300 python tools/kernel_bench.py --config 2 --iters 20 --only lookup,lookup_tiled,lookup_tiled_nhwc,build_fused > $OUT/kb_c2.jsonl 2>&1; cat $OUT/kb_c2.jsonl
timeout 300 python tools/kernel_bench.py --config 2 --iters 20 --smooth --only lookup_tiled,lookup_tiled_nhwc > $OUT/kb_c2_smooth.jsonl 2>&1; cat $OUT/kb_c2_smooth.jsonl
timeout 120 tools/mb/mb_gather > $OUT/mb_gather.txt 2>&1; cat $OUT/mb_gather.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_sectors.sum,l1tex__m_l1tex2xbar_write_sectors.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__registers_per_thread,launch__grid_size
timeout 600 ncu --metrics $M --clock-control none -k regex:"lookup" --csv --log-file $OUT/ncu_lookup_counters.csv python tools/kernel_bench.py --config 2 --iters 1 --warmup 0 --only lookup,lookup_tiled,lookup_tiled_nhwc > $OUT/ncu_lookup.log 2>&1; echo "ncu lookup exit=$?"
timeout 300 ncu --metrics $M --clock-control none --csv --log-file $OUT/ncu_mb_gather_counters.csv tools/mb/mb_gather > $OUT/ncu_mb.log 2>&1; echo "ncu mb exit=$?"
ls -la $OUT
