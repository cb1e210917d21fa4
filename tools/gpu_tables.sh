#!/bin/bash
# Final measurement tables (no profiler): per-kernel timings at the SURVEY shapes, training-path timings, bench line.
cd /root/repo
OUT=gpurun_out/${1:-tables_r1c}; mkdir -p $OUT
python tools/kernel_bench.py --config 2 --pwc --stock > $OUT/kernel_bench_c2.jsonl 2>$OUT/kb2.err
python tools/kernel_bench.py --config 2 --smooth --only lookup,lookup_tiled > $OUT/kernel_bench_c2_smooth.jsonl 2>>$OUT/kb2.err
python tools/kernel_bench.py --config 4 > $OUT/kernel_bench_c4.jsonl 2>$OUT/kb4.err
python tools/kernel_bench.py --config 1 > $OUT/kernel_bench_c1.jsonl 2>$OUT/kb1.err
python tools/kernel_bench.py --config 2 --all-precisions --only volume,volume_tiled,build_fused > $OUT/kernel_bench_c2_precisions.jsonl 2>>$OUT/kb2.err
python tools/bwd_bench.py --pwc > $OUT/bwd_bench.jsonl 2>$OUT/bwd.err
tools/mb/mb_gather > $OUT/mb_gather.txt 2>&1
tools/mb/mb_scatter_write > $OUT/mb_scatter_write.txt 2>&1
python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
cut -c1-400 $OUT/bench.json; cat $OUT/bench_reference.json | cut -c1-300
wc -l $OUT/*.jsonl
