#!/bin/bash
# targeted captures (no full launch list): lookup / volume / pyramid / prepass of the default (tiled) path
set -u
OUT=gpurun_out/${1:-profiles_r1b}; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $OUT/bench_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"lookup_tiled" -s 14 -c 2 -o $OUT/lookup_full $CMD > $OUT/ncu_lookup.log 2>&1
echo "lookup capture exit=$?"
$CMD > $OUT/bench_plain2.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"volume_gemm|pyramid_tiled|operand_prepass" -s 3 -c 3 -o $OUT/build_full $CMD > $OUT/ncu_build.log 2>&1
echo "build capture exit=$?"
$CMD > $OUT/bench_plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lookup_|volume_gemm|pyramid_|operand_prepass" --csv --log-file $OUT/launches_ffcorr.csv $CMD > $OUT/ncu_l.log 2>&1
echo "ffcorr launch list exit=$?"
python tools/kernel_bench.py --config 2 --pwc > $OUT/kernel_bench_c2.jsonl 2>&1
python tools/kernel_bench.py --config 4 > $OUT/kernel_bench_c4.jsonl 2>&1
python tools/kernel_bench.py --config 1 > $OUT/kernel_bench_c1.jsonl 2>&1
python bench.py --steps 5 --warmup 3 > $OUT/bench_r1.json 2> $OUT/bench_r1.err
cat $OUT/bench_r1.json | cut -c1-300
