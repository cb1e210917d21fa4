#!/bin/bash
# Round profile artefacts: launch list of one bench step + full ncu capture of the top hot-path kernel.
set -u
OUT=gpurun_out/${1:-profiles_r1}; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > $OUT/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "launch list exit=$?"
$CMD > $OUT/bench_plain2.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"lookup_kernel" -s 14 -c 2 -o $OUT/lookup_full $CMD > $OUT/ncu_lookup.log 2>&1
echo "lookup capture exit=$?"
$CMD > $OUT/bench_plain3.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"volume_gemm|pyramid_bulk|operand_prepass" -s 3 -c 3 -o $OUT/build_full $CMD > $OUT/ncu_build.log 2>&1
echo "build capture exit=$?"
python bench.py --steps 5 --warmup 3 > $OUT/bench_r1.json 2> $OUT/bench_r1.err
cat $OUT/bench_r1.json
ls -la $OUT
