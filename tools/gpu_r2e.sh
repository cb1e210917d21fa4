#!/bin/bash
# round 2, GPU call E: ncu evidence.  (1) launch list of the bench command, (2) --set full of one steady-state lookup launch
# and of the fused build inside bench.py, (3) --set full with source of the PWC level-2 kernel, (4) SASS mnemonic counts.
cd /root/repo
OUT=gpurun_out/r2e; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cudnn-benchmark --no-stock --no-pwc"
$CMD > $OUT/bench_plain.json 2> $OUT/bench_plain.err || { echo "plain bench failed"; tail -5 $OUT/bench_plain.err; exit 1; }
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cudnn-benchmark --no-stock --no-pwc > $OUT/bench_no_autotune.json 2> $OUT/bench_no_autotune.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1; echo "launch list exit=$?"
ncu --set full --import-source on --clock-control none -k regex:"lookup_tiled" -s 14 -c 1 -o $OUT/lookup_full $CMD > $OUT/ncu_lookup.log 2>&1; echo "lookup full exit=$?"
ncu --set full --import-source on --clock-control none -k regex:"volume_gemm|operand_" -s 3 -c 3 -o $OUT/build_full $CMD > $OUT/ncu_build.log 2>&1; echo "build full exit=$?"
python tools/pwc_level.py 16 32 112 256 > $OUT/pwc_l2_plain.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:"pwc81" -s 3 -c 1 -o $OUT/pwc_l2_full python tools/pwc_level.py 16 32 112 256 > $OUT/ncu_pwc.log 2>&1; echo "pwc full exit=$?"
for c in "16 64 56 128" "16 96 28 64" "16 128 14 32" "16 196 7 16"; do python tools/pwc_level.py $c 2>&1 | tail -1; done | tee $OUT/pwc_levels_back_to_back.txt
cat $OUT/pwc_l2_plain.log | tail -1
cuobjdump -sass focusflow_official_b200/libffcorr.so | grep -E "Function|UTCHMMA|LDTM|UTMALDG|UTMASTG|UBLKCP|LDGSTS|SYNCS|UCGABAR|UTMAPF" | awk '/Function/{f=$0} !/Function/{split($0,a," "); for(i in a) if (a[i] ~ /^(UTCHMMA|LDTM|UTMALDG|UTMASTG|UBLKCP|LDGSTS|SYNCS|UCGABAR)/) {split(a[i],b,"."); c[f" "b[1]]++}} END{for(k in c) print c[k], k}' | sort -k2 | cut -c1-200 > $OUT/sass_counts.txt; wc -l $OUT/sass_counts.txt
ls -la $OUT | head -30
