#!/bin/bash
# round 2, GPU call E: PWC after FFMA2 (tests + level timings), full GPU suite, ncu evidence:
# (1) launch list of the bench command, (2) --set full of one steady-state lookup launch and of the fused build inside
# bench.py, (3) --set full with source of the PWC level-2 kernel, (4) SASS mnemonic counts.
cd /root/repo
OUT=gpurun_out/r2e; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit=$?"; grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest.log | cut -c1-200 | head
for c in "16 32 112 256" "16 64 56 128" "16 96 28 64" "16 128 14 32" "16 196 7 16"; do timeout 120 python tools/pwc_level.py $c 2>&1 | tail -1; done | tee $OUT/pwc_levels_back_to_back.txt
timeout 200 python tools/bwd_bench.py --pwc 2>&1 | tail -4 | cut -c1-300 | tee $OUT/bwd_bench.txt
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cudnn-benchmark --no-stock --no-pwc"
timeout 300 $CMD > $OUT/bench_plain.json 2> $OUT/bench_plain.err || { echo "plain bench failed"; tail -5 $OUT/bench_plain.err; }
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cudnn-benchmark --no-stock > $OUT/bench_no_autotune.json 2> $OUT/bench_no_autotune.err; cut -c1-400 $OUT/bench_no_autotune.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1; echo "launch list exit=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"lookup_tiled" -s 14 -c 1 -o $OUT/lookup_full $CMD > $OUT/ncu_lookup.log 2>&1; echo "lookup full exit=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"volume_gemm|operand_" -s 3 -c 3 -o $OUT/build_full $CMD > $OUT/ncu_build.log 2>&1; echo "build full exit=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"pwc81" -s 3 -c 1 -o $OUT/pwc_l2_full python tools/pwc_level.py 16 32 112 256 > $OUT/ncu_pwc.log 2>&1; echo "pwc full exit=$?"
cuobjdump -sass focusflow_official_b200/libffcorr.so > $OUT/sass.txt 2>/dev/null; python - <<'PY'
import re, collections
cur=None; c=collections.OrderedDict()
for line in open("gpurun_out/r2e/sass.txt"):
    m=re.search(r"Function : (\S+)", line)
    if m: cur=m.group(1); c[cur]=collections.Counter(); continue
    m=re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur: c[cur][m.group(1)]+=1
keys=["UTCHMMA","LDTM","UTMALDG","UTMASTG","UBLKCP","LDGSTS","SYNCS","UCGABAR","FFMA2","FFMA","HMMA"]
with open("gpurun_out/r2e/sass_counts.txt","w") as f:
    f.write("# cuobjdump -sass focusflow_official_b200/libffcorr.so: mnemonic counts per kernel\n")
    for k,v in c.items():
        row=" ".join(f"{m}={v[m]}" for m in keys if v[m])
        if row: f.write(f"{k[:110]}: {row}\n")
PY
rm -f $OUT/sass.txt; wc -l $OUT/sass_counts.txt; ls -la $OUT | head -30
