#!/bin/bash
# round 2, GPU call F (N GPUs): final verification.  N=1: smoke(), the whole GPU suite, the two bench arms.
# N>=2: the two-GPU device-guard test and config 4 (strong scaling) at N.
cd /root/repo
N=${1:-1}
OUT=gpurun_out/r2f; mkdir -p $OUT
if [ $N -eq 1 ]; then
  timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 $OUT/smoke.log | cut -c1-300
  timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit=$?"; grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest.log | cut -c1-200 | head
  timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref exit=$?"; cut -c1-300 $OUT/bench_ref.json
  timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?"; cut -c1-600 $OUT/bench.json
  timeout 300 python bench.py --config 3 --steps 5 --warmup 3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; echo "c3 exit=$?"; cut -c1-300 $OUT/bench_c3.json
else
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_launches_follow or channels_last" > $OUT/pytest_${N}gpu.log 2>&1; echo "device-guard test exit=$?"; tail -3 $OUT/pytest_${N}gpu.log | cut -c1-300
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --config 4 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/config4_${N}gpu.json 2> $OUT/config4_${N}gpu.err; echo "config4 N=$N exit=$?"; cut -c1-400 $OUT/config4_${N}gpu.json
fi
