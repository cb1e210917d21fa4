#!/bin/bash
# round 2, GPU call R (2 GPUs): the two-GPU device-guard tests and the default bench under torchrun
cd /root/repo
OUT=gpurun_out/r2r; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_conv.py -q -m gpu -k "test_launches_follow or channels_last or lookup_conv_is" > $OUT/pytest_2gpu.log 2>&1; echo "2-GPU tests exit=$?"; tail -2 $OUT/pytest_2gpu.log | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err; echo "bench N=2 exit=$?"; cut -c1-260 $OUT/bench_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $OUT/bench_ref_2gpu.json 2> $OUT/bench_ref_2gpu.err; echo "ref N=2 exit=$?"; cut -c1-200 $OUT/bench_ref_2gpu.json
