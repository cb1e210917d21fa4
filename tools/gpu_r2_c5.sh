#!/bin/bash
# round 2: config 5 (DDP training step) on N GPUs of one box + the two-GPU device-guard test.  usage: gpu_r2_c5.sh N
cd /root/repo
N=$1
OUT=gpurun_out/r2_scale; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config 5 --steps 4 --warmup 2 > $OUT/config5_${N}gpu.json 2> $OUT/config5_${N}gpu.err
echo "config5 N=$N exit=$?"; cut -c1-1800 $OUT/config5_${N}gpu.json; grep -iE "error|timeout|Traceback" $OUT/config5_${N}gpu.err | head -5 | cut -c1-300
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_launches_follow" > $OUT/pytest_2gpu.log 2>&1; echo "device-guard test exit=$?"; tail -3 $OUT/pytest_2gpu.log | cut -c1-300
