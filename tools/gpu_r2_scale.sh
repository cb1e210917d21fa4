#!/bin/bash
# round 2: multi-GPU runs.  usage: gpu_r2_scale.sh N  (run under `gpurun --gpus N`)
cd /root/repo
N=$1
OUT=gpurun_out/r2_scale; mkdir -p $OUT
run() {  # run <tag> <bench args...>
  tag=$1; shift
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 "$@" > $OUT/${tag}_${N}gpu.json 2> $OUT/${tag}_${N}gpu.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > $OUT/${tag}_${N}gpu.json 2> $OUT/${tag}_${N}gpu.err
  fi
  echo "$tag N=$N exit=$?"; cut -c1-700 $OUT/${tag}_${N}gpu.json; tail -2 $OUT/${tag}_${N}gpu.err | cut -c1-300
}
run config4 --config 4 --steps 2 --warmup 1 --no-cpu-baseline
run config5 --config 5 --steps 4 --warmup 2
run config2 --config 2 --steps 4 --warmup 3 --no-cpu-baseline --no-stock --no-pwc
if [ $N -ge 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tensors_device" > $OUT/pytest_2gpu.log 2>&1; echo "2-GPU device-guard test exit=$?"; tail -3 $OUT/pytest_2gpu.log
fi
