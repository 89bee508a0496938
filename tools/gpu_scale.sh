#!/bin/bash
# bench.py on N GPUs of one box the way the driver launches it (torchrun, one rank per GPU).  Usage: gpu_scale.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 6 --warmup 3 \
  > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc=$?"; grep '"metric"' gpurun_out/bench_${N}gpu.json | tail -1 | cut -c1-1500
