#!/bin/bash
# The CLI sharding frames over all GPUs of the box (one host thread per GPU, frames page-locked once): decisions only.
# Keep the frame count modest: the decisions log of EVERY frame is formatted after the timed window (726 300 lines per
# 1080p frame), and an N-GPU gpurun call is charged N times its wall time.
set -u
N=${1:-8}
mkdir -p gpurun_out
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, "vvc-mip-gpu_b200")
from mipb200 import frames
fs = np.stack([frames.natural_frame(1920, 1080, 100 + i) for i in range(8)]).astype("<u2").tobytes()
with open("/tmp/in1080.u16", "wb") as f:
    for i in range(64):
        f.write(fs)
PY
M=vvc-mip-gpu_b200/bin/mipb200_main
F="--UseAlternativeSamples=1 --FilterType=filterFrame_2d_float_5x5_quarterCtu --KernelIdx=2"
: > gpurun_out/cli_multi_gpu.txt
for g in 1 $N; do
  echo "== mipb200_main -f 512 -s 1920x1080 --NumGpus=$g (decisions)" | tee -a gpurun_out/cli_multi_gpu.txt
  timeout 600 $M -f 512 -s 1920x1080 -o /tmp/in1080.u16 --InputFormat=u16 --NoLog --DecisionsLog=/dev/null --Energy --StageStamps=0 --NumGpus=$g $F 2>&1 \
    | grep -E "Throughput|Energy per frame|power|Elapsed" | tee -a gpurun_out/cli_multi_gpu.txt
done
