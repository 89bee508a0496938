#!/bin/bash
# CLI throughput and energy per frame on the GPU box: 512 x 1080p frames (raw u16 input), full-table readback and decisions only,
# original samples and the bench configuration.  Output: gpurun_out/cli_energy_1080p.txt
set -u
mkdir -p gpurun_out
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, "vvc-mip-gpu_b200")
from mipb200 import frames
fs = np.stack([frames.natural_frame(1920, 1080, 100 + i) for i in range(8)])
with open("/tmp/in1080.u16", "wb") as f:
    for i in range(64):
        f.write(fs.astype("<u2").tobytes())
PY
M=vvc-mip-gpu_b200/bin/mipb200_main
C="-f 512 -s 1920x1080 -o /tmp/in1080.u16 --InputFormat=u16 --NoLog --Energy --StageStamps=0"
F="--UseAlternativeSamples=1 --FilterType=filterFrame_2d_float_5x5_quarterCtu --KernelIdx=2"
: > gpurun_out/cli_energy_1080p.txt
for mode in "" "--DecisionsLog=/dev/null" "$F" "$F --DecisionsLog=/dev/null"; do
  echo "== mipb200_main $C $mode" | tee -a gpurun_out/cli_energy_1080p.txt
  timeout 300 $M $C $mode 2>&1 | grep -E "Throughput|Energy|power|Elapsed" | tee -a gpurun_out/cli_energy_1080p.txt
done
