#!/bin/bash
# GPU-box pass: full parity suite, tensor-core microbenchmark, CLI energy report, kernel timing.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
true
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, "vvc-mip-gpu_b200")
from mipb200 import frames
fs = np.stack([frames.natural_frame(1920, 1080, 100 + i) for i in range(8)])
with open("/tmp/in1080.u16", "wb") as f:
    for i in range(64):
        f.write(fs.astype("<u2").tobytes())
PY
timeout 300 vvc-mip-gpu_b200/bin/mipb200_main -f 512 -s 1920x1080 -o /tmp/in1080.u16 --InputFormat=u16 --NoLog --Energy --StageStamps=0 2>&1 | grep -v "Current frame" | tail -12 | tee gpurun_out/cli_energy_1080p.txt
timeout 300 vvc-mip-gpu_b200/bin/mipb200_main -f 512 -s 1920x1080 -o /tmp/in1080.u16 --InputFormat=u16 --NoLog --DecisionsLog=/dev/null --Energy --StageStamps=0 2>&1 | grep -v "Current frame" | tail -8 | tee -a gpurun_out/cli_energy_1080p.txt
python tools/profile_run.py 20 2>&1 | tail -2
