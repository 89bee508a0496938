#!/bin/bash
# The final kernel of the round on all 8 GPUs of one box, kept short (8x charge): in-process CLI (config 5 and 1080p,
# decisions, no hashing) and bench.py under torchrun with 8 ranks.
set -u
mkdir -p gpurun_out
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, "vvc-mip-gpu_b200")
from mipb200 import frames
base = [frames.natural_frame(1920, 1080, 100 + i) for i in range(19)]
with open("/dev/shm/in1080.u16", "wb") as f:
    for i in range(16):
        f.write(base[i].astype("<u2").tobytes())
with open("/dev/shm/in4320.u16", "wb") as f:
    for i in range(4):
        f.write(np.ascontiguousarray(np.block([[base[(i + 4 * r + c) % 19] for c in range(4)] for r in range(4)])).astype("<u2").tobytes())
PY
M=vvc-mip-gpu_b200/bin/mipb200_main
F="--UseAlternativeSamples=1 --FilterType=filterFrame_2d_float_5x5_quarterCtu --KernelIdx=2"
KEEP="Throughput|Energy per frame|Peak host|ERROR"
out=gpurun_out/r02_cli_final_kernel_8gpu.txt
: > $out
echo "== config 5: -f 2048 -s 7680x4320 (4-frame pool cycled), decisions to the host, no per-result hashing, --NumGpus=8" | tee -a $out
timeout 120 $M -f 2048 -s 7680x4320 -o /dev/shm/in4320.u16 --InputFormat=u16 --InputFrames=4 --NoLog --DecisionsBin=/dev/null --Energy --StageStamps=0 --NumGpus=8 $F 2>&1 | grep -E "$KEEP" | tee -a $out
echo "== 1080p decisions, -f 32768, no per-result hashing, --NumGpus=8" | tee -a $out
timeout 120 $M -f 32768 -s 1920x1080 -o /dev/shm/in1080.u16 --InputFormat=u16 --InputFrames=16 --NoLog --DecisionsBin=/dev/null --Energy --StageStamps=0 --NumGpus=8 $F 2>&1 | grep -E "$KEEP" | tee -a $out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_8gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e_costs"]["value"], d["e2e_costs"]["d2h_link"]["ceiling"], d["e2e_costs_compact"]["value"], [(s["value"], s["e2e"]["value"]) for s in d["sizes"]], d["shard_check"])
PY
