#!/bin/bash
# The final kernel of the round on 1 and 2 GPUs of one box: in-process CLI (config 5 and 1080p, decisions, no hashing) and
# bench.py under torchrun with 2 ranks.
set -u
mkdir -p gpurun_out
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, "vvc-mip-gpu_b200")
from mipb200 import frames
base = [frames.natural_frame(1920, 1080, 100 + i) for i in range(31)]
with open("/dev/shm/in1080.u16", "wb") as f:
    for i in range(16):
        f.write(base[i].astype("<u2").tobytes())
with open("/dev/shm/in4320.u16", "wb") as f:
    for i in range(16):
        f.write(np.ascontiguousarray(np.block([[base[i + 4 * r + c] for c in range(4)] for r in range(4)])).astype("<u2").tobytes())
PY
M=vvc-mip-gpu_b200/bin/mipb200_main
F="--UseAlternativeSamples=1 --FilterType=filterFrame_2d_float_5x5_quarterCtu --KernelIdx=2"
KEEP="Throughput|Energy per frame|Peak host|ERROR"
out=gpurun_out/r02_cli_final_kernel_1_2gpu.txt
: > $out
for g in 1 2; do
  echo "== config 5: -f $((256 * g)) -s 7680x4320, decisions to the host, no per-result hashing (--DecisionsBin=/dev/null), --NumGpus=$g" | tee -a $out
  timeout 300 $M -f $((256 * g)) -s 7680x4320 -o /dev/shm/in4320.u16 --InputFormat=u16 --InputFrames=16 --NoLog --DecisionsBin=/dev/null --Energy --StageStamps=0 --NumGpus=$g $F 2>&1 | grep -E "$KEEP" | tee -a $out
done
for g in 1 2; do
  echo "== 1080p decisions, no per-result hashing, --NumGpus=$g" | tee -a $out
  timeout 300 $M -f $((4096 * g)) -s 1920x1080 -o /dev/shm/in1080.u16 --InputFormat=u16 --InputFrames=16 --NoLog --DecisionsBin=/dev/null --Energy --StageStamps=0 --NumGpus=$g $F 2>&1 | grep -E "$KEEP" | tee -a $out
done
for g in 1 2; do
  echo "== 1080p full int32 tables (--BinaryLog=/dev/null) and compact tables (--CompactLog=/dev/null), --NumGpus=$g" | tee -a $out
  timeout 300 $M -f $((1024 * g)) -s 1920x1080 -o /dev/shm/in1080.u16 --InputFormat=u16 --InputFrames=16 --NoLog --BinaryLog=/dev/null --StageStamps=0 --NumGpus=$g $F 2>&1 | grep -E "Throughput|ERROR" | tee -a $out
  timeout 300 $M -f $((1024 * g)) -s 1920x1080 -o /dev/shm/in1080.u16 --InputFormat=u16 --InputFrames=16 --NoLog --CompactLog=/dev/null --StageStamps=0 --NumGpus=$g $F 2>&1 | grep -E "Throughput|ERROR" | tee -a $out
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_2gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e_costs"]["value"], d["e2e_costs"]["d2h_link"]["ceiling"], d["e2e_costs_compact"]["value"], [(s["value"], s["e2e"]["value"]) for s in d["sizes"]], d["shard_check"])
PY
