#!/bin/bash
# round 2, the 8-GPU call (gpurun --gpus 8): host-link ceiling on 1/2/4/8 GPUs, BASELINE config 5 through the product at
# 1/2/4/8 GPUs with G-independence checks, config 4, full-table scaling, one 8-rank bench line.
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | sed -n 2p
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, "vvc-mip-gpu_b200")
from mipb200 import frames
base = [frames.natural_frame(1920, 1080, 100 + i) for i in range(31)]
with open("/dev/shm/in1080.u16", "wb") as f:
    for i in range(16):
        f.write(base[i].astype("<u2").tobytes())
with open("/dev/shm/in4320.u16", "wb") as f:          # 16 distinct 8K frames: 4x4 mosaics of different 1080p frames
    for i in range(16):
        f.write(np.ascontiguousarray(np.block([[base[i + 4 * r + c] for c in range(4)] for r in range(4)])).astype("<u2").tobytes())
with open("/dev/shm/in2160.u16", "wb") as f:
    for i in range(16):
        f.write(np.ascontiguousarray(np.block([[base[i + 2 * r + c] for c in range(2)] for r in range(2)])).astype("<u2").tobytes())
PY
tools/bin/d2h_bench 2 > gpurun_out/r02_d2h_ceiling.jsonl 2>&1; cat gpurun_out/r02_d2h_ceiling.jsonl
M=vvc-mip-gpu_b200/bin/mipb200_main
F="--UseAlternativeSamples=1 --FilterType=filterFrame_2d_float_5x5_quarterCtu --KernelIdx=2"
KEEP="Frame ring|Throughput|Energy per frame|Average power|Elapsed|Peak host|ERROR"
for g in 8 4 2 1; do
  out=gpurun_out/r02_cli_4320p_${g}gpu.txt
  echo "== mipb200_main -f 2048 -s 7680x4320 (16 distinct frames cycled) alternative samples, decisions + digest, --NumGpus=$g" | tee $out
  timeout 300 $M -f 2048 -s 7680x4320 -o /dev/shm/in4320.u16 --InputFormat=u16 --InputFrames=16 --NoLog --Digest=gpurun_out/r02_digest_4320p_${g}gpu.csv --Energy --StageStamps=0 --NumGpus=$g $F 2>&1 | grep -E "$KEEP" | tee -a $out
done
for g in 4 2 1; do cmp gpurun_out/r02_digest_4320p_8gpu.csv gpurun_out/r02_digest_4320p_${g}gpu.csv && echo "digests of 2048 frames: 8 GPUs == $g GPU(s)" | tee -a gpurun_out/r02_cli_4320p_8gpu.txt; done
for g in 1 8; do timeout 300 $M -f 32 -s 7680x4320 -o /dev/shm/in4320.u16 --InputFormat=u16 --InputFrames=16 --RingFrames=8 --NoLog --DecisionsBin=/dev/shm/dec_${g}.bin --StageStamps=0 --NumGpus=$g $F 2>&1 | grep -E "Frame ring|ERROR"; done
cmp /dev/shm/dec_1.bin /dev/shm/dec_8.bin && echo "decisions files (32 frames, 1.76 GB): 1 GPU == 8 GPUs, byte for byte (ring streamed at 1 GPU)" | tee -a gpurun_out/r02_cli_4320p_8gpu.txt
rm -f /dev/shm/dec_*.bin
out=gpurun_out/r02_cli_2160p_1gpu.txt
echo "== BASELINE config 4: mipb200_main -f 256 -s 3840x2160, original samples, 1 GPU" | tee $out
timeout 300 $M -f 256 -s 3840x2160 -o /dev/shm/in2160.u16 --InputFormat=u16 --InputFrames=16 --NoLog --Digest=/dev/null --Energy --StageStamps=0 2>&1 | grep -E "$KEEP" | tee -a $out
out=gpurun_out/r02_cli_1080p_full_tables.txt
: > $out
for g in 1 2 4 8; do
  echo "== 1080p full int32 tables to the host (--BinaryLog=/dev/null), --NumGpus=$g" | tee -a $out
  timeout 300 $M -f $((1024 * g)) -s 1920x1080 -o /dev/shm/in1080.u16 --InputFormat=u16 --InputFrames=16 --NoLog --BinaryLog=/dev/null --StageStamps=0 --NumGpus=$g $F 2>&1 | grep -E "Throughput|ERROR" | tee -a $out
  echo "== 1080p decisions, --NumGpus=$g" | tee -a $out
  timeout 300 $M -f $((4096 * g)) -s 1920x1080 -o /dev/shm/in1080.u16 --InputFormat=u16 --InputFrames=16 --NoLog --Digest=/dev/null --StageStamps=0 --NumGpus=$g $F 2>&1 | grep -E "Throughput|ERROR" | tee -a $out
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "bench8 rc=$?"
tail -c 3000 gpurun_out/r02_bench_8gpu.json
