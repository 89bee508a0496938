#!/bin/bash
# persistent two-queue kernel: parity, timing (lone frame / steady state), other sizes
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6
rm -f gpurun_out/persist_timing.jsonl
for w in "1,1" "1" "1,1,1" "3,2"; do MIPB200_CHUNK_WEIGHTS=$w timeout 300 python tools/chunk_sweep.py 1920x1080 96 | tee -a gpurun_out/persist_timing.jsonl; done
MIPB200_CHUNK_WEIGHTS=1,1 timeout 300 python tools/chunk_sweep.py 3840x2160 32 | tee -a gpurun_out/persist_timing.jsonl
MIPB200_CHUNK_WEIGHTS=1,1 timeout 300 python tools/chunk_sweep.py 7680x4320 12 | tee -a gpurun_out/persist_timing.jsonl
MIPB200_CHUNK_WEIGHTS=1,1 timeout 300 python tools/chunk_sweep.py 416x240 200 | tee -a gpurun_out/persist_timing.jsonl
