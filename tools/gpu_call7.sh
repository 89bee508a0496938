#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -4
for w in "1,1,1" "1,1,1,1" "4,3,2,1"; do MIPB200_CHUNK_WEIGHTS=$w python tools/chunk_sweep.py 1920x1080 96 | tee -a gpurun_out/units2_timing.jsonl; done
for nt in 384 256; do timeout 120 tools/bin/microbench_tcgen05_$nt gpurun_out/microbench_tcgen05_$nt.json; done
