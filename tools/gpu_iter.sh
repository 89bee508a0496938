#!/bin/bash
# kernel iteration: parity subset + steady-state / lone-frame timing for a list of throughput splits
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -4
rm -f gpurun_out/iter_timing.jsonl
for w in ${SPLITS:-"1,1,1" "1,1" "1" "3,2" "1,1,1,1"}; do MODE=throughput MIPB200_CHUNK_WEIGHTS=$w timeout 300 python tools/chunk_sweep.py 1920x1080 96 | tee -a gpurun_out/iter_timing.jsonl; done
MODE=latency timeout 300 python tools/chunk_sweep.py 1920x1080 96 | tee -a gpurun_out/iter_timing.jsonl
