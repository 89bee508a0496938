#!/bin/bash
# final evidence of the round: driver-style run (tests, smoke, both bench arms), ncu captures, soak
set -u
mkdir -p gpurun_out
bash tools/gpu_round.sh
bash tools/gpu_ncu_r02.sh
python tools/soak.py 12 gpurun_out/r02_soak_decisions; echo "soak rc=$?"
for m in throughput latency; do MODE=$m timeout 300 python tools/chunk_sweep.py 1920x1080 96; done | tee gpurun_out/r02_launch_modes_final.jsonl
