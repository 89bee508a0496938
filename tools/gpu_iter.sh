#!/bin/bash
# Quick iteration on the GPU box: parity tests, then kernel timing of the hot path.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
python tools/profile_run.py 20 2>&1 | tail -2
python tools/profile_run.py 20 3840x2160 2>&1 | tail -1
if [ "${1:-}" = "mb" ]; then tools/bin/microbench gpurun_out/microbench.json | tail -14; fi
