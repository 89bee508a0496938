#!/bin/bash
# round 2, call 1: baseline bench, sanitizer, chunk sweep, single-GPU D2H ceiling
set -u
mkdir -p gpurun_out
nvidia-smi -L; nproc; free -g | head -2
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_base.json 2> gpurun_out/r02_bench_base.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_base.json
tools/bin/d2h_bench 2 > gpurun_out/d2h_1gpu.jsonl 2>&1; cat gpurun_out/d2h_1gpu.jsonl
bash tools/gpu_chunk_sweep.sh
bash tools/gpu_sanitize.sh
