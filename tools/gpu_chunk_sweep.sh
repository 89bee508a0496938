#!/bin/bash
# Sweep of the chunk split (CTAs per CTU half and their cost shares): lone-frame and steady-state ms per 1080p frame.
set -u
mkdir -p gpurun_out
out=gpurun_out/chunk_sweep.jsonl
: > $out
for w in "1,1,1" "1,1,1,1" "4,3,2,1" "3,3,2,1" "2,2,1,1" "4,4,3,2,1" "3,3,3,2,1" "2,2,2,1,1" "5,4,3,2,1,1" "3,2,1"; do
  MIPB200_CHUNK_WEIGHTS=$w python tools/chunk_sweep.py 1920x1080 96 2>&1 | tail -1 | tee -a $out
done
