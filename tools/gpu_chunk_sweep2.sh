#!/bin/bash
# Fine sweep of the two chunk splits around their defaults with the final kernel of the round.
set -u
mkdir -p gpurun_out
out=gpurun_out/chunk_sweep2.jsonl
: > $out
for w in "1,1,1" "1.06,1,1" "0.94,1,1" "1,1,1.06" "1,1,0.94" "1,1.06,1" "1,1,1,1"; do
  MODE=throughput STREAMS=3 MIPB200_CHUNK_WEIGHTS=$w python tools/chunk_sweep.py 1920x1080 96 2>&1 | tail -1 | tee -a $out
done
for w in "4,3,2,1" "5,4,3,2,1" "4,3,2,1,1" "9,7,5,3" "4,3,2,1.5" "5,3,2,1" "8,6,4,3,2,1"; do
  MODE=latency STREAMS=1 MIPB200_CHUNK_WEIGHTS_LONE=$w python tools/chunk_sweep.py 1920x1080 96 2>&1 | tail -1 | tee -a $out
done
