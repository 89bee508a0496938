#!/bin/bash
# chunk-count sweep at several resolutions (3 streams)
set -u
for res in 1920x1080 3840x2160 7680x4320; do
  for ch in 1 2 3 4 6; do
    n=30; [ $res = 3840x2160 ] && n=12; [ $res = 7680x4320 ] && n=6
    printf "%-10s chunks=%-2s " $res $ch
    STREAMS=3 MIPB200_CHUNKS=$ch python tools/profile_run.py $n $res 2>&1 | tail -1
  done
done
