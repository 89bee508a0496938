#!/bin/bash
# A/B timing of build variants: MIPB200_LIB selects the library, MIPB200_CHUNKS the work split, STREAMS the overlap.
set -u
L=vvc-mip-gpu_b200/lib
CH=${CHUNKS:-4 8}
NS=${NSTREAMS:-1 3}
for lib in "$@"; do
  for ch in $CH; do
    for ns in $NS; do
      printf "%-22s chunks=%-3s " "$lib" "$ch"
      MIPB200_VERBOSE=1 STREAMS=$ns MIPB200_LIB=$PWD/$L/$lib MIPB200_CHUNKS=$ch python tools/profile_run.py 30 2>&1 | tail -2 | tr '\n' ' '; echo
    done
  done
done
