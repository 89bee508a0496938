#!/bin/bash
set -u
mkdir -p gpurun_out
for nt in 768 384 256; do timeout 120 tools/bin/microbench_tcgen05_$nt gpurun_out/microbench_tcgen05_$nt.json; done
