#!/bin/bash
# One full ncu capture (with source) of the fused kernel, after a plain run of the same command.
set -u
mkdir -p gpurun_out
python tools/profile_run.py 2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mip_cost_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_cost python tools/profile_run.py 2 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; cat gpurun_out/plain2.log
