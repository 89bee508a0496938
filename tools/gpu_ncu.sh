#!/bin/bash
# One full ncu capture of the fused cost kernel (after the same command ran clean without ncu).
set -u
mkdir -p gpurun_out
python tools/profile_run.py 2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mip_cost_kernel -s 1 -c 1 -f -o gpurun_out/prof_cost python tools/profile_run.py 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; cat gpurun_out/plain2.log; tail -3 gpurun_out/ncu_full.log
