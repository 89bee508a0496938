#!/bin/bash
# Round-end evidence: launch list of a short bench run + one full capture of the fused kernel.
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
python tools/profile_run.py 2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mip_cost_kernel -s 1 -c 1 -f -o gpurun_out/prof_cost python tools/profile_run.py 2 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
