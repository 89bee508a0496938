#!/bin/bash
# Profiling evidence of the round: launch list of a short bench run, one full capture of the fused kernel, and the
# steady-state DRAM traffic of 30 back-to-back launches (range replay).  Every ncu run follows a plain run of the same command.
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sizes > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sizes > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -c 300 gpurun_out/plain_bench.log
python tools/profile_run.py 2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mip_cost_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_cost python tools/profile_run.py 2 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; cat gpurun_out/plain2.log
python tools/traffic_run.py 30 > gpurun_out/plain3.log 2>&1 && \
ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r02_traffic_range.csv python tools/traffic_run.py 30 > gpurun_out/ncu_range.log 2>&1
echo "range rc=$?"; cat gpurun_out/plain3.log; tail -5 gpurun_out/r02_traffic_range.csv
