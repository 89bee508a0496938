#!/bin/bash
# One gpurun call: bench line, pipe micro-benchmarks, ncu launch list, one full ncu capture of the cost kernel.
set -u
mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
tools/bin/microbench gpurun_out/microbench.json > gpurun_out/microbench.log 2>&1; echo "microbench rc=$?"
cat gpurun_out/microbench.log
python tools/profile_run.py 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_run.py 3 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"; cat gpurun_out/plain.log; tail -20 gpurun_out/launches.csv
python tools/profile_run.py 1 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mip_cost_kernel -s 1 -c 1 -f -o gpurun_out/prof_cost python tools/profile_run.py 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -5 gpurun_out/ncu_full.log
