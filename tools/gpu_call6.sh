#!/bin/bash
# unit-based kernel: parity, timing, traffic
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6
for w in "1,1,1" "1,1,1,1" "4,3,2,1" "1,1"; do MIPB200_CHUNK_WEIGHTS=$w python tools/chunk_sweep.py 1920x1080 96 | tee -a gpurun_out/units_timing.jsonl; done
python tools/traffic_run.py 30 > gpurun_out/plain3.log 2>&1 && \
ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/r02_traffic_range.csv python tools/traffic_run.py 30 > gpurun_out/ncu_range.log 2>&1
echo "range rc=$?"; cat gpurun_out/plain3.log; tail -6 gpurun_out/r02_traffic_range.csv
