#!/bin/bash
# A/B of kernel variants built as vvc-mip-gpu_b200/lib/<name>.so: parity tests against PARITY, steady-state and lone-frame
# timing of every name in VARIANTS (tools/chunk_sweep.py), and the executed warp instructions of one launch of each (ncu).
set -u
mkdir -p gpurun_out
out=gpurun_out/ab2_timing.jsonl
rm -f $out gpurun_out/ab2_inst.txt
lib() { echo $PWD/vvc-mip-gpu_b200/lib/$1.so; }
if [ -n "${PARITY:-}" ]; then MIPB200_LIB=$(lib $PARITY) timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/ab2_parity.txt; fi
for rep in ${REPS:-1 2}; do
  for v in ${VARIANTS:-base}; do
    for m in throughput latency; do
      echo -n "$v $m: " | tee -a $out
      MIPB200_LIB=$(lib $v) MODE=$m timeout 300 python tools/chunk_sweep.py 1920x1080 96 | tee -a $out
    done
  done
done
for v in ${VARIANTS:-base}; do
  echo "== $v" >> gpurun_out/ab2_inst.txt
  MIPB200_LIB=$(lib $v) MODE=throughput STREAMS=1 timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none \
    -k regex:mip_cost_kernel -s 8 -c 1 python tools/chunk_sweep.py 1920x1080 4 2>&1 | grep -E "inst_executed|time_duration" >> gpurun_out/ab2_inst.txt
done
cat gpurun_out/ab2_inst.txt
