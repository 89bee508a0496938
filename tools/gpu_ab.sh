#!/bin/bash
# A/B timing of kernel variants built as vvc-mip-gpu_b200/lib/var_<name>.so (MIPB200_LIB selects the library).
# VARIANTS="name[:weights] ..." ; PARITY=name runs the parity tests against that library first.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/ab_timing.jsonl
if [ -n "${PARITY:-}" ]; then MIPB200_LIB=$PWD/vvc-mip-gpu_b200/lib/var_$PARITY.so timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -3; fi
for rep in ${REPS:-1 2}; do
for v in ${VARIANTS:-base}; do
  n=${v%%:*}; w=""; [ "$v" != "$n" ] && w=${v#*:}
  echo -n "$v: " | tee -a gpurun_out/ab_timing.jsonl
  if [ -n "$w" ]; then MIPB200_CHUNK_WEIGHTS=$w MIPB200_LIB=$PWD/vvc-mip-gpu_b200/lib/var_$n.so MODE=throughput timeout 300 python tools/chunk_sweep.py 1920x1080 96 | tee -a gpurun_out/ab_timing.jsonl
  else MIPB200_LIB=$PWD/vvc-mip-gpu_b200/lib/var_$n.so MODE=throughput timeout 300 python tools/chunk_sweep.py 1920x1080 96 | tee -a gpurun_out/ab_timing.jsonl; fi
done
done
