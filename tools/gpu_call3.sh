#!/bin/bash
# round 2, call 3: packed-clamp kernel (parity + timing), full-table transport timeline, bench re-run
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -4
MIPB200_CHUNK_WEIGHTS=1,1,1 python tools/chunk_sweep.py 1920x1080 96 | tee gpurun_out/t7_timing.json
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, "vvc-mip-gpu_b200")
from mipb200 import frames
with open("/dev/shm/in1080.u16", "wb") as f:
    for i in range(16):
        f.write(frames.natural_frame(1920, 1080, 100 + i).astype("<u2").tobytes())
PY
M=vvc-mip-gpu_b200/bin/mipb200_main
F="--UseAlternativeSamples=1 --FilterType=filterFrame_2d_float_5x5_quarterCtu --KernelIdx=2"
for sl in 3 4 6; do
  echo "== full tables, --Slots=$sl"
  rm -f /dev/shm/trace.gpu0
  MIPB200_TRACE=/dev/shm/trace $M -f 600 -s 1920x1080 -o /dev/shm/in1080.u16 --InputFormat=u16 --InputFrames=16 --NoLog --BinaryLog=/dev/null --StageStamps=0 --Slots=$sl $F 2>&1 | grep -E "Throughput|ERROR"
  cp /dev/shm/trace.gpu0 gpurun_out/trace_costs_slots$sl.txt
done
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_v3.json 2> gpurun_out/r02_bench_v3.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_v3.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["e2e_costs"]["value"], [(s["value"], s["e2e"]["value"]) for s in d["sizes"]], d["roofline"]["frac"], d["roofline"]["frac_timed_region"], d["shard_check"])
PY
