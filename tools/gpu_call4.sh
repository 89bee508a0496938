#!/bin/bash
# round 2, call 4: in-order read-back (tests + full-table throughput), tcgen05 SATD micro-benchmark
set -u
mkdir -p gpurun_out
timeout 120 tools/bin/microbench_tcgen05 gpurun_out/microbench_tcgen05.json; echo "tcgen05 microbench rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py tests/test_cli.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -4
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
for sl in 2 3 4; do
  echo "== full tables, --Slots=$sl"
  rm -f /dev/shm/trace.gpu0
  MIPB200_TRACE=/dev/shm/trace $M -f 800 -s 1920x1080 -o /dev/shm/in1080.u16 --InputFormat=u16 --InputFrames=16 --NoLog --BinaryLog=/dev/null --StageStamps=0 --Slots=$sl $F 2>&1 | grep -E "Throughput|ERROR"
  cp /dev/shm/trace.gpu0 gpurun_out/trace_inorder_slots$sl.txt
done
echo "== decisions, --Slots=3"
$M -f 4096 -s 1920x1080 -o /dev/shm/in1080.u16 --InputFormat=u16 --InputFrames=16 --NoLog --Digest=/dev/null --StageStamps=0 $F 2>&1 | grep -E "Throughput|ERROR"
python bench.py --steps 10 --warmup 3 --no-sizes > gpurun_out/r02_bench_v4.json 2> gpurun_out/r02_bench_v4.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_v4.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["e2e_costs"]["value"], d["roofline"]["frac"], d["roofline"]["frac_timed_region"])
PY
