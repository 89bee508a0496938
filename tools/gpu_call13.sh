#!/bin/bash
# compact cost transport: parity, CLI rate, bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py tests/test_cli.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -4
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
echo "== compact tables"; $M -f 1600 -s 1920x1080 -o /dev/shm/in1080.u16 --InputFormat=u16 --InputFrames=16 --NoLog --CompactLog=/dev/null --StageStamps=0 --Energy $F 2>&1 | grep -E "Throughput|ERROR|Energy per"
echo "== int32 tables"; $M -f 1000 -s 1920x1080 -o /dev/shm/in1080.u16 --InputFormat=u16 --InputFrames=16 --NoLog --BinaryLog=/dev/null --StageStamps=0 --Energy $F 2>&1 | grep -E "Throughput|ERROR|Energy per"
MODE=throughput timeout 300 python tools/chunk_sweep.py 1920x1080 96
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sizes > gpurun_out/r02_bench_v7.json 2> gpurun_out/r02_bench_v7.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_v7.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_v7.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["e2e_costs"]["value"], d["e2e_costs_compact"]["value"], d["roofline"]["frac"])
PY
