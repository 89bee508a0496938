#!/bin/bash
# round 2, call 2: full GPU test suite on the reworked engine / CLI, then the streaming CLI at BASELINE sizes on one GPU
set -u
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.txt; fi
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, "vvc-mip-gpu_b200")
from mipb200 import frames
for (w, h, n, name) in ((7680, 4320, 4, "/dev/shm/in4320.u16"), (1920, 1080, 16, "/dev/shm/in1080.u16")):
    with open(name, "wb") as f:
        for i in range(n):
            f.write(frames.natural_frame(w, h, 100 + i).astype("<u2").tobytes())
PY
M=vvc-mip-gpu_b200/bin/mipb200_main
F="--UseAlternativeSamples=1 --FilterType=filterFrame_2d_float_5x5_quarterCtu --KernelIdx=2"
out=gpurun_out/r02_cli_1gpu.txt
: > $out
run() { echo "== $*" | tee -a $out; "$@" 2>&1 | grep -E "Frame ring|Throughput|Energy per frame|Average power|Elapsed|Peak host|ERROR" | tee -a $out; }
run $M -f 256 -s 7680x4320 -o /dev/shm/in4320.u16 --InputFormat=u16 --InputFrames=4 --NoLog --Digest=gpurun_out/dig4320_1gpu.csv --Energy --StageStamps=0 $F
run $M -f 64 -s 7680x4320 -o /dev/shm/in4320.u16 --InputFormat=u16 --InputFrames=4 --RingFrames=3 --NoLog --Digest=gpurun_out/dig4320_1gpu_streamed.csv --StageStamps=0 $F
run $M -f 4096 -s 1920x1080 -o /dev/shm/in1080.u16 --InputFormat=u16 --InputFrames=16 --NoLog --Digest=gpurun_out/dig1080_1gpu.csv --Energy --StageStamps=0 $F
run $M -f 1024 -s 1920x1080 -o /dev/shm/in1080.u16 --InputFormat=u16 --InputFrames=16 --NoLog --BinaryLog=/dev/null --Energy --StageStamps=0 $F
head -3 gpurun_out/dig4320_1gpu.csv; cmp <(head -65 gpurun_out/dig4320_1gpu.csv) gpurun_out/dig4320_1gpu_streamed.csv && echo "streamed == resident digests"
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; tail -c 1800 gpurun_out/r02_bench_ref.json
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_v2.json 2> gpurun_out/r02_bench_v2.err; echo "bench rc=$?"; tail -c 4000 gpurun_out/r02_bench_v2.json; tail -3 gpurun_out/r02_bench_v2.err
