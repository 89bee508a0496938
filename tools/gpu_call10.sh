#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py tests/test_cli.py -m gpu -x -q 2>&1 | tail -4
rm -f gpurun_out/modes_timing.jsonl
for m in throughput latency; do MODE=$m timeout 300 python tools/chunk_sweep.py 1920x1080 96 | tee -a gpurun_out/modes_timing.jsonl; done
MODE=latency MIPB200_CHUNK_WEIGHTS_LONE=5,4,3,2,1 timeout 300 python tools/chunk_sweep.py 1920x1080 96 | tee -a gpurun_out/modes_timing.jsonl
MODE=latency MIPB200_CHUNK_WEIGHTS_LONE=3,2,1 timeout 300 python tools/chunk_sweep.py 1920x1080 96 | tee -a gpurun_out/modes_timing.jsonl
MODE=latency MIPB200_CHUNK_WEIGHTS_LONE=8,6,4,3,2,1 timeout 300 python tools/chunk_sweep.py 1920x1080 96 | tee -a gpurun_out/modes_timing.jsonl
for nt in 768 384 256; do timeout 120 tools/bin/microbench_tcgen05_$nt gpurun_out/microbench_tcgen05_$nt.json > /dev/null; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_v6.json 2> gpurun_out/r02_bench_v6.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_v6.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["e2e_costs"]["value"], [(s["value"], s["e2e"]["value"]) for s in d["sizes"]], d["roofline"]["frac"], d["roofline"]["frac_timed_region"], d["roofline"]["kernel_ms_per_frame"])
PY
