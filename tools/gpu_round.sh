#!/bin/bash
# What the driver runs at round end, in one call: the GPU test suite, smoke(), both bench arms.
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
r = json.loads(open("gpurun_out/r02_bench_ref.json").read().strip().splitlines()[-1])
d = json.loads(open("gpurun_out/r02_bench_1gpu.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["e2e"]["value"], r["e2e_overlapped"], r["steps"], r["warmup"], r["config"]["workload"] == d["config"]["workload"])
print({k: d[k] for k in ("value", "ms_per_step", "steps", "warmup", "gpu_launches")}, "e2e", d["e2e"]["value"], "costs", d["e2e_costs"]["value"])
print([(s["value"], s["e2e"]["value"], round(s["frac_timed_region"], 3)) for s in d["sizes"]], d["shard_check"]["status"])
print("roofline", d["roofline"]["frac"], d["roofline"]["lone_frame"], d["roofline"]["traffic"], d["clocks"], d["cpu_baseline"]["value"])
print("ratios: e2e", d["e2e"]["value"] / r["e2e"]["value"], "value", d["value"] / r["value"], "like-for-like", d["e2e_costs"]["value"] / r["e2e"]["value"])
PY
