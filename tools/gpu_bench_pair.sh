#!/bin/bash
# Both bench arms back to back, as the driver runs them, plus the bench contract test.
set -u
mkdir -p gpurun_out
python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_1gpu.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/r02_bench_ref.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "costs", d["e2e_costs"]["value"], d["e2e_costs"]["d2h_link"], "compact", d["e2e_costs_compact"]["value"])
print("ref", r["value"], r["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["lone_frame"]["frac"])
PY
timeout 900 python -m pytest tests/test_bench_contract.py -x -q -m gpu 2>&1 | tail -2
