#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 tools/bin/microbench_tcgen05 gpurun_out/microbench_tcgen05.json; echo "tcgen05 microbench rc=$?"
python tools/soak.py 12 gpurun_out/r02_soak_decisions; echo "soak rc=$?"
python tools/soak.py 10 gpurun_out/r02_soak_full_tables --costs; echo "soak costs rc=$?"
bash tools/gpu_ncu_r02.sh
python bench.py --steps 10 --warmup 3 --no-sizes --no-cpu-baseline > gpurun_out/r02_bench_v5.json 2> gpurun_out/r02_bench_v5.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_v5.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["e2e_costs"]["value"], d["roofline"]["frac"], d["roofline"]["frac_timed_region"])
PY
