"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line (here it falls back to
the CPU port because no OpenCL runtime exists in the build container) and the GPU arm refuses to run without a device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "1080p frames/s" and line["unit"] == "frames/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    assert "workload" in line["config"]


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_gpu_arm_json_line():
    """One short run of the real arm: every key of the bench contract is present and sane."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["metric"] == "1080p frames/s" and line["unit"] == "frames/s" and line["n_gpus"] == 1 and line["scaling"] == "weak"
    assert line["value"] > 500 and line["gpu_launches"] == 32 and line["vs_baseline"] is None and line["data"] == "synthetic"
    e = line["e2e"]
    assert 0 < e["value"] <= line["value"] * 1.05 and e["h2d_bytes_per_step"] == 32 * 1920 * 1080 * 2 and e["d2h_bytes_per_step"] == 32 * 135 * 5380 * 5
    rf = line["roofline"]
    assert set(rf) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and rf["bound"] == "int32" and 0.3 < rf["frac"] < 1.0
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and set(rf["hbm"]) >= {"achieved", "peak", "unit", "frac"}
    assert 0.3 < rf["lone_frame"]["frac"] <= rf["frac"] * 1.02 and rf["traffic"] > rf["hbm"]["algorithmic_bytes_per_launch"]
    assert line["shard_check"]["status"] == "ok" and len(line["sizes"]) == 2 and all(s["value"] > 50 for s in line["sizes"])
    assert line["e2e_like_for_like"]["value"] == line["e2e_costs"]["value"] > 300
    link = line["e2e_costs"]["d2h_link"]          # the box's read-back ceiling, measured in the same run
    assert link["gbs_per_gpu"] > 10 and 0.5 < link["frac"] <= 1.02 and abs(link["frac"] - line["e2e_costs"]["value"] / link["ceiling"]) < 1e-9
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and "workload" in line["config"]
