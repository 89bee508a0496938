#!/usr/bin/env python3
"""Sustained run of the product's host path with a clock / power / throttle record.

Runs the CLI (streaming host, page-locked ring, 3 frames in flight) over a pool of 16 distinct 1080p frames cycled for
`seconds` (default 12 s, about 28 000 frames) in the bench configuration, with --Digest: one 64-bit hash per frame of the
decisions that reached the host.  Meanwhile this process samples NVML at 20 Hz: SM clock, board power, throttle reasons.
Checks: every frame's digest equals the digest of the same pool frame's first occurrence (bit-identity over the whole
run), no hw / thermal slowdown reason, frames/s.  Writes <out>.json (summary) and <out>.csv (the 20 Hz trace).

Usage: soak.py [seconds] [out_prefix] [--costs]      (--costs: full int32 tables instead of decisions only)"""
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vvc-mip-gpu_b200"))
import numpy as np

import mipb200
from mipb200 import frames

W, H, P = 1920, 1080, 16
args = [a for a in sys.argv[1:] if not a.startswith("--")]
seconds = float(args[0]) if args else 12.0
out = args[1] if len(args) > 1 else os.path.join(ROOT, "gpurun_out", "soak")
costs = "--costs" in sys.argv
rate = 900.0 if costs else 2250.0
n_frames = int(seconds * rate)

trace, stop = [], threading.Event()


def sampler():
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(0)
    names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
             nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
    t0 = time.perf_counter()
    while not stop.is_set():
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        trace.append((time.perf_counter() - t0, nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                      nv.nvmlDeviceGetTemperature(h, nv.NVML_TEMPERATURE_GPU), nv.nvmlDeviceGetUtilizationRates(h).gpu,
                      "+".join(nm for bit, nm in names.items() if r & bit) or "none"))
        time.sleep(0.05)


with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    pool = os.path.join(d, "pool.u16")
    with open(pool, "wb") as f:
        for i in range(P):
            f.write((frames.natural_frame(W, H, 500 + i) if i % 2 else frames.noise_frame(W, H, 500 + i)).astype("<u2").tobytes())
    dig = os.path.join(d, "digest.csv")
    cmd = [mipb200.CLI_PATH, "-f", str(n_frames), "-s", f"{W}x{H}", "-o", pool, "--InputFormat=u16", f"--InputFrames={P}", "--NoLog",
           f"--Digest={dig}", "--StageStamps=0", "--Energy", "--UseAlternativeSamples=1", "--FilterType=filterFrame_2d_float_5x5_quarterCtu", "--KernelIdx=2"]
    if costs:
        cmd.append("--BinaryLog=/dev/null")
    th = threading.Thread(target=sampler, daemon=True)
    th.start()
    time.sleep(0.5)
    r = subprocess.run(cmd, capture_output=True, text=True)
    time.sleep(0.3)
    stop.set()
    th.join(timeout=2)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:])
        sys.exit(2)
    lines = open(dig).read().splitlines()
    hdr, rows = lines[0], [ln.split(",", 1) for ln in lines[1:]]
    assert [int(a) for a, _ in rows] == list(range(n_frames)), "digest file is not in POC order"
    bad = sum(1 for i in range(P, n_frames) if rows[i][1] != rows[i % P][1])
    distinct = len({b for _, b in rows[:P]})

info = {ln.split(",")[0].strip(): ln.split(",")[1].strip() for ln in r.stdout.splitlines() if "," in ln and ln[0].isalpha()}
fps = float([ln for ln in r.stdout.splitlines() if ln.startswith("Throughput:")][0].split()[1])
busy = [t for t in trace if t[4] >= 50]
clk = sorted(t[1] for t in busy) or [0]
summary = {
    "what": f"mipb200_main, {n_frames} 1080p frames ({P} distinct, cycled), bench configuration, " + ("full int32 tables + decisions" if costs else "decisions") + " to the host, digest of every frame",
    "frames": n_frames, "frames_per_s": fps, "elapsed_ms": float(info.get("Elapsed time (ms) from writing samples to reading distortion (%dx)" % n_frames, "nan")),
    "digest_columns": hdr, "distinct_pool_digests": distinct, "frames_differing_from_first_occurrence": bad,
    "joules_per_frame": float(info.get("Energy per frame (J)", "nan")), "average_power_w": float(info.get("Average power (W)", "nan")),
    "peak_host_memory_mb": float(info.get("Peak host memory (MB)", "nan")),
    "nvml_20hz": {"samples": len(trace), "samples_under_load": len(busy), "sm_mhz_median_under_load": clk[len(clk) // 2], "sm_mhz_min_under_load": clk[0],
                  "power_w_max": max((t[2] for t in trace), default=0.0), "temperature_c_max": max((t[3] for t in trace), default=0),
                  "throttle_reasons_seen": sorted({t[5] for t in busy})},
}
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(summary, open(out + ".json", "w"), indent=1)
with open(out + ".csv", "w") as f:
    f.write("t_s,sm_mhz,power_w,temp_c,gpu_util,throttle\n")
    for t in trace:
        f.write("%.3f,%d,%.1f,%d,%d,%s\n" % t)
print(json.dumps(summary))
sys.exit(1 if bad or distinct != P else 0)
