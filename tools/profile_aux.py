#!/usr/bin/env python3
"""Times the stand-alone kernels (top-k shortlist, argmin, filter) on a 1080p frame with CUDA events and reports their HBM
fraction (these are the byte-bound kernels of the library; the cost kernel is INT32-bound).  Usage: profile_aux.py [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vvc-mip-gpu_b200"))
import numpy as np
import torch

import mipb200
from mipb200 import frames

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
W, H = 1920, 1080
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6536.7) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6536.7
eng = mipb200.Engine(W, H, filter_type=8, kernel_idx=2, slots=1, emit=mipb200.EMIT_COSTS | mipb200.EMIT_DECISIONS)
f = torch.from_numpy(frames.natural_frame(W, H, 1).view(np.int16)).cuda()
n = eng.n_ctus
NB = 4   # rotate buffers so that the 52.8 MB tables do not sit in the 126 MB L2 between repetitions
cost = torch.empty((NB, n, mipb200.COSTS_PER_CTU), dtype=torch.int32, device="cuda")
bm = torch.empty((n, mipb200.CUS_PER_CTU), dtype=torch.uint8, device="cuda")
bc = torch.empty((n, mipb200.CUS_PER_CTU), dtype=torch.int32, device="cuda")
st = torch.cuda.Stream()
for b in range(NB):
    eng.run_device(f.data_ptr(), cost[b].data_ptr(), stream=st.cuda_stream)
st.synchronize()
out = {}


def timed(name, fn, bytes_per_call):
    for i in range(3):
        fn(i)
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(reps):
        fn(i)
    e1.record(st)
    st.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = bytes_per_call / (ms * 1e-3) / 1e9
    out[name] = {"ms": round(ms, 4), "algorithmic_bytes": bytes_per_call, "GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 3)}
    print(f"{name:24s} {ms:8.4f} ms  {gbs:8.1f} GB/s  {gbs / peak:5.1%} of {peak} GB/s")


ncost = 4 * n * mipb200.COSTS_PER_CTU
ncu5 = 5 * n * mipb200.CUS_PER_CTU
timed("decide (argmin)", lambda i: eng.decide_device(cost[i % NB].data_ptr(), bm.data_ptr(), bc.data_ptr(), stream=st.cuda_stream), ncost + ncu5)
for k in (1, 3, 12):
    tm = torch.empty((n, mipb200.CUS_PER_CTU, k), dtype=torch.uint8, device="cuda")
    tc = torch.empty((n, mipb200.CUS_PER_CTU, k), dtype=torch.int32, device="cuda")
    timed(f"topk k={k}", lambda i: eng.topk_device(cost[i % NB].data_ptr(), k, tm.data_ptr(), tc.data_ptr(), stream=st.cuda_stream), ncost + ncu5 * k)
g = torch.empty_like(f)
timed("filter 5x5 2d", lambda i: eng.filter_device(f.data_ptr(), g.data_ptr(), stream=st.cuda_stream), 4 * W * H)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "aux_kernels.json"), "w"), indent=1)
