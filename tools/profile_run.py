#!/usr/bin/env python3
"""Short, deterministic run of the hot path for ncu: N 1080p frames, alternative samples
(filter -> fused MIP cost kernel -> decisions), device resident.  Usage: profile_run.py [frames] [WxH]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vvc-mip-gpu_b200"))
import numpy as np
import torch

import mipb200
from mipb200 import frames

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
W, H = (int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1920x1080").split("x"))
eng = mipb200.Engine(W, H, filter_type=8, kernel_idx=2, slots=1, emit=mipb200.EMIT_COSTS | mipb200.EMIT_DECISIONS)
pool = torch.from_numpy(np.stack([frames.natural_frame(W, H, i) for i in range(2)]).view(np.int16)).cuda()
cost = torch.empty((eng.n_ctus, mipb200.COSTS_PER_CTU), dtype=torch.int32, device="cuda")
bm = torch.empty((eng.n_ctus, mipb200.CUS_PER_CTU), dtype=torch.uint8, device="cuda")
bc = torch.empty((eng.n_ctus, mipb200.CUS_PER_CTU), dtype=torch.int32, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
st = stream.cuda_stream
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(n + 1):
    if i == 1:
        ev0.record()
    eng.run_device(pool[i % 2].data_ptr(), cost.data_ptr(), d_best_mode=bm.data_ptr(), d_best_cost=bc.data_ptr(), stream=st)
ev1.record()
torch.cuda.synchronize()
print(f"{n} frames {W}x{H}: {ev0.elapsed_time(ev1) / n:.3f} ms/frame, checksum {int(cost.to(torch.int64).clamp(min=0).sum())}")
