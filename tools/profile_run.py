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
ns = int(os.environ.get("STREAMS", "1"))   # frames round-robin over this many streams (independent frames overlap)
streams = [torch.cuda.Stream() for _ in range(ns)]
engs = [eng] + [mipb200.Engine(W, H, filter_type=8, kernel_idx=2, slots=1, emit=mipb200.EMIT_COSTS | mipb200.EMIT_DECISIONS) for _ in range(ns - 1)]
outs = [(cost, bm, bc)] + [(torch.empty_like(cost), torch.empty_like(bm), torch.empty_like(bc)) for _ in range(ns - 1)]
torch.cuda.set_stream(streams[0])
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(n + ns):
    if i == ns:
        torch.cuda.synchronize()
        ev0.record(streams[0])
        for s_ in streams[1:]:
            s_.wait_stream(streams[0])
    k = i % ns
    c_, m_, b_ = outs[k]
    engs[k].run_device(pool[i % 2].data_ptr(), c_.data_ptr(), d_best_mode=m_.data_ptr(), d_best_cost=b_.data_ptr(), stream=streams[k].cuda_stream)
for s_ in streams[1:]:
    streams[0].wait_stream(s_)
ev1.record(streams[0])
torch.cuda.synchronize()
print(f"{n} frames {W}x{H} on {ns} stream(s): {ev0.elapsed_time(ev1) / n:.3f} ms/frame, checksum {int(cost.to(torch.int64).clamp(min=0).sum())}")
