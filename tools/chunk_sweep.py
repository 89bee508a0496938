#!/usr/bin/env python3
"""One configuration of the chunk split (MIPB200_CHUNK_WEIGHTS / MIPB200_CHUNK_WEIGHTS_LONE and MODE=auto|throughput|latency
in the environment): ms per 1080p
frame of the fused kernel (bench configuration: filter 8 / k 2, costs + decisions) with launches back to back on one
stream (what a lone frame costs) and with frames round-robin over three streams (steady state).  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vvc-mip-gpu_b200"))
import numpy as np
import torch

import mipb200
from mipb200 import frames

W, H = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1920x1080").split("x"))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 96
emit = mipb200.EMIT_COSTS | mipb200.EMIT_DECISIONS
FT = int(os.environ.get("FILTER", "8"))      # 8 = the bench configuration (2-D 5x5), 0 = original samples
pool = torch.from_numpy(np.stack([frames.natural_frame(W, H, i) for i in range(4)]).view(np.int16)).cuda()
mode = {"auto": mipb200.LAUNCH_AUTO, "throughput": mipb200.LAUNCH_THROUGHPUT, "latency": mipb200.LAUNCH_LATENCY}[os.environ.get("MODE", "auto")]
res = {"weights": os.environ.get("MIPB200_CHUNK_WEIGHTS"), "weights_lone": os.environ.get("MIPB200_CHUNK_WEIGHTS_LONE"), "mode": os.environ.get("MODE", "auto"), "filter": FT, "size": f"{W}x{H}"}
for ns in tuple(int(v) for v in os.environ.get('STREAMS', '1,3').split(',')):
    engs = [mipb200.Engine(W, H, filter_type=FT, kernel_idx=2 if FT >= 5 else 0, slots=1, emit=emit) for _ in range(ns)]
    n = engs[0].n_ctus
    for e_ in engs:
        e_.set_launch_mode(mode)
    outs = [(torch.empty((n, mipb200.COSTS_PER_CTU), dtype=torch.int32, device="cuda"), torch.empty((n, mipb200.CUS_PER_CTU), dtype=torch.uint8, device="cuda"),
             torch.empty((n, mipb200.CUS_PER_CTU), dtype=torch.int32, device="cuda")) for _ in range(3)]
    streams = [torch.cuda.Stream() for _ in range(ns)]

    def go(cnt):
        for i in range(cnt):
            k = i % ns
            c_, m_, b_ = outs[i % 3]
            engs[k].run_device(pool[i % 4].data_ptr(), c_.data_ptr(), d_best_mode=m_.data_ptr(), d_best_cost=b_.data_ptr(), stream=streams[k].cuda_stream)

    best = 1e9
    for rep in range(3):
        go(6)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for s_ in streams[1:]:
            s_.wait_stream(streams[0])
        go(N)
        for s_ in streams[1:]:
            streams[0].wait_stream(s_)
        e1.record(streams[0])
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / N)
    res["ms_1stream" if ns == 1 else f"ms_{ns}streams"] = round(best, 4)
    res["checksum"] = int(outs[0][0].to(torch.int64).clamp(min=0).sum()) ^ int(outs[0][2].to(torch.int64).sum())
    for e in engs:
        e.close()
print(json.dumps(res), flush=True)
