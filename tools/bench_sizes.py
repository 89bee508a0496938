#!/usr/bin/env python3
"""Throughput of the other BASELINE configurations on one GPU (bench.py covers the headline 1080p configuration):
2160p with original samples (config 4) and 4320p with alternative samples (config 5).  Same method as bench.py:
`value` = device-resident frames on 3 streams, CUDA-event timed; `e2e` = the host path (pinned frames in, decisions out).
Usage: bench_sizes.py [out.json]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vvc-mip-gpu_b200"))
import numpy as np
import torch

import mipb200
from mipb200 import frames

# algorithmic INT32 ops per frame: SURVEY.md 8(d)
CASES = [("2160p, original samples", 3840, 2160, 0, 0, 16, 5.2766e10), ("4320p, filterFrame_2d_float_5x5_quarterCtu k=2", 7680, 4320, 8, 2, 6, 2.1155e11)]
NS, STEPS = 3, 4
peak = json.load(open(os.path.join(ROOT, "profiles", "int32_peak.json")))["int32_tops"]
out = []
for name, W, H, ft, kidx, B, ops in CASES:
    pool = [frames.natural_frame(W, H, 700 + i) for i in range(B)]
    d_pool = torch.from_numpy(np.stack(pool).view(np.int16)).cuda()
    engs = [mipb200.Engine(W, H, filter_type=ft, kernel_idx=kidx, slots=1, emit=mipb200.EMIT_DECISIONS) for _ in range(NS)]
    n = engs[0].n_ctus
    bm = torch.empty((NS, n, mipb200.CUS_PER_CTU), dtype=torch.uint8, device="cuda")
    bc = torch.empty((NS, n, mipb200.CUS_PER_CTU), dtype=torch.int32, device="cuda")
    cost = torch.empty((NS, n, mipb200.COSTS_PER_CTU), dtype=torch.int32, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(NS)]

    def step():
        for i in range(B):
            k = i % NS
            engs[k].run_device(d_pool[i].data_ptr(), cost[k].data_ptr(), d_best_mode=bm[k].data_ptr(), d_best_cost=bc[k].data_ptr(), stream=streams[k].cuda_stream)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for s in streams[1:]:
        s.wait_stream(streams[0])
    for _ in range(STEPS):
        step()
    for s in streams[1:]:
        streams[0].wait_stream(s)
    e1.record(streams[0])
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (STEPS * B)
    for e in engs:
        e.close()
    del d_pool, cost
    # host path: pinned frames -> decisions
    pin = torch.empty((B, H, W), dtype=torch.int16, pin_memory=True)
    pin.numpy()[...] = np.stack(pool).view(np.int16)
    host = [pin[i].numpy().view(np.uint16) for i in range(B)]
    with mipb200.Engine(W, H, filter_type=ft, kernel_idx=kidx, slots=3, emit=mipb200.EMIT_DECISIONS) as eng:
        def run(nf):
            sub = got = 0
            chk = 0
            while got < nf:
                while sub < nf and eng.in_flight() < 3:
                    eng.submit(host[sub % B], sub)
                    sub += 1
                r = eng.collect()
                chk += int(r.best_cost[0, 0])
                got += 1
            return chk
        run(B)
        t0 = time.perf_counter()
        run(STEPS * B)
        e2e = STEPS * B / (time.perf_counter() - t0)
    rec = {"workload": name, "width": W, "height": H, "ms_per_frame": round(ms, 4), "value_fps": round(1e3 / ms, 1), "e2e_fps_decisions": round(e2e, 1),
           "int32_frac_timed_region": round(ops / (ms * 1e-3) / 1e12 / peak, 3), "h2d_bytes_per_frame": 2 * W * H, "d2h_bytes_per_frame": 5 * n * mipb200.CUS_PER_CTU}
    print(json.dumps(rec), flush=True)
    out.append(rec)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
