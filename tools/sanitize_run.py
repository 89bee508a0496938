#!/usr/bin/env python3
"""Small run of every device entry point for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
a 256x184 frame (4 CTUs, two of them cut by the bottom frame edge) through the host path with costs + SAD/SATD +
decisions + a top-3 shortlist, for the original samples and one filter of each kind (1-D/2-D, 3x3/5x5), plus the
stand-alone filter / argmin / shortlist kernels.  Prints one checksum line per configuration: the lines of a run under
the sanitizer must equal those of a plain run.  Usage: sanitize_run.py [filter types, default 0,1,3,5,7]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vvc-mip-gpu_b200"))
import numpy as np
import torch

import mipb200
from mipb200 import frames

W, H = 256, 184
types = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,1,3,5,7").split(",")]
fs = [frames.natural_frame(W, H, 11), frames.noise_frame(W, H, 12)]
for ft in types:
    kidx = 0 if ft == 0 else (2 if ft >= 5 else 3)
    with mipb200.Engine(W, H, filter_type=ft, kernel_idx=kidx, slots=2, top_k=3,
                        emit=mipb200.EMIT_COSTS | mipb200.EMIT_SAD_SATD | mipb200.EMIT_DECISIONS) as eng:
        chk = 0
        for poc, f in enumerate(fs):
            eng.submit(f, poc)
        for _ in fs:
            r = eng.collect()
            for a in (r.cost, r.sad, r.satd, r.best_cost, r.topk_cost):
                chk = (chk * 1000003 + int(a.astype(np.int64).sum())) % (1 << 61)
            chk = (chk * 1000003 + int(r.best_mode.astype(np.int64).sum()) + int(r.topk_mode.astype(np.int64).sum())) % (1 << 61)
        # device-resident entry points
        d_f = torch.from_numpy(fs[0].view(np.int16)).cuda()
        d_cost = torch.empty((eng.n_ctus, mipb200.COSTS_PER_CTU), dtype=torch.int32, device="cuda")
        d_bm = torch.empty((eng.n_ctus, mipb200.CUS_PER_CTU), dtype=torch.uint8, device="cuda")
        d_bc = torch.empty((eng.n_ctus, mipb200.CUS_PER_CTU), dtype=torch.int32, device="cuda")
        d_tm = torch.empty((eng.n_ctus, mipb200.CUS_PER_CTU, 2), dtype=torch.uint8, device="cuda")
        d_tc = torch.empty((eng.n_ctus, mipb200.CUS_PER_CTU, 2), dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        eng.run_device(d_f.data_ptr(), d_cost.data_ptr(), stream=st)
        eng.decide_device(d_cost.data_ptr(), d_bm.data_ptr(), d_bc.data_ptr(), stream=st)
        eng.topk_device(d_cost.data_ptr(), 2, d_tm.data_ptr(), d_tc.data_ptr(), stream=st)
        eng.run_device(d_f.data_ptr(), 0, d_best_mode=d_bm.data_ptr(), d_best_cost=d_bc.data_ptr(), stream=st)   # decisions only
        if ft:
            d_o = torch.empty_like(d_f)
            eng.filter_device(d_f.data_ptr(), d_o.data_ptr(), stream=st)
            chk = (chk * 1000003 + int(d_o.to(torch.int64).sum())) % (1 << 61)
        torch.cuda.synchronize()
        chk = (chk * 1000003 + int(d_cost.to(torch.int64).sum()) + int(d_bc.to(torch.int64).sum()) + int(d_tc.to(torch.int64).sum())) % (1 << 61)
    print(f"filter_type {ft} kernel_idx {kidx}: checksum {chk}", flush=True)
