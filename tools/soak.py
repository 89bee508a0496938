#!/usr/bin/env python3
"""Soak / determinism run of the host path: a pool of distinct frames cycled through a pipelined engine many times;
every result must equal, bit for bit, the first result for the same frame.  A race in the kernel (task counter,
shared-memory argmin, TMA staging) or in the slot ring would show up as a mismatch sooner or later.
Usage: soak.py [frames_decisions] [frames_costs]"""
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vvc-mip-gpu_b200"))
import numpy as np

import mipb200
from mipb200 import frames


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def soak(n, emit, fields, label):
    W, H, P = 1920, 1080, 8
    pool = [frames.natural_frame(W, H, 500 + i) if i % 2 else frames.noise_frame(W, H, 500 + i) for i in range(P)]
    want = {}
    bad = 0
    t0 = time.time()
    with mipb200.Engine(W, H, filter_type=7, kernel_idx=1, slots=3, emit=emit) as eng:
        sub = got = 0
        while got < n:
            while sub < n and eng.in_flight() < 3:
                eng.submit(pool[sub % P], sub)
                sub += 1
            r = eng.collect()
            d = digest(*[getattr(r, f) for f in fields])
            k = r.poc % P
            if k not in want:
                want[k] = d
            elif want[k] != d:
                bad += 1
            assert r.poc == got
            got += 1
    dt = time.time() - t0
    print(f"{label}: {n} frames, {len(want)} distinct, {bad} mismatches, {n / dt:.0f} frames/s incl. hashing")
    return bad


if __name__ == "__main__":
    n_dec = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
    n_cost = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    bad = soak(n_dec, mipb200.EMIT_DECISIONS, ("best_mode", "best_cost"), "decisions")
    bad += soak(n_cost, mipb200.EMIT_COSTS | mipb200.EMIT_SAD_SATD | mipb200.EMIT_DECISIONS, ("cost", "sad", "satd", "best_mode", "best_cost"), "full tables")
    sys.exit(1 if bad else 0)
