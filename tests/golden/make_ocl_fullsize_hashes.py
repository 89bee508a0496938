#!/usr/bin/env python3
"""Generate tests/golden/ocl_b200_fullsize_hashes.json: per-CTU SHA-256 prefixes of the minSadHad tables that the
reference's OWN, UNMODIFIED OpenCL kernels produce for whole 1080p / 2160p frames (BASELINE configurations) on a B200.

Same runner as make_ocl_fixtures.py (oracle/_ref/mipref_ocl); full tables are 53 / 200 MB, so only hashes are kept:
entries of CUs that are not fully inside the frame (garbage in the reference's buffer) are set to -1 first, then every
CTU's 97 840 little-endian int32 values are hashed (first 16 hex digits of SHA-256).

    gpurun -- 'python tests/golden/make_ocl_fullsize_hashes.py gpurun_out/ocl_fullsize_hashes.json'
    cp gpurun_out/ocl_fullsize_hashes.json tests/golden/ocl_b200_fullsize_hashes.json      # then commit
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vvc-mip-gpu_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mipb200 import frames, tables  # noqa: E402
from make_ocl_fixtures import run  # noqa: E402

CASES = [("natural", 1920, 1080, 0, 0, 0), ("natural", 1920, 1080, 1, 8, 2), ("noise", 1920, 1080, 2, 3, 1),
         ("natural", 3840, 2160, 3, 0, 0)]


def cost_mask(width, height):
    """bool [nCTU][97840]: the (CU, mode) entries of CUs fully inside the frame."""
    cu = tables.in_frame_mask(width, height)
    out = np.zeros((cu.shape[0], tables.COSTS_PER_CTU), dtype=bool)
    for t in tables.TYPES:
        m = cu[:, tables.CU_OFFSETS[t.idx]:tables.CU_OFFSETS[t.idx + 1]]
        out[:, tables.COST_OFFSETS[t.idx]:tables.COST_OFFSETS[t.idx + 1]] = np.repeat(m, t.modes, axis=1)
    return out


def ctu_hashes(cost, width, height):
    c = np.where(cost_mask(width, height), cost, -1).astype(np.int32)
    return [hashlib.sha256(np.ascontiguousarray(c[i], dtype="<i4").tobytes()).hexdigest()[:16] for i in range(c.shape[0])]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ocl_fullsize_hashes.json"
    res = []
    for kind, w, h, seed, ft, kidx in CASES:
        f = frames.noise_frame(w, h, seed) if kind == "noise" else frames.natural_frame(w, h, seed)
        cost, _, info = run(f, ft, kidx)
        res.append({"frame_kind": kind, "width": w, "height": h, "seed": seed, "filter_type": ft, "kernel_idx": kidx,
                    "device": info.get("device"), "opencl_lib": info.get("opencl_lib"), "sha256_16_per_ctu": ctu_hashes(cost, w, h)})
        print(kind, w, h, seed, ft, kidx, "ok", flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    json.dump(res, open(out, "w"), indent=0)


if __name__ == "__main__":
    main()
