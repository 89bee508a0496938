#!/usr/bin/env python3
"""Generate tests/golden/ocl_b200_*.npz: outputs of the reference's OWN, UNMODIFIED OpenCL kernels.

Runs on the GPU box (NVIDIA OpenCL driver on the B200) through oracle/_ref/mipref_ocl, which
embeds the reference's .cl sources at build time (oracle/ocl_ref/Makefile; built in the build
container where /root/reference is mounted, shipped to the box as a git-ignored artefact).

    gpurun -- 'python tests/golden/make_ocl_fixtures.py gpurun_out/ocl_fixtures'
    cp gpurun_out/ocl_fixtures/*.npz tests/golden/          # then commit

Each .npz holds the frame recipe (kind/seed/size), the filter configuration, the int32-narrowed
minSadHad buffer of frame 0 and, for filtered runs, the filtered frame.  Also prints the
reference's throughput on 1080p (kernel time and end-to-end) for BASELINE purposes.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vvc-mip-gpu_b200"))
from mipb200 import frames, tables  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "mipref_ocl")


def make_frame(kind, w, h, seed):
    return {"kat": lambda: frames.kat_frame(w, h), "noise": lambda: frames.noise_frame(w, h, seed),
            "natural": lambda: frames.natural_frame(w, h, seed)}[kind]()


def run(frame, ft, kidx, reps=1):
    h, w = frame.shape
    with tempfile.TemporaryDirectory() as d:
        fp = os.path.join(d, "f.u16")
        frame.astype("<u2").tofile(fp)
        name = tables.FILTER_NAMES[ft - 1] if ft else "none"
        r = subprocess.run([BIN, fp, str(w), str(h), name, str(kidx), os.path.join(d, "out"), str(reps)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"mipref_ocl rc={r.returncode}\n{r.stdout}\n{r.stderr}")
        info = json.loads(r.stdout.strip().splitlines()[-1])
        cost = np.fromfile(os.path.join(d, "out.cost.i32"), dtype="<i4").reshape(-1, 97840)
        filt = np.fromfile(os.path.join(d, "out.filt.u16"), dtype="<u2").reshape(h, w) if ft else None
    return cost, filt, info


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ocl_fixtures"
    os.makedirs(out, exist_ok=True)
    log = []
    cost_cases = [("kat", 256, 256, 0, 0, 0), ("noise", 128, 184, 11, 0, 0), ("natural", 128, 184, 12, 8, 2),
                  ("noise", 128, 184, 13, 1, 1), ("natural", 128, 184, 14, 5, 2), ("noise", 128, 184, 15, 3, 4)]
    for kind, w, h, seed, ft, kidx in cost_cases:
        f = make_frame(kind, w, h, seed)
        cost, filt, info = run(f, ft, kidx)
        path = os.path.join(out, f"ocl_b200_cost_{kind}_{w}x{h}_s{seed}_f{ft}k{kidx}.npz")
        kw = dict(frame_kind=kind, width=w, height=h, seed=seed, filter_type=ft, kernel_idx=kidx, cost=cost)
        if filt is not None:
            kw["filtered"] = filt
        np.savez_compressed(path, **kw)
        log.append({"case": os.path.basename(path), **info})
        print("wrote", path, os.path.getsize(path), flush=True)
    # filters only: every type x kernelIdx on a small natural frame (all 9/25 border classes, partial bottom tile)
    f = make_frame("natural", 128, 88, 21)
    filt_all = {}
    for ft in range(1, 9):
        for kidx in range(tables.num_kernel_idx(ft)):
            _, filt, _ = run(f, ft, kidx)
            filt_all[f"f{ft}k{kidx}"] = filt
    np.savez_compressed(os.path.join(out, "ocl_b200_filters_natural_128x88_s21.npz"), frame_kind="natural", width=128, height=88, seed=21, **filt_all)
    # reference throughput on 1080p (original samples, and BASELINE config 2)
    f = frames.natural_frame(1920, 1080, 0)
    for ft, kidx in ((0, 0), (8, 2)):
        _, _, info = run(f, ft, kidx, reps=5)
        log.append({"case": f"1080p_f{ft}k{kidx}", **info})
        print(json.dumps(info), flush=True)
    json.dump(log, open(os.path.join(out, "ocl_b200_runs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
