"""Seeded synthetic luma frames and the reference's CSV frame format.

The reference's only sample input (data/original_frames_0_1.csv) is missing from its
checkout, so every workload here is synthetic and deterministic.
The arithmetic of the path is data independent; content only steers parity coverage
(clamps, ties, the DC rule, filter border classes).
"""
from __future__ import annotations

import numpy as np


def kat_frame(width: int = 256, height: int = 256) -> np.ndarray:
    """Integer-defined frame of the survey's known-answer vectors."""
    y, x = np.mgrid[0:height, 0:width].astype(np.int64)
    v = (7 * x + 13 * y + 29 * ((x * y) % 31) + 97 * (((x // 8) + (y // 8)) % 5)) % 1024
    return v.astype(np.uint16)


def noise_frame(width: int, height: int, seed: int, bits: int = 10) -> np.ndarray:
    rng = np.random.default_rng(0xB200 + seed)
    return rng.integers(0, 1 << bits, size=(height, width), dtype=np.uint16)


def natural_frame(width: int, height: int, seed: int, bits: int = 10) -> np.ndarray:
    """Low-frequency cosines + 8x8 block texture + small noise (video-like statistics)."""
    rng = np.random.default_rng(0xB200 + seed)
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    peak = (1 << bits) - 1
    img = np.full((height, width), 0.5 * peak)
    for _ in range(3):
        fx, fy = rng.uniform(0.5, 6.0, 2) * 2 * np.pi / max(width, height)
        ph = rng.uniform(0, 2 * np.pi)
        img += 0.12 * peak * np.cos(fx * x + fy * y + ph)
    blocks = rng.uniform(-0.08, 0.08, size=((height + 7) // 8, (width + 7) // 8)) * peak
    img += np.kron(blocks, np.ones((8, 8)))[:height, :width]
    img += rng.integers(-8, 9, size=(height, width))
    lo, hi = (16 << (bits - 8)), (235 << (bits - 8))
    return np.clip(np.rint(img), lo, hi).astype(np.uint16)


def extreme_frame(width: int, height: int, kind: int) -> np.ndarray:
    """0: all 0, 1: all 1023, 2: 1-px checkerboard 0/1023, 3: vertical ramp."""
    if kind == 0:
        return np.zeros((height, width), np.uint16)
    if kind == 1:
        return np.full((height, width), 1023, np.uint16)
    if kind == 2:
        y, x = np.mgrid[0:height, 0:width]
        return (((x + y) & 1) * 1023).astype(np.uint16)
    y = np.arange(height, dtype=np.int64)[:, None]
    return np.broadcast_to((y * 1023) // max(height - 1, 1), (height, width)).astype(np.uint16)


def impulse_frame(width: int, height: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(0xB200 + seed)
    f = np.full((height, width), 512, np.uint16)
    n = max(16, width * height // 512)
    f[rng.integers(0, height, n), rng.integers(0, width, n)] = rng.integers(0, 1024, n).astype(np.uint16)
    return f


def sweep_frames(width: int, height: int):
    """The 16 frames of BASELINE config 3: 4 content classes x 4 seeds."""
    out = []
    for s in range(4):
        out.append(noise_frame(width, height, s))
    for s in range(4):
        out.append(natural_frame(width, height, 100 + s))
    for k in range(4):
        out.append(extreme_frame(width, height, k))
    for s in range(4):
        out.append(impulse_frame(width, height, 200 + s))
    return out


def write_csv(path: str, frames) -> None:
    """Reference input format (main.cpp:364-384): one text line per frame row, comma
    separated decimal samples, frames stacked vertically without separator."""
    with open(path, "w") as f:
        for fr in frames:
            for row in np.asarray(fr):
                f.write(",".join(map(str, row.tolist())))
                f.write("\n")


def read_csv(path: str, width: int, height: int, n_frames: int) -> np.ndarray:
    out = np.empty((n_frames, height, width), np.uint16)
    with open(path) as f:
        for i in range(n_frames):
            for r in range(height):
                vals = f.readline().split(",")[:width]
                out[i, r] = np.array(vals, dtype=np.int64).astype(np.uint16)
    return out


def read_cost_dump(path: str):
    """Reads a `mipb200_main --BinaryLog` file -> (header dict, int32 costs [frames][nCTU][97840]).
    Layout: 64-byte header = "MIPB200C", then little-endian u32 version, width, height, frames, CTUs, costs per CTU,
    bit depth, filter type, kernel index (rest zero); then the tables frame by frame in POC order."""
    with open(path, "rb") as f:
        raw = f.read(64)
        if len(raw) != 64 or raw[:8] != b"MIPB200C":
            raise ValueError(f"{path} is not a mipb200 cost dump")
        v = np.frombuffer(raw[8:48], dtype="<u4")
        hdr = dict(version=int(v[0]), width=int(v[1]), height=int(v[2]), frames=int(v[3]), n_ctus=int(v[4]), costs_per_ctu=int(v[5]),
                   bit_depth=int(v[6]), filter_type=int(v[7]), kernel_idx=int(v[8]))
        data = np.fromfile(f, dtype="<i4")
    return hdr, data.reshape(hdr["frames"], hdr["n_ctus"], hdr["costs_per_ctu"])


def read_decisions_dump(path: str):
    """Reads a `mipb200_main --DecisionsBin` file -> (header dict, uint8 modes [frames][nCTU][5380][k], int32 costs, same shape).
    Layout: 64-byte header = "MIPB200D", then little-endian u32 version, width, height, frames, CTUs, CUs per CTU, bit depth,
    filter type, kernel index, k (entries per CU: 1, or --TopK); then one record per frame in POC order: the modes of every
    CU (uint8, 0xFF = CU not inside the frame) followed by their costs (int32, -1), both in [CTU][CU][k] order."""
    with open(path, "rb") as f:
        raw = f.read(64)
        if len(raw) != 64 or raw[:8] != b"MIPB200D":
            raise ValueError(f"{path} is not a mipb200 decisions dump")
        v = np.frombuffer(raw[8:48], dtype="<u4")
        hdr = dict(version=int(v[0]), width=int(v[1]), height=int(v[2]), frames=int(v[3]), n_ctus=int(v[4]), cus_per_ctu=int(v[5]),
                   bit_depth=int(v[6]), filter_type=int(v[7]), kernel_idx=int(v[8]), k=int(v[9]))
        n = hdr["n_ctus"] * hdr["cus_per_ctu"] * hdr["k"]
        body = np.fromfile(f, dtype=np.uint8)
    rec = body.reshape(hdr["frames"], 5 * n)
    shape = (hdr["frames"], hdr["n_ctus"], hdr["cus_per_ctu"], hdr["k"])
    modes = rec[:, :n].reshape(shape)
    costs = np.ascontiguousarray(rec[:, n:]).view("<i4").reshape(shape)
    return hdr, modes, costs


def read_compact_dump(path: str):
    """Reads a `mipb200_main --CompactLog` file -> (header dict, uint8 records [frames][nCTU][276672]).
    Layout: 64-byte header = "MIPB200K", then little-endian u32 version, width, height, frames, CTUs, bytes per CTU, bit
    depth, filter type, kernel index; then one compact record per CTU and frame in POC order (csrc/mip_compact.h: uint16
    entries for CU types of at most 32 samples, int32 for the others; mipb200.expand_costs() gives the int32 table)."""
    with open(path, "rb") as f:
        raw = f.read(64)
        if len(raw) != 64 or raw[:8] != b"MIPB200K":
            raise ValueError(f"{path} is not a mipb200 compact cost dump")
        v = np.frombuffer(raw[8:48], dtype="<u4")
        hdr = dict(version=int(v[0]), width=int(v[1]), height=int(v[2]), frames=int(v[3]), n_ctus=int(v[4]), bytes_per_ctu=int(v[5]),
                   bit_depth=int(v[6]), filter_type=int(v[7]), kernel_idx=int(v[8]))
        data = np.fromfile(f, dtype=np.uint8)
    return hdr, data.reshape(hdr["frames"], hdr["n_ctus"], hdr["bytes_per_ctu"])
