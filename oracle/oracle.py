"""ctypes binding of the CPU oracle (oracle/libmip_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under vvc-mip-gpu_b200/ may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmip_oracle.so")
COSTS_PER_CTU = 97840
CUS_PER_CTU = 5380
SKIPPED = -1

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mip_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u16p = ctypes.POINTER(ctypes.c_uint16)
        i32p = ctypes.POINTER(ctypes.c_int32)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        ip = ctypes.POINTER(ctypes.c_int)
        L.mipo_filter_frame.argtypes = [u16p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, u16p]
        L.mipo_filter_frame.restype = ctypes.c_int
        L.mipo_frame_costs.argtypes = [u16p, u16p, ctypes.c_int, ctypes.c_int, i32p, i32p, i32p, ctypes.c_int]
        L.mipo_frame_costs.restype = ctypes.c_int
        L.mipo_run_frame.argtypes = [u16p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p, i32p, i32p, ctypes.c_int]
        L.mipo_run_frame.restype = ctypes.c_int
        L.mipo_set_bit_depth.argtypes = [ctypes.c_int]
        L.mipo_set_bit_depth.restype = ctypes.c_int
        L.mipo_decisions.argtypes = [i32p, ctypes.c_int, u8p, i32p]
        L.mipo_decisions.restype = None
        L.mipo_satd4x4.argtypes = [ip]
        L.mipo_satd4x4.restype = ctypes.c_int
        L.mipo_reduced_prediction.argtypes = [ctypes.c_int, ctypes.c_int, ip, ip, ip]
        L.mipo_upsample.argtypes = [ip, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, ip, ip]
        L.mipo_cu_boundaries.argtypes = [u16p] + [ctypes.c_int] * 6 + [ip, ip, ip, ip]
        _lib = L
    return _lib


def _u16(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16))


def _i32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if a is not None else None


def _ci(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def num_ctus(w: int, h: int) -> int:
    return ((w + 127) // 128) * ((h + 127) // 128)


def filter_frame(frame: np.ndarray, filter_type: int, kernel_idx: int) -> np.ndarray:
    frame = np.ascontiguousarray(frame, dtype=np.uint16)
    h, w = frame.shape
    out = np.empty_like(frame)
    rc = lib().mipo_filter_frame(_u16(frame), w, h, filter_type, kernel_idx, _u16(out))
    if rc != 0:
        raise ValueError(f"mipo_filter_frame rc={rc}")
    return out


def run_frame(frame: np.ndarray, filter_type: int = 0, kernel_idx: int = 0, want_sad_satd: bool = False, threads: int = 0,
              bit_depth: int = 10):
    """-> cost[nCTU, 97840] int32 (and sad, satd when asked).  bit_depth 10 is the reference; 8 / 12 are the extension."""
    frame = np.ascontiguousarray(frame, dtype=np.uint16)
    h, w = frame.shape
    n = num_ctus(w, h)
    cost = np.empty((n, COSTS_PER_CTU), dtype=np.int32)
    sad = np.empty_like(cost) if want_sad_satd else None
    satd = np.empty_like(cost) if want_sad_satd else None
    if lib().mipo_set_bit_depth(bit_depth) != 0:
        raise ValueError(f"bit_depth {bit_depth} not in (8, 10, 12)")
    try:
        rc = lib().mipo_run_frame(_u16(frame), w, h, filter_type, kernel_idx, _i32(cost), _i32(sad), _i32(satd), threads)
    finally:
        lib().mipo_set_bit_depth(10)
    if rc != 0:
        raise ValueError(f"mipo_run_frame rc={rc}")
    return (cost, sad, satd) if want_sad_satd else cost


def decisions(cost: np.ndarray):
    cost = np.ascontiguousarray(cost, dtype=np.int32)
    n = cost.shape[0]
    bm = np.empty((n, CUS_PER_CTU), dtype=np.uint8)
    bc = np.empty((n, CUS_PER_CTU), dtype=np.int32)
    lib().mipo_decisions(_i32(cost), n, bm.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _i32(bc))
    return bm, bc


def topk(cost: np.ndarray, k: int):
    """The k cheapest modes of every CU in ascending (cost, mode) order (stable sort over the mode axis):
    modes uint8 [n][5380][k] (255 for skipped CUs), costs int32 [n][5380][k].  An extension: the reference stops at the
    cost log (main_aux_functions.h:735-798); entry 0 is decisions()."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "vvc-mip-gpu_b200"))
    from mipb200 import tables as T
    cost = np.ascontiguousarray(cost, dtype=np.int32)
    n = cost.shape[0]
    tm = np.empty((n, CUS_PER_CTU, k), dtype=np.uint8)
    tc = np.empty((n, CUS_PER_CTU, k), dtype=np.int32)
    for t in T.TYPES:
        c = cost[:, T.COST_OFFSETS[t.idx]:T.COST_OFFSETS[t.idx + 1]].reshape(n, t.n, t.modes)
        order = np.argsort(c, axis=2, kind="stable")[:, :, :k]
        sel = np.take_along_axis(c, order, axis=2)
        skipped = c[:, :, :1] == -1
        tm[:, T.CU_OFFSETS[t.idx]:T.CU_OFFSETS[t.idx + 1]] = np.where(skipped, 255, order).astype(np.uint8)
        tc[:, T.CU_OFFSETS[t.idx]:T.CU_OFFSETS[t.idx + 1]] = np.where(skipped, -1, sel)
    return tm, tc


def satd4x4(diff16) -> int:
    d = np.ascontiguousarray(diff16, dtype=np.intc).reshape(16)
    return int(lib().mipo_satd4x4(_ci(d)))


def reduced_prediction(size_id: int, mode: int, redT, redL) -> np.ndarray:
    r = 8 if size_id == 2 else 4
    t = np.ascontiguousarray(redT, dtype=np.intc)
    l = np.ascontiguousarray(redL, dtype=np.intc)
    out = np.zeros(r * r, dtype=np.intc)
    lib().mipo_reduced_prediction(size_id, mode, _ci(t), _ci(l), _ci(out))
    return out.reshape(r, r)


def upsample(red: np.ndarray, w: int, h: int, refT, refL) -> np.ndarray:
    red = np.ascontiguousarray(red, dtype=np.intc)
    r = red.shape[0]
    t = np.ascontiguousarray(refT, dtype=np.intc)
    l = np.ascontiguousarray(refL, dtype=np.intc)
    out = np.zeros(w * h, dtype=np.intc)
    lib().mipo_upsample(_ci(red), r, w, h, _ci(t), _ci(l), _ci(out))
    return out.reshape(h, w)


def cu_boundaries(F: np.ndarray, X: int, Y: int, w: int, h: int, b: int):
    F = np.ascontiguousarray(F, dtype=np.uint16)
    refT = np.zeros(w, dtype=np.intc)
    refL = np.zeros(h, dtype=np.intc)
    redT = np.zeros(b, dtype=np.intc)
    redL = np.zeros(b, dtype=np.intc)
    lib().mipo_cu_boundaries(_u16(F), F.shape[1], X, Y, w, h, b, _ci(refT), _ci(refL), _ci(redT), _ci(redL))
    return refT, refL, redT, redL


def fnv64(values: np.ndarray) -> int:
    """FNV-1a 64 over int32 values fed as 4 little-endian bytes each."""
    data = np.ascontiguousarray(values, dtype="<i4").tobytes()
    h = 0xCBF29CE484222325
    for byte in data:
        h = ((h ^ byte) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h
