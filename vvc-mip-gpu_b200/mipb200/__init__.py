"""mipb200 -- thin ctypes view of the C ABI in include/mipb200.h (libmipb200.so).

The product is the CUDA library and the C++ CLI (csrc/); this module only lets the tests
and bench.py call the same entry points.  There is no CPU path here: importing works
without a GPU (so that symbol/loader tests can run), creating an Engine does not.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

from . import tables  # noqa: F401

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("MIPB200_LIB") or os.path.join(_PKG, "lib", "libmipb200.so")   # override: A/B builds only
CLI_PATH = os.path.join(_PKG, "bin", "mipb200_main")
CSRC = os.path.join(_PKG, "csrc")

COSTS_PER_CTU = 97840
CUS_PER_CTU = 5380
SKIPPED = -1

EMIT_COSTS = 1
EMIT_SAD_SATD = 2
EMIT_DECISIONS = 4
EMIT_COSTS_COMPACT = 8
COMPACT_BYTES_PER_CTU = 276672
TOPK_MAX = 12
LAUNCH_AUTO, LAUNCH_THROUGHPUT, LAUNCH_LATENCY = 0, 1, 2

# every symbol include/mipb200.h declares (tests/test_abi.py checks the header against this)
ABI_SYMBOLS = (
    "mipb200_create", "mipb200_destroy", "mipb200_next_input", "mipb200_submit", "mipb200_collect",
    "mipb200_in_flight", "mipb200_set_launch_mode", "mipb200_num_ctus", "mipb200_device_count", "mipb200_run_device", "mipb200_filter_device",
    "mipb200_decide_device", "mipb200_topk_device", "mipb200_compact_bytes_per_ctu", "mipb200_expand_costs", "mipb200_kernel_launches", "mipb200_device_energy_mj", "mipb200_pin_host", "mipb200_pin_host_on", "mipb200_unpin_host", "mipb200_sync", "mipb200_last_error",
    "mipb200_version",
)


class MipError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mipb200 error {code}: {msg}")
        self.code = code


class Config(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int), ("height", ctypes.c_int), ("device", ctypes.c_int),
        ("filter_type", ctypes.c_int), ("kernel_idx", ctypes.c_int), ("slots", ctypes.c_int),
        ("emit", ctypes.c_uint), ("top_k", ctypes.c_int), ("bit_depth", ctypes.c_int),
    ]


class Result(ctypes.Structure):
    _fields_ = [
        ("poc", ctypes.c_int64), ("n_ctus", ctypes.c_int),
        ("cost", ctypes.POINTER(ctypes.c_int32)), ("sad", ctypes.POINTER(ctypes.c_int32)),
        ("satd", ctypes.POINTER(ctypes.c_int32)), ("best_mode", ctypes.POINTER(ctypes.c_uint8)),
        ("best_cost", ctypes.POINTER(ctypes.c_int32)), ("gpu_ms", ctypes.c_float),
        ("top_k", ctypes.c_int), ("topk_mode", ctypes.POINTER(ctypes.c_uint8)), ("topk_cost", ctypes.POINTER(ctypes.c_int32)),
        ("cost_compact", ctypes.POINTER(ctypes.c_uint8)),
    ]


def build(verbose: bool = False) -> None:
    """Compile libmipb200.so and the CLI for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libmipb200.so failed:\n" + r.stdout + r.stderr)


_lib = None


def lib() -> ctypes.CDLL:
    """Load the CUDA library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32p, u16p, u8p = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_uint8)
        L.mipb200_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(Config)]
        L.mipb200_create.restype = ctypes.c_int
        L.mipb200_destroy.argtypes = [vp]
        L.mipb200_destroy.restype = None
        L.mipb200_next_input.argtypes = [vp]
        L.mipb200_next_input.restype = vp
        L.mipb200_submit.argtypes = [vp, vp, ctypes.c_int64]
        L.mipb200_submit.restype = ctypes.c_int
        L.mipb200_collect.argtypes = [vp, ctypes.POINTER(Result)]
        L.mipb200_collect.restype = ctypes.c_int
        L.mipb200_in_flight.argtypes = [vp]
        L.mipb200_in_flight.restype = ctypes.c_int
        L.mipb200_set_launch_mode.argtypes = [vp, ctypes.c_int]
        L.mipb200_set_launch_mode.restype = ctypes.c_int
        L.mipb200_num_ctus.argtypes = [ctypes.c_int, ctypes.c_int]
        L.mipb200_num_ctus.restype = ctypes.c_int
        L.mipb200_device_count.argtypes = []
        L.mipb200_device_count.restype = ctypes.c_int
        L.mipb200_run_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
        L.mipb200_run_device.restype = ctypes.c_int
        L.mipb200_filter_device.argtypes = [vp, vp, vp, vp]
        L.mipb200_filter_device.restype = ctypes.c_int
        L.mipb200_decide_device.argtypes = [vp, vp, vp, vp, vp]
        L.mipb200_decide_device.restype = ctypes.c_int
        L.mipb200_topk_device.argtypes = [vp, vp, ctypes.c_int, vp, vp, vp]
        L.mipb200_topk_device.restype = ctypes.c_int
        L.mipb200_compact_bytes_per_ctu.argtypes = []
        L.mipb200_compact_bytes_per_ctu.restype = ctypes.c_size_t
        L.mipb200_expand_costs.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int]
        L.mipb200_expand_costs.restype = ctypes.c_int
        L.mipb200_kernel_launches.argtypes = [vp]
        L.mipb200_kernel_launches.restype = ctypes.c_longlong
        L.mipb200_device_energy_mj.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]
        L.mipb200_device_energy_mj.restype = ctypes.c_int
        L.mipb200_pin_host.argtypes = [vp, ctypes.c_size_t]
        L.mipb200_pin_host.restype = ctypes.c_int
        L.mipb200_pin_host_on.argtypes = [ctypes.c_int, vp, ctypes.c_size_t]
        L.mipb200_pin_host_on.restype = ctypes.c_int
        L.mipb200_unpin_host.argtypes = [vp]
        L.mipb200_unpin_host.restype = ctypes.c_int
        L.mipb200_sync.argtypes = [vp]
        L.mipb200_sync.restype = ctypes.c_int
        L.mipb200_last_error.argtypes = []
        L.mipb200_last_error.restype = ctypes.c_char_p
        L.mipb200_version.argtypes = []
        L.mipb200_version.restype = ctypes.c_char_p
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise MipError(rc, lib().mipb200_last_error().decode())


def device_energy_mj(device: int = 0) -> int:
    """Board energy counter in millijoules (NVML); raises MipError where NVML or the counter is missing."""
    v = ctypes.c_ulonglong()
    _check(lib().mipb200_device_energy_mj(device, ctypes.byref(v)))
    return int(v.value)


_ARRAY_TYPES = {}


class FrameResult:
    """Host views (numpy, zero-copy over the engine's pinned ring) of one frame's results.
    Valid until the next collect() on the same engine; copy() what must outlive that."""

    def __init__(self, r: Result):
        n = r.n_ctus
        self.poc = int(r.poc)
        self.n_ctus = n
        self.gpu_ms = float(r.gpu_ms)

        def view(p, shape, dt):
            if not p:
                return None
            n = int(np.prod(shape))
            key = (dt, n)
            arr_t = _ARRAY_TYPES.get(key)
            if arr_t is None:       # building a ctypes array type per call costs ~0.1 ms for a 13 M-element table
                arr_t = _ARRAY_TYPES[key] = (ctypes.c_int32 if dt == np.int32 else ctypes.c_uint8) * n
            return np.frombuffer(arr_t.from_address(ctypes.addressof(p.contents)), dtype=dt).reshape(shape)

        self.cost = view(r.cost, (n, COSTS_PER_CTU), np.int32)
        self.sad = view(r.sad, (n, COSTS_PER_CTU), np.int32)
        self.satd = view(r.satd, (n, COSTS_PER_CTU), np.int32)
        self.best_mode = view(r.best_mode, (n, CUS_PER_CTU), np.uint8)
        self.best_cost = view(r.best_cost, (n, CUS_PER_CTU), np.int32)
        self.top_k = int(r.top_k)
        self.topk_mode = view(r.topk_mode, (n, CUS_PER_CTU, self.top_k), np.uint8) if self.top_k else None
        self.topk_cost = view(r.topk_cost, (n, CUS_PER_CTU, self.top_k), np.int32) if self.top_k else None
        self.cost_compact = view(r.cost_compact, (n, COMPACT_BYTES_PER_CTU), np.uint8)

    def expand_costs(self, threads: int = 1) -> np.ndarray:
        """The compact table (EMIT_COSTS_COMPACT) as the int32 table [n_ctus][97840]."""
        return expand_costs(self.cost_compact, threads)


def expand_costs(compact: np.ndarray, threads: int = 1) -> np.ndarray:
    """mipb200_expand_costs(): uint8 [n_ctus][COMPACT_BYTES_PER_CTU] -> int32 [n_ctus][97840]."""
    compact = np.ascontiguousarray(compact, dtype=np.uint8).reshape(-1, COMPACT_BYTES_PER_CTU)
    out = np.empty((compact.shape[0], COSTS_PER_CTU), dtype=np.int32)
    _check(lib().mipb200_expand_costs(compact.ctypes.data, compact.shape[0], out.ctypes.data, threads))
    return out


class Engine:
    """One GPU's MIP engine (mipb200_create .. mipb200_destroy)."""

    def __init__(self, width: int, height: int, device: int = 0, filter_type: int = 0, kernel_idx: int = 0,
                 slots: int = 3, emit: int = EMIT_COSTS, top_k: int = 0, bit_depth: int = 0):
        self._h = ctypes.c_void_p()
        self.cfg = Config(width, height, device, filter_type, kernel_idx, slots, emit, top_k, bit_depth)
        _check(lib().mipb200_create(ctypes.byref(self._h), ctypes.byref(self.cfg)))
        self.width, self.height = width, height
        self.n_ctus = lib().mipb200_num_ctus(width, height)
        self._pending = []              # frames in flight, oldest first

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            lib().mipb200_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- host path (pinned ring, async copies)
    def next_input(self) -> Optional[np.ndarray]:
        p = lib().mipb200_next_input(self._h)
        if not p:
            return None
        buf = (ctypes.c_uint16 * (self.width * self.height)).from_address(p)
        return np.frombuffer(buf, dtype=np.uint16).reshape(self.height, self.width)

    def submit(self, frame: np.ndarray, poc: int = 0) -> None:
        if frame.dtype != np.uint16 or frame.shape != (self.height, self.width) or not frame.flags.c_contiguous:
            raise ValueError("frame must be C-contiguous uint16 [height, width]")
        _check(lib().mipb200_submit(self._h, frame.ctypes.data, poc))
        self._pending.append(frame)     # a page-locked frame is DMA'd in place: keep it alive until it has been collected

    def collect(self) -> FrameResult:
        r = Result()
        _check(lib().mipb200_collect(self._h, ctypes.byref(r)))
        if self._pending:
            self._pending.pop(0)
        return FrameResult(r)

    def in_flight(self) -> int:
        return lib().mipb200_in_flight(self._h)

    def set_launch_mode(self, mode: int) -> None:
        """LAUNCH_AUTO / LAUNCH_THROUGHPUT / LAUNCH_LATENCY: how a frame is cut into thread blocks (see include/mipb200.h)."""
        _check(lib().mipb200_set_launch_mode(self._h, mode))

    def run(self, frame: np.ndarray) -> FrameResult:
        self.submit(np.ascontiguousarray(frame, dtype=np.uint16))
        return self.collect()

    # ---- device-resident path (raw device pointers, e.g. torch tensors' data_ptr())
    def run_device(self, d_frame: int, d_cost: int, d_sad: int = 0, d_satd: int = 0, d_best_mode: int = 0,
                   d_best_cost: int = 0, stream: int = 0) -> None:
        _check(lib().mipb200_run_device(self._h, d_frame, d_cost or None, d_sad or None, d_satd or None,
                                        d_best_mode or None, d_best_cost or None, stream or None))

    def filter_device(self, d_frame: int, d_out: int, stream: int = 0) -> None:
        _check(lib().mipb200_filter_device(self._h, d_frame, d_out, stream or None))

    def decide_device(self, d_cost: int, d_best_mode: int, d_best_cost: int, stream: int = 0) -> None:
        _check(lib().mipb200_decide_device(self._h, d_cost, d_best_mode, d_best_cost, stream or None))

    def topk_device(self, d_cost: int, k: int, d_modes: int, d_costs: int, stream: int = 0) -> None:
        _check(lib().mipb200_topk_device(self._h, d_cost, k, d_modes, d_costs, stream or None))

    def kernel_launches(self) -> int:
        return int(lib().mipb200_kernel_launches(self._h))

    def sync(self) -> None:
        _check(lib().mipb200_sync(self._h))
