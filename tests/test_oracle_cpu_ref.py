"""The oracle's 2-D integer filters against the reference's OWN CPU implementation of them
(parallelOptFilterCpuInt_3x3 / _5x5, reference main_aux_functions.h:1323-1773), compiled from the reference checkout by
oracle/cpu_ref/Makefile into oracle/_ref/libref_cpu_filters.so.  A second, GPU-free pin next to the OpenCL fixtures
(SURVEY.md 8(c) item 2).  Skipped where neither the built library nor /root/reference exists."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_cpu_filters.so")


@pytest.fixture(scope="module")
def cpu_ref():
    if not os.path.exists(LIB):
        if not os.path.isdir("/root/reference"):
            pytest.skip("oracle/_ref/libref_cpu_filters.so not built and /root/reference not mounted")
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle", "cpu_ref")], check=True, capture_output=True)
    lib = ctypes.CDLL(LIB)
    u16p = ctypes.POINTER(ctypes.c_uint16)
    lib.mipref_cpu_filter_int.argtypes = [u16p, u16p] + [ctypes.c_int] * 5
    lib.mipref_cpu_filter_int.restype = None

    def run(frame, taps, kidx, threads=4):
        frame = np.ascontiguousarray(frame, dtype=np.uint16)
        out = np.zeros_like(frame)
        lib.mipref_cpu_filter_int(frame.ctypes.data_as(u16p), out.ctypes.data_as(u16p), frame.shape[1], frame.shape[0], taps, kidx, threads)
        return out

    return run


@pytest.mark.parametrize("taps,ft,kidxs", [(3, 3, range(5)), (5, 7, range(3))])
def test_2d_int_filters_equal_the_references_cpu_filters(oracle, cpu_ref, taps, ft, kidxs):
    from mipb200 import frames
    cases = [frames.noise_frame(256, 136, 5), frames.natural_frame(384, 128, 6), frames.extreme_frame(128, 128, 2),
             frames.impulse_frame(192, 72, 7), frames.natural_frame(1920, 1080, 8)]
    for f in cases:
        for k in kidxs:
            want = cpu_ref(f, taps, k)
            got = oracle.filter_frame(f, ft, k)
            assert np.array_equal(got, want), (taps, k, f.shape, int(np.abs(got.astype(int) - want.astype(int)).max()))


def test_thread_count_does_not_matter(cpu_ref):
    from mipb200 import frames
    f = frames.noise_frame(128, 64, 1)
    assert np.array_equal(cpu_ref(f, 5, 1, threads=1), cpu_ref(f, 5, 1, threads=8))
