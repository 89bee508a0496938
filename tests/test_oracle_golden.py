"""Pins of the CPU oracle.

1. survey_kat.json -- known-answer vectors generated during the survey from a literal
   re-enactment of the reference kernels' index arithmetic.
2. ocl_b200_*.npz   -- outputs of the reference's own unmodified OpenCL kernels executed on a
   B200 through NVIDIA's OpenCL driver by oracle/ocl_ref (see tests/golden/README.md); present
   only once that run has been made.
"""
import glob
import json
import os

import numpy as np
import pytest

from mipb200 import frames, tables as T

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KAT = json.load(open(os.path.join(GOLD, "survey_kat.json")))


def test_kat_frame_definition():
    f = frames.kat_frame(256, 256)
    assert f[0, :5].tolist() == KAT["frame"]["row0"]
    assert f[255, 251:].tolist() == KAT["frame"]["row255_tail"]
    assert int(f.sum()) == KAT["frame"]["sum"]


@pytest.mark.parametrize("case", KAT["filtered"], ids=lambda c: f"ft{c['filter_types']}k{c['kernel_idx']}")
def test_kat_filtered_frames(oracle, case):
    f = frames.kat_frame(256, 256)
    for ft in case["filter_types"]:
        g = oracle.filter_frame(f, ft, case["kernel_idx"])
        assert g[0, :5].tolist() == case["row0"]
        assert int(g.sum()) == case["sum"]


@pytest.fixture(scope="module")
def kat_runs(oracle):
    f = frames.kat_frame(256, 256)
    out = {}
    for ft, k in ((0, 0), (7, 2), (8, 2), (1, 1), (2, 1)):
        cost, sad, satd = oracle.run_frame(f, ft, k, want_sad_satd=True)
        bm, bc = oracle.decisions(cost)
        out[(ft, k)] = (cost, sad, satd, bm)
    return out


@pytest.mark.parametrize("case", KAT["ctu_sums"], ids=lambda c: f"ft{c['filter_types'][0]}k{c['kernel_idx']}ctu{c['ctu']}")
def test_kat_ctu_sums_and_hashes(oracle, kat_runs, case):
    for ft in case["filter_types"]:
        cost, sad, satd, bm = kat_runs[(ft, case["kernel_idx"])]
        c = case["ctu"]
        assert int(cost[c].astype(np.int64).sum()) == case["sum_cost"]
        assert int(sad[c].astype(np.int64).sum()) == case["sum_sad"]
        assert int(satd[c].astype(np.int64).sum()) == case["sum_satd"]
        assert f"{oracle.fnv64(cost[c]):016x}" == case["fnv_cost"]
        assert int(bm[c].astype(np.int64).sum()) == case["sum_best_mode"]
        assert f"{oracle.fnv64(bm[c].astype(np.int32)):016x}" == case["fnv_best_mode"]


def test_kat_individual_entries(kat_runs):
    cols = {"orig_ctu0": ((0, 0), 0), "orig_ctu3": ((0, 0), 3), "f7k2_ctu3": ((7, 2), 3), "f1k1_ctu3": ((1, 1), 3)}
    for e in KAT["entries"]:
        t = T.TYPES[e["t"]]
        idx = T.COST_OFFSETS[t.idx] + e["cu"] * t.modes + e["mode"]
        for key, (run, ctu) in cols.items():
            cost, sad, satd, _ = kat_runs[run]
            assert [int(sad[ctu, idx]), int(satd[ctu, idx]), int(cost[ctu, idx])] == e[key], (e, key)


def test_kat_per_type_sums(kat_runs):
    cost = kat_runs[(0, 0)][0]
    got = [int(cost[3, T.COST_OFFSETS[t.idx]:T.COST_OFFSETS[t.idx + 1]].astype(np.int64).sum()) for t in T.TYPES]
    assert got == KAT["per_type_sum_cost_orig_ctu3"]


OCL = sorted(glob.glob(os.path.join(GOLD, "ocl_b200_cost_*.npz")))
OCL_FILT = sorted(glob.glob(os.path.join(GOLD, "ocl_b200_filters_*.npz")))


@pytest.mark.skipif(not OCL, reason="no reference-OpenCL-on-B200 fixtures committed yet")
@pytest.mark.parametrize("path", OCL, ids=os.path.basename)
def test_oracle_equals_reference_opencl_kernels_on_b200(oracle, path):
    """minSadHad of the reference's untouched .cl kernels (frame 0) == oracle, on in-frame CUs."""
    z = np.load(path)
    frame = frames_from_fixture(z)
    ft, kidx = int(z["filter_type"]), int(z["kernel_idx"])
    want = z["cost"]                              # int32 [nCTU][97840] as read back from minSadHad
    got = oracle.run_frame(frame, ft, kidx)
    ok = got != -1                                # CUs fully inside the frame; the rest is garbage in the reference
    assert ok.any()
    assert np.array_equal(got[ok], want[ok]), f"{int((got[ok] != want[ok]).sum())} of {int(ok.sum())} in-frame costs differ"
    if "filtered" in z.files:
        assert np.array_equal(oracle.filter_frame(frame, ft, kidx), z["filtered"])


def frames_from_fixture(z):
    kind = str(z["frame_kind"])
    w, h, seed = int(z["width"]), int(z["height"]), int(z["seed"])
    return {"kat": lambda: frames.kat_frame(w, h), "noise": lambda: frames.noise_frame(w, h, seed),
            "natural": lambda: frames.natural_frame(w, h, seed)}[kind]()


@pytest.mark.skipif(not OCL_FILT, reason="no reference-OpenCL-on-B200 filter fixtures committed yet")
@pytest.mark.parametrize("path", OCL_FILT, ids=os.path.basename)
def test_oracle_filters_equal_reference_opencl_kernels_on_b200(oracle, path):
    """All 8 filterFrame_* kernels x every KernelIdx, run unmodified on the B200 == oracle, every pixel."""
    z = np.load(path)
    frame = frames_from_fixture(z)
    n = 0
    for ft in range(1, 9):
        for kidx in range(T.num_kernel_idx(ft)):
            want = z[f"f{ft}k{kidx}"]
            got = oracle.filter_frame(frame, ft, kidx)
            assert np.array_equal(got, want), f"filter_type={ft} kernel_idx={kidx}: {int((got != want).sum())} pixels differ"
            n += 1
    assert n == 32


# ---- whole BASELINE-size frames: per-CTU hashes of the reference's OpenCL output on a B200 (make_ocl_fullsize_hashes.py)
def _fullsize_cases():
    p = os.path.join(GOLD, "ocl_b200_fullsize_hashes.json")
    return json.load(open(p)) if os.path.exists(p) else []


def fullsize_helpers():
    import importlib.util
    import sys
    sys.path.insert(0, GOLD)
    spec = importlib.util.spec_from_file_location("make_ocl_fullsize_hashes", os.path.join(GOLD, "make_ocl_fullsize_hashes.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def fullsize_frame(case):
    w, h, seed = case["width"], case["height"], case["seed"]
    return frames.noise_frame(w, h, seed) if case["frame_kind"] == "noise" else frames.natural_frame(w, h, seed)


@pytest.mark.parametrize("case", _fullsize_cases(), ids=lambda c: f"{c['width']}x{c['height']}_f{c['filter_type']}k{c['kernel_idx']}")
def test_oracle_equals_reference_opencl_on_whole_frames(oracle, case):
    """1080p (original samples, the bench configuration, a 3x3 filter) and 2160p: every CTU of the oracle's table hashes to
    what the reference's unmodified kernels produced on the B200 (CUs not fully inside the frame masked to -1)."""
    m = fullsize_helpers()
    cost = oracle.run_frame(fullsize_frame(case), case["filter_type"], case["kernel_idx"])
    got = m.ctu_hashes(cost, case["width"], case["height"])
    bad = [i for i, (a, b) in enumerate(zip(got, case["sha256_16_per_ctu"])) if a != b]
    assert len(got) == len(case["sha256_16_per_ctu"]) and not bad, f"CTUs that differ from the reference: {bad[:20]}"
