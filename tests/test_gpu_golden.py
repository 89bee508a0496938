"""CUDA engine vs the reference's OWN kernels (fixtures from the unmodified intra.cl run on a B200 through NVIDIA's
OpenCL driver, tests/golden/ocl_b200_*.npz) -- no oracle in between."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _frame(z):
    from mipb200 import frames
    kind, w, h, seed = str(z["frame_kind"]), int(z["width"]), int(z["height"]), int(z["seed"])
    return {"kat": lambda: frames.kat_frame(w, h), "noise": lambda: frames.noise_frame(w, h, seed),
            "natural": lambda: frames.natural_frame(w, h, seed)}[kind]()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ocl_b200_cost_*.npz"))), ids=os.path.basename)
def test_costs_equal_reference_opencl(mip, path):
    z = np.load(path)
    f = _frame(z)
    h, w = f.shape
    with mip.Engine(w, h, filter_type=int(z["filter_type"]), kernel_idx=int(z["kernel_idx"]), slots=1, emit=mip.EMIT_COSTS) as eng:
        got = eng.run(f).cost.copy()
    want = z["cost"]
    ok = got != -1            # CUs fully inside the frame (the reference leaves garbage in the others)
    assert ok.sum() > 0.5 * ok.size
    assert np.array_equal(got[ok], want[ok]), f"{int((got[ok] != want[ok]).sum())} of {int(ok.sum())} in-frame costs differ from the reference kernels"


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ocl_b200_filters_*.npz"))), ids=os.path.basename)
def test_filters_equal_reference_opencl(mip, path):
    import torch
    from mipb200 import tables
    z = np.load(path)
    f = _frame(z)
    h, w = f.shape
    d_in = torch.from_numpy(f.view(np.int16)).cuda()
    d_out = torch.empty_like(d_in)
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    n = 0
    for ft in range(1, 9):
        for kidx in range(tables.num_kernel_idx(ft)):
            with mip.Engine(w, h, filter_type=ft, kernel_idx=kidx, slots=1) as eng:
                eng.filter_device(d_in.data_ptr(), d_out.data_ptr(), st.cuda_stream)
                st.synchronize()
            got = d_out.cpu().numpy().view(np.uint16)
            assert np.array_equal(got, z[f"f{ft}k{kidx}"]), f"filter_type={ft} kernel_idx={kidx}"
            n += 1
    assert n == 32


def _fullsize():
    import test_oracle_golden as G
    return G


@pytest.mark.gpu
@pytest.mark.parametrize("case", __import__("test_oracle_golden")._fullsize_cases(),
                         ids=lambda c: f"{c['width']}x{c['height']}_f{c['filter_type']}k{c['kernel_idx']}")
def test_engine_equals_reference_opencl_on_whole_frames(mip, case):
    """The CUDA engine against the reference's own OpenCL output for whole 1080p / 2160p frames (per-CTU hashes), no oracle
    in between."""
    G = _fullsize()
    m = G.fullsize_helpers()
    f = G.fullsize_frame(case)
    with mip.Engine(case["width"], case["height"], filter_type=case["filter_type"], kernel_idx=case["kernel_idx"], slots=1) as eng:
        got = m.ctu_hashes(eng.run(f).cost, case["width"], case["height"])
    bad = [i for i, (a, b) in enumerate(zip(got, case["sha256_16_per_ctu"])) if a != b]
    assert len(got) == len(case["sha256_16_per_ctu"]) and not bad, f"CTUs that differ from the reference: {bad[:20]}"
