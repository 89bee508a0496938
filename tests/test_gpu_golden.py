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
