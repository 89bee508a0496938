"""CUDA engine vs the CPU oracle through the C ABI: bit-exact costs (integer work).

Sizes are chosen so that the oracle finishes in seconds: 256x256 (all four frame-edge
rules, 4 CTUs), 256x184 (partial bottom CTU row, 56 valid rows like 1080p), 384x128.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(mip, frame, ft=0, kidx=0, emit=None):
    h, w = frame.shape
    emit = emit if emit is not None else (mip.EMIT_COSTS | mip.EMIT_SAD_SATD | mip.EMIT_DECISIONS)
    with mip.Engine(w, h, filter_type=ft, kernel_idx=kidx, slots=2, emit=emit) as eng:
        r = eng.run(frame)
        return {k: (None if getattr(r, k) is None else getattr(r, k).copy()) for k in ("cost", "sad", "satd", "best_mode", "best_cost")}


def _assert_same(got, want, what):
    if not np.array_equal(got, want):
        bad = np.argwhere(got != want)
        raise AssertionError(f"{what}: {len(bad)} of {want.size} differ; first {bad[:5].tolist()} got {got[tuple(bad[0])]} want {want[tuple(bad[0])]}")


@pytest.mark.parametrize("shape", [(256, 256), (184, 256), (128, 384)])
@pytest.mark.parametrize("content", ["kat", "noise", "natural", "checker", "zeros", "ones"])
def test_costs_original_samples(mip, oracle, shape, content):
    from mipb200 import frames
    h, w = shape
    f = {"kat": lambda: frames.kat_frame(w, h), "noise": lambda: frames.noise_frame(w, h, 1),
         "natural": lambda: frames.natural_frame(w, h, 2), "checker": lambda: frames.extreme_frame(w, h, 2),
         "zeros": lambda: frames.extreme_frame(w, h, 0), "ones": lambda: frames.extreme_frame(w, h, 1)}[content]()
    got = _run(mip, f)
    cost, sad, satd = oracle.run_frame(f, want_sad_satd=True)
    _assert_same(got["cost"], cost, "cost")
    _assert_same(got["sad"], sad, "sad")
    _assert_same(got["satd"], satd, "satd")
    bm, bc = oracle.decisions(cost)
    _assert_same(got["best_mode"], bm, "best_mode")
    _assert_same(got["best_cost"], bc, "best_cost")


@pytest.mark.parametrize("ft", [1, 2, 3, 4, 5, 6, 7, 8])
def test_costs_all_filters_all_kernel_idx(mip, oracle, ft):
    from mipb200 import frames, tables
    f = frames.noise_frame(256, 184, 7)
    for kidx in range(tables.num_kernel_idx(ft)):
        got = _run(mip, f, ft, kidx, emit=mip.EMIT_COSTS)
        _assert_same(got["cost"], oracle.run_frame(f, ft, kidx), f"cost ft={ft} kidx={kidx}")


@pytest.mark.parametrize("ft", [1, 3, 5, 7])
def test_filter_only(mip, oracle, ft):
    import torch
    from mipb200 import frames, tables
    f = frames.natural_frame(256, 88, 3)
    for kidx in range(tables.num_kernel_idx(ft)):
        with mip.Engine(256, 88, filter_type=ft, kernel_idx=kidx, slots=1) as eng:
            d_in = torch.from_numpy(f.view(np.int16)).cuda()
            d_out = torch.empty_like(d_in)
            torch.cuda.synchronize()
            st = torch.cuda.Stream()
            eng.filter_device(d_in.data_ptr(), d_out.data_ptr(), st.cuda_stream)
            st.synchronize()
            got = d_out.cpu().numpy().view(np.uint16)
        _assert_same(got, oracle.filter_frame(f, ft, kidx), f"filter ft={ft} kidx={kidx}")


def test_pipelined_frames_fifo_and_determinism(mip, oracle):
    """3 slots, 7 frames in flight-order; results come back FIFO and equal the oracle."""
    from mipb200 import frames
    fs = [frames.noise_frame(256, 128, 10 + i) for i in range(7)]
    want = [oracle.run_frame(f) for f in fs]
    with mip.Engine(256, 128, slots=3, emit=mip.EMIT_COSTS) as eng:
        got, sub = [], 0
        while len(got) < len(fs):
            while sub < len(fs) and eng.in_flight() < 3:
                buf = eng.next_input()
                buf[...] = fs[sub]
                eng.submit(buf, poc=sub)
                sub += 1
            r = eng.collect()
            assert r.poc == len(got)
            got.append(r.cost.copy())
    for g, w in zip(got, want):
        _assert_same(g, w, "pipelined cost")


def test_device_resident_path(mip, oracle):
    import torch
    from mipb200 import frames
    f = frames.natural_frame(256, 256, 5)
    with mip.Engine(256, 256, slots=1) as eng:
        d_in = torch.from_numpy(f.view(np.int16)).cuda()
        d_cost = torch.empty((eng.n_ctus, mip.COSTS_PER_CTU), dtype=torch.int32, device="cuda")
        d_bm = torch.empty((eng.n_ctus, mip.CUS_PER_CTU), dtype=torch.uint8, device="cuda")
        d_bc = torch.empty((eng.n_ctus, mip.CUS_PER_CTU), dtype=torch.int32, device="cuda")
        n0 = eng.kernel_launches()
        torch.cuda.synchronize()
        st = torch.cuda.Stream()
        eng.run_device(d_in.data_ptr(), d_cost.data_ptr(), d_best_mode=d_bm.data_ptr(), d_best_cost=d_bc.data_ptr(),
                       stream=st.cuda_stream)
        st.synchronize()                      # the work ran on the stream we passed, nowhere else
        assert eng.kernel_launches() - n0 == 1          # filter, costs and the per-CU argmin are one fused kernel
        want = oracle.run_frame(f)
        _assert_same(d_cost.cpu().numpy(), want, "device cost")
        bm, bc = oracle.decisions(want)
        _assert_same(d_bm.cpu().numpy(), bm, "device best_mode")
        _assert_same(d_bc.cpu().numpy(), bc, "device best_cost")
        # decisions without the cost table, and the stand-alone argmin over an existing table
        d_bm2, d_bc2 = torch.empty_like(d_bm), torch.empty_like(d_bc)
        eng.run_device(d_in.data_ptr(), 0, d_best_mode=d_bm2.data_ptr(), d_best_cost=d_bc2.data_ptr(), stream=st.cuda_stream)
        st.synchronize()
        _assert_same(d_bm2.cpu().numpy(), bm, "decisions-only best_mode")
        _assert_same(d_bc2.cpu().numpy(), bc, "decisions-only best_cost")
        d_bm2.zero_(); d_bc2.zero_(); torch.cuda.synchronize()
        eng.decide_device(d_cost.data_ptr(), d_bm2.data_ptr(), d_bc2.data_ptr(), stream=st.cuda_stream)
        st.synchronize()
        _assert_same(d_bm2.cpu().numpy(), bm, "decide_device best_mode")
        _assert_same(d_bc2.cpu().numpy(), bc, "decide_device best_cost")


@pytest.mark.parametrize("k", [1, 3, 12])
def test_topk_shortlist(mip, oracle, k):
    """k cheapest modes per CU in (cost, mode) order: the engine's result ring and the device entry point vs a stable sort
    of the oracle's table; ties (flat frame: every mode of a CU costs the same) must come out in mode order."""
    import torch
    from mipb200 import frames
    for f, ft in ((frames.natural_frame(256, 136, 31), 0), (np.full((128, 128), 700, np.uint16), 0), (frames.noise_frame(384, 128, 5), 7)):
        h, w = f.shape
        want_cost = oracle.run_frame(f, ft, 1)
        tm, tc = oracle.topk(want_cost, k)
        with mip.Engine(w, h, filter_type=ft, kernel_idx=1, emit=mip.EMIT_DECISIONS, top_k=k) as eng:
            r = eng.run(f)
            assert r.cost is None
            bm, bc = oracle.decisions(want_cost)
            _assert_same(r.best_mode, bm, "best_mode")
            _assert_same(r.best_cost, bc, "best_cost")
            if k > 1:
                assert r.top_k == k
                _assert_same(r.topk_mode, tm, "topk_mode")
                _assert_same(r.topk_cost, tc, "topk_cost")
            else:
                assert r.top_k == 0 and r.topk_mode is None
            d_cost = torch.from_numpy(want_cost).cuda()
            d_m = torch.zeros((eng.n_ctus, mip.CUS_PER_CTU, k), dtype=torch.uint8, device="cuda")
            d_c = torch.zeros((eng.n_ctus, mip.CUS_PER_CTU, k), dtype=torch.int32, device="cuda")
            st = torch.cuda.Stream()
            torch.cuda.synchronize()
            eng.topk_device(d_cost.data_ptr(), k, d_m.data_ptr(), d_c.data_ptr(), stream=st.cuda_stream)
            st.synchronize()
            _assert_same(d_m.cpu().numpy(), tm, "topk_device modes")
            _assert_same(d_c.cpu().numpy(), tc, "topk_device costs")
    with pytest.raises(mip.MipError):
        mip.Engine(128, 128, emit=mip.EMIT_COSTS, top_k=4)        # a shortlist needs EMIT_DECISIONS
    with pytest.raises(mip.MipError):
        mip.Engine(128, 128, emit=mip.EMIT_DECISIONS, top_k=13)


@pytest.mark.parametrize("bits", [8, 12])
def test_other_bit_depths(mip, oracle, bits):
    """8- and 12-bit pipelines (default sample 1 << (bits - 1), clamp (1 << bits) - 1): an extension of the reference's
    hard-wired 10 bits, checked against the oracle with the same generalisation; incl. saturated frames and a filter."""
    from mipb200 import frames
    top = (1 << bits) - 1
    cases = [(frames.noise_frame(256, 136, 3, bits=bits), 0, 0), (frames.natural_frame(384, 128, 4, bits=bits), 8, 2),
             (np.full((128, 136), top, np.uint16), 0, 0), (frames.noise_frame(136, 64, 9, bits=bits), 1, 4)]
    chk = np.indices((128, 128)).sum(axis=0) % 2 * top                        # 1-px checkerboard 0 / max
    cases.append((chk.astype(np.uint16), 3, 1))
    for f, ft, kidx in cases:
        h, w = f.shape
        assert int(f.max()) <= top
        want = oracle.run_frame(f, ft, kidx, want_sad_satd=True, bit_depth=bits)
        with mip.Engine(w, h, filter_type=ft, kernel_idx=kidx, emit=mip.EMIT_COSTS | mip.EMIT_SAD_SATD | mip.EMIT_DECISIONS,
                        bit_depth=bits) as eng:
            r = eng.run(f)
            _assert_same(r.cost, want[0], f"{bits}-bit cost")
            _assert_same(r.sad, want[1], f"{bits}-bit sad")
            _assert_same(r.satd, want[2], f"{bits}-bit satd")
            bm, bc = oracle.decisions(want[0])
            _assert_same(r.best_mode, bm, f"{bits}-bit best_mode")
            _assert_same(r.best_cost, bc, f"{bits}-bit best_cost")
    if bits == 12:   # the depth matters: the same 12-bit frame through the 10-bit pipeline clamps differently
        f = cases[0][0]
        assert not np.array_equal(oracle.run_frame(f, bit_depth=12), oracle.run_frame(f, bit_depth=10))
    with pytest.raises(mip.MipError):
        mip.Engine(128, 128, bit_depth=9)


def test_energy_counter(mip):
    """NVML energy counter through the ABI: monotonic, and it moves while the GPU works."""
    from mipb200 import frames
    try:
        e0 = mip.device_energy_mj(0)
    except mip.MipError as ex:
        pytest.skip(f"no NVML energy counter on this box: {ex}")
    f = frames.natural_frame(1920, 1080, 1)
    with mip.Engine(1920, 1080, emit=mip.EMIT_DECISIONS) as eng:
        for _ in range(200):
            eng.run(f)
    e1 = mip.device_energy_mj(0)
    assert e1 > e0


def test_errors(mip):
    with pytest.raises(mip.MipError):
        mip.Engine(250, 128)
    with pytest.raises(mip.MipError):
        mip.Engine(256, 128, filter_type=9)
    with pytest.raises(mip.MipError):
        mip.Engine(256, 128, filter_type=5, kernel_idx=3)
    with mip.Engine(256, 128, slots=1) as eng:
        with pytest.raises(mip.MipError):
            eng.collect()


def test_submit_from_callers_pinned_memory(mip, oracle):
    """A frame that already lives in page-locked memory is DMA'd in place (no staging copy); same result."""
    import torch
    from mipb200 import frames
    fs = [frames.noise_frame(256, 128, 70 + i) for i in range(4)]
    pin = torch.empty((4, 128, 256), dtype=torch.int16, pin_memory=True)
    pin.numpy()[...] = np.stack(fs).view(np.int16)
    with mip.Engine(256, 128, slots=2, emit=mip.EMIT_COSTS) as eng:
        got = []
        for i in range(4):
            eng.submit(pin[i].numpy().view(np.uint16), poc=i)
            if eng.in_flight() == 2:
                got.append(eng.collect().cost.copy())
        while eng.in_flight():
            got.append(eng.collect().cost.copy())
    for g, f in zip(got, fs):
        _assert_same(g, oracle.run_frame(f), "pinned-source cost")


@pytest.mark.parametrize("shape", [(4, 128), (8, 128), (12, 256), (60, 128), (64, 256), (68, 128), (132, 384)])
@pytest.mark.parametrize("ft,kidx", [(0, 0), (7, 1), (2, 3), (5, 0)])
def test_ragged_heights(mip, oracle, shape, ft, kidx):
    """Heights that leave one, a few or no valid rows in a tile half (skipped halves, 4-row frames, 1 row past a half)."""
    from mipb200 import frames
    h, w = shape
    f = frames.noise_frame(w, h, 90 + h)
    got = _run(mip, f, ft, kidx, emit=mip.EMIT_COSTS | mip.EMIT_DECISIONS)
    want = oracle.run_frame(f, ft, kidx)
    _assert_same(got["cost"], want, f"cost {w}x{h} ft={ft}")
    bm, bc = oracle.decisions(want)
    _assert_same(got["best_mode"], bm, "best_mode")
    _assert_same(got["best_cost"], bc, "best_cost")


@pytest.mark.parametrize("shape", [(240, 416), (480, 832), (64, 8), (72, 136), (128, 248)])
@pytest.mark.parametrize("ft,kidx", [(0, 0), (8, 2), (1, 4)])
def test_widths_that_are_not_multiples_of_128(mip, oracle, shape, ft, kidx):
    """The reference whitelists 832x480 and 416x240 but has no x-guards there; here CUs crossing the right edge are skipped."""
    from mipb200 import frames, tables as T
    h, w = shape
    f = frames.natural_frame(w, h, 17)
    got = _run(mip, f, ft, kidx, emit=mip.EMIT_COSTS | mip.EMIT_DECISIONS)
    want = oracle.run_frame(f, ft, kidx)
    _assert_same(got["cost"], want, f"cost {w}x{h} ft={ft}")
    bm, bc = oracle.decisions(want)
    _assert_same(got["best_mode"], bm, "best_mode")
    assert np.array_equal(got["best_mode"] != 0xFF, T.in_frame_mask(w, h))


def test_random_geometries(mip, oracle):
    """Seeded sweep over odd-but-legal geometries (W % 8 == 0, H % 4 == 0: partial right and bottom CTUs, single partial
    CTU, one-row-of-4 frames), contents, filters and bit depths; cost, SAD, SATD and decisions against the oracle."""
    from mipb200 import frames, tables as T
    rng = np.random.default_rng(20261018)
    for it in range(14):
        w = int(rng.integers(1, 61)) * 8
        h = int(rng.integers(1, 81)) * 4
        ft = int(rng.integers(0, 9))
        kidx = int(rng.integers(0, T.num_kernel_idx(ft))) if ft else 0
        bits = (10, 10, 8, 12)[it % 4]
        kind = it % 3
        f = (frames.noise_frame(w, h, 100 + it, bits=bits) if kind == 0 else
             frames.natural_frame(w, h, 100 + it, bits=bits) if kind == 1 else
             (rng.integers(0, 2, size=(h, w)) * ((1 << bits) - 1)).astype(np.uint16))
        want = oracle.run_frame(f, ft, kidx, want_sad_satd=True, bit_depth=bits)
        with mip.Engine(w, h, filter_type=ft, kernel_idx=kidx, slots=1, bit_depth=bits,
                        emit=mip.EMIT_COSTS | mip.EMIT_SAD_SATD | mip.EMIT_DECISIONS) as eng:
            r = eng.run(f)
            tag = f"{w}x{h} ft={ft} k={kidx} bits={bits} kind={kind}"
            _assert_same(r.cost, want[0], tag + " cost")
            _assert_same(r.sad, want[1], tag + " sad")
            _assert_same(r.satd, want[2], tag + " satd")
            bm, bc = oracle.decisions(want[0])
            _assert_same(r.best_mode, bm, tag + " best_mode")
            _assert_same(r.best_cost, bc, tag + " best_cost")


@pytest.mark.parametrize("shape", [(8, 4), (8, 8), (16, 4), (8, 64), (64, 4), (128, 4), (8, 128), (136, 132), (24, 20)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_tiny_frames(mip, oracle, shape):
    """Frames smaller than the TMA box (144x69) down to the smallest legal one (8x4: two 4x4 CUs and one 8x4 CU fit)."""
    from mipb200 import frames
    w, h = shape
    for ft in (0, 3, 6):
        f = frames.noise_frame(w, h, w * 1000 + h)
        want = oracle.run_frame(f, ft, 1)
        assert (want != -1).any()
        with mip.Engine(w, h, filter_type=ft, kernel_idx=1, slots=1, emit=mip.EMIT_COSTS | mip.EMIT_DECISIONS) as eng:
            r = eng.run(f)
            _assert_same(r.cost, want, f"{w}x{h} ft={ft} cost")
            bm, bc = oracle.decisions(want)
            _assert_same(r.best_mode, bm, "best_mode")
            _assert_same(r.best_cost, bc, "best_cost")
