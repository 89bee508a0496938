"""Engine contract on the GPU (include/mipb200.h): result lifetime, device restore, pointer kinds, in-place DMA -- and the
memory-safety checks that stand in for compute-sanitizer (closed on this GPU pool, see profiles/r02_sanitizer.txt): every
output buffer sits between canary zones and starts poisoned; after the kernels ran, the canaries are intact (nothing was
written out of bounds), no poison is left where a result belongs (nothing was left uninitialised) and the results equal the
oracle (nothing raced: the TMA box / scratch overlay, the shared-memory argmin and the task counter all feed every value)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_results_stay_valid_across_the_next_submit(mip, oracle):
    """collect -> submit -> read: the header promises that result pointers live until the next collect()."""
    from mipb200 import frames
    W, H = 256, 136
    fs = [frames.noise_frame(W, H, 400 + i) for i in range(5)]
    want = [oracle.run_frame(f) for f in fs]
    for slots in (1, 2, 3):
        with mip.Engine(W, H, slots=slots, emit=mip.EMIT_COSTS | mip.EMIT_DECISIONS) as eng:
            sub = 0
            while sub < min(slots, 5):
                eng.submit(fs[sub], sub)
                sub += 1
            for got in range(5):
                r = eng.collect()                  # views into the ring, not copies
                while sub < 5 and eng.in_flight() < slots:
                    eng.submit(fs[sub], sub)       # refill every free slot BEFORE looking at the result
                    sub += 1
                eng.sync()                         # everything queued has run: a slot shared with `r` would be overwritten by now
                assert r.poc == got
                assert np.array_equal(r.cost, want[got]), (slots, got)
                bm, bc = oracle.decisions(want[got])
                assert np.array_equal(r.best_mode, bm) and np.array_equal(r.best_cost, bc)


def test_current_device_is_restored_and_device_pointers_are_refused(mip):
    import torch
    from mipb200 import frames
    n = torch.cuda.device_count()
    other = n - 1                                   # with one GPU this is device 0 itself: still must not change
    torch.cuda.set_device(other)
    f = frames.noise_frame(128, 128, 1)
    with mip.Engine(128, 128, device=0, slots=2, emit=mip.EMIT_DECISIONS) as eng:
        assert torch.cuda.current_device() == other
        eng.submit(f, 0)
        assert torch.cuda.current_device() == other
        eng.collect()
        d = torch.zeros((128, 128), dtype=torch.int16, device="cuda:0")
        with pytest.raises(mip.MipError) as ei:
            mip._check(mip.lib().mipb200_submit(eng._h, d.data_ptr(), 1))
        assert ei.value.code == -1 and "mipb200_run_device" in str(ei.value)
        assert eng.in_flight() == 0
    assert torch.cuda.current_device() == other
    torch.cuda.set_device(0)


def test_pinned_frames_are_used_in_place_and_kept_alive(mip, oracle):
    """A page-locked frame is DMA'd where it lies; the Python wrapper holds a reference until the frame is collected, so a
    temporary may be dropped right after submit()."""
    import torch
    from mipb200 import frames
    W, H = 256, 128
    fs = [frames.natural_frame(W, H, 40 + i) for i in range(3)]
    with mip.Engine(W, H, filter_type=5, kernel_idx=1, slots=3, emit=mip.EMIT_COSTS) as eng:
        for i, f in enumerate(fs):
            t = torch.from_numpy(f.view(np.int16)).pin_memory()
            eng.submit(t.numpy().view(np.uint16), i)
            del t
        torch.cuda.empty_cache()
        _ = [torch.empty(1 << 20, dtype=torch.uint8).pin_memory() for _ in range(4)]      # would recycle freed pinned blocks
        for i, f in enumerate(fs):
            assert np.array_equal(eng.collect().cost, oracle.run_frame(f, 5, 1)), i
    # mipb200_pin_host_on(): caller-owned pageable memory, page-locked in place
    buf = np.ascontiguousarray(np.stack(fs))
    mip._check(mip.lib().mipb200_pin_host_on(0, buf.ctypes.data, buf.nbytes))
    try:
        with mip.Engine(W, H, slots=2, emit=mip.EMIT_COSTS) as eng:
            eng.submit(buf[1], 1)
            assert np.array_equal(eng.collect().cost, oracle.run_frame(fs[1]))
    finally:
        mip._check(mip.lib().mipb200_unpin_host(buf.ctypes.data))


POISON32 = -1515870811          # 0xA5A5A5A5: never a legal cost (>= -1), mode byte 0xA5 = 165 is never a legal mode (< 32 or 0xFF)
GUARD = 4096                    # canary elements before and after every buffer


def _guarded(torch, n, dtype, fill):
    t = torch.full((n + 2 * GUARD,), fill, dtype=dtype, device="cuda")
    return t, t[GUARD:GUARD + n]


def _canaries_intact(t, fill):
    return bool((t[:GUARD] == fill).all()) and bool((t[-GUARD:] == fill).all())


@pytest.mark.parametrize("size", [(256, 184), (136, 72), (384, 260)])
@pytest.mark.parametrize("ft,kidx", [(0, 0), (1, 3), (3, 4), (5, 2), (7, 0), (8, 2)])
def test_no_out_of_bounds_no_uninitialised_outputs(mip, oracle, size, ft, kidx):
    import torch
    from mipb200 import frames, tables as T
    W, H = size
    f = frames.noise_frame(W, H, 7 * ft + W)
    want, wsad, wsatd = oracle.run_frame(f, ft, kidx, want_sad_satd=True)
    wbm, wbc = oracle.decisions(want)
    wtm, wtc = oracle.topk(want, 3)
    with mip.Engine(W, H, filter_type=ft, kernel_idx=kidx, slots=1, emit=mip.EMIT_COSTS) as eng:
        n = eng.n_ctus
        # the frame too: TMA reads a box that overhangs the frame on every side; the hardware must clip it
        ft_all, ft_in = _guarded(torch, W * H, torch.int16, 0x2AAA)
        ft_in.copy_(torch.from_numpy(f.view(np.int16)).reshape(-1))
        bufs = {k: _guarded(torch, n * mip.COSTS_PER_CTU, torch.int32, POISON32) for k in ("cost", "sad", "satd")}
        bufs["bc"] = _guarded(torch, n * mip.CUS_PER_CTU, torch.int32, POISON32)
        bufs["bm"] = _guarded(torch, n * mip.CUS_PER_CTU, torch.uint8, 0xA5)
        bufs["tm"] = _guarded(torch, n * mip.CUS_PER_CTU * 3, torch.uint8, 0xA5)
        bufs["tc"] = _guarded(torch, n * mip.CUS_PER_CTU * 3, torch.int32, POISON32)
        st = torch.cuda.current_stream().cuda_stream
        p = {k: v[1].data_ptr() for k, v in bufs.items()}
        eng.run_device(ft_in.data_ptr(), p["cost"], d_sad=p["sad"], d_satd=p["satd"], d_best_mode=p["bm"], d_best_cost=p["bc"], stream=st)
        eng.topk_device(p["cost"], 3, p["tm"], p["tc"], stream=st)
        torch.cuda.synchronize()
        for k, (whole, _) in bufs.items():
            assert _canaries_intact(whole, 0xA5 if whole.dtype == torch.uint8 else POISON32), f"{k}: written out of bounds"
        got = {k: v[1].cpu().numpy() for k, v in bufs.items()}
        for k in ("cost", "sad", "satd", "bc", "tc"):
            assert not (got[k] == POISON32).any(), f"{k}: {int((got[k] == POISON32).sum())} elements never written"
        for k in ("bm", "tm"):
            assert not (got[k] == 0xA5).any(), f"{k}: elements never written"
        assert np.array_equal(got["cost"].reshape(n, -1), want) and np.array_equal(got["sad"].reshape(n, -1), wsad) and np.array_equal(got["satd"].reshape(n, -1), wsatd)
        assert np.array_equal(got["bm"].reshape(n, -1), wbm) and np.array_equal(got["bc"].reshape(n, -1), wbc)
        assert np.array_equal(got["tm"].reshape(n, -1, 3), wtm) and np.array_equal(got["tc"].reshape(n, -1, 3), wtc)
        assert np.array_equal(ft_in.cpu().numpy().view(np.uint16).reshape(H, W), f) and _canaries_intact(ft_all, 0x2AAA)
        # decisions-only launch (no table), then the stand-alone argmin and filter kernels
        bufs2 = {"bc": _guarded(torch, n * mip.CUS_PER_CTU, torch.int32, POISON32), "bm": _guarded(torch, n * mip.CUS_PER_CTU, torch.uint8, 0xA5)}
        eng.run_device(ft_in.data_ptr(), 0, d_best_mode=bufs2["bm"][1].data_ptr(), d_best_cost=bufs2["bc"][1].data_ptr(), stream=st)
        torch.cuda.synchronize()
        assert _canaries_intact(bufs2["bm"][0], 0xA5) and _canaries_intact(bufs2["bc"][0], POISON32)
        assert np.array_equal(bufs2["bm"][1].cpu().numpy().reshape(n, -1), wbm) and np.array_equal(bufs2["bc"][1].cpu().numpy().reshape(n, -1), wbc)
        bufs2["bm"][1].fill_(0xA5); bufs2["bc"][1].fill_(POISON32)
        eng.decide_device(p["cost"], bufs2["bm"][1].data_ptr(), bufs2["bc"][1].data_ptr(), stream=st)
        torch.cuda.synchronize()
        assert _canaries_intact(bufs2["bm"][0], 0xA5) and _canaries_intact(bufs2["bc"][0], POISON32)
        assert np.array_equal(bufs2["bm"][1].cpu().numpy().reshape(n, -1), wbm) and np.array_equal(bufs2["bc"][1].cpu().numpy().reshape(n, -1), wbc)
        if ft:
            fo_all, fo = _guarded(torch, W * H, torch.int16, 0x2AAA)
            fo.fill_(-1)
            eng.filter_device(ft_in.data_ptr(), fo.data_ptr(), stream=st)
            torch.cuda.synchronize()
            assert _canaries_intact(fo_all, 0x2AAA)
            assert np.array_equal(fo.cpu().numpy().view(np.uint16).reshape(H, W), oracle.filter_frame(f, ft, kidx))


def test_repeated_launches_are_bit_identical(mip):
    """A race would show as run-to-run variation: 200 launches over 3 streams (CTAs of different frames share SMs), 2 frames."""
    import torch
    from mipb200 import frames
    W, H = 640, 360
    fs = torch.from_numpy(np.stack([frames.natural_frame(W, H, 5), frames.noise_frame(W, H, 6)]).view(np.int16)).cuda()
    with mip.Engine(W, H, filter_type=8, kernel_idx=2, slots=1, emit=mip.EMIT_COSTS) as eng:
        n = eng.n_ctus
        outs = [(torch.empty(n * mip.COSTS_PER_CTU, dtype=torch.int32, device="cuda"), torch.empty(n * mip.CUS_PER_CTU, dtype=torch.uint8, device="cuda"),
                 torch.empty(n * mip.CUS_PER_CTU, dtype=torch.int32, device="cuda")) for _ in range(6)]
        streams = [torch.cuda.Stream() for _ in range(3)]
        ref = {}
        for it in range(200):
            k = it % 6
            if it >= 6:
                streams[k % 3].synchronize()
                sig = (int(outs[k][0].to(torch.int64).sum()), int((outs[k][0].to(torch.int64) * 31 % 1000003).sum()), int(outs[k][1].to(torch.int64).sum()), int(outs[k][2].to(torch.int64).sum()))
                assert ref.setdefault((it - 6) % 2, sig) == sig, f"launch {it - 6} differs"
            c, m, b = outs[k]
            eng.run_device(fs[it % 2].data_ptr(), c.data_ptr(), d_best_mode=m.data_ptr(), d_best_cost=b.data_ptr(), stream=streams[k % 3].cuda_stream)
        torch.cuda.synchronize()


def test_launch_modes_give_identical_results(mip, oracle):
    """The throughput and the lone-frame split of a frame's work (mipb200_set_launch_mode) are two schedules of the same
    arithmetic: tables and decisions are identical, for the host path (AUTO picks either, by what is in flight) and for the
    device path."""
    import torch
    from mipb200 import frames
    W, H = 384, 200
    fs = [frames.noise_frame(W, H, 900 + i) for i in range(4)]
    want = [oracle.run_frame(f, 7, 1) for f in fs]
    for mode in (mip.LAUNCH_AUTO, mip.LAUNCH_THROUGHPUT, mip.LAUNCH_LATENCY):
        with mip.Engine(W, H, filter_type=7, kernel_idx=1, slots=3, emit=mip.EMIT_COSTS | mip.EMIT_DECISIONS) as eng:
            eng.set_launch_mode(mode)
            for i, f in enumerate(fs[:3]):
                eng.submit(f, i)               # AUTO: frame 0 enters an empty pipeline, 1 and 2 do not
            for i in range(3):
                r = eng.collect()
                bm, bc = oracle.decisions(want[i])
                assert np.array_equal(r.cost, want[i]) and np.array_equal(r.best_mode, bm) and np.array_equal(r.best_cost, bc), (mode, i)
            d_f = torch.from_numpy(fs[3].view(np.int16)).cuda()
            d_c = torch.empty((eng.n_ctus, mip.COSTS_PER_CTU), dtype=torch.int32, device="cuda")
            eng.run_device(d_f.data_ptr(), d_c.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert np.array_equal(d_c.cpu().numpy(), want[3]), mode
    with pytest.raises(mip.MipError):
        with mip.Engine(W, H, slots=1) as eng:
            eng.set_launch_mode(7)


@pytest.mark.parametrize("ft,kidx", [(0, 0), (8, 2)])
def test_compact_cost_table(mip, oracle, ft, kidx):
    """MIPB200_EMIT_COSTS_COMPACT: 71 % of the bytes, nothing lost -- expanded it equals the int32 table, for noise (the
    largest costs a 10-bit frame can produce in the 16-bit entries), extremes and a frame with partial CTUs."""
    from mipb200 import frames
    W, H = 384, 200
    fs = [frames.noise_frame(W, H, 61), frames.extreme_frame(W, H, 2), frames.natural_frame(W, H, 62)]
    with mip.Engine(W, H, filter_type=ft, kernel_idx=kidx, slots=3, emit=mip.EMIT_COSTS_COMPACT | mip.EMIT_DECISIONS) as eng:
        for i, f in enumerate(fs):
            eng.submit(f, i)
        for i, f in enumerate(fs):
            r = eng.collect()
            want = oracle.run_frame(f, ft, kidx)
            assert r.cost is None and r.cost_compact.shape == (eng.n_ctus, mip.COMPACT_BYTES_PER_CTU)
            assert np.array_equal(r.expand_costs(threads=3), want), i
            bm, bc = oracle.decisions(want)
            assert np.array_equal(r.best_mode, bm) and np.array_equal(r.best_cost, bc)
    assert mip.lib().mipb200_compact_bytes_per_ctu() == mip.COMPACT_BYTES_PER_CTU
    for kw in (dict(emit=mip.EMIT_COSTS_COMPACT | mip.EMIT_COSTS), dict(emit=mip.EMIT_COSTS_COMPACT, bit_depth=12),
               dict(emit=mip.EMIT_COSTS_COMPACT | mip.EMIT_DECISIONS, top_k=3), dict(emit=mip.EMIT_COSTS_COMPACT | mip.EMIT_SAD_SATD)):
        with pytest.raises(mip.MipError) as ei:
            mip.Engine(W, H, **kw)
        assert ei.value.code == -1


def test_device_timeline_trace(mip, tmp_path, monkeypatch):
    """MIPB200_TRACE=<file>: one line per collected frame, upload start <= kernel start <= kernel end <= results on the host."""
    from mipb200 import frames
    monkeypatch.setenv("MIPB200_TRACE", str(tmp_path / "trace"))
    f = frames.noise_frame(256, 128, 5)
    with mip.Engine(256, 128, slots=2, emit=mip.EMIT_DECISIONS) as eng:
        for i in range(4):
            eng.submit(f, i)
            if eng.in_flight() == 2:
                eng.collect()
        while eng.in_flight():
            eng.collect()
    lines = [ln for ln in open(str(tmp_path / "trace") + ".gpu0").read().splitlines() if not ln.startswith("#")]
    assert [int(ln.split()[0]) for ln in lines] == [0, 1, 2, 3]
    for ln in lines:
        t = [float(v) for v in ln.split()[1:]]
        assert t == sorted(t) and t[3] - t[0] < 1000.0
