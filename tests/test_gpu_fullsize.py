"""BASELINE.json sizes on the GPU: full comparison with the oracle where it finishes in seconds (1080p, 2160p) and
size-independent properties at 4320p (decisions == argmin of the table, min(2*SAD, SATD) identity, skipped rows,
crop invariance of interior CTUs, shard-count independence)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _eq(a, b, what):
    if not np.array_equal(a, b):
        raise AssertionError(f"{what}: {int((a != b).sum())} of {a.size} differ")


@pytest.mark.parametrize("ft,kidx", [(0, 0), (8, 2), (1, 1)])
def test_1080p_equals_oracle(mip, oracle, ft, kidx):
    from mipb200 import frames
    f = frames.natural_frame(1920, 1080, 31 + ft)
    with mip.Engine(1920, 1080, filter_type=ft, kernel_idx=kidx, slots=1, emit=mip.EMIT_COSTS | mip.EMIT_DECISIONS) as eng:
        r = eng.run(f)
        cost, bm, bc = r.cost.copy(), r.best_mode.copy(), r.best_cost.copy()
    want = oracle.run_frame(f, ft, kidx)
    _eq(cost, want, "1080p cost")
    wbm, wbc = oracle.decisions(want)
    _eq(bm, wbm, "1080p best_mode")
    _eq(bc, wbc, "1080p best_cost")
    assert int((cost != -1).sum()) == 12_359_520       # in-frame (CU, mode) pairs of a 1080p frame


def test_2160p_10bit_equals_oracle(mip, oracle):
    from mipb200 import frames
    f = frames.natural_frame(3840, 2160, 77)
    with mip.Engine(3840, 2160, slots=1, emit=mip.EMIT_COSTS) as eng:
        cost = eng.run(f).cost.copy()
    _eq(cost, oracle.run_frame(f), "2160p cost")
    assert int((cost != -1).sum()) == 49_494_960


def test_4320p_properties(mip, oracle):
    from mipb200 import frames, tables as T
    W, H = 7680, 4320
    f = frames.natural_frame(W, H, 5)
    with mip.Engine(W, H, filter_type=7, kernel_idx=2, slots=1, emit=mip.EMIT_COSTS | mip.EMIT_SAD_SATD | mip.EMIT_DECISIONS) as eng:
        r = eng.run(f)
        cost, sad, satd, bm, bc = r.cost.copy(), r.sad.copy(), r.satd.copy(), r.best_mode.copy(), r.best_cost.copy()
    assert cost.shape == (2040, 97840)
    assert int((cost != -1).sum()) == 198_125_280
    ok = cost != -1
    _eq(cost[ok], np.minimum(2 * sad[ok].astype(np.int64), satd[ok]).astype(np.int32), "cost == min(2*SAD, SATD)")
    # decisions == argmin of the table, type by type
    for t in T.TYPES:
        c = cost[:, T.COST_OFFSETS[t.idx]:T.COST_OFFSETS[t.idx + 1]].reshape(2040, t.n, t.modes)
        live = c[:, :, 0] != -1
        _eq(bm[:, T.CU_OFFSETS[t.idx]:T.CU_OFFSETS[t.idx + 1]][live], c.argmin(axis=2).astype(np.uint8)[live], f"best_mode {t.name}")
        _eq(bc[:, T.CU_OFFSETS[t.idx]:T.CU_OFFSETS[t.idx + 1]][live], c.min(axis=2)[live], f"best_cost {t.name}")
    # last CTU row: 96 valid rows -> exactly the CUs with y + h <= 96 are live
    mask = T.in_frame_mask(W, H)
    assert np.array_equal((bm != 0xFF), mask)
    # crop invariance: an interior CTU only sees its own samples, the row above, the column left and the filter halo
    for (cx, cy) in ((7, 5), (31, 20), (58, 32)):
        x0, y0 = 128 * cx - 128, 128 * cy - 128
        crop = np.ascontiguousarray(f[y0:y0 + 384, x0:x0 + 384])
        want = oracle.run_frame(crop, 7, 2)[4]          # centre CTU of the 3x3 crop
        _eq(cost[cy * 60 + cx], want, f"CTU ({cx},{cy}) vs oracle on its 384x384 neighbourhood")


def test_shard_count_does_not_change_results(mip):
    """Two engines on the same device standing in for two GPUs (frames poc % 2) == one engine."""
    from mipb200 import frames, shard
    fs = [frames.noise_frame(256, 184, 300 + i) for i in range(6)]
    one = {}
    with mip.Engine(256, 184, filter_type=3, kernel_idx=1, slots=3, emit=mip.EMIT_COSTS) as eng:
        shard.run_pipelined(eng, fs, list(range(6)), lambda poc, r: one.__setitem__(poc, r.cost.copy()))
    per_rank = []
    for g in range(2):
        d = {}
        with mip.Engine(256, 184, filter_type=3, kernel_idx=1, slots=2, emit=mip.EMIT_COSTS) as eng:
            shard.run_pipelined(eng, fs, shard.frames_for_rank(6, g, 2), lambda poc, r: d.__setitem__(poc, r.cost.copy()))
        per_rank.append(d)
    merged = shard.merge_in_poc_order(per_rank, 6)
    for poc in range(6):
        _eq(merged[poc], one[poc], f"poc {poc}")


def test_filter_fusion_exactness_sweep_1080p(mip, oracle):
    """BASELINE config 3: every filterFrame_* type x every KernelIdx (32 combinations) over the 16 synthetic 1080p frames
    (4 content classes x 4 seeds: noise, natural, extremes, impulses).  Each frame is evaluated under two combinations,
    every combination once; full 1080p cost tables compared with the oracle bit for bit."""
    from mipb200 import frames, tables as T
    fs = frames.sweep_frames(1920, 1080)
    combos = [(ft, k) for ft in range(1, 9) for k in range(T.num_kernel_idx(ft))]
    assert len(fs) == 16 and len(combos) == 32
    for i, (ft, k) in enumerate(combos):
        f = fs[i % 16]
        with mip.Engine(1920, 1080, filter_type=ft, kernel_idx=k, slots=1, emit=mip.EMIT_COSTS) as eng:
            got = eng.run(f).cost.copy()
        _eq(got, oracle.run_frame(f, ft, k), f"frame {i % 16} filter_type={ft} kernel_idx={k}")


@pytest.mark.parametrize("bits,ft,kidx", [(12, 7, 2), (8, 2, 3)])
def test_1080p_other_bit_depths_with_shortlists(mip, oracle, bits, ft, kidx):
    """The 8- and 12-bit pipelines and the top-5 shortlist at the BASELINE size: cost, SAD, SATD, decisions and shortlist
    against the oracle (noise uses the whole sample range, so the clamps at 0 and (1 << bits) - 1 are exercised)."""
    from mipb200 import frames
    f = frames.noise_frame(1920, 1080, 900 + bits, bits=bits)
    with mip.Engine(1920, 1080, filter_type=ft, kernel_idx=kidx, slots=1, bit_depth=bits, top_k=5,
                    emit=mip.EMIT_COSTS | mip.EMIT_SAD_SATD | mip.EMIT_DECISIONS) as eng:
        r = eng.run(f)
        got = [a.copy() for a in (r.cost, r.sad, r.satd, r.best_mode, r.best_cost, r.topk_mode, r.topk_cost)]
    cost, sad, satd = oracle.run_frame(f, ft, kidx, want_sad_satd=True, bit_depth=bits)
    _eq(got[0], cost, f"{bits}-bit 1080p cost")
    _eq(got[1], sad, f"{bits}-bit 1080p sad")
    _eq(got[2], satd, f"{bits}-bit 1080p satd")
    bm, bc = oracle.decisions(cost)
    _eq(got[3], bm, "best_mode")
    _eq(got[4], bc, "best_cost")
    tm, tc = oracle.topk(cost, 5)
    _eq(got[5], tm, "topk_mode")
    _eq(got[6], tc, "topk_cost")
    ok = cost != -1
    assert np.array_equal(cost[ok], np.minimum(2 * sad[ok], satd[ok]))          # intra.cl:1166
