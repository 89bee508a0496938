"""Unit tests of the oracle's stages against independent numpy restatements of the
arithmetic specification and hand-computed cases."""
import numpy as np
import pytest

from mipb200 import frames, tables as T


def hadamard4():
    h2 = np.array([[1, 1], [1, -1]])
    return np.kron(h2, h2)


def satd_matrix_form(d):
    """(sum |H d H^T| with DC replaced by |DC| >> 2, +1) >> 1  (kernel_aux_functions.cl:238-246)."""
    h = hadamard4()
    c = h @ np.asarray(d).reshape(4, 4) @ h.T
    a = np.abs(c)
    s = int(a.sum() - a[0, 0] + (int(a[0, 0]) >> 2))
    return (s + 1) >> 1


def test_satd_known_answers(oracle):
    assert oracle.satd4x4(np.zeros(16)) == 0
    imp = np.zeros(16); imp[0] = 1023          # impulse: every coefficient is +-1023
    assert oracle.satd4x4(imp) == (15 * 1023 + (1023 >> 2) + 1) >> 1
    dc = np.full(16, 7)                         # DC only: 16*7 = 112 -> 112 >> 2 = 28 -> (28+1)>>1
    assert oracle.satd4x4(dc) == 14
    assert oracle.satd4x4(-dc) == 14


def test_satd_random_vs_matrix_form(oracle):
    rng = np.random.default_rng(1)
    for _ in range(500):
        d = rng.integers(-1023, 1024, 16)
        assert oracle.satd4x4(d) == satd_matrix_form(d)


def test_boundaries_frame_edge_rules(oracle):
    f = frames.noise_frame(256, 256, 3)
    # interior CU
    rT, rL, redT, redL = oracle.cu_boundaries(f, 64, 32, 16, 8, 4)
    assert rT.tolist() == f[31, 64:80].tolist() and rL.tolist() == f[32:40, 63].tolist()
    assert redT.tolist() == [(int(f[31, 64 + 4 * q:68 + 4 * q].sum()) + 2) >> 2 for q in range(4)]
    assert redL.tolist() == [(int(f[32 + 2 * q:34 + 2 * q, 63].sum()) + 1) >> 1 for q in range(4)]
    # top-left corner: DC
    rT, rL, redT, redL = oracle.cu_boundaries(f, 0, 0, 8, 16, 4)
    assert set(rT.tolist()) == {512} and set(rL.tolist()) == {512} and redT.tolist() == [512] * 4
    # top edge, X > 0: F[0][X-1] replicated; left comes from the frame
    rT, rL, _, _ = oracle.cu_boundaries(f, 32, 0, 8, 8, 4)
    assert set(rT.tolist()) == {int(f[0, 31])} and rL.tolist() == f[0:8, 31].tolist()
    # left edge, Y > 0: F[Y-1][0] replicated
    rT, rL, _, _ = oracle.cu_boundaries(f, 0, 64, 8, 8, 4)
    assert set(rL.tolist()) == {int(f[63, 0])} and rT.tolist() == f[63, 0:8].tolist()
    # factor-1 identity (4x4 with b=2 has factor 2; 4x8 sizeId1: top factor 1)
    rT, rL, redT, redL = oracle.cu_boundaries(f, 8, 8, 4, 8, 4)
    assert redT.tolist() == rT.tolist()


def _ref_matrices():
    """MIP matrices parsed from the generated header: one little-endian word per row, tap i = byte i."""
    import re, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "vvc-mip-gpu_b200", "csrc", "mip_matrices.h")).read()
    out = {}
    for name, shape in (("MIP_MAT_ID2", (6, 64, 8)), ("MIP_MAT_ID1", (8, 16, 8)), ("MIP_MAT_ID0", (16, 16, 4))):
        body = hdr[hdr.index(name + "_W"):]
        words = [int(w, 16) for w in re.findall(r"0x([0-9a-f]+)u", body[body.index("=") + 1: body.index("};")])]
        out[name] = np.array([(w >> (8 * i)) & 0xFF for w in words for i in range(shape[2])]).reshape(shape)
    return out


@pytest.mark.parametrize("size_id", [0, 1, 2])
def test_reduced_prediction_vs_numpy(oracle, size_id):
    mats = _ref_matrices()[f"MIP_MAT_ID{size_id}"]
    M, r, b = mats.shape[0], (8 if size_id == 2 else 4), (2 if size_id == 0 else 4)
    rng = np.random.default_rng(size_id)
    for trial in range(40):
        hi = 1024 if trial % 3 else 2          # also near-constant inputs (clamps, first-sample offsets)
        redT, redL = rng.integers(0, hi, b) * (1023 if hi == 2 else 1), rng.integers(0, hi, b) * (1023 if hi == 2 else 1)
        for mode in range(2 * M):
            tr, mat = mode >= M, mode % M
            bd = np.concatenate([redL, redT] if tr else [redT, redL]).astype(np.int64)
            first = bd[0]
            inp = bd - first
            inp[0] = 0 if size_id == 2 else 512 - first
            v = (mats[mat][:, :2 * b] @ inp + 32 - 32 * inp.sum()) >> 6
            want = np.clip(v + first, 0, 1023).reshape(r, r)
            if tr:
                want = want.T
            assert np.array_equal(oracle.reduced_prediction(size_id, mode, redT, redL), want)


@pytest.mark.parametrize("w,h,r", [(64, 64, 8), (32, 8, 8), (8, 32, 8), (16, 16, 8), (8, 16, 8), (32, 4, 4), (4, 32, 4), (8, 8, 4), (4, 8, 4), (16, 4, 4)])
def test_upsample_vs_numpy(oracle, w, h, r):
    rng = np.random.default_rng(w * 100 + h)
    red = rng.integers(0, 1024, (r, r))
    refT, refL = rng.integers(0, 1024, w), rng.integers(0, 1024, h)
    uH, uV = w // r, h // r
    hor = np.zeros((r, w), np.int64)
    for j in range(r):
        y = j * uV + uV - 1
        for x in range(w):
            o = x % uH + 1
            after = red[j, x // uH]
            before = refL[y] if x < uH else red[j, x // uH - 1]
            hor[j, x] = ((uH - o) * before + o * after + (uH >> 1)) >> (uH.bit_length() - 1)
    want = np.zeros((h, w), np.int64)
    for y in range(h):
        o, j = y % uV + 1, y // uV
        before = refT if j == 0 else hor[j - 1]
        want[y] = ((uV - o) * before + o * hor[j] + (uV >> 1)) >> (uV.bit_length() - 1)
    assert np.array_equal(oracle.upsample(red, w, h, refT, refL), want)


def test_float_rounding_equals_int():
    """round(N/S) in fp32 (float filter kernels, intra.cl:1794,2046,2507,2819) == (N + S/2)/S for every
    denominator and numerator the filters can produce (S <= 81, N <= 1023*S)."""
    for s in range(1, 82):
        n = np.arange(0, 1023 * s + 1, dtype=np.int64)
        q = (n.astype(np.float32) / np.float32(s))
        flt = np.where(q - np.floor(q) >= 0.5, np.floor(q) + 1, np.floor(q)).astype(np.int64)   # round half away from zero
        assert np.array_equal(flt, (n + s // 2) // s), s


def test_filters_vs_numpy_2d(oracle):
    f = frames.noise_frame(256, 88, 5).astype(np.int64)
    H, W = f.shape
    k5 = lambda idx: {0: np.ones((5, 5), int), 1: np.ones((5, 5), int) + 4 * (np.arange(25).reshape(5, 5) == 12),
                      2: np.outer([1, 2, 3, 2, 1], [1, 2, 3, 2, 1])}[idx]
    k3 = lambda idx: np.array({0: [1, 1, 1, 1, 1, 1, 1, 1, 1], 1: [1, 2, 1, 2, 3, 2, 1, 2, 1], 2: [1, 2, 1, 2, 12, 2, 1, 2, 1],
                               3: [1, 1, 1, 1, 8, 1, 1, 1, 1], 4: [1, 2, 1, 2, 4, 2, 1, 2, 1]}[idx]).reshape(3, 3)
    for ft, R, ks, n in ((3, 1, k3, 5), (7, 2, k5, 3)):
        for idx in range(n):
            K = ks(idx)
            pad = np.pad(f, R)
            ones = np.pad(np.ones_like(f), R)
            num = sum(K[dy, dx] * pad[dy:dy + H, dx:dx + W] for dy in range(2 * R + 1) for dx in range(2 * R + 1))
            den = sum(K[dy, dx] * ones[dy:dy + H, dx:dx + W] for dy in range(2 * R + 1) for dx in range(2 * R + 1))
            want = (num + den // 2) // den
            assert np.array_equal(oracle.filter_frame(f.astype(np.uint16), ft, idx), want), (ft, idx)
            assert np.array_equal(oracle.filter_frame(f.astype(np.uint16), ft + 1, idx), want), (ft + 1, idx)


def test_filters_1d_position_classes(oracle):
    """1-D types: separable numerator with zero padding, denominator by position class (App. A.6)."""
    f = frames.noise_frame(128, 32, 9).astype(np.int64)
    H, W = f.shape
    # 3x3, idx 1: k = (1,2,1): interior 4+8+4 = 16, edge 2+6+4 = 12, corner 1+4+4 = 9
    k = np.array([1, 2, 1])
    pad = np.pad(f, 1)
    num = sum(k[dy] * k[dx] * pad[dy:dy + H, dx:dx + W] for dy in range(3) for dx in range(3))
    den = np.full((H, W), 16)
    den[0, :] = den[-1, :] = 12
    den[:, 0] = den[:, -1] = 12
    den[0, 0] = den[0, -1] = den[-1, 0] = den[-1, -1] = 9
    assert np.array_equal(oracle.filter_frame(f.astype(np.uint16), 1, 1), (num + den // 2) // den)
    # 5x5, idx 1: k = (1,1,1,1,1) but the denominators come from the 2-D table with centre 5
    k = np.ones(5, int)
    K = np.ones((5, 5), int); K[2, 2] = 5
    pad = np.pad(f, 2)
    num = sum(pad[dy:dy + H, dx:dx + W] for dy in range(5) for dx in range(5))
    got = oracle.filter_frame(f.astype(np.uint16), 5, 1)
    y, x = 10, 10
    assert got[y, x] == (num[y, x] + 29 // 2) // 29                      # interior: sum K = 29
    assert got[0, 0] == (num[0, 0] + K[2:, 2:].sum() // 2) // K[2:, 2:].sum()   # outer corner
    assert got[1, 1] == (num[1, 1] + K[1:, 1:].sum() // 2) // K[1:, 1:].sum()   # inner corner
    assert got[0, 1] == (num[0, 1] + K[1:, 2:].sum() // 2) // K[1:, 2:].sum()   # interface
    assert got[0, 10] == (num[0, 10] + K[:, 2:].sum() // 2) // K[:, 2:].sum()   # outer edge
    assert got[10, 1] == (num[10, 1] + K[:, 1:].sum() // 2) // K[:, 1:].sum()   # inner edge
    assert np.array_equal(got, oracle.filter_frame(f.astype(np.uint16), 6, 1))


def test_skipped_cus_and_decisions(oracle):
    f = frames.noise_frame(128, 72, 4)          # one partial CTU: 72 valid rows
    cost = oracle.run_frame(f)
    mask = T.in_frame_mask(128, 72)
    bm, bc = oracle.decisions(cost)
    for t in T.TYPES:
        c = cost[0, T.COST_OFFSETS[t.idx]:T.COST_OFFSETS[t.idx + 1]].reshape(t.n, t.modes)
        m = mask[0, T.CU_OFFSETS[t.idx]:T.CU_OFFSETS[t.idx + 1]]
        assert ((c == -1).all(axis=1) == ~m).all() and (c[m] >= 0).all()
        assert np.array_equal(bm[0, T.CU_OFFSETS[t.idx]:T.CU_OFFSETS[t.idx + 1]][m], c[m].argmin(axis=1))
    assert (bm[0][~mask[0]] == 0xFF).all()


def test_thread_count_does_not_change_results(oracle):
    f = frames.natural_frame(256, 128, 8)
    assert np.array_equal(oracle.run_frame(f, 5, 2, threads=1), oracle.run_frame(f, 5, 2, threads=4))


def test_bit_depth_generalisation(oracle):
    """bit_depth only moves the three constants the reference hard-wires (intra.cl:61, 446, 482): 10 is the default; a
    flat frame predicts itself at any depth; at 8 / 12 bits the first CU's default boundary sample is 128 / 2048."""
    from mipb200 import frames
    f = frames.noise_frame(128, 128, 11, bits=8)
    assert np.array_equal(oracle.run_frame(f), oracle.run_frame(f, bit_depth=10))
    with pytest.raises(ValueError):
        oracle.run_frame(f, bit_depth=9)
    for bits in (8, 12):
        flat = np.full((128, 128), 1 << (bits - 1), np.uint16)              # equals the default sample: every boundary is flat
        assert not oracle.run_frame(flat, bit_depth=bits).any()
        assert oracle.run_frame(flat, bit_depth=10).any() == (bits != 10)
    hi = np.full((128, 128), 4095, np.uint16)
    c12 = oracle.run_frame(hi, bit_depth=12)
    c10 = oracle.run_frame(hi, bit_depth=10)
    assert c12[0, 0] < c10[0, 0]                                              # 10-bit clamp at 1023 cannot reach 4095
    assert not np.array_equal(oracle.run_frame(f, bit_depth=8), oracle.run_frame(f, bit_depth=10))
