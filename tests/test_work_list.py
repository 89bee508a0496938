"""The fused kernel's work list (csrc/mip_work_list.h: lane records, CU ordinals, chunk splits), checked on the CPU against the
CU tables: every (CU, mode) of a CTU is computed exactly once, at the right place, by the code of its shape, and written
to its place in the reference's cost order (main_aux_functions.h:585-630 reads the table in that order)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from mipb200 import tables as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vvc-mip-gpu_b200", "csrc")

REC_INRANGE, REC_WRITER, REC_G64, REC_GS1, REC_GA32, REC_G4x4 = (1 << 26, 1 << 27, 1 << 28, 1 << 29, 1 << 30, 1 << 31)
GS1 = [(8, 8), (16, 4), (4, 16), (32, 4), (4, 32)]
GA32 = [(8, 4), (4, 8)]
G12 = [(32, 32), (32, 16), (16, 32), (32, 8), (8, 32), (16, 16), (16, 8), (8, 16)]


@pytest.fixture(scope="module")
def dump_exe(tmp_path_factory):
    exe = tmp_path_factory.mktemp("wl") / "work_list_dump"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-I", CSRC, "-o", str(exe), os.path.join(ROOT, "tests", "work_list_dump.cpp")], check=True)
    return str(exe)


def _dump(exe, *args):
    r = subprocess.run([exe, *args], capture_output=True, timeout=120)
    return r.returncode, r.stdout


def _parse(buf):
    off = 0

    def take(fmt_dtype, n):
        nonlocal off
        a = np.frombuffer(buf, dtype=fmt_dtype, count=n, offset=off)
        off += a.nbytes
        return a

    halves = []
    for _ in range(2):
        ntasks = int(take("<u4", 1)[0])
        recs = take("<u4", ntasks * 64).reshape(ntasks, 32, 2)
        ncu = int(take("<u4", 1)[0])
        halves.append((recs, take("<u2", ncu)))
    splits = []
    for _ in range(2):
        chunks = int(take("<u4", 1)[0])
        per_half = []
        for _ in range(2):
            per_half.append((take("<i4", chunks + 1), take("<u2", chunks + 1)))
        splits.append((chunks, per_half))
    assert off == len(buf)
    return halves, splits


@pytest.fixture(scope="module")
def work(dump_exe):
    rc, out = _dump(dump_exe, "3", "4", "-", "4,3,2,1")      # the engine's defaults: 1:1:1 and 4:3:2:1
    assert rc == 0
    return _parse(out)


def _type_of_cost():
    owner = np.empty(T.COSTS_PER_CTU, dtype=np.int32)
    for t in T.TYPES:
        owner[T.COST_OFFSETS[t.idx]:T.COST_OFFSETS[t.idx] + t.n * t.modes] = t.idx
    return owner


def test_every_cu_mode_pair_exactly_once_and_in_place(work):
    halves, _ = work
    owner = _type_of_cost()
    written = np.zeros(T.COSTS_PER_CTU, dtype=np.int32)
    computed = np.zeros(T.COSTS_PER_CTU, dtype=np.int32)       # strip groups that compute each cost (4 for 64x64)
    for hf, (recs, ord2cu) in enumerate(halves):
        assert recs.shape[0] <= 1700 - 2                        # the table's last row is the end mark
        x, y = recs[..., 0].astype(np.int64), recs[..., 1].astype(np.int64)
        assert not (x & 0x80).any()                             # bit 7 of .x: the draw's "always zero" bits
        assert not (x == 0xFFFFFFFF).any()                      # the end mark is no record
        cux, cuy, mode, part = x & 0xFF, (x >> 8) & 0xFF, (x >> 16) & 0xFF, (x >> 24) & 3
        inr, wr = (x & REC_INRANGE) != 0, (x & REC_WRITER) != 0
        coff, slot, sub = y & 0x1FFFF, (y >> 17) & 0xFFF, y >> 29
        assert (wr == (inr & (part == 0))).all()
        for task in range(recs.shape[0]):
            tys = set(owner[coff[task]])
            assert len(tys) == 1                                # a warp task = one CU type
            t = T.TYPES[tys.pop()]
            g = x[task] & (REC_G64 | REC_GS1 | REC_GA32 | REC_G4x4)
            assert len(set(g)) == 1 and len(set(sub[task])) == 1
            g, sb = int(g[0]), int(sub[task][0])
            if (t.w, t.h) == (4, 4):
                assert g == REC_G4x4 and inr[task].all() and len(set(slot[task])) == 1        # the warp is one CU
            elif (t.w, t.h) in GA32:
                assert g == REC_GA32 and GA32[sb] == (t.w, t.h)
            elif (t.w, t.h) in GS1:
                assert g == REC_GS1 and GS1[sb] == (t.w, t.h)
            elif (t.w, t.h) == (64, 64):
                assert g == REC_G64
            else:
                assert g == 0 and G12[sb] == (t.w, t.h)
            if t.modes == 16:                                   # each half warp is one CU, in range as a whole
                for h in (slice(0, 16), slice(16, 32)):
                    assert len(set(slot[task][h])) == 1 and len(set(inr[task][h])) == 1
                    assert list(mode[task][h]) == list(range(16))
            assert (part[task] == 0).all() or (t.w, t.h) == (64, 64)
            for lane in range(32):
                if not inr[task, lane]:
                    continue
                c = int(coff[task, lane])
                cu, m = divmod(c - T.COST_OFFSETS[t.idx], t.modes)
                px, py = t.pos(cu)
                assert m == mode[task, lane] and px == cux[task, lane] and py == cuy[task, lane] + 64 * hf
                assert 0 <= cuy[task, lane] and cuy[task, lane] + t.h <= 64            # no CU crosses the halves
                assert int(ord2cu[slot[task, lane]]) == T.CU_OFFSETS[t.idx] + cu        # the decision lands at the CU's place
                computed[c] += 1
                written[c] += int(wr[task, lane])
    assert (written == 1).all()
    is64 = _type_of_cost() == 0
    assert (computed[is64] == 4).all() and (computed[~is64] == 1).all()
    assert sorted(np.concatenate([h[1] for h in halves]).tolist()) == list(range(T.CUS_PER_CTU))


def test_chunks_partition_the_list_and_never_split_a_cu(work):
    halves, splits = work
    for chunks, per_half in splits:
        for hf, (begin, chunk_ord) in enumerate(per_half):
            recs, ord2cu = halves[hf]
            assert begin[0] == 0 and begin[-1] == recs.shape[0] and (np.diff(begin) > 0).all()
            assert chunk_ord[0] == 0 and chunk_ord[-1] == len(ord2cu) and (np.diff(chunk_ord.astype(int)) <= 2048).all()
            y = recs[..., 1].astype(np.int64)
            inr = (recs[..., 0].astype(np.int64) & REC_INRANGE) != 0
            slot = (y >> 17) & 0xFFF
            for k in range(chunks):                              # the CUs of chunk k are exactly its ordinal range
                s = slot[begin[k]:begin[k + 1]][inr[begin[k]:begin[k + 1]]]
                assert s.min() == chunk_ord[k] and s.max() == chunk_ord[k + 1] - 1


def test_shares_shape_the_split(dump_exe):
    rc, out = _dump(dump_exe, "3", "4", "-", "4,3,2,1")
    _, splits = _parse(out)
    b = splits[1][1][0][0]                                       # lone-frame split, half 0
    assert np.diff(b)[0] < np.diff(b)[-1]                        # cheap short tasks come last: the 1-share chunk holds the most tasks
    rc, out = _dump(dump_exe, "1", "1")                          # one chunk per half: 2690 CUs > the 2048 of the decision table
    assert rc == 3
