"""Generated CU tables == the reference's own headers (parsed, never copied)."""
import os
import re

import pytest

from mipb200 import tables as T

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not mounted (it never is on the GPU box)")


def _c_array(txt, name):
    m = re.search(name + r"\s*(\[[^=]*\])+\s*=\s*\{(.*?)\};", txt, re.S)
    assert m, name
    body = re.sub(r"/\*.*?\*/", "", m.group(2), flags=re.S)
    body = re.sub(r"//[^\n]*", "", body)
    return body


def _ints(body):
    # evaluates the small arithmetic expressions the reference uses in its initialisers
    out = []
    for item in body.replace("{", " ").replace("}", " ").split(","):
        item = item.strip()
        if item:
            out.append(int(eval(item, {"__builtins__": {}})))  # noqa: S307 - digits and + * ( ) only
    return out


def test_totals():
    assert T.NUM_TYPES == 47
    assert T.CUS_PER_CTU == 5380 and T.COSTS_PER_CTU == 97840
    assert T.COST_OFFSETS[28] == 12 * 1156 and T.COST_OFFSETS[46] == 13872 + 16 * 3200
    assert sum(t.n for t in T.TYPES if t.size_id == 2) == 1156
    assert sum(t.n for t in T.TYPES if t.size_id == 1) == 3200


def test_every_cu_is_distinct_and_on_the_4x4_grid():
    seen = set()
    for t in T.TYPES:
        for cu in range(t.n):
            x, y = t.pos(cu)
            assert x % 4 == 0 and y % 4 == 0 and x + t.w <= 128 and y + t.h <= 128
            seen.add((x, y, t.w, t.h))
    assert len(seen) == 5380
    assert len({(t.w, t.h) for t in T.TYPES}) == 17


@needs_ref
def test_geometry_matches_reference_constants_h():
    txt = open(os.path.join(REF, "constants.h")).read()
    assert _ints(_c_array(txt, "ALL_widths")) == [t.w for t in T.TYPES]
    assert _ints(_c_array(txt, "ALL_heights")) == [t.h for t in T.TYPES]
    assert _ints(_c_array(txt, "ALL_cusPerCtu")) == [t.n for t in T.TYPES]
    assert _ints(_c_array(txt, "ALL_cuColumnsPerCtu")) == [t.cols for t in T.TYPES]
    assert _ints(_c_array(txt, "ALL_cuRowsPerCtu")) == [t.rows for t in T.TYPES]
    assert _ints(_c_array(txt, "ALL_numPredModes")) == [t.num_matrices for t in T.TYPES]
    assert _ints(_c_array(txt, "ALL_reducedPredSizes")) == [t.red_size for t in T.TYPES]
    assert _ints(_c_array(txt, "ALL_reducedBoundarySizes")) == [t.bdry_size for t in T.TYPES]
    assert _ints(_c_array(txt, "ALL_stridedDistortionsPerCtu")) == list(T.COST_OFFSETS)
    assert _ints(_c_array(txt, "ALL_stridedCusPerCtu")) == list(T.CU_OFFSETS)


@needs_ref
def test_positions_match_reference_all_x_pos_y_pos():
    txt = open(os.path.join(REF, "constants.h")).read()
    for name, axis in (("ALL_X_POS", 0), ("ALL_Y_POS", 1)):
        body = _c_array(txt, "const unsigned char " + name)
        rows = re.findall(r"\{([^{}]*)\}", body)
        assert len(rows) == 46
        for t, row in zip(T.TYPES[:46], rows):
            vals = [int(v) for v in row.split(",") if v.strip()]
            assert vals[: t.n] == [t.pos(cu)[axis] for cu in range(t.n)], (name, t.name)


@needs_ref
def test_type_names_match_reference_log_names():
    txt = open(os.path.join(REF, "main_aux_functions.h")).read()
    start = txt.index("translateCuSizeIdx_ALL")
    names = re.findall(r'return "(ALL_[^"]+)"', txt[start:start + 6000])
    assert names[:47] == [t.name for t in T.TYPES]


@needs_ref
def test_filter_names_and_coefficients_match_reference():
    txt = open(os.path.join(REF, "constants.h")).read()
    block = txt[txt.index("availableFilters = {"): txt.index("availableFilters_arm")]
    assert re.findall(r'"(filterFrame_[^"]+)"', block) == list(T.FILTER_NAMES)
    import ctypes, subprocess, tempfile
    src = r'''
    #include "mip_filters.h"
    int k3(int i,int dy,int dx){return mip_k3(i,dy,dx);} int k5(int i,int dy,int dx){return mip_k5(i,dy,dx);}
    '''
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vvc-mip-gpu_b200", "csrc")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "f.c"), "w").write(src)
        subprocess.run(["gcc", "-shared", "-fPIC", "-I", csrc, "-o", os.path.join(d, "f.so"), os.path.join(d, "f.c")], check=True)
        L = ctypes.CDLL(os.path.join(d, "f.so"))
        k3 = _ints(_c_array(txt, "const unsigned short convKernelLib"))
        k5 = _ints(_c_array(txt, "const unsigned short convKernelLib_5x5"))
        assert k3 == [L.k3(i, dy, dx) for i in range(5) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
        assert k5 == [L.k5(i, dy, dx) for i in range(3) for dy in range(-2, 3) for dx in range(-2, 3)]


@needs_ref
def test_matrices_match_reference_mip_matrix_cl():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_tables", os.path.join(root, "tools", "gen_tables.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    mats = g.parse_matrices(REF)
    hdr = open(os.path.join(root, "vvc-mip-gpu_b200", "csrc", "mip_matrices.h")).read()
    for name, ref_name, k_src, pad, nk in (("MIP_MAT_ID2_W", "mipMatrix16x16", 7, 1, 8), ("MIP_MAT_ID1_W", "mipMatrix8x8", 8, 0, 8), ("MIP_MAT_ID0_W", "mipMatrix4x4", 4, 0, 4)):
        body = hdr[hdr.index(name):]
        words = [int(w, 16) for w in re.findall(r"0x([0-9a-f]+)u", body[body.index("=") + 1: body.index("};")])]
        ours = [(w >> (8 * i)) & 0xFF for w in words for i in range(nk)]       # tap i = byte i (little endian)
        (nm, np_, _), vals = mats[ref_name]
        want = []
        for r in range(nm * np_):
            want += [0] * pad + vals[r * k_src:(r + 1) * k_src]
        assert ours == want, name
