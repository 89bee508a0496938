"""CU geometry of the MIP pre-analysis: the 47 CU "types" of a 128x128 CTU.

Every type is a raster grid of equally shaped CUs: CU index ``k`` sits at
``(xs[k % cols], ys[k // cols])`` relative to the CTU origin.  The grids are
*generated* here from (start, step, count) rules; they reproduce the reference's
``ALL_X_POS / ALL_Y_POS / ALL_widths / ALL_heights / ALL_cusPerCtu`` tables
(reference ``constants.h:799-973, 1116-1352``) and its enumeration order, which
is also the row order of the cost log (``main_aux_functions.h:735-798``) and the
layout of the ``minSadHad`` buffer (``constants.h:1558-1631``,
``intra.cl:1144-1148``).  ``tests/test_tables.py`` checks the equality against a
parse of the reference headers when ``/root/reference`` is mounted.

sizeId follows VVC MIP: 0 for 4x4, 1 for blocks with a side of 4 or 8x8, else 2.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

CTU = 128


def _lin(start: int, step: int, count: int) -> Tuple[int, ...]:
    return tuple(start + step * i for i in range(count))


def _alt8() -> Tuple[int, ...]:
    # 0,24,32,56,64,88,96,120 : the two 8-sample slots of every 32-sample period
    # that are not covered by the G1/G3/G5 groups (reference constants.h:1277,1339)
    return tuple(32 * (i // 2) + 24 * (i % 2) for i in range(8))


@dataclass(frozen=True)
class CuType:
    idx: int
    name: str
    w: int
    h: int
    xs: Tuple[int, ...]
    ys: Tuple[int, ...]

    @property
    def cols(self) -> int:
        return len(self.xs)

    @property
    def rows(self) -> int:
        return len(self.ys)

    @property
    def n(self) -> int:
        return self.cols * self.rows

    @property
    def size_id(self) -> int:
        if self.w == 4 and self.h == 4:
            return 0
        if self.w == 4 or self.h == 4 or (self.w == 8 and self.h == 8):
            return 1
        return 2

    @property
    def num_matrices(self) -> int:
        return (16, 8, 6)[self.size_id]

    @property
    def modes(self) -> int:
        """Modes evaluated per CU: every matrix, normal and transposed."""
        return 2 * self.num_matrices

    @property
    def red_size(self) -> int:
        """Side of the reduced prediction (8 for sizeId 2, else 4)."""
        return 8 if self.size_id == 2 else 4

    @property
    def bdry_size(self) -> int:
        """Reduced boundary samples per side (2 for sizeId 0, else 4)."""
        return 2 if self.size_id == 0 else 4

    def pos(self, cu: int) -> Tuple[int, int]:
        return self.xs[cu % self.cols], self.ys[cu // self.cols]


def _grid(w: int, h: int, xs, ys):
    return (w, h, tuple(xs), tuple(ys))


def _aligned(w: int, h: int):
    return _grid(w, h, _lin(0, w, CTU // w), _lin(0, h, CTU // h))


_SPEC: List[Tuple[str, Tuple]] = [
    # ---- sizeId 2, aligned (ids 0..8)
    ("ALL_AL_64x64", _aligned(64, 64)),
    ("ALL_AL_32x32", _aligned(32, 32)),
    ("ALL_AL_32x16", _aligned(32, 16)),
    ("ALL_AL_16x32", _aligned(16, 32)),
    ("ALL_AL_32x8", _aligned(32, 8)),
    ("ALL_AL_8x32", _aligned(8, 32)),
    ("ALL_AL_16x16", _aligned(16, 16)),
    ("ALL_AL_16x8", _aligned(16, 8)),
    ("ALL_AL_8x16", _aligned(8, 16)),
    # ---- sizeId 2, half-aligned / unaligned (ids 9..27)
    ("ALL_NA_32x16", _grid(32, 16, _lin(0, 32, 4), _lin(8, 32, 4))),
    ("ALL_NA_16x32", _grid(16, 32, _lin(8, 32, 4), _lin(0, 32, 4))),
    ("ALL_NA_32x8_G1", _grid(32, 8, _lin(0, 32, 4), _lin(4, 16, 8))),
    ("ALL_NA_32x8_G2", _grid(32, 8, _lin(0, 32, 4), _lin(12, 32, 4))),
    ("ALL_NA_8x32_G1", _grid(8, 32, _lin(4, 16, 8), _lin(0, 32, 4))),
    ("ALL_NA_8x32_G2", _grid(8, 32, _lin(12, 32, 4), _lin(0, 32, 4))),
    ("ALL_NA_16x16_G1", _grid(16, 16, _lin(8, 32, 4), _lin(0, 16, 8))),
    ("ALL_NA_16x16_G2", _grid(16, 16, _lin(0, 16, 8), _lin(8, 32, 4))),
    ("ALL_NA_16x16_G3", _grid(16, 16, _lin(8, 32, 4), _lin(8, 32, 4))),
    ("ALL_NA_16x8_G1", _grid(16, 8, _lin(8, 32, 4), _lin(0, 8, 16))),
    ("ALL_NA_16x8_G2", _grid(16, 8, _lin(0, 16, 8), _lin(4, 16, 8))),
    ("ALL_NA_16x8_G3", _grid(16, 8, _lin(0, 16, 8), _lin(12, 32, 4))),
    ("ALL_NA_16x8_G4", _grid(16, 8, _lin(8, 32, 4), _lin(4, 16, 8))),
    ("ALL_NA_16x8_G5", _grid(16, 8, _lin(8, 32, 4), _lin(12, 32, 4))),
    ("ALL_NA_8x16_G1", _grid(8, 16, _lin(4, 16, 8), _lin(0, 16, 8))),
    ("ALL_NA_8x16_G2", _grid(8, 16, _lin(0, 8, 16), _lin(8, 32, 4))),
    ("ALL_NA_8x16_G3", _grid(8, 16, _lin(12, 32, 4), _lin(0, 16, 8))),
    ("ALL_NA_8x16_G4", _grid(8, 16, _lin(12, 32, 4), _lin(8, 32, 4))),
    ("ALL_NA_8x16_G5", _grid(8, 16, _lin(4, 16, 8), _lin(8, 32, 4))),
    # ---- sizeId 1, aligned (ids 28..36)
    ("ALL_AL_32x4", _aligned(32, 4)),
    ("ALL_AL_4x32", _aligned(4, 32)),
    ("ALL_AL_16x4", _aligned(16, 4)),
    ("ALL_AL_4x16", _aligned(4, 16)),
    ("ALL_AL_8x8", _aligned(8, 8)),
    ("ALL_AL_8x4_1half", _grid(8, 4, _lin(0, 8, 16), _lin(0, 4, 16))),
    ("ALL_AL_8x4_2half", _grid(8, 4, _lin(0, 8, 16), _lin(64, 4, 16))),
    ("ALL_AL_4x8_1half", _grid(4, 8, _lin(0, 4, 32), _lin(0, 8, 8))),
    ("ALL_AL_4x8_2half", _grid(4, 8, _lin(0, 4, 32), _lin(64, 8, 8))),
    # ---- sizeId 1, half-aligned / unaligned (ids 37..45)
    ("ALL_NA_16x4_G123", _grid(16, 4, _lin(8, 32, 4), _lin(0, 4, 32))),
    ("ALL_NA_4x16_G123", _grid(4, 16, _lin(0, 4, 32), _lin(8, 32, 4))),
    ("ALL_NA_8x8_G1", _grid(8, 8, _lin(4, 16, 8), _lin(0, 8, 16))),
    ("ALL_NA_8x8_G2", _grid(8, 8, _lin(12, 32, 4), _alt8())),
    ("ALL_NA_8x8_G3", _grid(8, 8, _lin(0, 8, 16), _lin(4, 16, 8))),
    ("ALL_NA_8x8_G4", _grid(8, 8, _alt8(), _lin(12, 32, 4))),
    ("ALL_NA_8x8_G5", _grid(8, 8, _lin(4, 16, 8), _lin(4, 16, 8))),
    ("ALL_NA_8x4_G1", _grid(8, 4, _lin(4, 16, 8), _lin(0, 4, 32))),
    ("ALL_NA_4x8_G1", _grid(4, 8, _lin(0, 4, 32), _lin(4, 16, 8))),
    # ---- sizeId 0 (id 46)
    ("ALL_AL_4x4", _aligned(4, 4)),
]

TYPES: Tuple[CuType, ...] = tuple(
    CuType(i, name, g[0], g[1], g[2], g[3]) for i, (name, g) in enumerate(_SPEC)
)
NUM_TYPES = len(TYPES)  # 47

# offset of each type inside a CTU's cost vector; COST_OFFSETS[47] = 97840
COST_OFFSETS: Tuple[int, ...] = tuple(
    sum(t.n * t.modes for t in TYPES[:i]) for i in range(NUM_TYPES + 1)
)
# offset of each type inside a CTU's per-CU vector; CU_OFFSETS[47] = 5380
CU_OFFSETS: Tuple[int, ...] = tuple(sum(t.n for t in TYPES[:i]) for i in range(NUM_TYPES + 1))

COSTS_PER_CTU = COST_OFFSETS[-1]
CUS_PER_CTU = CU_OFFSETS[-1]

FILTER_NAMES: Tuple[str, ...] = (
    # filter_type 1..8 = this order (reference constants.h:25-34); 0 = original samples
    "filterFrame_1d_int",
    "filterFrame_1d_float",
    "filterFrame_2d_int_quarterCtu",
    "filterFrame_2d_float_quarterCtu",
    "filterFrame_1d_int_5x5",
    "filterFrame_1d_float_5x5",
    "filterFrame_2d_int_5x5_quarterCtu",
    "filterFrame_2d_float_5x5_quarterCtu",
)


def filter_type_of(name: str) -> int:
    """CLI filter name -> engine filter_type (1..8); '' / None -> 0."""
    if not name:
        return 0
    return FILTER_NAMES.index(name) + 1


def filter_is_5x5(filter_type: int) -> bool:
    return filter_type >= 5


def filter_is_2d(filter_type: int) -> bool:
    return filter_type in (3, 4, 7, 8)


def num_kernel_idx(filter_type: int) -> int:
    """Valid --KernelIdx range of a filter type (5 tables for 3x3, 3 for 5x5)."""
    return 3 if filter_is_5x5(filter_type) else 5


def ctu_grid(width: int, height: int) -> Tuple[int, int]:
    return (width + CTU - 1) // CTU, (height + CTU - 1) // CTU


def num_ctus(width: int, height: int) -> int:
    c, r = ctu_grid(width, height)
    return c * r


def in_frame_mask(width: int, height: int):
    """bool[nCTU][5380]: CU fully inside the frame (bottom edge: reference intra.cl:96,232,717; right edge: ours)."""
    import numpy as np

    cols, rows = ctu_grid(width, height)
    m = np.zeros((cols * rows, CUS_PER_CTU), dtype=bool)
    for ctu in range(cols * rows):
        x0, y0 = CTU * (ctu % cols), CTU * (ctu // cols)
        for t in TYPES:
            for cu in range(t.n):
                x, y = t.pos(cu)
                m[ctu, CU_OFFSETS[t.idx] + cu] = (y0 + y + t.h) <= height and (x0 + x + t.w) <= width
    return m
