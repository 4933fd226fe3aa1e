"""Regenerates the pixel rows of jumper's compass disc on a non-integer rect from the real Qt 6 raster
engine (the constant tables in oracle/shim/qt_raster.cpp and procgen_b200/csrc/pg_raster.cuh).
Three such rects exist under the 64x64 contract: easy mode with the agent-centred view (visibility 12,
compass_dim 3), and the whole-world views of center_agent = false (easy: 20 cells, hard: 40 cells;
basic-abstract-game.cpp:819-838). Hard mode's centred view lands on an integer rect (midpoint algorithm)
and memory mode draws no compass. Needs the Qt 6 backed oracle (oracle/_ref/libenv_ref_qt6.so):
python tests/tools/qt6_compass_mask.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import qt6_support  # noqa: E402
from oracle.ref_env import REF_LIB_QT6  # noqa: E402

lib = C.CDLL(REF_LIB_QT6, handle=qt6_support.lazy_dlopen(REF_LIB_QT6))
f = np.float32


def rect_of(visibility, compass_dim):
    """prepare_for_drawing + get_abs_rect(view_dim - compass_dim - .25, .25, compass_dim, compass_dim) in the
    reference's float arithmetic (basic-abstract-game.cpp:803-805, 832-835; jumper.cpp:138)."""
    raw_unit = f(64) / f(visibility)
    unit = f(raw_unit * f(64.0 / 64.0))
    view_dim = f(64.0 / float(raw_unit))
    x = f(float(f(view_dim - f(compass_dim))) - .25)
    return float(f(x * unit)), float(f(f(.25) * unit)), float(f(f(compass_dim) * unit))


def rows_of(x, y, w):
    bg = 0xff102030
    dst = np.full((64, 64), bg, np.uint32)
    lib.shim_test_draw_ellipse(dst.ctypes.data_as(C.c_void_p), 64, 64, C.c_double(x), C.c_double(y), C.c_double(w), C.c_double(w),
                               168, 166, 158, 255, 1)
    rows = []
    for yy in range(64):
        xs = np.nonzero(dst[yy] != bg)[0]
        if len(xs):
            assert xs.max() - xs.min() + 1 == len(xs), "row is not one span"
            rows.append((yy, int(xs.min()), int(xs.max()) + 1))
    assert [r[0] for r in rows] == list(range(rows[0][0], rows[0][0] + len(rows)))
    return rows


for name, vis, dim in (("easy, agent-centred", 12, 3), ("easy, whole world", 20, 3), ("hard, whole world", 40, 2)):
    x, y, w = rect_of(vis, dim)
    rows = rows_of(x, y, w)
    print(f"// {name}: rect {x!r} {y!r} {w!r}, first row {rows[0][0]}")
    print("{" + ", ".join("{%d, %d}" % (a, b) for _, a, b in rows) + "}")
