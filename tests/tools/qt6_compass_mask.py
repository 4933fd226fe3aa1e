"""Regenerates the pixel rows of jumper easy mode's compass disc from the real Qt 6 raster engine
(the constant table in oracle/shim/qt_raster.cpp and procgen_b200/csrc/games/jumper.cuh).
Needs the Qt 6 backed oracle (oracle/_ref/libenv_ref_qt6.so): python tests/tools/qt6_compass_mask.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import qt6_support  # noqa: E402
from oracle.ref_env import REF_LIB_QT6  # noqa: E402

lib = C.CDLL(REF_LIB_QT6, handle=qt6_support.lazy_dlopen(REF_LIB_QT6))
f = np.float32
unit = f(64) / f(12)                                  # 64 px over visibility 12 (jumper.cpp:220-222)
x, y, w = float(f(8.75) * unit), float(f(.25) * unit), float(f(3) * unit)  # get_abs_rect(view_dim - compass_dim - .25, .25, 3, 3)
bg = 0xff102030
dst = np.full((64, 64), bg, np.uint32)
lib.shim_test_draw_ellipse(dst.ctypes.data_as(C.c_void_p), 64, 64, C.c_double(x), C.c_double(y), C.c_double(w), C.c_double(w),
                           168, 166, 158, 255, 1)
print("rect", repr(x), repr(y), repr(w))
rows = []
for yy in range(64):
    xs = np.nonzero(dst[yy] != bg)[0]
    if len(xs):
        assert xs.max() - xs.min() + 1 == len(xs), "row is not one span"
        rows.append((yy, int(xs.min()), int(xs.max()) + 1))
print(rows)
