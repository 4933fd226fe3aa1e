"""TEST INFRASTRUCTURE — builds oracle/_ref/libenv_ref.so.

Compiles the reference's own C++ sources IN PLACE (unmodified, read from /root/reference, nothing
copied) against the stand-in headers in oracle/shim/ (Qt subset + libenv.h) and links them with the
CPU raster restatement oracle/shim/qt_raster.cpp.  Flags mirror procgen/CMakeLists.txt:30-36
(-O2 -g -fno-omit-frame-pointer, C++17) plus -ffp-contract=off so the result has the float
semantics of the published wheels, which are built -march=ivybridge, i.e. without FMA
(CMakeLists.txt:30).  The reference's own build system (cmake + Qt5 + gym3) is not used.

Output goes only into oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).
Usage: python oracle/build_ref.py [--qt6]   (--qt6 additionally builds libenv_ref_qt6.so whose
QPainter forwards to the real Qt 6.6.3 raster engine bundled with Nsight Compute).
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/procgen/src"
OUT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "shim")
CXXFLAGS = ["-std=c++17", "-O2", "-g", "-fno-omit-frame-pointer", "-ffp-contract=off", "-fPIC",
            "-w", "-I" + SHIM, "-I" + REF_SRC]


def _compile(src, obj):
    if os.path.exists(obj) and os.path.getmtime(obj) > max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(SHIM, "qt_shim.h")),
            os.path.getmtime(os.path.join(SHIM, "libenv.h")), os.path.getmtime(os.path.join(SHIM, "qt_raster.cpp"))):
        return
    subprocess.check_call(["g++", *CXXFLAGS, "-c", src, "-o", obj])


def reference_available():
    return os.path.isdir(REF_SRC)


def build(qt6=False, verbose=False):
    if not reference_available():
        raise RuntimeError("reference sources not present; oracle/_ref must be prebuilt")
    objdir = os.path.join(OUT, "obj")
    os.makedirs(objdir, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(REF_SRC, "*.cpp")) + glob.glob(os.path.join(REF_SRC, "games", "*.cpp")))
    jobs = [(s, os.path.join(objdir, "ref_" + os.path.basename(s).replace(".cpp", ".o"))) for s in srcs]
    backends = [("qt_raster.cpp", "libenv_ref.so", [])]
    if qt6:
        from oracle import qt6_support  # noqa
        backends.append(("qt6_backend.cpp", "libenv_ref_qt6.so", qt6_support.link_flags()))
    hook_obj = os.path.join(objdir, "shim_test_hook.o")
    jobs.append((os.path.join(SHIM, "shim_test_hook.cpp"), hook_obj))
    shim_objs = {}
    for bsrc, _, _ in backends:
        o = os.path.join(objdir, "shim_" + bsrc.replace(".cpp", ".o"))
        jobs.append((os.path.join(SHIM, bsrc), o))
        shim_objs[bsrc] = o
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(lambda j: _compile(*j), jobs))
    ref_objs = [o for s, o in jobs if os.path.basename(o).startswith("ref_")]
    outs = []
    for bsrc, libname, extra in backends:
        lib = os.path.join(OUT, libname)
        subprocess.check_call(["g++", "-shared", "-o", lib, *ref_objs, shim_objs[bsrc], hook_obj, "-lz", "-lpthread", *extra])
        outs.append(lib)
        if verbose:
            print("built", lib)
    return outs


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    build(qt6="--qt6" in sys.argv, verbose=True)
