"""TEST INFRASTRUCTURE — locating and linking the Qt 6.6.3 libraries that ship (header-less) with
Nsight Compute, used only to pin the raster restatement (oracle/shim/qt6_backend.cpp)."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
STUBS = os.path.join(HERE, "_ref", "qt6_stubs")
# DT_NEEDED entries of libQt6Gui/Core that are absent in this image; none is used by QImage/QPainter
# raster drawing, so empty libraries with the right sonames satisfy the loader.
MISSING = ["libEGL.so.1", "libGLX.so.0", "libOpenGL.so.0", "libX11.so.6", "libdbus-1.so.3", "libfontconfig.so.1",
           "libglib-2.0.so.0", "libgthread-2.0.so.0", "libxkbcommon.so.0"]


def qt_dir():
    c = sorted(glob.glob("/opt/nvidia/nsight-compute/*/host/linux-desktop-glibc_2_11_3-x64/libQt6Gui.so.6"))
    return os.path.dirname(c[-1]) if c else None


def available():
    return qt_dir() is not None


def make_stubs():
    os.makedirs(STUBS, exist_ok=True)
    empty = os.path.join(STUBS, "empty.c")
    open(empty, "w").write("")
    for name in MISSING:
        out = os.path.join(STUBS, name)
        if not os.path.exists(out):
            subprocess.check_call(["gcc", "-shared", "-fPIC", f"-Wl,-soname,{name}", empty, "-o", out])
    return STUBS


def link_flags():
    q = qt_dir()
    if q is None:
        raise RuntimeError("Qt6 libraries (Nsight Compute) not found")
    stubs = make_stubs()
    # DT_RPATH (not RUNPATH) so the transitive DT_NEEDED of the Qt libs resolve through the stubs too
    return [f"-L{q}", "-l:libQt6Gui.so.6", "-l:libQt6Core.so.6", "-Wl,--allow-shlib-undefined", "-Wl,--disable-new-dtags",
            f"-Wl,-rpath,{q}:{stubs}"]


def lazy_dlopen(path):
    """dlopen(path, RTLD_LAZY | RTLD_GLOBAL). ctypes always adds RTLD_NOW, which would force every
    import of Qt's unused GUI dependencies (dbus, glib, X11, EGL ...) to resolve; with lazy binding
    the empty stub libraries are enough, exactly as for a normally linked executable."""
    import ctypes as C

    helper = os.path.join(STUBS, "liblazy_dlopen.so")
    if not os.path.exists(helper):
        os.makedirs(STUBS, exist_ok=True)
        src = os.path.join(STUBS, "lazy_dlopen.c")
        open(src, "w").write('#include <dlfcn.h>\nvoid *pg_lazy_dlopen(const char *p) { return dlopen(p, RTLD_LAZY | RTLD_GLOBAL); }\n'
                             'const char *pg_dlerror(void) { return dlerror(); }\n')
        subprocess.check_call(["gcc", "-shared", "-fPIC", src, "-o", helper, "-ldl"])
    h = C.CDLL(helper)
    h.pg_lazy_dlopen.restype = C.c_void_p
    h.pg_lazy_dlopen.argtypes = [C.c_char_p]
    h.pg_dlerror.restype = C.c_char_p
    for name in MISSING:
        if not h.pg_lazy_dlopen(os.path.join(make_stubs(), name).encode()):
            raise OSError(h.pg_dlerror().decode())
    handle = h.pg_lazy_dlopen(path.encode())
    if not handle:
        raise OSError(h.pg_dlerror().decode())
    return handle
