"""The C-ABI library loads and exports every symbol include/procgen_b200.h declares (no compute)."""
import ctypes as C
import os
import re

from procgen_b200 import libenv as L


def _declared_symbols():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "procgen_b200.h")).read()
    return re.findall(r"LIBENV_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)


def test_header_symbols_match_binding():
    assert sorted(_declared_symbols()) == sorted(L.EXPORTS)


def test_product_library_exports_everything(product_lib):
    lib = C.CDLL(product_lib)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} not exported"
    lib.libenv_version.restype = C.c_int
    assert lib.libenv_version() == 1
    lib.pgb200_is_device_build.restype = C.c_int
    assert lib.pgb200_is_device_build() == 1


def test_struct_layouts_match_oracle_shim():
    """include/procgen_b200.h and oracle/shim/libenv.h restate the same gym3 header."""
    from oracle import ref_env as R

    for a, b in [(L.TensorType, R.TensorType), (L.Option, R.Option), (L.Options, R.Options), (L.Buffers, R.Buffers)]:
        assert C.sizeof(a) == C.sizeof(b)
        assert [(n, getattr(a, n).offset) for n, _ in a._fields_] == [(n, getattr(b, n).offset) for n, _ in b._fields_]
    assert C.sizeof(L.TensorType) == 128 + 4 + 4 + 64 + 4 + 4 + 4
    assert C.sizeof(L.Option) == 128 + 4 + 4 + 8


def test_package_refuses_non_cuda_build(hostsim_lib):
    """The CPU debug harness must never be usable as the product."""
    import pytest

    from procgen_b200 import ProcgenGym3Env

    lib = L.bind(C.CDLL(hostsim_lib))
    assert lib.pgb200_is_device_build() == 0
    saved = L.LIB_PATH
    try:
        L.LIB_PATH = hostsim_lib
        L._lib = None
        with pytest.raises(RuntimeError):
            L.load()
    finally:
        L.LIB_PATH = saved
        L._lib = None


def test_compiler_is_a_validated_release():
    """The warp-lockstep execution of the per-env scalar logic (pg_engine.cuh) is proven per compiler by the
    bit-exact GPU suite; build.py lists the nvcc releases that suite ran against and warns for any other."""
    from procgen_b200 import build

    ok = build.check_toolchain()
    assert ok is not False, "nvcc release not in build.VALIDATED_NVCC: run pytest -m gpu, then add it"
