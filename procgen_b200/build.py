"""Build libprocgen_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libprocgen_b200.so")
REPO_ROOT = os.path.dirname(PKG_DIR)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",          # the reference wheels are built without FMA (CMakeLists.txt:30)
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu*")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  glob.glob(os.path.join(CSRC, "games", "*.cuh")) +
                  [os.path.join(REPO_ROOT, "include", "procgen_b200.h")])


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build_variant(name, extra_flags):
    """A tuning variant of the product library (same sources, extra -D flags) next to it; selected at
    run time with PROCGEN_B200_LIB. Used by tools/ for A/B kernel experiments only."""
    out = os.path.join(PKG_DIR, f"libprocgen_b200_{name}.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc, *NVCC_FLAGS, *extra_flags, os.path.join(CSRC, "pg_runtime.cu"), "-o", out, "-lz", "-ldl"])
    return out


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PG_NVCC_EXTRA", "").split()
    tmp = LIB_PATH + ".building"
    cmd = [nvcc, *NVCC_FLAGS, *extra, os.path.join(CSRC, "pg_runtime.cu"), "-o", tmp, "-lz", "-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    os.replace(tmp, LIB_PATH)  # never leave a half-written library where a snapshot could pick it up
    return LIB_PATH


def build_hostsim(out_dir=None, force=False):
    """CPU debug harness for tests only (same sources, kernels as loops). Never used by the package."""
    out_dir = out_dir or os.path.join(REPO_ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libprocgen_hostsim.so")
    if not force and os.path.exists(out) and all(os.path.getmtime(s) <= os.path.getmtime(out) for s in _sources()):
        return out
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-g", "-ffp-contract=off", "-fPIC", "-shared", "-DPG_HOSTSIM",
                           "-x", "c++", os.path.join(CSRC, "pg_runtime.cu"), "-o", out, "-lz", "-ldl"])
    return out
