"""Build libprocgen_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

One translation unit per game (csrc/games_tu/tu_<game>.cu: that game's logic / render kernels) plus
csrc/pg_runtime.cu (host runtime + C ABI); the units compile in parallel into csrc/_obj/ and are
linked into one shared library."""
from __future__ import annotations

import glob
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libprocgen_b200.so")
REPO_ROOT = os.path.dirname(PKG_DIR)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",          # the reference wheels are built without FMA (CMakeLists.txt:30)
    "-Xcompiler", "-fPIC", "-Xfatbin", "-compress-all",
]
HOSTSIM_FLAGS = ["-std=c++17", "-O2", "-g", "-ffp-contract=off", "-fPIC", "-DPG_HOSTSIM", "-x", "c++"]


def _units():
    return [os.path.join(CSRC, "pg_runtime.cu")] + sorted(glob.glob(os.path.join(CSRC, "games_tu", "tu_*.cu")))


def _headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  glob.glob(os.path.join(CSRC, "games", "*.cuh")) + [os.path.join(REPO_ROOT, "include", "procgen_b200.h")])


def _sources():
    return _units() + _headers()


def _deps(unit):
    """Headers a unit depends on: everything shared, and only its own game's header for a game unit."""
    base = os.path.basename(unit)
    deps = [unit]
    for h in _headers():
        in_games = os.path.basename(os.path.dirname(h)) == "games"
        if in_games and base.startswith("tu_") and os.path.basename(h) != base[3:-3] + ".cuh":
            continue
        deps.append(h)
    return deps


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def _compile_all(obj_dir, compiler_cmd, tag, force, verbose):
    """Compile every unit whose object is older than one of its dependencies; returns the objects."""
    os.makedirs(obj_dir, exist_ok=True)
    stamp = os.path.join(obj_dir, "flags.txt")
    flags_txt = " ".join(compiler_cmd)
    if not os.path.exists(stamp) or open(stamp).read() != flags_txt:
        force = True
    jobs = []
    objs = []
    for u in _units():
        o = os.path.join(obj_dir, os.path.basename(u)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or any(os.path.getmtime(d) > os.path.getmtime(o) for d in _deps(u)):
            jobs.append((u, o))

    def run(job):
        u, o = job
        cmd = [*compiler_cmd, "-c", u, "-o", o + ".tmp"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"{tag}: compiling {os.path.basename(u)} failed:\n{r.stdout[-4000:]}\n{r.stderr[-8000:]}")
        os.replace(o + ".tmp", o)

    if jobs:
        workers = int(os.environ.get("PG_BUILD_JOBS", str(min(len(jobs), os.cpu_count() or 4))))
        with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:
            list(ex.map(run, jobs))
    open(stamp, "w").write(flags_txt)
    return objs


def _nvcc():
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


VALIDATED_NVCC = ("release 12.9",)


def check_toolchain():
    """The logic kernel runs each env's scalar step on all 32 lanes of a warp in lockstep (pg_engine.cuh); the
    bit-exact GPU suite is what proves a given compiler keeps the lanes converged between the __syncwarp points.
    A compiler outside the validated list still builds, but says so: re-run `pytest -m gpu` before trusting it."""
    import sys

    try:
        txt = subprocess.run([_nvcc(), "--version"], capture_output=True, text=True, check=True).stdout
    except (OSError, subprocess.CalledProcessError):
        return None
    if not any(v in txt for v in VALIDATED_NVCC):
        print(f"procgen_b200: nvcc is not one of the validated releases {VALIDATED_NVCC}; run the GPU parity suite "
              "(pytest -m gpu) before using this build", file=sys.stderr)
        return False
    return True


def build_variant(name, extra_flags):
    """A tuning variant of the product library (same sources, extra -D flags) next to it; selected at
    run time with PROCGEN_B200_LIB. Used by tools/ for A/B kernel experiments only."""
    out = os.path.join(PKG_DIR, f"libprocgen_b200_{name}.so")
    objs = _compile_all(os.path.join(CSRC, "_obj", name), [_nvcc(), *NVCC_FLAGS, *extra_flags], name, False, False)
    subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", out, "-lz", "-ldl"])
    return out


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    extra = os.environ.get("PG_NVCC_EXTRA", "").split()
    check_toolchain()
    objs = _compile_all(os.path.join(CSRC, "_obj", "product"), [_nvcc(), *NVCC_FLAGS, *extra], "product", force, verbose)
    tmp = LIB_PATH + ".building"
    subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", tmp, "-lz", "-ldl"])
    os.replace(tmp, LIB_PATH)  # never leave a half-written library where a snapshot could pick it up
    return LIB_PATH


def build_hostsim(out_dir=None, force=False):
    """CPU debug harness for tests only (same sources, kernels as loops). Never used by the package."""
    out_dir = out_dir or os.path.join(REPO_ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libprocgen_hostsim.so")
    if not force and os.path.exists(out) and all(os.path.getmtime(s) <= os.path.getmtime(out) for s in _sources()):
        return out
    objs = _compile_all(os.path.join(out_dir, "obj"), ["g++", *HOSTSIM_FLAGS], "hostsim", force, False)
    subprocess.check_call(["g++", "-shared", *objs, "-o", out + ".tmp", "-lz", "-ldl"])
    os.replace(out + ".tmp", out)
    return out
