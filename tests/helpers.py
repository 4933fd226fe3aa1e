"""Shared test helpers: lockstep comparison of any libenv-ABI implementation against the oracle."""
import os

import numpy as np

from oracle.ref_env import RefVecEnv, default_pack, mt19937_actions


def make_pair(lib_path, num, env_name, extra_options=None, **kw):
    ref = RefVecEnv(num, env_name, **kw)
    dut = RefVecEnv(num, env_name, lib_path=lib_path, resource_root=default_pack(), extra_options=extra_options, **kw)
    return ref, dut


def assert_same_observation(ref, dut, t, rgb_tol=0):
    r1, o1, f1 = ref.observe()
    r2, o2, f2 = dut.observe()
    assert np.array_equal(r1, r2), f"step {t}: rew differs at envs {np.nonzero(r1 != r2)[0][:8]}"
    assert np.array_equal(f1, f2), f"step {t}: first differs at envs {np.nonzero(f1 != f2)[0][:8]}"
    for k in ref.info:
        assert np.array_equal(ref.info[k], dut.info[k]), f"step {t}: info[{k}] differs"
    if rgb_tol == 0:
        if not np.array_equal(o1["rgb"], o2["rgb"]):
            d = np.abs(o1["rgb"].astype(int) - o2["rgb"].astype(int))
            bad = np.nonzero(d.reshape(d.shape[0], -1).sum(1))[0]
            raise AssertionError(f"step {t}: rgb differs in envs {bad[:8]}, {int((d.sum(-1) > 0).sum())} px, max |d| {d.max()}")
    else:
        d = np.abs(o1["rgb"].astype(int) - o2["rgb"].astype(int))
        assert d.max() <= rgb_tol, f"step {t}: rgb max |d| {d.max()} > {rgb_tol}"


def run_lockstep(ref, dut, steps, seed=0, rgb_tol=0):
    acts = mt19937_actions(seed, ref.num, steps)
    assert_same_observation(ref, dut, -1, rgb_tol)
    for t in range(steps):
        ref.act(acts[t])
        dut.act(acts[t])
        assert_same_observation(ref, dut, t, rgb_tol)
    if hasattr(dut.lib, "pgb200_get_errors"):
        import ctypes as C

        dut.lib.pgb200_get_errors.restype = C.c_uint32
        err = dut.lib.pgb200_get_errors(C.c_void_p(dut.h), None)
        assert err == 0, f"device latched error bits {err:#x} (capacity overflow / unsupported feature)"


def run_state_roundtrip(make_ref, make_dut, num, steps, check_every=25):
    """Full-state parity through the reference's own wire format (vecgame.cpp:437-457), following the
    shape of the reference's state_test.py: (1) the blobs of both implementations are byte-identical
    along a lockstep run, (2) reference blobs loaded into FRESH envs of both implementations (built
    with another seed) give the same frame and info immediately and the same trajectory afterwards,
    (3) which is also the trajectory the original envs continue on."""
    ref, dut = make_ref(0), make_dut(0)
    acts = mt19937_actions(0, num, 2 * steps)
    for t in range(steps):
        if t % check_every == 0:
            for e in range(num):
                assert ref.get_state(e) == dut.get_state(e), f"step {t} env {e}: state blobs differ"
        ref.act(acts[t])
        dut.act(acts[t])
        ref.observe()
        dut.observe()
    blobs = [ref.get_state(e) for e in range(num)]
    assert blobs == [dut.get_state(e) for e in range(num)]
    ref2, dut2 = make_ref(5), make_dut(5)
    for e in range(num):
        ref2.set_state(e, blobs[e])
        dut2.set_state(e, blobs[e])
    assert_same_observation(ref2, dut2, "after set_state")
    for t in range(steps, 2 * steps):
        ref.act(acts[t])
        ref2.act(acts[t])
        dut2.act(acts[t])
        assert_same_observation(ref2, dut2, t)
        r0, o0, f0 = ref.observe()
        r2, o2, f2 = ref2.observe()
        assert np.array_equal(o0["rgb"], o2["rgb"]) and np.array_equal(r0, r2) and np.array_equal(f0, f2)
    assert [ref2.get_state(e) for e in range(num)] == [dut2.get_state(e) for e in range(num)]
    for env in (ref, dut, ref2, dut2):
        env.close()


SNAP_SCRIPT = r"""
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from helpers import make_pair, run_lockstep
for name, mode in [("coinrun", "hard"), ("maze", "hard"), ("bigfish", "hard"), ("heist", "hard"), ("fruitbot", "hard")]:
    ref, dut = make_pair({lib!r}, 8, name, extra_options={{"snap_target_rect": False}}, distribution_mode=mode,
                         num_levels=200, start_level=0, rand_seed=0)
    run_lockstep(ref, dut, 200)
    ref.close(); dut.close()
print("SNAP_OFF_OK")
"""


def run_snap_off_lockstep(lib):
    """The Qt-5-style un-snapped target rect (the known Qt 5 / Qt 6 risk) is a switch in both
    implementations; the oracle reads it from the environment once per process, hence the subprocess."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, QT_SHIM_SNAP="0")
    out = subprocess.run([sys.executable, "-c", SNAP_SCRIPT.format(root=root, lib=lib)], env=env, capture_output=True, text=True)
    assert "SNAP_OFF_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
