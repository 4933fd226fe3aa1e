"""Shared test helpers: lockstep comparison of any libenv-ABI implementation against the oracle."""
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
