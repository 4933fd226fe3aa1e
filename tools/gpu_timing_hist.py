"""Per-env logic duration histogram (needs PGB200_DEBUG_TIMING=1)."""
import os, sys
os.environ["PGB200_DEBUG_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from procgen_b200 import ProcgenGym3Env
game = sys.argv[1] if len(sys.argv) > 1 else "coinrun"
mode = sys.argv[2] if len(sys.argv) > 2 else "easy"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
env = ProcgenGym3Env(n, game, distribution_mode=mode, num_levels=0, rand_seed=0)
g = torch.Generator(device="cuda").manual_seed(0)
for t in range(60):
    env.act(torch.randint(0, 15, (n,), device="cuda", dtype=torch.int32, generator=g))
rew, ob, first = env.observe()
torch.cuda.synchronize()
cyc = np.zeros(n, np.uint32)
assert env._lib.pgb200_debug_cycles(env._h, cyc.ctypes.data) == 0
first = first.cpu().numpy()
print("envs", n, "resets this step", int(first.sum()))
for name, sel in [("no-reset", ~first), ("reset", first)]:
    c = cyc[sel]
    if len(c):
        print(name, "n", len(c), "mean %.0f" % c.mean(), "p50 %.0f p90 %.0f p99 %.0f max %.0f cycles" % tuple(np.percentile(c, [50, 90, 99, 100])))
print("sum cycles / (148 SMs * 48 warps) = %.0f cycles ~ %.2f ms at 1.9 GHz" % (cyc.sum() / (148 * 48), cyc.sum() / (148 * 48) / 1.9e6))
if os.environ.get("PG_PHASES"):
    import ctypes as C, struct
    env._lib.pgb200_debug_read_env.restype = C.c_int
    buf = (C.c_ubyte * 1024)()
    idx = np.nonzero(first)[0][:64] if not os.environ.get("PG_PHASES_ALL") else np.arange(0, n, max(1, n // 256))
    acc = np.zeros(12)
    for e in idx:
        env._lib.pgb200_debug_read_env(env._h, int(e), buf, None, 0)
        acc += np.array(struct.unpack_from("<12I", bytes(buf), 240), dtype=np.float64)
    if len(idx):
        print("mean cycles per marked phase over", len(idx), "reset envs:", [int(v) for v in acc / len(idx)])

if os.environ.get("PG_SLOWEST"):
    import ctypes as C, struct
    env._lib.pgb200_debug_read_env.restype = C.c_int
    buf = (C.c_ubyte * 1024)()
    order = np.argsort(-cyc.astype(np.int64))
    print("slowest non-reset envs: cycles | phases [pre, step_entities, collisions, erase] n_ents")
    shown = 0
    for e in order:
        if first[e]:
            continue
        env._lib.pgb200_debug_read_env(env._h, int(e), buf, None, 0)
        ph = struct.unpack_from("<12I", bytes(buf), 240)
        print(int(cyc[e]), ph[:5])
        shown += 1
        if shown >= 12:
            break
    med = np.argsort(cyc)[len(cyc) // 2]
    env._lib.pgb200_debug_read_env(env._h, int(med), buf, None, 0)
    print("median env:", int(cyc[med]), struct.unpack_from("<12I", bytes(buf), 240)[:5])
