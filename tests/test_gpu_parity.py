"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs — bit-exact for rew / first / info AND rgb (the oracle's raster restatement and the
device rasteriser implement the same integer rules, so the tolerance is 0)."""
import hashlib
import os

import numpy as np
import pytest

from helpers import make_pair, run_lockstep

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,mode,n,steps", [
    ("coinrun", "easy", 64, 1000),   # BASELINE.json configs[0]
    ("coinrun", "hard", 64, 1000),
    ("bigfish", "hard", 64, 1000),   # configs[2] game, small N
    ("bigfish", "easy", 32, 600),
    ("maze", "hard", 64, 800),       # configs[3] games, small N
    ("maze", "easy", 32, 500),
    ("heist", "hard", 64, 800),      # configs[3]; exercises the rotated-sprite raster paths
    ("heist", "easy", 32, 500),
    ("miner", "hard", 32, 600),
    ("leaper", "hard", 32, 600),
    ("plunder", "hard", 32, 800),
    ("chaser", "hard", 32, 600),
    ("climber", "hard", 32, 600),
    ("ninja", "hard", 32, 800),
    ("fruitbot", "hard", 32, 600),
    ("caveflyer", "hard", 32, 600),
    ("bossfight", "hard", 32, 800),
    ("dodgeball", "hard", 32, 600),
    ("dodgeball", "memory", 16, 400),
    ("starpilot", "hard", 32, 800),
    ("starpilot", "extreme", 16, 400),
    ("jumper", "hard", 32, 600),
    ("jumper", "easy", 32, 600),
])
def test_libenv_host_buffers_bit_exact(ref_lib, product_lib, name, mode, n, steps):
    ref, dut = make_pair(product_lib, n, name, distribution_mode=mode, num_levels=200, start_level=0, rand_seed=0)
    run_lockstep(ref, dut, steps)
    ref.close()
    dut.close()


@pytest.mark.parametrize("name", ["coinrun", "bossfight", "chaser", "starpilot", "jumper", "plunder", "heist", "leaper"])
def test_state_blobs_byte_identical_and_portable(ref_lib, product_lib, name):
    from helpers import run_state_roundtrip
    from oracle.ref_env import RefVecEnv, default_pack

    kw = dict(distribution_mode="hard", num_levels=200, start_level=0)
    run_state_roundtrip(lambda seed: RefVecEnv(8, name, rand_seed=seed, **kw),
                        lambda seed: RefVecEnv(8, name, rand_seed=seed, lib_path=product_lib, resource_root=default_pack(), **kw),
                        8, 150)


def test_python_api_state_roundtrip(product_lib):
    """ProcgenGym3Env.get_state / set_state (env.py:140-153): restoring a snapshot replays the same frames."""
    import torch

    from procgen_b200 import ProcgenGym3Env

    env = ProcgenGym3Env(16, "coinrun", distribution_mode="hard", num_levels=0, start_level=0, rand_seed=7)
    gen = torch.Generator(device="cuda").manual_seed(3)
    acts = torch.randint(0, 15, (40, 16), device="cuda", dtype=torch.int32, generator=gen)
    for t in range(10):
        env.act(acts[t])
        env.observe()
    states = env.callmethod("get_state")
    frames = []
    for t in range(10, 40):
        env.act(acts[t])
        frames.append(env.observe()[1]["rgb"].clone())
    env.callmethod("set_state", states)
    for t in range(10, 40):
        env.act(acts[t])
        assert torch.equal(env.observe()[1]["rgb"], frames[t - 10])
    assert env.errors() == 0
    env.close()


def test_sixteen_game_list_bit_exact(ref_lib, product_lib):
    """BASELINE.json configs[4] shape on one GPU: env n plays game n % 16."""
    names = "bigfish,bossfight,caveflyer,chaser,climber,coinrun,dodgeball,fruitbot,heist,jumper,leaper,maze,miner,ninja,plunder,starpilot"
    ref, dut = make_pair(product_lib, 64, names, distribution_mode="hard", num_levels=200, start_level=0, rand_seed=0)
    run_lockstep(ref, dut, 500)
    ref.close()
    dut.close()


@pytest.mark.parametrize("fixture", sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz")))
def test_device_api_reproduces_golden(product_lib, fixture):
    import torch

    from procgen_b200 import ProcgenGym3Env

    g = np.load(os.path.join(GOLDEN, fixture), allow_pickle=False)
    env = ProcgenGym3Env(int(g["num"]), str(g["env_name"]), distribution_mode=str(g["mode"]),
                         num_levels=int(g["num_levels"]), start_level=0, rand_seed=int(g["rand_seed"]))
    acts = g["actions"]
    for t in range(acts.shape[0]):
        env.act(torch.as_tensor(acts[t], device="cuda"))
        rew, ob, first = env.observe()
        assert np.array_equal(rew.cpu().numpy(), g["rew"][t])
        assert np.array_equal(first.cpu().numpy(), g["first"][t].astype(bool))
        info = env.get_info_tensors()
        assert np.array_equal(info["level_seed"].cpu().numpy(), g["level_seed"][t])
        assert hashlib.sha256(ob["rgb"].cpu().numpy().tobytes()).hexdigest() == str(g["rgb_sha256"][t])
    assert env.errors() == 0
    env.close()


ALL16 = "bigfish,bossfight,caveflyer,chaser,climber,coinrun,dodgeball,fruitbot,heist,jumper,leaper,maze,miner,ninja,plunder,starpilot"


@pytest.mark.parametrize("name,mode,n_big,steps", [
    ("coinrun", "easy", 65536, 60),    # BASELINE configs[1]
    ("bigfish", "hard", 65536, 40),    # configs[2]
    ("maze", "hard", 32768, 40),       # configs[3]
    ("heist", "hard", 32768, 40),      # configs[3]
    (ALL16, "hard", 32768, 40),        # configs[4], one GPU's share
    ("bossfight", "hard", 16384, 60),  # hundreds of entities per env: the parallel list compaction under load
    ("jumper", "hard", 16384, 40),     # warp-parallel level generation under load
])
def test_full_size_properties(ref_lib, product_lib, name, mode, n_big, steps):
    """Benchmark-size runs: size-independent properties.
    (a) prefix property: envs [0,64) of the big run == the 64-env oracle run (per-env independence +
        sequential seed chain); (b) run-to-run determinism via a checksum of all observations;
    (c) no env latched an error bit."""
    import torch

    from oracle.ref_env import RefVecEnv, mt19937_actions
    from procgen_b200 import ProcgenGym3Env

    n_small = 64
    kw = dict(distribution_mode=mode, num_levels=0, start_level=0, rand_seed=0)
    acts_small = mt19937_actions(7, n_small, steps)
    gen = torch.Generator(device="cuda").manual_seed(0)
    acts_big = torch.randint(0, 15, (steps, n_big), device="cuda", dtype=torch.int32, generator=gen)
    acts_big[:, :n_small] = torch.as_tensor(acts_small, device="cuda")

    def run():
        env = ProcgenGym3Env(n_big, name, **kw)
        digest = hashlib.sha256()
        heads = []
        for t in range(steps):
            env.act(acts_big[t])
            rew, ob, first = env.observe()
            heads.append((rew[:n_small].cpu().numpy().copy(), ob["rgb"][:n_small].cpu().numpy().copy(),
                          first[:n_small].cpu().numpy().copy()))
            if t % 10 == 9:
                digest.update(ob["rgb"].cpu().numpy().tobytes())
                digest.update(rew.cpu().numpy().tobytes())
        assert env.errors() == 0
        env.close()
        return digest.hexdigest(), heads

    d1, heads = run()
    d2, _ = run()
    assert d1 == d2
    ref = RefVecEnv(n_small, name, **kw)
    ref.observe()
    for t in range(steps):
        ref.act(acts_small[t])
        rew, ob, first = ref.observe()
        assert np.array_equal(heads[t][0], rew), f"step {t}"
        assert np.array_equal(heads[t][1], ob["rgb"]), f"step {t}"
        assert np.array_equal(heads[t][2], first.astype(bool)), f"step {t}"
    ref.close()
