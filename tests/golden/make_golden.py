"""Generates tests/golden/*.npz from oracle/_ref (the reference's own game logic compiled unmodified
+ the raster restatement). Run in the build container: python tests/golden/make_golden.py
Each fixture: seeded actions, per-step rew/first/info, per-step sha256 of the rgb batch, final frame."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_env import RefVecEnv, mt19937_actions  # noqa: E402

CASES = [("coinrun", "easy", 8, 96), ("coinrun", "hard", 8, 96), ("bigfish", "hard", 8, 96), ("maze", "hard", 8, 96), ("heist", "hard", 8, 96),
         ("bossfight", "hard", 8, 96), ("caveflyer", "hard", 8, 96), ("chaser", "hard", 8, 96), ("climber", "hard", 8, 96),
         ("dodgeball", "hard", 8, 96), ("fruitbot", "hard", 8, 96), ("jumper", "hard", 8, 96), ("jumper", "easy", 8, 96),
         ("leaper", "hard", 8, 96), ("miner", "hard", 8, 96), ("ninja", "hard", 8, 96), ("plunder", "hard", 8, 96),
         ("starpilot", "hard", 8, 96), ("maze", "memory", 4, 64), ("dodgeball", "extreme", 4, 64)]


def main():
    for name, mode, n, steps in CASES:
        env = RefVecEnv(n, name, distribution_mode=mode, num_levels=200, start_level=0, rand_seed=0)
        acts = mt19937_actions(1, n, steps)
        env.observe()
        rew, first, seed, prev_seed, prev_complete, sha = [], [], [], [], [], []
        for t in range(steps):
            env.act(acts[t])
            r, ob, f = env.observe()
            rew.append(r.copy()); first.append(f.copy())
            seed.append(env.info["level_seed"].copy()); prev_seed.append(env.info["prev_level_seed"].copy())
            prev_complete.append(env.info["prev_level_complete"].copy())
            sha.append(hashlib.sha256(ob["rgb"].tobytes()).hexdigest())
        np.savez_compressed(os.path.join(HERE, f"{name}_{mode}.npz"), env_name=name, mode=mode, num=n, num_levels=200,
                            rand_seed=0, actions=acts, rew=np.array(rew), first=np.array(first), level_seed=np.array(seed),
                            prev_level_seed=np.array(prev_seed), prev_level_complete=np.array(prev_complete),
                            rgb_sha256=np.array(sha), last_rgb=ob["rgb"].copy())
        env.close()
        print("wrote", name, mode)


if __name__ == "__main__":
    main()
