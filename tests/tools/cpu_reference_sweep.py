"""CPU rows of SURVEY §8(d) on this machine: the reference's game logic (compiled unmodified) with
(ii) the restated raster and (i) a free rasteriser (QT_SHIM_NODRAW), sweeping num_threads and N.
usage: python tests/tools/cpu_reference_sweep.py [game] [mode]   -> JSON lines on stdout"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, time, json
sys.path.insert(0, %r)
from oracle.ref_env import RefVecEnv, mt19937_actions
game, mode, n, threads, budget = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
env = RefVecEnv(n, game, distribution_mode=mode, num_levels=0, start_level=0, rand_seed=0, num_threads=threads)
acts = mt19937_actions(0, n, 64)
for t in range(5):
    env.act(acts[t]); env.observe()
t0 = time.perf_counter(); steps = 0
while time.perf_counter() - t0 < budget:
    env.act(acts[steps %% 64]); env.observe(); steps += 1
el = time.perf_counter() - t0
print(json.dumps({"steps_per_s": n * steps / el, "vec_steps": steps, "seconds": el}))
''' % ROOT


def main():
    game = sys.argv[1] if len(sys.argv) > 1 else "coinrun"
    mode = sys.argv[2] if len(sys.argv) > 2 else "easy"
    ncpu = os.cpu_count()
    for nodraw in (0, 1):
        for n in (64, 4096):
            for threads in (0, 4, ncpu):
                env = dict(os.environ)
                if nodraw:
                    env["QT_SHIM_NODRAW"] = "1"
                out = subprocess.run([sys.executable, "-c", CHILD, game, mode, str(n), str(threads), "4"], env=env, capture_output=True, text=True)
                line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else "{}"
                d = json.loads(line)
                d.update({"game": game, "mode": mode, "num_envs": n, "num_threads": threads, "host_cpus": ncpu,
                          "row": "logic only (free rasteriser)" if nodraw else "logic + restated raster"})
                print(json.dumps(d), flush=True)


if __name__ == "__main__":
    main()
