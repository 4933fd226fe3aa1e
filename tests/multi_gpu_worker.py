"""Worker of tests/test_multi_gpu.py — one process per GPU (torchrun, NCCL). Each rank owns a shard of
ONE logical VecGame (ProcgenGym3Env(shard=(rank, world))); every step all shards are gathered to
rank 0 — through the peer-write path (symmetric memory + mirror copies) and through the plain NCCL
gather — and rank 0 compares the gathered frames with the unsharded oracle run, bit for bit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist


def main():
    game, per, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    from oracle.ref_env import RefVecEnv, mt19937_actions
    from procgen_b200 import ProcgenGym3Env

    kw = dict(distribution_mode="hard", num_levels=200, start_level=0, rand_seed=0)
    total = per * world
    acts = mt19937_actions(0, total, steps)
    ref = RefVecEnv(total, game, **kw) if rank == 0 else None
    results = {}
    for how in ("peer", "nccl"):
        env = ProcgenGym3Env(per, game, shard=(rank, world), **kw)
        used_peer = env.enable_peer_gather(0) if how == "peer" else False
        if rank == 0:
            ref_env = RefVecEnv(total, game, **kw)
            ref_env.observe()
        bad = 0
        for t in range(steps):
            env.act(torch.as_tensor(acts[t][rank * per:(rank + 1) * per], device="cuda"))
            env.observe()
            g = env.gather_observations(0)
            if rank == 0:
                ref_env.act(acts[t])
                _, ob, _ = ref_env.observe()
                if not np.array_equal(g.cpu().numpy(), ob["rgb"]):
                    bad += 1
        if rank == 0:
            ref_env.close()
        results[how] = (bad, used_peer, env.gather_how())
        assert env.errors() == 0
        env.close()
    if rank == 0:
        for how, (bad, used_peer, desc) in results.items():
            print(f"MULTI_GPU {how}: mismatching steps {bad}/{steps}; peer path active: {used_peer}; {desc}")
        print("MULTI_GPU_OK" if all(b == 0 for b, _, _ in results.values()) else "MULTI_GPU_FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
