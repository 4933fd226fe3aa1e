"""The N>1 path on CPU: two ranks (torch.distributed, gloo, 127.0.0.1), each owning the contiguous env
range [r*N/2, (r+1)*N/2) of ONE logical VecGame. Every rank replays the global seed
chain for its range (env_index_offset / env_index_total), steps its shard with the shared action
stream, and the only collective on the path — the observation gather to rank 0 — reassembles
frames that must equal the unsharded reference run bit for bit. The device engine runs in its host
debug build here (no GPU in the loop); bench.py --gpus N is the same layout over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_ENVS, STEPS, WORLD = 16, 120, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, hostsim_lib, game, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from oracle.ref_env import RefVecEnv, default_pack, mt19937_actions

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = N_ENVS // world
    kw = dict(distribution_mode="hard", num_levels=200, start_level=0, rand_seed=0)
    shard = RefVecEnv(per, game, lib_path=hostsim_lib, resource_root=default_pack(),
                      extra_options={"env_index_offset": rank * per, "env_index_total": N_ENVS}, **kw)
    acts = mt19937_actions(0, N_ENVS, STEPS)
    frames, rews, firsts = [], [], []
    for t in range(STEPS):
        shard.act(acts[t][rank * per:(rank + 1) * per].copy())
        rew, ob, first = shard.observe()
        mine = torch.from_numpy(ob["rgb"].copy())
        mine_r = torch.from_numpy(rew.copy())
        mine_f = torch.from_numpy(first.astype(np.uint8).copy())
        if rank == 0:
            parts = [torch.empty_like(mine) for _ in range(world)]
            parts_r = [torch.empty_like(mine_r) for _ in range(world)]
            parts_f = [torch.empty_like(mine_f) for _ in range(world)]
            dist.gather(mine, parts, dst=0)
            dist.gather(mine_r, parts_r, dst=0)
            dist.gather(mine_f, parts_f, dst=0)
            if t % 10 == 9 or t == STEPS - 1:
                frames.append(torch.cat(parts).numpy())
            rews.append(torch.cat(parts_r).numpy())
            firsts.append(torch.cat(parts_f).numpy())
        else:
            dist.gather(mine, None, dst=0)
            dist.gather(mine_r, None, dst=0)
            dist.gather(mine_f, None, dst=0)
    dist.barrier()
    if rank == 0:
        np.savez(out_path, frames=np.array(frames), rew=np.array(rews), first=np.array(firsts))
    dist.destroy_process_group()


@pytest.mark.parametrize("game", ["bigfish,plunder"])
def test_two_ranks_gathered_equal_unsharded_reference(ref_lib, hostsim_lib, tmp_path, game):
    import torch.multiprocessing as mp

    from oracle.ref_env import RefVecEnv, mt19937_actions

    out = str(tmp_path / "gathered.npz")
    import torch  # noqa: F401  (imported before the fork so the ranks do not pay for it again)

    mp.start_processes(_rank_main, args=(WORLD, _free_port(), hostsim_lib, game, out), nprocs=WORLD, join=True, start_method="fork")
    got = np.load(out)
    ref = RefVecEnv(N_ENVS, game, distribution_mode="hard", num_levels=200, start_level=0, rand_seed=0)
    acts = mt19937_actions(0, N_ENVS, STEPS)
    k = 0
    for t in range(STEPS):
        ref.act(acts[t])
        rew, ob, first = ref.observe()
        assert np.array_equal(rew, got["rew"][t]), f"step {t}: rew"
        assert np.array_equal(first.astype(np.uint8), got["first"][t]), f"step {t}: first"
        if t % 10 == 9 or t == STEPS - 1:
            assert np.array_equal(ob["rgb"], got["frames"][k]), f"step {t}: gathered rgb != unsharded reference"
            k += 1
    ref.close()
