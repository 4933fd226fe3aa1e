#!/usr/bin/env python
"""bench.py — env-steps/sec of the B200 Procgen hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--game coinrun] ...

A "step" = one act()+observe() pass over one batch of `--envs-per-gpu` environments (default 65536,
BASELINE configs[1]: coinrun, distribution_mode=easy) with synthetic uniform-random actions.
  value   whole-job env-steps/s, actions already resident in HBM, observations left in HBM
  e2e     same metric through the libenv C ABI with HOST buffers (pinned): actions H2D and
          rgb/rew/first/info D2H inside the timed region — the drop-in path gym3's CEnv would drive
  roofline  HBM bound; algorithmic bytes = 12288 B rgb write per env-step
  cpu_baseline  oracle/_ref (reference game logic compiled unmodified + restated Qt raster) on the
          host cores, bounded sample
With --impl reference the reference's CPU implementation (oracle/_ref, all host threads) is timed
instead and reported on the same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_ENV_STEP = 64 * 64 * 3
METRIC = "env-steps/sec"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(game, mode, budget_s=15.0, envs_per_worker=64, workers=None):
    """Times oracle/_ref on ALL host cores on a bounded sample of the same workload.

    The reference scales across cores by running independent VecGames (its own thread pool is one
    mutex + notify_all per game and gets slower with threads for cheap steps), so the
    baseline is `workers` = nproc independent reference VecGames (num_threads=0, 64 envs each),
    each driven by its own host thread (ctypes releases the GIL inside libenv_act/observe)."""
    import numpy as np

    from oracle.ref_env import RefVecEnv

    cores = os.cpu_count() or 1
    workers = cores if workers is None else workers
    n = envs_per_worker
    envs = [RefVecEnv(n, game, distribution_mode=mode, num_levels=0, start_level=0, rand_seed=w, num_threads=0)
            for w in range(workers)]
    counts = [0] * workers
    start = threading.Barrier(workers + 1)
    t_end = [0.0]

    def work(w):
        env = envs[w]
        rng = np.random.RandomState(w)
        acts = rng.randint(0, 15, size=(64, n)).astype(np.int32)
        env.observe()
        for i in range(2):
            env.act(acts[i])
            env.observe()
        start.wait()
        i = 0
        while time.perf_counter() < t_end[0]:
            env.act(acts[i & 63])
            env.observe()
            i += 1
        counts[w] = i

    threads = [threading.Thread(target=work, args=(w,)) for w in range(workers)]
    for t in threads:
        t.start()
    t_end[0] = time.perf_counter() + budget_s + 3600.0
    start.wait()
    t0 = time.perf_counter()
    t_end[0] = t0 + budget_s
    for t in threads:
        t.join()
    el = time.perf_counter() - t0
    for e in envs:
        e.close()
    total = n * sum(counts)
    return total / el, {"cores": workers, "sample": f"{workers} independent reference VecGames x {n} envs, {sum(counts)} "
                        f"vec-steps total in {el:.1f}s (num_threads=0 each, one host thread per VecGame)",
                        "seconds": el, "env_steps": total}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    per_step_budget = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    rates = []
    detail = None
    for _ in range(args.warmup):
        cpu_reference_rate(args.game, args.mode, budget_s=min(per_step_budget, 2.0))
    for _ in range(args.steps):
        r, detail = cpu_reference_rate(args.game, args.mode, budget_s=per_step_budget)
        rates.append(r)
    value = sum(rates) / len(rates)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * args.envs_per_gpu * args.gpus / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u8", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": detail["cores"], "kind": "reference",
                         "sample": detail["sample"] + " per bench step; reference game logic compiled unmodified, "
                                   "Qt raster restated on CPU (Qt itself is not installable here)"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(out))


def workload_config(args):
    return {"workload": f"{args.game} distribution_mode={args.mode} num_envs={args.envs_per_gpu}/GPU num_levels=0 rand_seed=0, "
                        "uniform random actions", "envs_per_gpu": args.envs_per_gpu, "game": args.game,
            "distribution_mode": args.mode,
            "parallelism": f"env-sharded x{args.gpus}, " + ("rgb gathered to rank 0 every step (NCCL)" if getattr(args, "gather", False)
                                                               else "no per-step collective"),
            "l2": "per-step working set (12 KiB obs + env state per env x num_envs) exceeds the 126 MB L2; no explicit flush"}


def run_ours(args):
    import numpy as np
    import torch

    from procgen_b200 import ProcgenGym3Env

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    n = args.envs_per_gpu
    K, W = args.steps, args.warmup

    env = ProcgenGym3Env(n, args.game, distribution_mode=args.mode, num_levels=0, start_level=0, rand_seed=0,
                         shard=(rank, world) if world > 1 else None)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    actions = torch.randint(0, 15, (W + K, n), device=dev, dtype=torch.int32, generator=gen)
    env.observe()
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for t in range(W):
        env.act(actions[t])
        env.observe()
    barrier()

    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    time.sleep(0.15)
    launches0 = env.kernel_launches()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    ev0.record()
    for t in range(K):
        env._ac.copy_(actions[W + t], non_blocking=True)      # act(): action tensor -> library buffer (D2D)
        kev[t][0].record()
        env._lib.pgb200_act_device(env._h)                      # the step+render kernel
        kev[t][1].record()
        rew, ob, first = env.observe()                           # aliases of HBM buffers; nothing to copy
        if args.gather and dist is not None:
            gathered = env.gather_observations(0)                # the one collective of SURVEY §8e
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    kernel_ms = [a.elapsed_time(b) for a, b in kev]
    launches = env.kernel_launches() - launches0
    clocks = sampler.stop()
    checksum = int(ob["rgb"].sum().item())
    errors = env.errors()
    if dist is not None:
        tmax = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    value = n * world * K / (elapsed_ms / 1000.0)

    # ---- roofline pass: the same steps with every kernel launched back to back on one stream (one
    # launch per game covering all its envs) and CUDA events recorded around each launch on that
    # stream, so a kernel's duration is its own. In the throughput loop above launches of different
    # env chunks overlap on the SMs and an event pair would also count the neighbours' time.
    Kr = max(5, min(20, K))
    env.set_launch_shape(chunks=1, serialize=True)
    for t in range(3):
        env.act(actions[t % (W + K)])
        env.observe()
    env.kernel_timing_begin(Kr * 64)
    for t in range(Kr):
        env.act(actions[(W + t) % (W + K)])
        env.observe()
    ktimes = env.kernel_timing_end()
    env.set_launch_shape(chunks=0, serialize=False)
    env.close()

    # ---- e2e: the reference-facing C ABI with host buffers, H2D/D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        Ke = max(3, min(args.e2e_steps, K))
        henv = ProcgenGym3Env(n, args.game, distribution_mode=args.mode, num_levels=0, start_level=0, rand_seed=0,
                              shard=(rank, world) if world > 1 else None, host_buffers=True)
        host_actions = actions[: W + Ke].cpu().numpy()
        for t in range(min(W, 3)):
            henv.act(host_actions[t])
            henv.observe()
        barrier()
        t0 = time.perf_counter()
        for t in range(Ke):
            henv.act(host_actions[W + t])
            rew_h, ob_h, first_h = henv.observe()
        barrier()
        el = time.perf_counter() - t0
        if dist is not None:
            tmax = torch.tensor([el], device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            el = float(tmax.item())
        e2e = {"value": n * world * Ke / el, "unit": "env-steps/s", "h2d_bytes_per_step": 4 * n * world,
               "d2h_bytes_per_step": (64 * 64 * 3 + 4 + 1 + 4 + 1 + 4) * n * world, "steps": Ke,
               "api": "libenv_act + libenv_observe (host numpy buffers)", "timer": "host perf_counter around the calls"}
        henv.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_kind = measured_peak_gbs()
    # Dominant kernel = render_kernel (writes the observations). One launch renders `envs_per_launch`
    # frames = 12 288 algorithmic bytes each; its duration comes from the roofline pass.
    pairs = max(1, ktimes["launch_pairs"])
    render_ms_avg = ktimes["render_ms"] / pairs
    logic_ms_avg = ktimes["logic_ms"] / pairs
    envs_per_launch = ktimes["env_steps"] / pairs
    algo_bytes_per_launch = ALGO_BYTES_PER_ENV_STEP * envs_per_launch
    achieved = algo_bytes_per_launch / (render_ms_avg / 1000.0) / 1e9 if render_ms_avg > 0 else 0.0
    k_avg_ms = sum(kernel_ms) / len(kernel_ms)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            # ncu dram bytes of one render_kernel launch, scaled to this run's envs per launch
            traffic = tj["render_kernel"]["dram_bytes_per_env"] * envs_per_launch
        except Exception:
            traffic = None
    cpu = None
    if not args.no_cpu_baseline:
        rate, detail = cpu_reference_rate(args.game, args.mode, budget_s=args.cpu_budget)
        cpu = {"value": rate, "unit": "env-steps/s", "cores": detail["cores"], "kind": "reference",
               "sample": detail["sample"] + "; reference game logic compiled unmodified + Qt raster restated on CPU"}
    out = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+u8", "data": "synthetic", "config": workload_config(args),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_kind": peak_kind, "kernel": f"render_kernel<{args.game}>",
                     "kernel_ms_avg": render_ms_avg, "launches_timed": ktimes["launch_pairs"], "envs_per_launch": envs_per_launch,
                     "algorithmic_bytes_per_launch": algo_bytes_per_launch,
                     "logic_kernel_ms_avg": logic_ms_avg, "step_ms_avg": k_avg_ms,
                     "how": "CUDA events around each launch, launches serialised on one stream (separate pass of %d steps)" % Kr,
                     "whole_step_achieved": ALGO_BYTES_PER_ENV_STEP * n / (k_avg_ms / 1000.0) / 1e9},
        "cpu_baseline": cpu, "obs_checksum": checksum, "env_error_bits": errors,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--game", default="coinrun")
    ap.add_argument("--mode", default="easy")
    ap.add_argument("--envs-per-gpu", type=int, default=65536)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", action="store_true",
                    help="BASELINE configs[4] variant: NCCL-gather every step's rgb shard to rank 0 inside the timed region")
    args = ap.parse_args()
    if args.game == "all16":  # BASELINE configs[4]: env n plays game n % 16
        args.game = ("bigfish,bossfight,caveflyer,chaser,climber,coinrun,dodgeball,fruitbot,heist,jumper,leaper,maze,"
                     "miner,ninja,plunder,starpilot")
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
