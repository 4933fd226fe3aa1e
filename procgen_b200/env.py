"""ProcgenGym3Env on B200 — host-side mirror of the reference's Python boundary.

Mirrors ``procgen/env.py``: ``BaseProcgenEnv`` (:66-200) + ``ProcgenGym3Env`` (:203-246) with the same
constructor keywords, defaults, option marshalling and the gym3 ``Env`` surface callers use
(``num``, ``ob_space``, ``ac_space``, ``observe()``, ``act()``, ``get_info()``, ``callmethod()``).
Differences, all on purpose:

* ``observe()`` returns ``torch.cuda`` tensors that alias the library's HBM buffers (no copy, no
  host round trip); ``act()`` accepts a CUDA tensor (stays on device) or anything array-like.
* ``host_buffers=True`` selects the reference's exact contract instead: numpy buffers owned by
  the caller, filled through ``libenv_set_buffers/act/observe`` like gym3's ``CEnv`` does.
* ``shard=(rank, world_size)`` makes this handle own envs ``[rank*num, (rank+1)*num)`` of one
  logical ``world_size*num``-env VecGame (same per-env seed chain, vecgame.cpp:301-314);
  ``gather_observations()`` is the single NCCL gather of SURVEY §8(e).

gym3 is not installable here, so the few space types used are defined below with gym3's names.
"""
from __future__ import annotations

import ctypes as C
import os
import random
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import libenv as L

MAX_STATE_SIZE = 2 ** 20  # env.py:12

ENV_NAMES = [  # env.py:14-31
    "bigfish", "bossfight", "caveflyer", "chaser", "climber", "coinrun", "dodgeball", "fruitbot",
    "heist", "jumper", "leaper", "maze", "miner", "ninja", "plunder", "starpilot",
]

EXPLORATION_LEVEL_SEEDS = {  # env.py:33-42
    "coinrun": 1949448038, "caveflyer": 1259048185, "leaper": 1318677581, "jumper": 1434825276,
    "maze": 158988835, "heist": 876640971, "climber": 1561126160, "ninja": 1123500215,
}

DISTRIBUTION_MODE_DICT = {"easy": 0, "hard": 1, "extreme": 2, "memory": 10, "exploration": 20}  # env.py:45-51


# ---- the slice of gym3.types the reference's callers touch
@dataclass(frozen=True)
class Discrete:
    n: int
    dtype_name: str = "int32"


@dataclass(frozen=True)
class TensorType:
    eltype: Discrete
    shape: Tuple[int, ...]


class DictType(dict):
    pass


def create_random_seed():
    """env.py:54-63 (mpi4py de-correlation becomes torch.distributed rank de-correlation)."""
    rand_seed = random.SystemRandom().randint(0, 2 ** 31 - 1)
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            rand_seed = rand_seed - (rand_seed % dist.get_world_size()) + dist.get_rank()
    except Exception:
        pass
    return rand_seed


def _broadcast_seed(seed):
    try:
        import torch
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            t = torch.tensor([seed], dtype=torch.int64, device=dev)
            dist.broadcast(t, src=0)
            return int(t.item())
    except Exception:
        pass
    raise ValueError("shard=(rank, world) needs an explicit rand_seed when torch.distributed is not initialised")


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so torch can alias library-owned HBM."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class BaseProcgenEnv:
    """env.py:66-200."""

    def __init__(self, num, env_name, options, debug=False, rand_seed=None, num_levels=0, start_level=0,
                 use_sequential_levels=False, debug_mode=0, resource_root=None, num_threads=4, render_mode=None,
                 host_buffers=False, device=None, shard=None, snap_target_rect=True, lib_path=None):
        self._lib = L.load(lib_path)
        self.combos = self.get_combos()
        if render_mode is None:
            render_human = False
        elif render_mode == "rgb_array":
            render_human = True
        else:
            raise Exception(f"invalid render mode {render_mode}")
        if render_human:
            raise NotImplementedError("render_mode='rgb_array' (512x512 antialiased info['rgb']) is out of scope")
        if rand_seed is None:
            rand_seed = create_random_seed()
            if shard is not None:
                # every shard of one logical VecGame must replay the SAME per-env seed chain
                # (vecgame.cpp:301-314): take rank 0's draw instead of de-correlating per rank
                rand_seed = _broadcast_seed(rand_seed)
        if resource_root is None:
            resource_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data") + os.sep

        options = dict(options)
        options.update({
            "env_name": env_name,
            "num_levels": num_levels,
            "start_level": start_level,
            "num_actions": len(self.combos),
            "use_sequential_levels": bool(use_sequential_levels),
            "debug_mode": debug_mode,
            "rand_seed": rand_seed,
            "num_threads": num_threads,
            "render_human": render_human,
            "resource_root": resource_root,
        })
        self._host_buffers = bool(host_buffers)
        self._torch = None
        if self._lib.pgb200_is_device_build():
            import torch

            self._torch = torch
            if not torch.cuda.is_available():
                raise RuntimeError("procgen_b200 needs a CUDA device (there is no CPU fallback)")
            if device is None:
                device = torch.cuda.current_device()
            device = torch.device(device).index if not isinstance(device, int) else device
            options["cuda_device"] = int(device)
        self.device_index = device
        if shard is not None:
            rank, world = shard
            options["env_index_offset"] = int(rank) * int(num)
            options["env_index_total"] = int(world) * int(num)
        self.shard = shard
        if not snap_target_rect:
            options["snap_target_rect"] = False
        self.options = options
        self.num = int(num)

        self._keep = []
        self._h = self._lib.libenv_make(self.num, L.make_options(self._keep, options))
        if not self._h:
            raise RuntimeError("libenv_make failed")

        # spaces (vecgame.cpp:212-268); the action space is unwrapped like env.py:138
        self.ob_space = DictType(rgb=TensorType(Discrete(256, "uint8"), (64, 64, 3)))
        self.ac_space = TensorType(Discrete(len(self.combos), "int32"), ())
        self._info_names = ["prev_level_seed", "prev_level_complete", "level_seed"]

        if self._host_buffers:
            self._setup_host_buffers()
        else:
            self._setup_device_buffers()

    # ------------------------------------------------------------------ buffer plumbing
    def _setup_host_buffers(self):
        n = self.num
        self._rgb = None
        if self._torch is not None:
            try:
                # page-locked from the start (cudaHostAlloc through torch): the library then DMAs straight
                # into it without having to register 12 KiB/env of pageable memory
                self._rgb_pinned = self._torch.zeros((n, 64, 64, 3), dtype=self._torch.uint8, pin_memory=True)
                self._rgb = self._rgb_pinned.numpy()
            except Exception:
                self._rgb = None
        if self._rgb is None:
            self._rgb = np.zeros((n, 64, 64, 3), np.uint8)
        self._rew = np.zeros(n, np.float32)
        self._first = np.zeros(n, np.uint8)
        self._ac = np.zeros(n, np.int32)
        self._info = {"prev_level_seed": np.zeros(n, np.int32), "prev_level_complete": np.zeros(n, np.uint8),
                      "level_seed": np.zeros(n, np.int32)}
        ob_ptrs = (C.c_void_p * n)(*[self._rgb.ctypes.data + e * 64 * 64 * 3 for e in range(n)])
        ac_ptrs = (C.c_void_p * n)(*[self._ac.ctypes.data + e * 4 for e in range(n)])
        info_ptrs = (C.c_void_p * (3 * n))()
        for si, name in enumerate(self._info_names):
            arr = self._info[name]
            for e in range(n):
                info_ptrs[si * n + e] = arr.ctypes.data + e * arr.itemsize
        self._bufs = L.Buffers(ob_ptrs, self._rew.ctypes.data_as(C.POINTER(C.c_float)),
                               self._first.ctypes.data_as(C.POINTER(C.c_uint8)), info_ptrs, ac_ptrs)
        self._keep += [ob_ptrs, ac_ptrs, info_ptrs]
        self._lib.libenv_set_buffers(self._h, C.byref(self._bufs))

    def _setup_device_buffers(self):
        torch = self._torch
        if torch is None:
            raise RuntimeError("device-resident buffers need the CUDA build")
        dev = torch.device("cuda", self.device_index)
        with torch.cuda.device(dev):
            self._stream_handle = torch.cuda.current_stream(dev).cuda_stream
            self._lib.pgb200_set_stream(self._h, C.c_void_p(self._stream_handle))
            db = L.DeviceBuffers()
            rc = self._lib.pgb200_get_device_buffers(self._h, C.byref(db))
            if rc != 0:
                raise RuntimeError("pgb200_get_device_buffers failed")
            n = self.num

            def alias(ptr, shape, typestr):
                return torch.as_tensor(_CudaArray(ptr, shape, typestr), device=dev)

            self._rgb = alias(db.rgb, (n, 64, 64, 3), "|u1")
            self._rew = alias(db.rew, (n,), "<f4")
            self._first = alias(db.first, (n,), "|u1")
            self._ac = alias(db.action, (n,), "<i4")
            self._info = {"prev_level_seed": alias(db.prev_level_seed, (n,), "<i4"),
                          "prev_level_complete": alias(db.prev_level_complete, (n,), "|u1"),
                          "level_seed": alias(db.level_seed, (n,), "<i4")}
        self._dev = dev
        self._pinned_ac = None

    # ------------------------------------------------------------------ gym3 Env surface
    def observe(self):
        """-> (rew f32[N], {"rgb": u8[N,64,64,3]}, first bool[N]); device tensors unless host_buffers."""
        if self._host_buffers:
            self._lib.libenv_observe(self._h)
            return self._rew, {"rgb": self._rgb}, self._first.astype(bool)
        return self._rew, {"rgb": self._rgb}, self._first.bool()

    def act(self, ac):
        """env.py:197-200: actions are cast to int32. Asynchronous, like VecGame::act."""
        if self._host_buffers:
            self._ac[:] = np.asarray(ac).astype(np.int32)
            self._lib.libenv_act(self._h)
            return
        torch = self._torch
        with torch.cuda.device(self._dev):
            # keep every launch on the caller's current stream
            cur = torch.cuda.current_stream(self._dev).cuda_stream
            if cur != self._stream_handle:
                self._lib.pgb200_set_stream(self._h, C.c_void_p(cur))
                self._stream_handle = cur
            if torch.is_tensor(ac) and ac.is_cuda:
                self._ac.copy_(ac.to(torch.int32), non_blocking=True)
            else:
                host = torch.as_tensor(np.asarray(ac).astype(np.int32))
                if self._pinned_ac is None:
                    # two pinned staging buffers, each guarded by an event recorded behind its H2D copy:
                    # act() never rewrites host memory a still-queued DMA is going to read
                    self._pinned_ac = [torch.empty(self.num, dtype=torch.int32, pin_memory=True) for _ in range(2)]
                    self._pinned_ev = [torch.cuda.Event(), torch.cuda.Event()]
                    self._pinned_used = [False, False]
                    self._pinned_i = 0
                i = self._pinned_i
                self._pinned_i = 1 - i
                if self._pinned_used[i]:
                    self._pinned_ev[i].synchronize()
                self._pinned_ac[i].copy_(host)
                self._ac.copy_(self._pinned_ac[i], non_blocking=True)
                self._pinned_ev[i].record(torch.cuda.current_stream(self._dev))
                self._pinned_used[i] = True
            self._lib.pgb200_act_device(self._h)

    def get_info(self) -> List[dict]:
        """gym3's list-of-dicts form (env.py:128-136). One D2H copy for all three columns; callers on
        the hot path should use get_info_tensors() (columns, no copy) instead."""
        if self._host_buffers:
            cols = [self._info[k].tolist() for k in self._info_names]
        else:
            torch = self._torch
            packed = torch.stack([self._info[k].to(torch.int32) for k in self._info_names]).cpu().numpy()
            cols = [packed[0].tolist(), packed[1].astype(np.uint8).tolist(), packed[2].tolist()]
        a, b, c3 = self._info_names
        return [{a: x, b: y, c3: z} for x, y, z in zip(*cols)]

    def get_info_tensors(self):
        """Column form of get_info() without a host copy."""
        return dict(self._info)

    def callmethod(self, method: str, *args, **kwargs):
        return getattr(self, method)(*args, **kwargs)

    def get_state(self):
        """One bytes blob per env in the reference's wire format (env.py:139-147, vecgame.cpp:437-445)."""
        import ctypes as C

        buf = C.create_string_buffer(MAX_STATE_SIZE)
        out = []
        for i in range(self.num):
            n = int(self._lib.get_state(self._h, i, buf, MAX_STATE_SIZE))
            out.append(bytes(buf.raw[:n]))
        return out

    def set_state(self, states):
        """Load one blob per env (env.py:149-153, vecgame.cpp:447-457); observations and info are refreshed."""
        assert len(states) == self.num
        for i, st in enumerate(states):
            self._lib.set_state(self._h, i, st, len(st))

    def get_combos(self):  # env.py:155-172
        return [("LEFT", "DOWN"), ("LEFT",), ("LEFT", "UP"), ("DOWN",), (), ("UP",), ("RIGHT", "DOWN"), ("RIGHT",),
                ("RIGHT", "UP"), ("D",), ("A",), ("W",), ("S",), ("Q",), ("E",)]

    def keys_to_act(self, keys_list: Sequence[Sequence[str]]) -> List[Optional[np.ndarray]]:  # env.py:174-195
        result = []
        for keys in keys_list:
            action = None
            max_len = -1
            for i, combo in enumerate(self.get_combos()):
                pressed = all(key in keys for key in combo)
                if pressed and (max_len < len(combo)):
                    action = i
                    max_len = len(combo)
            if action is not None:
                action = np.array([action])
            result.append(action)
        return result

    # ------------------------------------------------------------------ B200 extras
    def sync(self):
        self._lib.pgb200_sync(self._h)

    def errors(self) -> int:
        """OR of the per-env sticky error bits (0 = healthy)."""
        return int(self._lib.pgb200_get_errors(self._h, None))

    def kernel_launches(self) -> int:
        return int(self._lib.pgb200_kernel_launches(self._h))

    def set_launch_shape(self, chunks: int = 0, serialize: bool = False) -> None:
        """Measurement knob (see pgb200_set_launch_shape): env chunks per step, launches back to back."""
        self._lib.pgb200_set_launch_shape(self._h, int(chunks), int(bool(serialize)))

    def kernel_timing_begin(self, max_launch_pairs: int) -> None:
        """Bracket every (logic, render) kernel pair with CUDA events until kernel_timing_end()."""
        self._lib.pgb200_kernel_timing_begin(self._h, int(max_launch_pairs))

    def kernel_timing_end(self) -> dict:
        import ctypes as C

        out = (C.c_double * 5)()
        pairs = int(self._lib.pgb200_kernel_timing_end(self._h, out))
        return {"logic_ms": out[0], "render_ms": out[1], "setup_ms": out[4], "launch_pairs": pairs, "env_steps": out[3]}

    def enable_peer_gather(self, dst: int = 0) -> bool:
        """Turn gather_observations() into peer writes: allocate the gathered array as torch symmetric
        memory (every rank maps every rank's copy over NVLink), hand the library the address of this
        shard's slot in `dst`'s copy, and let every render launch be followed by a copy of its frames
        to that slot (pgb200_set_rgb_mirror) — the transfer then overlaps the rest of the step and
        gather_observations() only has to run one cross-rank barrier. Collective (call on all ranks).
        Returns False (and keeps the NCCL gather) when symmetric memory cannot be set up."""
        import torch.distributed as dist

        torch = self._torch
        world, rank = dist.get_world_size(), dist.get_rank()
        ok = torch.zeros(1, device=self._dev, dtype=torch.int32)
        try:
            import torch.distributed._symmetric_memory as symm_mem

            n = self.num
            frame = 64 * 64 * 3
            buf = symm_mem.empty((2 * world * n * frame,), dtype=torch.uint8, device=self._dev)
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            ptr = int(hdl.buffer_ptrs[dst])
            slots = [ptr + (b * world * n + rank * n) * frame for b in (0, 1)]
            self._peer = {"buf": buf, "hdl": hdl, "dst": dst, "slots": slots,
                          "views": [buf[b * world * n * frame:(b + 1) * world * n * frame].view(world * n, 64, 64, 3) for b in (0, 1)]}
            ok += 1
        except Exception as e:  # noqa: BLE001 - any failure means: keep the collective
            self._peer_error = repr(e)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) != 1:
            self._peer = None
            return False
        with torch.cuda.device(self._dev):
            self._lib.pgb200_set_rgb_mirror(self._h, C.c_void_p(self._peer["slots"][0]), C.c_void_p(self._peer["slots"][1]))
        return True

    def gather_observations(self, dst: int = 0):
        """The only cross-rank exchange on this path: every rank's rgb shard on `dst`.
        Returns u8[world*num,64,64,3] on dst, None elsewhere. After enable_peer_gather() the frames
        were already written into dst's memory behind each render launch and this is one barrier;
        otherwise it is one NCCL gather."""
        import torch.distributed as dist

        torch = self._torch
        world = dist.get_world_size()
        peer = getattr(self, "_peer", None)
        if peer is not None and peer["dst"] == dst:
            peer["hdl"].barrier(channel=0)   # on the current stream: every rank's copies of this step have landed
            if dist.get_rank() == dst:
                return peer["views"][int(self._lib.pgb200_mirror_parity(self._h))]
            return None
        if dist.get_rank() == dst:
            out = getattr(self, "_gather_buf", None)
            if out is None or out.shape[0] != world * self.num:
                out = self._gather_buf = torch.empty((world * self.num, 64, 64, 3), dtype=torch.uint8, device=self._dev)
            dist.gather(self._rgb, list(out.chunk(world, dim=0)), dst=dst)
            return out
        dist.gather(self._rgb, None, dst=dst)
        return None

    # ------------------------------------------------------------------ consumer epilogue
    def enable_consumer_output(self, dtype=None, frames: int = 1):
        """Have the render kernel also write what a learner feeds its network: rgb / 255 as float16 or
        bfloat16, planar CHW, `frames` frames stacked along the channel axis with baselines'
        VecFrameStack reset rule (an env that starts an episode sees zeros for the older frames).
        consumer_observation() then returns [num, 3*frames, 64, 64] without any further kernel."""
        torch = self._torch
        dtype = dtype or torch.float16
        code = {torch.float16: 1, torch.bfloat16: 2}[dtype]
        slots = 1 if frames == 1 else 2 * frames
        with torch.cuda.device(self._dev):
            self._consumer = torch.zeros((self.num, slots, 3, 64, 64), dtype=dtype, device=self._dev)
            self._consumer_k = int(frames)
            torch.cuda.current_stream(self._dev).synchronize()
            rc = self._lib.pgb200_set_consumer_output(self._h, C.c_void_p(self._consumer.data_ptr()), code, int(frames))
        if rc != 0:
            raise ValueError("pgb200_set_consumer_output rejected the arguments")

    def consumer_observation(self):
        """[num, 3*frames, 64, 64] view (oldest frame first) of the consumer output; valid until the next act()."""
        k = self._consumer_k
        if k == 1:
            return self._consumer[:, 0]
        s = int(self._lib.pgb200_consumer_slot(self._h))
        return self._consumer[:, s + 1:s + 1 + k].reshape(self.num, 3 * k, 64, 64)

    def gather_how(self) -> str:
        if getattr(self, "_peer", None) is not None:
            return ("peer writes: each render launch is followed by a copy of its frames into rank-0 symmetric memory over "
                    "NVLink (pgb200_set_rgb_mirror); gather = one symmetric-memory barrier per step")
        return "torch.distributed.gather (NCCL) of the rgb shard after the step" + (
            f" (peer path unavailable: {self._peer_error})" if getattr(self, "_peer_error", None) else "")

    def close(self):
        if getattr(self, "_consumer", None) is not None and getattr(self, "_h", None):
            self._lib.pgb200_set_consumer_output(self._h, None, 0, 0)
            self._consumer = None
        if getattr(self, "_peer", None) is not None and getattr(self, "_h", None):
            self._lib.pgb200_set_rgb_mirror(self._h, None, None)
            self._peer = None
        if getattr(self, "_h", None):
            self._lib.libenv_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ProcgenGym3Env(BaseProcgenEnv):
    """env.py:203-246 — same keywords and defaults."""

    def __init__(self, num, env_name, center_agent=True, use_backgrounds=True, use_monochrome_assets=False,
                 restrict_themes=False, use_generated_assets=False, paint_vel_info=False, distribution_mode="hard",
                 **kwargs):
        assert distribution_mode in DISTRIBUTION_MODE_DICT, f'"{distribution_mode}" is not a valid distribution mode.'
        if distribution_mode == "exploration":
            assert env_name in EXPLORATION_LEVEL_SEEDS, f"{env_name} does not support exploration mode"
            distribution_mode = DISTRIBUTION_MODE_DICT["hard"]
            assert "num_levels" not in kwargs, "exploration mode overrides num_levels"
            kwargs["num_levels"] = 1
            assert "start_level" not in kwargs, "exploration mode overrides start_level"
            kwargs["start_level"] = EXPLORATION_LEVEL_SEEDS[env_name]
        else:
            distribution_mode = DISTRIBUTION_MODE_DICT[distribution_mode]
        options = {
            "center_agent": bool(center_agent),
            "use_generated_assets": bool(use_generated_assets),
            "use_monochrome_assets": bool(use_monochrome_assets),
            "restrict_themes": bool(restrict_themes),
            "use_backgrounds": bool(use_backgrounds),
            "paint_vel_info": bool(paint_vel_info),
            "distribution_mode": distribution_mode,
        }
        super().__init__(num, env_name, options, **kwargs)
