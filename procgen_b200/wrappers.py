"""Caller-side adapters of the reference (procgen/env.py:249-265, procgen/gym_registration.py:6-34):
``ProcgenEnv`` / ``ToBaselinesVecEnv`` (the baselines VecEnv surface train-procgen drives) and the
``procgen-<name>-v0`` single-env gym interface. gym3 and gym are not installable in this image, so
the two gym3 adapters the reference composes (``gym3.ToBaselinesVecEnv``, ``gym3.ToGymEnv`` over
``ExtractDictObWrapper(key="rgb")``) are restated here with the same call semantics, and the space
objects are minimal stand-ins with gym's attribute names (``shape``, ``dtype``, ``n``, ``low``, ``high``).
"""
from __future__ import annotations

import numpy as np

from .env import ENV_NAMES, ProcgenGym3Env


class Box:
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


class DiscreteSpace:
    def __init__(self, n):
        self.n, self.shape, self.dtype = int(n), (), np.dtype(np.int32)

    def __repr__(self):
        return f"Discrete({self.n})"


class DictSpace(dict):
    @property
    def spaces(self):
        return self


def _to_numpy(x):
    return x if isinstance(x, np.ndarray) else x.detach().cpu().numpy()


class ToBaselinesVecEnv:
    """gym3.ToBaselinesVecEnv + the reference's render() override (env.py:249-261).

    ``reset()`` returns the current observation (gym3 envs reset themselves), ``step_async`` = ``act``,
    ``step_wait`` = ``observe`` -> ``(obs, rews, dones, infos)`` with ``dones`` = gym3's ``first``.
    Observations are numpy dicts by default (what baselines' wrappers expect); ``device_tensors=True``
    keeps the ``torch.cuda`` aliases of the library's HBM buffers instead (no copy)."""

    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 15}
    reward_range = (-float("inf"), float("inf"))
    spec = None

    def __init__(self, env, device_tensors=False):
        self.env = env
        self.num_envs = env.num
        self.device_tensors = bool(device_tensors)
        self.observation_space = DictSpace(rgb=Box(0, 255, (64, 64, 3), np.uint8))
        self.action_space = DiscreteSpace(env.ac_space.eltype.n)
        self._closed = False

    def _convert(self, rew, ob, first):
        if self.device_tensors:
            return dict(ob), rew, first
        return {k: _to_numpy(v) for k, v in ob.items()}, _to_numpy(rew), _to_numpy(first).astype(bool)

    def reset(self):
        rew, ob, first = self.env.observe()
        ob, _, first = self._convert(rew, ob, first)
        return ob

    def step_async(self, actions):
        self.env.act(actions)

    def step_wait(self):
        rew, ob, first = self.env.observe()
        ob, rew, first = self._convert(rew, ob, first)
        return ob, rew, first, self.env.get_info()

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def render(self, mode="human"):
        info = self.env.get_info()[0]
        _, ob, _ = self.env.observe()
        if mode == "rgb_array":
            if "rgb" in info:
                return info["rgb"]
            return _to_numpy(ob["rgb"][0])

    def seed(self, seed=None):
        raise NotImplementedError("procgen environments are seeded at construction (rand_seed=...)")

    def close(self):
        if not self._closed:
            self._closed = True
            self.env.close()

    @property
    def unwrapped(self):
        return self


def ProcgenEnv(num_envs, env_name, **kwargs):
    """env.py:264-265."""
    device_tensors = kwargs.pop("device_tensors", False)
    return ToBaselinesVecEnv(ProcgenGym3Env(num=num_envs, env_name=env_name, **kwargs), device_tensors=device_tensors)


class ToGymEnv:
    """gym3.ToGymEnv(ExtractDictObWrapper(env, key="rgb")) for a 1-env ProcgenGym3Env: the classic
    ``reset() -> ob`` / ``step(ac) -> (ob, rew, done, info)`` interface. As in gym3, an episode end
    is reported with the first observation of the next episode (the env resets itself)."""

    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 15}
    reward_range = (-float("inf"), float("inf"))

    def __init__(self, env):
        assert env.num == 1
        self.env = env
        self.observation_space = Box(0, 255, (64, 64, 3), np.uint8)
        self.action_space = DiscreteSpace(env.ac_space.eltype.n)

    def reset(self):
        _, ob, _ = self.env.observe()
        return _to_numpy(ob["rgb"])[0]

    def step(self, ac):
        self.env.act(np.array([ac], dtype=np.int32))
        rew, ob, first = self.env.observe()
        return _to_numpy(ob["rgb"])[0], float(_to_numpy(rew)[0]), bool(_to_numpy(first)[0]), self.env.get_info()[0]

    def render(self, mode="rgb_array"):
        _, ob, _ = self.env.observe()
        return _to_numpy(ob["rgb"])[0]

    def close(self):
        self.env.close()


def make_env(render_mode=None, render=False, **kwargs):
    """gym_registration.py:6-26. Human rendering (gym3's ViewerWrapper window) and the 512x512
    rgb_array info are out of scope here: both raise through ProcgenGym3Env."""
    if render:
        render_mode = "human"
    if render_mode == "human":
        render_mode = "rgb_array"
    kwargs["render_mode"] = render_mode
    env = ProcgenGym3Env(num=1, num_threads=0, **kwargs)
    return ToGymEnv(env)


REGISTRY = {}


def register_environments():
    """gym_registration.py:29-34: ids ``procgen-<name>-v0``. Registered with gym / gymnasium when one
    is importable, and always in this module's REGISTRY (``procgen_b200.make(id, **kwargs)``)."""
    for env_name in ENV_NAMES:
        REGISTRY[f"procgen-{env_name}-v0"] = {"env_name": env_name}
    for modname in ("gym", "gymnasium"):
        try:
            reg = __import__(modname + ".envs.registration", fromlist=["register"]).register
        except Exception:
            continue
        for env_id, kw in REGISTRY.items():
            try:
                reg(id=env_id, entry_point="procgen_b200.wrappers:make_env", kwargs=dict(kw))
            except Exception:
                pass


def make(env_id, **kwargs):
    kw = dict(REGISTRY[env_id])
    kw.update(kwargs)
    return make_env(**kw)
