"""procgen_b200 — B200-native vectorised Procgen hot path behind the reference's ProcgenGym3Env API.

Mirrors procgen/__init__.py: the public names are ProcgenGym3Env and ENV_NAMES-style helpers.
Importing the package never touches CUDA; constructing an env requires the sm_100a library and a
GPU (there is no CPU fallback).
"""
from .env import ProcgenGym3Env, BaseProcgenEnv, ENV_NAMES, EXPLORATION_LEVEL_SEEDS, DISTRIBUTION_MODE_DICT  # noqa: F401
from .wrappers import ProcgenEnv, ToBaselinesVecEnv, ToGymEnv, make, make_env, register_environments  # noqa: F401

register_environments()  # procgen/__init__.py:10

__all__ = ["ProcgenEnv", "ProcgenGym3Env", "BaseProcgenEnv", "ENV_NAMES", "ToBaselinesVecEnv", "make"]
