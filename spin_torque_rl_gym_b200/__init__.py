"""spin_torque_rl_gym_b200 — B200-native batched LLGS hot path of Spin-Torque RL-Gym.

Host layer (Python/PyTorch) mirroring the reference's env / solver / device API for the per-step
Landau-Lifshitz-Gilbert-Slonczewski integration, on top of hand-written sm_100a CUDA kernels behind the C-ABI of
include/stg.h (libstg.so). See DESIGN.md.
"""
__version__ = "0.1.0"

from . import _lib, build, params  # noqa: F401
from .envs import SpinTorqueArrayVectorEnv, SpinTorqueVectorEnv  # noqa: F401
