"""`torch.library` registration of the path's entry points ("optionally registered as torch.library custom ops so
torch.compile graphs don't break"). Importing this module defines the `stg::` operator namespace:

    torch.ops.stg.vec3_cross(a, b)            [N,3] x [N|1,3] -> [N,3]     (K4, NumPy-ordered FP64, as VectorizedMagneticsOperations)
    torch.ops.stg.vec3_dot(a, b)              -> [N]
    torch.ops.stg.vec3_normalize(a)           -> [N,3]
    torch.ops.stg.stt_env_step(handle, action, m, target, total_energy, last_action, step_count, episode)
                                              -> (obs [N,12] f32, reward [N] f64, terminated [N] bool, truncated [N] bool)   (K1)

Every op is CUDA-only (`device_types="cuda"`: there is no CPU kernel to dispatch to) and carries a fake (meta) implementation, so
`torch.compile(..., fullgraph=True)` traces through a rollout step without a graph break; the compiled graph calls the same C-ABI
entry points as the eager classes (this module adds no kernels and no torch.compile-generated code to the path).

`stt_env_step` is the functional face of `SpinTorqueVectorEnv.step`: the env's FP64 state planes are passed explicitly and declared
as mutated, which is what keeps successive steps ordered inside a traced graph (the kernel works on the planes it is given:
normally `state_tensors(env)`, under functionalisation their copies); `handle` (an int from `register_env`) names the
env whose parameter table, RNG seed and output buffers the launch uses. The returned tensors are fresh copies (a custom op may
not return views of persistent buffers); the zero-copy path is `env.step` itself.
"""
from __future__ import annotations

import weakref
from typing import Tuple

import torch
from torch.library import custom_op

from .physics.vectorized_operations import _vec3_op

_ENVS: "weakref.WeakValueDictionary[int, object]" = weakref.WeakValueDictionary()
_NEXT = [1]


def _rows(a: torch.Tensor) -> int:
    return a.numel() // 3


@custom_op("stg::vec3_cross", mutates_args=(), device_types="cuda")
def vec3_cross(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return _vec3_op("cross", a, b)


@vec3_cross.register_fake
def _(a, b):
    return a.new_empty((_rows(a), 3), dtype=torch.float64)


@custom_op("stg::vec3_dot", mutates_args=(), device_types="cuda")
def vec3_dot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return _vec3_op("dot", a, b)


@vec3_dot.register_fake
def _(a, b):
    return a.new_empty((_rows(a),), dtype=torch.float64)


@custom_op("stg::vec3_normalize", mutates_args=(), device_types="cuda")
def vec3_normalize(a: torch.Tensor) -> torch.Tensor:
    return _vec3_op("normalize", a)


@vec3_normalize.register_fake
def _(a):
    return a.new_empty((_rows(a), 3), dtype=torch.float64)


# ---- K1 env step ---------------------------------------------------------------------------------------------------------------
def register_env(env) -> int:
    """Give `env` (a SpinTorqueVectorEnv with device outputs) an integer handle for `torch.ops.stg.stt_env_step`. The registry
    holds a weak reference: the handle dies with the env."""
    if getattr(env, "host_outputs", False):
        raise ValueError("stg::stt_env_step returns device tensors: build the env with host_outputs=False")
    h = getattr(env, "_torch_op_handle", None)
    if h is None or _ENVS.get(h) is not env:
        h = _NEXT[0]
        _NEXT[0] += 1
        _ENVS[h] = env
        env._torch_op_handle = h
    return h


def state_tensors(env) -> Tuple[torch.Tensor, ...]:
    """The env's state planes in the argument order of `stg::stt_env_step`."""
    return (env._m, env._target, env._total_energy, env._last_action, env._step_count, env._episode)


@custom_op("stg::stt_env_step",
           mutates_args=("m", "target", "total_energy", "last_action", "step_count", "episode"), device_types="cuda")
def stt_env_step(handle: int, action: torch.Tensor, m: torch.Tensor, target: torch.Tensor, total_energy: torch.Tensor,
                 last_action: torch.Tensor, step_count: torch.Tensor, episode: torch.Tensor
                 ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    env = _ENVS.get(handle)
    if env is None:
        raise RuntimeError(f"stg::stt_env_step: no live env behind handle {handle} (register_env)")
    given = (m, target, total_energy, last_action, step_count, episode)
    for g, own in zip(given, state_tensors(env)):
        if g.shape != own.shape or g.dtype != own.dtype or g.device != own.device or not g.is_contiguous():
            raise RuntimeError("stg::stt_env_step: state tensors must have the layout of state_tensors(env) "
                               f"(got {tuple(g.shape)} {g.dtype} on {g.device}, expected {tuple(own.shape)} {own.dtype} on {own.device})")
    # The launch reads and writes the planes it is GIVEN (under functionalisation these are copies of the env's own planes that
    # the graph copies back afterwards), so the state pointers of the env's argument block are swapped for the call.
    a = env._step_args()
    saved = [getattr(a.state, f) for f, _ in a.state._fields_]
    for (f, _), g in zip(a.state._fields_, given):
        setattr(a.state, f, g.data_ptr())
    try:
        obs, reward, terminated, truncated, _ = env.step(action)
        return obs.clone(), reward.clone(), terminated.clone(), truncated.clone()
    finally:
        for (f, _), v in zip(a.state._fields_, saved):
            setattr(a.state, f, v)


@stt_env_step.register_fake
def _(handle, action, m, target, total_energy, last_action, step_count, episode):
    n = action.shape[0]
    return (action.new_empty((n, 12), dtype=torch.float32), action.new_empty((n,), dtype=torch.float64),
            action.new_empty((n,), dtype=torch.bool), action.new_empty((n,), dtype=torch.bool))


def env_step(env, action: torch.Tensor):
    """`torch.ops.stg.stt_env_step` on `env` (registered on first use): traceable equivalent of `env.step(action)[:4]`."""
    return torch.ops.stg.stt_env_step(register_env(env), action, *state_tensors(env))
