"""Stable-Baselines3 `VecEnv` adapter and a GPU-resident rollout collector for SpinTorqueVectorEnv.

SB3 is not installed in the build image; the adapter implements the documented VecEnv protocol (num_envs, observation_space,
action_space, reset, step_async/step_wait/step, close, env_is_wrapped, get_attr/set_attr/env_method, seed) with NumPy in/out and
same-step auto-reset, `infos[i]['terminal_observation']` and `infos[i]['TimeLimit.truncated']` like SB3's DummyVecEnv, and
subclasses stable_baselines3's VecEnv when it can be imported, so `PPO('MlpPolicy', SB3VecEnvAdapter(env))` works unchanged.

For env counts where SB3's host-side rollout buffer cannot exist (n_steps=2048 x 1M envs x 12 floats = 100 TB), RolloutCollector
keeps a [T, N, ...] ring of observations / actions / rewards / dones on the GPU with SB3-shaped tensors."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np

from .. import _lib
from ..parallel import all_reduce_stats

try:  # pragma: no cover - stable_baselines3 is not in the build image
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
except Exception:  # noqa: BLE001
    _VecEnvBase = object


class SB3VecEnvAdapter(_VecEnvBase):
    def __init__(self, env):
        if not env.autoreset:
            raise ValueError("the SB3 adapter needs an env constructed with autoreset=True")
        self.env = env
        if _VecEnvBase is not object:          # the real SB3 base class: let it set num_envs / spaces / render bookkeeping
            super().__init__(env.num_envs, env.single_observation_space, env.single_action_space)
        self.num_envs = env.num_envs
        self.observation_space = env.single_observation_space
        self.action_space = env.single_action_space
        self.render_mode = None
        self._actions = None
        self.reset_infos: List[Dict[str, Any]] = [{} for _ in range(self.num_envs)]
        self._seed = None

    def reset(self) -> np.ndarray:
        obs, _ = self.env.reset(seed=self._seed)
        self._seed = None
        obs_np = obs.cpu().numpy()
        # host_outputs env: the tensor IS the pinned buffer the next step overwrites, and SB3 keeps reset()'s array as _last_obs
        return obs_np.copy() if obs.device.type == "cpu" else obs_np

    def seed(self, seed: Optional[int] = None) -> List[Optional[int]]:
        self._seed = seed
        return [None if seed is None else seed + i for i in range(self.num_envs)]

    def step_async(self, actions: np.ndarray) -> None:
        self._actions = np.asarray(actions, dtype=np.float32)

    def step_wait(self):
        obs, rew, term, trunc, info = self.env.step(self._actions)
        obs_np = obs.cpu().numpy()
        if obs.device.type == "cpu":          # host_outputs env: the tensor IS the pinned buffer the next step overwrites
            obs_np = obs_np.copy()
        term_np, trunc_np = term.cpu().numpy(), trunc.cpu().numpy()
        dones = term_np | trunc_np
        infos: List[Dict[str, Any]] = [{} for _ in range(self.num_envs)]
        if dones.any():
            fin = info["final_observation"].cpu().numpy()
            for i in np.nonzero(dones)[0]:
                infos[i]["terminal_observation"] = fin[i].copy()
                infos[i]["TimeLimit.truncated"] = bool(trunc_np[i] and not term_np[i])
                infos[i]["is_success"] = bool(term_np[i])
        return obs_np, rew.cpu().numpy().astype(np.float32), dones, infos

    def step(self, actions: np.ndarray):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        self.env.close()

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        n = self.num_envs if indices is None else len(list(indices))
        return [getattr(self.env, attr_name)] * n

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        setattr(self.env, attr_name, value)

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> List[Any]:
        n = self.num_envs if indices is None else len(list(indices))
        return [getattr(self.env, method_name)(*args, **kwargs)] * n

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        n = self.num_envs if indices is None else len(list(indices))
        return [False] * n

    def get_images(self) -> Sequence[Optional[np.ndarray]]:
        return [None] * self.num_envs


class RolloutCollector:
    """GPU-resident rollout of `n_steps` env steps: everything stays in HBM, one kernel launch per step plus the policy.
    `policy(obs[N,12] f32 cuda) -> actions[N,2] f32 cuda` (any torch module / callable)."""

    def __init__(self, env, n_steps: int, store_observations: bool = True):
        torch = _lib.require_cuda()
        self.env, self.n_steps = env, int(n_steps)
        N, dev = env.num_envs, env.device
        self.store_observations = store_observations
        if store_observations:
            self.observations = torch.empty(self.n_steps, N, _lib.OBS_DIM, dtype=torch.float32, device=dev)
        self.actions = torch.empty(self.n_steps, N, 2, dtype=torch.float32, device=dev)
        self.rewards = torch.empty(self.n_steps, N, dtype=torch.float32, device=dev)
        self.dones = torch.empty(self.n_steps, N, dtype=torch.bool, device=dev)
        self._last_obs = None

    def collect(self, policy: Callable) -> Dict[str, Any]:
        env = self.env
        if self._last_obs is None:
            self._last_obs, _ = env.reset()
        obs = self._last_obs
        for t in range(self.n_steps):
            if self.store_observations:
                self.observations[t].copy_(obs)
            act = policy(obs)
            self.actions[t].copy_(act)
            obs, rew, term, trunc, _ = env.step(self.actions[t])
            self.rewards[t].copy_(rew)
            self.dones[t].copy_(term | trunc)
        self._last_obs = obs
        return all_reduce_stats(env.stats_tensor())       # one collective per rollout: episode statistics (K5)
