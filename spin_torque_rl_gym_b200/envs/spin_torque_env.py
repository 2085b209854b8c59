"""Single-environment façade with the reference's gym.Env API (spin_torque_gym/envs/spin_torque_env.py:26-745) on top of the
batched CUDA env (num_envs=1), plus `make()` with the reference's ids, so `gym.make('SpinTorque-v0', device_type=..., ...)`
call sites can switch by changing the import."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .. import _lib
from .array_env import SpinTorqueArrayVectorEnv
from .vector_env import SpinTorqueVectorEnv

try:  # pragma: no cover
    import gymnasium as _gym
    _EnvBase = _gym.Env
except Exception:  # noqa: BLE001
    _gym = None
    _EnvBase = object


class SpinTorqueEnv(_EnvBase):
    """One SpinTorque-v0 environment. Same kwargs as the reference; extra kwargs (device, dtype, integrator, rng_seed) are
    forwarded to SpinTorqueVectorEnv. step() returns NumPy / Python scalars and the reference's info keys."""

    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 30}

    def __init__(self, device_type: str = "stt_mram", device_params: Optional[Dict[str, Any]] = None, **kwargs):
        seed = kwargs.get("seed")
        kwargs.setdefault("autoreset", False)
        kwargs.setdefault("dtype", _lib.require_cuda().float64)       # single env: bit-faithful FP64 stages by default
        self._vec = SpinTorqueVectorEnv(num_envs=1, device_type=device_type, device_params=device_params, **kwargs)
        v = self._vec
        self.device_type = device_type
        self.max_steps, self.max_current, self.max_duration = v.max_steps, v.max_current, v.max_duration
        self.temperature, self.include_thermal = v.temperature, v.include_thermal
        self.action_mode, self.observation_mode = v.action_mode, v.observation_mode
        self.success_threshold, self.energy_penalty_weight = v.success_threshold, v.energy_penalty_weight
        self.render_mode = kwargs.get("render_mode")
        self.target_states = v.target_states
        self.action_space = v.single_action_space
        self.observation_space = v.single_observation_space
        self.episode_history: List[Dict[str, Any]] = []
        self._seed = seed
        self._np_random = np.random.default_rng(seed)
        self.current_magnetization = None
        self.target_magnetization = None
        self.step_count = 0
        self.total_energy = 0.0
        self.last_action = np.zeros(2)

    def seed(self, seed: Optional[int] = None) -> List[Optional[int]]:
        self._seed = seed
        self._np_random = np.random.default_rng(seed)
        return [seed]

    def _sync_state(self):
        v = self._vec
        self.current_magnetization = v.magnetization[0].cpu().numpy()
        self.target_magnetization = v.target[0].cpu().numpy()
        self.step_count = int(v._step_count[0])
        self.total_energy = float(v._total_energy[0])
        self.last_action = v._last_action[:, 0].cpu().numpy()

    def _get_info(self) -> Dict[str, Any]:
        align = float(np.dot(self.current_magnetization, self.target_magnetization))
        return {"step_count": self.step_count, "total_energy": self.total_energy, "current_alignment": align,
                "is_success": align >= self.success_threshold, "target_reached": align >= self.success_threshold,
                "magnetization_magnitude": float(np.linalg.norm(self.current_magnetization)),
                "device_type": self.device_type, "episode_history": self.episode_history.copy()}

    def reset(self, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None):
        """(obs[12] f32, info). Without options the start state / target come from the env's Philox stream keyed by `seed`."""
        obs, _ = self._vec.reset(seed=seed, options=options)
        self.episode_history = []
        self._sync_state()
        return obs[0].cpu().numpy(), self._get_info()

    def step(self, action) -> Tuple[np.ndarray, float, bool, bool, Dict[str, Any]]:
        if self.current_magnetization is None:
            raise RuntimeError("Environment must be reset before calling step")
        a = np.asarray(action, dtype=np.float32).reshape(-1)
        if a.shape != (2,):                                    # SafetyWrapper.validate_action (utils/monitoring.py:297-299)
            a = np.array([0.0, 1e-12], dtype=np.float32)
        prev_align = float(np.dot(self.current_magnetization, self.target_magnetization))
        obs, rew, term, trunc, inf = self._vec.step(a[None])
        self._sync_state()
        align = float(np.dot(self.current_magnetization, self.target_magnetization))
        reward = float(rew[0])
        info = self._get_info()
        info.update({"final_magnetization": self.current_magnetization.copy(),
                     "energy_consumed": float(inf["step_energy"][0]), "pulse_duration": float(self.last_action[1]),
                     "current_density": float(self.last_action[0]), "simulation_success": int(inf["status"][0]) == 0,
                     "is_success": align >= self.success_threshold, "step_energy": float(inf["step_energy"][0]),
                     "alignment_improvement": align - prev_align, "current_alignment": align})
        self.episode_history.append({"step": self.step_count, "action": [float(self.last_action[0]), float(self.last_action[1])],
                                     "magnetization": self.current_magnetization.copy(), "reward": reward,
                                     "energy": info["energy_consumed"], "alignment": align})
        return obs[0].cpu().numpy(), reward, bool(term[0]), bool(trunc[0]), info

    def analyze_episode(self) -> Dict[str, Any]:
        """envs/spin_torque_env.py:720-745."""
        if not self.episode_history:
            return {}
        total_energy = sum(h["energy"] for h in self.episode_history)
        final_alignment = self.episode_history[-1]["alignment"]
        switching_step = next((i + 1 for i, h in enumerate(self.episode_history)
                               if h["alignment"] >= self.success_threshold), None)
        return {"episode_length": len(self.episode_history), "total_energy": total_energy,
                "final_alignment": final_alignment, "success": final_alignment >= self.success_threshold,
                "switching_step": switching_step,
                "average_reward": float(np.mean([h["reward"] for h in self.episode_history])),
                "energy_efficiency": final_alignment / total_energy if total_energy > 0 else 0,
                "history": self.episode_history.copy()}

    def get_device_info(self) -> Dict[str, Any]:
        p = self._vec.device_params
        return {"device_type": self.device_type, "volume": p.get("volume"), "thickness": p.get("thickness", 1e-9),
                "saturation_magnetization": p.get("saturation_magnetization"), "parameters": dict(p)}

    def render(self, mode: Optional[str] = None):
        return None

    def close(self):
        self._vec.close()


class SpinTorqueArrayEnv(_EnvBase):
    """One SpinTorqueArray-v0 environment (envs/array_env.py:20-755) on the K3 kernel."""

    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 10}

    def __init__(self, array_size=(4, 4), **kwargs):
        kwargs.setdefault("autoreset", False)
        self._vec = SpinTorqueArrayVectorEnv(num_envs=1, array_size=array_size, **kwargs)
        v = self._vec
        self.array_size, self.n_rows, self.n_cols, self.n_devices = v.array_size, v.n_rows, v.n_cols, v.n_devices
        self.action_space, self.observation_space = v.single_action_space, v.single_observation_space
        self.coupling_matrix = v.coupling_matrix
        self.target_pattern = v.target_pattern
        self.max_steps, self.success_threshold = v.max_steps, v.success_threshold
        self.current_pattern = None
        self.step_count, self.total_energy = 0, 0.0

    def _sync(self):
        v = self._vec
        self.current_pattern = v.current_pattern[0].cpu().numpy()
        self.step_count, self.total_energy = int(v._step_count[0]), float(v._total_energy[0])

    def reset(self, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None):
        obs, _ = self._vec.reset(seed=seed, options=options)
        self._sync()
        sim = float(np.mean(np.sum(self.current_pattern * self._vec._target[0].cpu().numpy().reshape(self.current_pattern.shape), -1)))
        return obs[0].cpu().numpy(), {"step_count": 0, "total_energy": 0.0, "pattern_similarity": sim,
                                      "is_success": sim >= self.success_threshold, "array_size": self.array_size}

    def step(self, action):
        if self.current_pattern is None:
            raise RuntimeError("Environment must be reset before calling step")
        a = np.asarray(action, dtype=np.float32).reshape(1, -1)
        obs, rew, term, trunc, inf = self._vec.step(a)
        self._sync()
        sim = float(inf["pattern_similarity"][0])
        info = {"step_count": self.step_count, "total_energy": self.total_energy, "pattern_similarity": sim,
                "is_success": sim >= self.success_threshold, "array_size": self.array_size,
                "energy_consumed": float(inf["step_energy"][0]), "step_energy": float(inf["step_energy"][0])}
        return obs[0].cpu().numpy(), float(rew[0]), bool(term[0]), bool(trunc[0]), info

    def close(self):
        self._vec.close()


# ---- registry with the reference's ids (spin_torque_gym/envs/__init__.py:14-33) ----------------------------------------------
_REGISTRY = {
    "SpinTorque-v0": (SpinTorqueEnv, {"device_type": "stt_mram"}),
    "SpinTorqueArray-v0": (SpinTorqueArrayEnv, {"array_size": (4, 4), "device_type": "stt_mram"}),
}


def make(env_id: str, num_envs: Optional[int] = None, **kwargs):
    """`make('SpinTorque-v0', **reference_kwargs)` -> single env with the reference API;
    `make('SpinTorque-v0', num_envs=N, ...)` -> SpinTorqueVectorEnv / SpinTorqueArrayVectorEnv with N envs on the GPU."""
    if env_id not in _REGISTRY:
        raise KeyError(f"unknown environment id {env_id!r}; available: {sorted(_REGISTRY)}")
    cls, defaults = _REGISTRY[env_id]
    kw = dict(defaults)
    kw.update(kwargs)
    if num_envs is None:
        return cls(**kw)
    if env_id == "SpinTorque-v0":
        return SpinTorqueVectorEnv(num_envs=num_envs, **kw)
    return SpinTorqueArrayVectorEnv(num_envs=num_envs, **kw)


def register_with_gymnasium() -> bool:
    """Register the ids with gymnasium when it is installed (entry points of this package)."""
    if _gym is None:
        return False
    from gymnasium.envs.registration import register
    register(id="SpinTorque-v0", entry_point="spin_torque_rl_gym_b200.envs.spin_torque_env:SpinTorqueEnv",
             max_episode_steps=100, kwargs={"device_type": "stt_mram"})
    register(id="SpinTorqueArray-v0", entry_point="spin_torque_rl_gym_b200.envs.spin_torque_env:SpinTorqueArrayEnv",
             max_episode_steps=200, kwargs={"array_size": (4, 4), "device_type": "stt_mram"})
    return True
