"""CPU oracle for the SpinTorque-v0 hot path (NumPy restatement of the reference algorithm).

TEST INFRASTRUCTURE ONLY — never imported by the product package `spin_torque_rl_gym_b200`.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it,
and only as the checker (or as the timed CPU baseline), never as a fallback for the CUDA path.

Parity status: PINNED. `tests/golden/stt_*.npz` were produced by importing the live reference
(/root/reference, through oracle/shims) with `oracle/gen_golden.py`; `tests/test_oracle_golden.py`
checks this restatement against them bit-for-bit (FP64).

Each function cites the reference lines it restates (paths relative to /root/reference/spin_torque_gym).
The arithmetic deliberately keeps NumPy's operation order (np.cross / np.dot on 3-vectors, Python-float
scalars) so that results are bit-identical to the reference in FP64.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

GAMMA = 2.21e5                      # physics/simple_solver.py:59
MU0 = 4 * np.pi * 1e-7              # physics/simple_solver.py:60
KB_SOLVER = 1.38e-23                # physics/simple_solver.py:377 (NOT 1.380649e-23)


def default_stt_params() -> Dict:
    """envs/spin_torque_env.py:156-173 (== devices/device_factory.py:129-144)."""
    return {
        'volume': 50e-9 * 100e-9 * 2e-9,
        'area': 50e-9 * 100e-9,
        'thickness': 2e-9,
        'aspect_ratio': 2.0,
        'saturation_magnetization': 800e3,
        'damping': 0.01,
        'uniaxial_anisotropy': 1.2e6,
        'exchange_constant': 20e-12,
        'polarization': 0.7,
        'resistance_parallel': 1e3,
        'resistance_antiparallel': 2e3,
        'easy_axis': np.array([0, 0, 1]),
        'reference_magnetization': np.array([0, 0, 1]),
    }


# --------------------------------------------------------------------------------------------------
# A1  action sanitising: utils/monitoring.py:288-315 then envs/spin_torque_env.py:409-433
# --------------------------------------------------------------------------------------------------
def parse_action(action, max_current: float, max_duration: float):
    a = np.array(action, dtype=np.float32) if not isinstance(action, np.ndarray) else action
    if a.shape != (2,):
        a = np.array([0.0, 1e-12], dtype=np.float32)
    # clips happen in the caller's dtype, in place (float32 for SB3 / action_space.sample())
    a[0] = np.clip(a[0], -1e8, 1e8)
    a[1] = np.clip(a[1], 1e-12, 1e-6)
    if np.any(np.isnan(a)) or np.any(np.isinf(a)):
        a = np.array([0.0, 1e-12], dtype=np.float32)
    j = float(a[0])
    t = float(a[1])
    j = float(np.clip(j, -max_current, max_current))
    t = float(np.clip(t, 1e-12, max_duration))
    return j, t


# --------------------------------------------------------------------------------------------------
# A2  step policy: physics/simple_solver.py:137-145
# --------------------------------------------------------------------------------------------------
def substep_plan(t_end: float, max_step: float = 1e-12):
    dt = min(max_step, t_end / 100)
    n = max(10, int(t_end / dt))
    dt = t_end / n
    return n, dt


def _guard_normalise(m):
    """physics/simple_solver.py:208-229."""
    if not np.isfinite(m).all():
        return np.array([0.0, 0.0, 1.0])
    mag = np.linalg.norm(m)
    if mag < 1e-12:
        return np.array([0.0, 0.0, 1.0])
    out = m / mag
    if not np.isfinite(out).all():
        return np.array([0.0, 0.0, 1.0])
    return out


# --------------------------------------------------------------------------------------------------
# A4/A5  RHS: physics/simple_solver.py:297-388
# --------------------------------------------------------------------------------------------------
def llgs_rhs(m, cur, p, xi, h_app=None):
    """dm/dt for magnetisation m with current density `cur` (already time-gated) and noise sample xi.

    p: dict with damping, saturation_magnetization, uniaxial_anisotropy, volume, easy_axis (normalised),
       polarization, h_th (thermal field strength or 0).
    """
    alpha, ms, ku, vol, e, pol = (p['damping'], p['saturation_magnetization'], p['uniaxial_anisotropy'],
                                  p['volume'], p['easy_axis_n'], p['polarization'])
    h_applied = np.zeros(3) if h_app is None else h_app
    h_k = (2 * ku) / (MU0 * ms)                                         # :368
    h_anis = h_k * np.dot(m, e) * e                                      # :369
    h_demag = -ms * m[2] * np.array([0, 0, 1])                          # :373
    if p['h_th'] > 0 and xi is not None:
        h_thermal = p['h_th'] * xi                                       # :381
    else:
        h_thermal = np.zeros(3)
    h_eff = h_applied + h_anis + h_demag + h_thermal                    # :385
    if abs(cur) > 1e-12:                                                 # :327
        tau = (pol * cur / (ms * vol)) * np.cross(m, np.cross(m, e))     # :330-331
    else:
        tau = np.zeros(3)
    gamma_eff = GAMMA / (1 + alpha ** 2)                                 # :338
    prec = np.cross(m, h_eff)
    damp = alpha * np.cross(m, prec)
    return -gamma_eff * (prec + damp) + tau                              # :343


def thermal_strength(p, temperature: float) -> float:
    """physics/simple_solver.py:375-380."""
    if temperature <= 0:
        return 0.0
    return float(np.sqrt(2 * p['damping'] * KB_SOLVER * temperature /
                         (MU0 * p['saturation_magnetization'] * p['volume'] * GAMMA)))


def prepare_params(device_params: Dict, thermal: bool, temperature: float) -> Dict:
    p = dict(device_params)
    e = np.asarray(p.get('easy_axis', np.array([0, 0, 1])))
    p['easy_axis_n'] = e / np.linalg.norm(e)                             # :319, :356
    p.setdefault('damping', 0.01)
    p.setdefault('saturation_magnetization', 800e3)
    p.setdefault('uniaxial_anisotropy', 1e6)
    p.setdefault('volume', 1e-24)
    p.setdefault('polarization', 0.7)
    p['h_th'] = thermal_strength(p, temperature if thermal else 0.0)
    return p


# --------------------------------------------------------------------------------------------------
# A2+A6  fixed-step integration: physics/simple_solver.py:137-179, :263-295
# --------------------------------------------------------------------------------------------------
def integrate(m0, j, t_pulse, p, method='rk4', noise=None, h_app=None, return_traj=False, t_end=None):
    """Integrate over (0, t_end) with the rectangular pulse current_func(t) = j if t <= t_pulse else 0
    (envs/spin_torque_env.py:442-443). `noise` is None or an array [n, 4, 3] (rk4) / [n, 1, 3] (euler) of
    N(0,1) samples consumed in the order (substep, stage, xyz), i.e. the order in which the reference calls
    np.random.normal(0, 1, 3) once per RHS evaluation.
    Returns (m_final_row, n, ok) where m_final_row is the last trajectory row (normalised once by the
    per-substep guard) and ok mirrors RobustLLGSSolver's output validation (A3).
    """
    if t_end is None:
        t_end = t_pulse
    m = _guard_normalise(np.asarray(m0, dtype=float))                    # :119
    n, dt = substep_plan(t_end)
    t = np.linspace(0.0, t_end, n + 1)                                   # :142
    traj = np.zeros((n + 1, 3)) if return_traj else None
    if return_traj:
        traj[0] = m

    def cur(tt):
        return j if tt <= t_pulse else 0.0

    def xi(i, s):
        return None if noise is None else noise[i, s]

    for i in range(n):
        ti = t[i]
        if method == 'euler':
            m_new = m + dt * llgs_rhs(m, cur(ti), p, xi(i, 0), h_app)
        else:
            k1 = dt * llgs_rhs(m, cur(ti), p, xi(i, 0), h_app)
            k2 = dt * llgs_rhs(m + k1 / 2, cur(ti + dt / 2), p, xi(i, 1), h_app)
            k3 = dt * llgs_rhs(m + k2 / 2, cur(ti + dt / 2), p, xi(i, 2), h_app)
            k4 = dt * llgs_rhs(m + k3, cur(ti + dt), p, xi(i, 3), h_app)
            m_new = m + (k1 + 2 * k2 + 2 * k3 + k4) / 6
        m = _guard_normalise(m_new)                                      # :168
        if return_traj:
            traj[i + 1] = m
    return (traj if return_traj else m), n, True


def stt_resistance(m, p):
    """devices/stt_mram.py:78-94 (validate_magnetization renormalises first, base_device.py:94-116)."""
    mag = np.linalg.norm(m)
    mm = m / mag
    ref = np.asarray(p.get('reference_magnetization', np.array([0, 0, 1])))
    ref = ref / np.linalg.norm(ref)                                      # devices/stt_mram.py:30
    r_p = p.get('resistance_parallel', 1e3)
    r_ap = p.get('resistance_antiparallel', 2e3)
    tmr = (r_ap - r_p) / r_p
    cos_theta = np.dot(mm, ref)
    r = r_p * (1 + tmr * (1 - cos_theta) / 2)
    return max(r, r_p * 0.5)


@dataclass
class SttOracleEnv:
    """Restates SpinTorqueEnv.reset/step for action_mode='continuous', observation_mode='vector'
    (envs/spin_torque_env.py:250-407) with the sanitisation of SURVEY.md §8c (no wall-clock timeout, no
    memo caches, noise injected explicitly)."""
    device_params: Dict = field(default_factory=default_stt_params)
    max_steps: int = 100
    max_current: float = 2e6
    max_duration: float = 5e-9
    temperature: float = 300.0
    include_thermal: bool = True
    success_threshold: float = 0.9
    energy_penalty_weight: float = 0.1
    method: str = 'rk4'

    def __post_init__(self):
        self.p = prepare_params(self.device_params, self.include_thermal, self.temperature)
        self.m = None
        self.target = None
        self.step_count = 0
        self.total_energy = 0.0
        self.last_action = np.zeros(2)

    def reset(self, initial_state, target_state):
        """envs/spin_torque_env.py:280-299 with explicit initial/target state (options path)."""
        self.step_count = 0
        self.total_energy = 0.0
        self.last_action = np.zeros(2)
        m = np.asarray(initial_state, dtype=float)
        self.m = m / np.linalg.norm(m)
        t = np.asarray(target_state, dtype=float)
        self.target = t / np.linalg.norm(t)
        return self.observation()

    def observation(self):
        """envs/spin_torque_env.py:500-520 (uncached)."""
        r = stt_resistance(self.m, self.p)
        r0 = self.p.get('resistance_parallel', 1e3)
        return np.array([
            *self.m, *self.target, r / r0, self.temperature / 300.0,
            (self.max_steps - self.step_count) / self.max_steps,
            self.total_energy / 1e-12,
            self.last_action[0] / self.max_current,
            self.last_action[1] / self.max_duration,
        ], dtype=np.float32)

    def step(self, action, noise=None):
        """envs/spin_torque_env.py:328-390. `noise`: None or [n,4,3] (see integrate)."""
        j, t_pulse = parse_action(action, self.max_current, self.max_duration)
        self.last_action = np.array([j, t_pulse])
        prev_align = np.dot(self.m, self.target)
        m_row, n, ok = integrate(self.m, j, t_pulse, self.p, self.method, noise)
        if ok:
            m_new = m_row / np.linalg.norm(m_row)                         # :464
        else:
            m_new = self.m
        if abs(j) > 1e-12:                                                # :474-480
            r = stt_resistance(self.m, self.p)                            # pre-step m
            area = self.p.get('area', 1e-14)
            v = j * r * area
            energy = v ** 2 / r * t_pulse
        else:
            energy = 0.0
        self.m = m_new
        self.total_energy += energy
        self.step_count += 1
        align = np.dot(self.m, self.target)
        improvement = align - prev_align
        success = bool(align >= self.success_threshold)
        obs = self.observation()
        # CompositeReward.compute over the four default components, in dict order
        # (envs/spin_torque_env.py:184-207, rewards/composite_reward.py:88-110)
        total = 0.0
        total += 10.0 * (10.0 if success else 0.0)
        total += (-self.energy_penalty_weight) * (-energy / 1e-12)
        total += 1.0 * improvement
        total += -2.0 * 0.0
        reward = float(total)
        if math.isnan(reward) or math.isinf(reward):                      # utils/monitoring.py:342-346
            reward = -1.0
        reward = float(np.clip(reward, -1e6, 1e6))
        if np.any(np.isnan(obs)) or np.any(np.isinf(obs)):                # utils/monitoring.py:326-328
            obs = np.nan_to_num(obs, nan=0.0, posinf=1e6, neginf=-1e6)
        terminated = success
        truncated = self.step_count >= self.max_steps
        info = dict(n_sub=n, energy=energy, alignment=align, j=j, t=t_pulse)
        return obs, reward, terminated, truncated, info
