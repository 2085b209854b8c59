"""CPU oracle for SpinTorqueArray-v0 (NumPy restatement of envs/array_env.py). TEST INFRASTRUCTURE ONLY.

Parity status: PINNED — bit-exact against tests/golden/array_env.npz (live reference, oracle/gen_golden.py array).
Restates: _compute_coupling_matrix :289-318, step :358-411, _apply_action :413-476 (incl. the `global` mode reading [J, T] as
[idx, J] and the resistance taken from the UPDATED magnetisation through the NumPy view), _compute_effective_field :478-495,
_simulate_device_dynamics :497-531, _compute_pattern_similarity :533-541, rewards :182-221, _get_observation :543-551.
"""
from __future__ import annotations

import numpy as np

from .stt_oracle import default_stt_params, stt_resistance
from . import devices_oracle as DO

GAMMA = 2.21e5
MU0 = 4 * np.pi * 1e-7


def coupling_matrix(n_rows, n_cols, strength=0.1, kind='dipolar'):
    n = n_rows * n_cols
    c = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            ir, ic = divmod(i, n_cols)
            jr, jc = divmod(j, n_cols)
            d = np.sqrt((ir - jr) ** 2 + (ic - jc) ** 2)
            if kind == 'dipolar':
                if d > 0:
                    c[i, j] = strength / (d ** 3)
            elif kind == 'exchange':
                if d == 1:
                    c[i, j] = strength
            elif kind == 'stray_field':
                if d > 0:
                    c[i, j] = strength / (d ** 2)
    return c


def checkerboard(n_rows, n_cols):
    p = np.zeros((n_rows, n_cols, 3))
    for i in range(n_rows):
        for j in range(n_cols):
            p[i, j] = [0, 0, 1] if (i + j) % 2 == 0 else [0, 0, -1]
    return p


class ArrayOracleEnv:
    def __init__(self, array_size=(4, 4), device_type='stt_mram', device_params=None, max_steps=200, max_current=2e6,
                 max_duration=5e-9, include_coupling=True, coupling_strength=0.1, coupling_type='dipolar',
                 action_mode='individual', success_threshold=0.9, energy_penalty_weight=0.1):
        self.n_rows, self.n_cols = array_size
        self.n_devices = self.n_rows * self.n_cols
        self.device_type = device_type
        self.p = dict(default_stt_params() if device_params is None else device_params)
        self.max_steps, self.max_current, self.max_duration = max_steps, max_current, max_duration
        self.include_coupling = include_coupling
        self.action_mode = action_mode
        self.success_threshold, self.energy_penalty_weight = success_threshold, energy_penalty_weight
        self.coupling = coupling_matrix(self.n_rows, self.n_cols, coupling_strength, coupling_type) if include_coupling else None
        self.target = checkerboard(self.n_rows, self.n_cols)
        self.pattern = None
        self.step_count = 0
        self.total_energy = 0.0

    def reset(self, initial_pattern, target_pattern=None):
        self.step_count, self.total_energy = 0, 0.0
        self.pattern = np.array(initial_pattern, dtype=float).copy()
        if target_pattern is not None:
            self.target = np.array(target_pattern, dtype=float).copy()
        return self.observation()

    def observation(self):
        return np.concatenate([self.pattern, self.target], axis=2).astype(np.float32)

    def similarity(self, pattern):
        s = []
        for i in range(self.n_rows):
            for j in range(self.n_cols):
                s.append(np.dot(pattern[i, j], self.target[i, j]))
        return np.mean(s)

    def _field(self, idx, m):
        if self.device_type == 'stt_mram':
            mm = m / np.linalg.norm(m)
            e = np.asarray(self.p.get('easy_axis', np.array([0, 0, 1])))
            h = np.zeros(3).astype(float).copy()
            h += (2 * self.p.get('uniaxial_anisotropy', 1e6) / (MU0 * self.p.get('saturation_magnetization', 800e3))) \
                * np.dot(mm, e) * e
        else:
            h = DO.effective_field(self.device_type, self.p, m, np.zeros(3))
        hc = np.zeros(3)
        if self.include_coupling:
            for j in range(self.n_devices):
                if j != idx:
                    jr, jc = divmod(j, self.n_cols)
                    hc += self.coupling[idx, j] * self.pattern[jr, jc]
        return h + hc

    @staticmethod
    def _dynamics(m0, cur, dur, h):
        if abs(cur) > 1e-12:
            p_hat = np.array([0, 0, 1])
            tau = 0.1 * cur * np.cross(m0, np.cross(m0, p_hat))
            dm = -GAMMA * np.cross(m0, h)
            dm += 0.01 * np.cross(m0, dm)
            dm += tau
            dt = dur / 10
            m = m0.copy()
            for _ in range(10):
                m += dm * dt
                m = m / np.linalg.norm(m)
            return m
        return m0

    def _resistance(self, m):
        if self.device_type == 'stt_mram':
            return stt_resistance(m, self.p)
        return float(DO.resistance(self.device_type, self.p, m))

    def step(self, action):
        prev_sim = self.similarity(self.pattern.copy())
        cur = float(action[1]) if len(action) > 1 else 0.0
        dur = float(action[2]) if len(action) > 2 else 1e-9
        cur = np.clip(cur, -self.max_current, self.max_current)
        dur = np.clip(dur, 1e-12, self.max_duration)
        nd, nc, nr = self.n_devices, self.n_cols, self.n_rows
        if self.action_mode == 'individual':
            aff = [int(np.clip(action[0], 0, nd - 1))]
        elif self.action_mode == 'row':
            r = int(np.clip(action[0], 0, nr - 1))
            aff = list(range(r * nc, (r + 1) * nc))
        elif self.action_mode == 'column':
            c = int(np.clip(action[0], 0, nc - 1))
            aff = list(range(c, nd, nc))
        else:
            aff = list(range(nd))
        energy_total = 0.0
        for idx in aff:
            r, c = divmod(idx, nc)
            m = self.pattern[r, c]                       # a view, like the reference
            h = self._field(idx, m)
            final = self._dynamics(m, cur, dur, h)
            self.pattern[r, c] = final
            res = self._resistance(m)                    # view => resistance of the UPDATED magnetisation
            area = self.p.get('area', 1e-14)
            if abs(cur) > 1e-12:
                v = cur * res * area
                energy_total += v ** 2 / res * dur
        self.total_energy += energy_total
        self.step_count += 1
        sim = self.similarity(self.pattern)
        improvement = sim - prev_sim
        success = bool(sim >= self.success_threshold)
        mags = np.linalg.norm(self.pattern, axis=2)
        total = 0.0
        total += 10.0 * (10.0 if success else sim * 5.0)
        total += (-self.energy_penalty_weight) * (-energy_total / 1e-12)
        total += 1.0 * improvement
        total += 2.0 * max(0, 1.0 - np.std(mags))
        return self.observation(), float(total), success, self.step_count >= self.max_steps, \
            dict(energy=energy_total, similarity=sim, current=float(cur), duration=float(dur), affected=aff)
