"""NumPy-backed driver of tests/hostsim/_hostsim.so (host build of the kernel bodies). Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from spin_torque_rl_gym_b200 import _lib, params as P

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_hostsim.so")
_H = None


def build(force=False):
    src = os.path.join(HERE, "hostsim.cpp")
    deps = [src] + [os.path.join(HERE, "..", "..", "spin_torque_rl_gym_b200", "csrc", f)
                    for f in ("llgs_core.cuh", "stt_env_core.cuh", "rk45_core.cuh", "array_core.cuh")] + [os.path.join(HERE, "..", "..", "include", "stg.h")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=fast", "-fPIC", "-shared", "-x", "c++", src,
                        "-o", SO], check=True)
    return SO


def lib():
    global _H
    if _H is None:
        _H = C.CDLL(build())
        _H.hostsim_stt_step.argtypes = [C.POINTER(_lib.StgSttStepArgs), C.c_int]
        _H.hostsim_stt_reset.argtypes = [C.POINTER(_lib.StgSttResetArgs)]
        _H.hostsim_stt_solve.argtypes = [C.POINTER(_lib.StgSttSolveArgs), C.c_int]
        _H.hostsim_llgs_rk45.argtypes = [C.POINTER(_lib.StgRk45Args)]
        _H.hostsim_array_step.argtypes = [C.POINTER(_lib.StgArrayStepArgs)]
    return _H


def _p(a):
    return None if a is None else a.ctypes.data


class HostSimEnv:
    """Same state layout / flags as SpinTorqueVectorEnv, NumPy arrays instead of CUDA tensors."""

    def __init__(self, num_envs, device_type="stt_mram", device_params=None, target_states=None, max_steps=100,
                 max_current=2e6, max_duration=5e-9, temperature=300.0, include_thermal_fluctuations=True,
                 success_threshold=0.9, energy_penalty_weight=0.1, f64=True, integrator="rk4", rng_seed=0,
                 env_offset=0, autoreset=False, force_general=False, pair=True, thermal_stream="xoshiro"):
        N = self.N = int(num_envs)
        if device_params is None:
            device_params = P.env_default_device_params(device_type)
        st = P.make_param_struct(device_type, device_params, max_steps=max_steps, max_current=max_current,
                                 max_duration=max_duration, temperature=temperature,
                                 thermal=include_thermal_fluctuations, success_threshold=success_threshold,
                                 energy_penalty_weight=energy_penalty_weight)
        self.table = P.fold([st])
        self.axis_z = P.all_axis_z(self.table) and not force_general
        self.f64 = f64
        self.integrator = integrator
        self.thermal = include_thermal_fluctuations and temperature > 0
        self.seed, self.env_offset, self.autoreset = rng_seed, env_offset, autoreset
        self.pair = pair
        self.thermal_stream = thermal_stream
        tt = np.array([[0, 0, 1.0], [0, 0, -1.0]]) if target_states is None else \
            np.array([np.asarray(t, float) / np.linalg.norm(t) for t in target_states])
        self.target_table = np.ascontiguousarray(tt)
        self.m = np.zeros((3, N)); self.m[2] = 1
        self.target = np.zeros((3, N)); self.target[2] = 1
        self.total_energy = np.zeros(N)
        self.last_action = np.zeros((2, N))
        self.step_count = np.zeros(N, np.int32)
        self.episode = np.zeros(N, np.int32)
        self.obs = np.zeros((N, 12), np.float32)
        self.final_obs = np.zeros((N, 12), np.float32)
        self.reward = np.zeros(N)
        self.terminated = np.zeros(N, np.uint8)
        self.truncated = np.zeros(N, np.uint8)
        self.step_energy = np.zeros(N)
        self.n_sub = np.zeros(N, np.int32)
        self.status = np.zeros(N, np.int32)

    def _state(self):
        s = _lib.StgSttState()
        s.m, s.target, s.total_energy = _p(self.m), _p(self.target), _p(self.total_energy)
        s.last_action, s.step_count, s.episode = _p(self.last_action), _p(self.step_count), _p(self.episode)
        return s

    def reset(self, m0=None, target=None, mask=None):
        a = _lib.StgSttResetArgs()
        a.d_table = _p(self.table); a.state = self._state()
        keep = []
        if m0 is not None:
            m0 = np.ascontiguousarray(np.broadcast_to(np.asarray(m0, float), (self.N, 3))); keep.append(m0); a.d_m0 = _p(m0)
        if target is not None:
            target = np.ascontiguousarray(np.broadcast_to(np.asarray(target, float), (self.N, 3))); keep.append(target)
            a.d_target0 = _p(target)
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8); a.d_mask = _p(mask)
        a.d_target_table = _p(self.target_table); a.n_targets = len(self.target_table)
        a.d_obs = _p(self.obs); a.seed = self.seed; a.env_offset = self.env_offset
        a.n_envs = self.N; a.n_sets = 1
        lib().hostsim_stt_reset(C.byref(a))
        return self.obs.copy()

    def step(self, actions, noise=None, perm=None):
        act = np.ascontiguousarray(np.asarray(actions, np.float32).reshape(self.N, 2))
        a = _lib.StgSttStepArgs()
        flags = 0
        if self.integrator == "euler": flags |= _lib.F_EULER
        if self.axis_z: flags |= _lib.F_AXIS_Z
        if self.autoreset: flags |= _lib.F_AUTORESET
        if not self.pair: flags |= _lib.F_NO_PAIR
        if noise is not None:
            noise = np.ascontiguousarray(noise, np.float64)
            flags |= _lib.F_THERMAL_INJECT; a.d_noise = _p(noise); a.noise_stride = noise.shape[1]
        elif self.thermal:
            flags |= _lib.F_THERMAL_PHILOX
            if self.thermal_stream == "philox": flags |= _lib.F_STREAM_PHILOX10
        if perm is not None:
            perm = np.ascontiguousarray(perm, np.int32); flags |= _lib.F_SORTED; a.d_perm = _p(perm)
        a.d_table = _p(self.table); a.state = self._state(); a.d_action = _p(act)
        o = a.out
        o.obs, o.reward, o.terminated, o.truncated = _p(self.obs), _p(self.reward), _p(self.terminated), _p(self.truncated)
        o.step_energy, o.n_sub, o.status, o.final_obs = _p(self.step_energy), _p(self.n_sub), _p(self.status), _p(self.final_obs)
        a.d_target_table = _p(self.target_table); a.n_targets = len(self.target_table)
        a.seed, a.env_offset, a.n_envs, a.n_sets, a.flags = self.seed, self.env_offset, self.N, 1, flags
        lib().hostsim_stt_step(C.byref(a), 1 if self.f64 else 0)
        return self.obs.copy(), self.reward.copy(), self.terminated.astype(bool), self.truncated.astype(bool)


def host_solve(m0, t_end, device_params, method="rk4", current=0.0, t_pulse=None, current_grid=None, field_grid=None,
               max_step=1e-12, f64=True):
    """SimpleLLGSSolver.solve_batch's kernel call on the host build of the kernel bodies: returns (trajectory [N, rows, 3],
    n_sub [N]). Grids as in StgSttSolveArgs ([n_sub,3] / [n_sub,3,3], shared by every trajectory)."""
    m0 = np.ascontiguousarray(np.asarray(m0, dtype=np.float64).reshape(-1, 3))
    n = m0.shape[0]
    te = np.broadcast_to(np.asarray(t_end, dtype=np.float64), (n,))
    tp = te if t_pulse is None else np.broadcast_to(np.asarray(t_pulse, dtype=np.float64), (n,))
    pulse = np.ascontiguousarray(np.stack([np.broadcast_to(np.asarray(current, dtype=np.float64), (n,)), tp, te], 1))
    st = P.make_param_struct("stt_mram", device_params, max_steps=1, max_current=1.0, max_duration=1.0, temperature=300.0,
                             thermal=False, success_threshold=0.9, energy_penalty_weight=0.1, applied_field=(0, 0, 0),
                             max_step=max_step)
    table = P.fold([st])
    rows = int(np.ceil(float(te.max()) / min(max_step, float(te.max()) / 100))) + 16
    traj = np.zeros((n, rows, 3))
    out, nsub, guard = np.zeros((n, 3)), np.zeros(n, np.int32), np.zeros(n, np.int32)
    a = _lib.StgSttSolveArgs()
    a.d_table, a.d_m0, a.d_pulse, a.d_m_out = table.ctypes.data, m0.ctypes.data, pulse.ctypes.data, out.ctypes.data
    a.d_traj, a.traj_stride, a.d_n_sub, a.d_guard = traj.ctypes.data, rows, nsub.ctypes.data, guard.ctypes.data
    keep = []
    for name, g in (("d_current_grid", current_grid), ("d_field_grid", field_grid)):
        if g is not None:
            g = np.ascontiguousarray(np.asarray(g, dtype=np.float64))
            keep.append(g)
            setattr(a, name, g.ctypes.data)
            a.grid_stride, a.grid_envs = g.shape[0], 1
    a.n_envs, a.n_sets = n, 1
    a.flags = (_lib.F_EULER if method == "euler" else 0) | (_lib.F_AXIS_Z if P.all_axis_z(table) else 0)
    lib().hostsim_stt_solve(C.byref(a), 1 if f64 else 0)
    return traj, nsub
