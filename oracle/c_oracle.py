"""ctypes driver of oracle/c/libstt_oracle.so (plain-C restatement, see oracle/c/stt_oracle.c). TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "c", "libstt_oracle.so")
_L = None


class OracleParams(C.Structure):
    _fields_ = [("damping", C.c_double), ("ms", C.c_double), ("ku", C.c_double), ("volume", C.c_double),
                ("polarization", C.c_double), ("easy_axis", C.c_double * 3), ("ref", C.c_double * 3),
                ("r_p", C.c_double), ("r_ap", C.c_double), ("area", C.c_double), ("temperature", C.c_double),
                ("max_current", C.c_double), ("max_duration", C.c_double), ("success_threshold", C.c_double),
                ("energy_weight", C.c_double), ("max_steps", C.c_int32), ("thermal", C.c_int32), ("euler", C.c_int32),
                ("pad", C.c_int32)]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "c", "stt_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(src) > os.path.getmtime(SO):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return SO


def lib():
    global _L
    if _L is None:
        _L = C.CDLL(build())
        _L.stt_oracle_step.restype = C.c_int64
    return _L


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class COracleEnv:
    """Batched SpinTorque-v0 (STT device, continuous/vector mode) on the CPU. Layout [n][3] row-major."""

    def __init__(self, n, device_params=None, max_steps=100, max_current=2e6, max_duration=5e-9, temperature=300.0,
                 include_thermal=True, success_threshold=0.9, energy_penalty_weight=0.1, method="rk4", nthreads=1):
        from .stt_oracle import default_stt_params
        dp = dict(default_stt_params() if device_params is None else device_params)
        p = OracleParams()
        p.damping = dp.get("damping", 0.01)
        p.ms = dp.get("saturation_magnetization", 800e3)
        p.ku = dp.get("uniaxial_anisotropy", 1e6)
        p.volume = dp.get("volume", 1e-24)
        p.polarization = dp.get("polarization", 0.7)
        p.easy_axis = (C.c_double * 3)(*np.asarray(dp.get("easy_axis", [0, 0, 1]), float))
        p.ref = (C.c_double * 3)(*np.asarray(dp.get("reference_magnetization", [0, 0, 1]), float))
        p.r_p = dp.get("resistance_parallel", 1e3)
        p.r_ap = dp.get("resistance_antiparallel", 2e3)
        p.area = dp.get("area", 1e-14)
        p.temperature = temperature
        p.max_current, p.max_duration = max_current, max_duration
        p.success_threshold, p.energy_weight = success_threshold, energy_penalty_weight
        p.max_steps, p.thermal, p.euler = max_steps, int(bool(include_thermal)), int(method == "euler")
        self.p, self.n, self.nthreads = p, int(n), int(nthreads)
        self.rng_seed = 12345
        n = self.n
        self.m = np.zeros((n, 3)); self.m[:, 2] = 1
        self.target = np.zeros((n, 3)); self.target[:, 2] = 1
        self.total_energy = np.zeros(n)
        self.step_count = np.zeros(n, np.int32)
        self.last_action = np.zeros((n, 2))
        self.obs = np.zeros((n, 12), np.float32)
        self.reward = np.zeros(n)
        self.terminated = np.zeros(n, np.uint8)
        self.truncated = np.zeros(n, np.uint8)
        self.step_energy = np.zeros(n)
        self.n_sub = np.zeros(n, np.int32)

    def reset(self, m0, target):
        m0 = np.broadcast_to(np.asarray(m0, float), (self.n, 3))
        target = np.broadcast_to(np.asarray(target, float), (self.n, 3))
        self.m[...] = m0 / np.linalg.norm(m0, axis=1, keepdims=True)
        self.target[...] = target / np.linalg.norm(target, axis=1, keepdims=True)
        self.total_energy[:] = 0
        self.step_count[:] = 0
        self.last_action[:] = 0
        lib().stt_oracle_obs(C.byref(self.p), C.c_int64(self.n), _p(self.m), _p(self.target), _p(self.total_energy),
                             _p(self.step_count), _p(self.last_action), _p(self.obs))
        return self.obs.copy()

    def step(self, actions, noise=None):
        act = np.ascontiguousarray(np.asarray(actions, np.float32).reshape(self.n, 2))
        stride = 0
        if noise is not None:
            noise = np.ascontiguousarray(noise, np.float64)
            stride = noise.shape[1]
        self.substeps = lib().stt_oracle_step(
            C.byref(self.p), C.c_int64(self.n), _p(self.m), _p(self.target), _p(self.total_energy), _p(self.step_count),
            _p(self.last_action), _p(act), _p(noise), C.c_int64(stride), _p(self.obs), _p(self.reward),
            _p(self.terminated), _p(self.truncated), _p(self.step_energy), _p(self.n_sub), C.c_int32(self.nthreads),
            C.c_uint64(self.rng_seed))
        return self.obs.copy(), self.reward.copy(), self.terminated.astype(bool), self.truncated.astype(bool)
