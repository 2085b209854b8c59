"""Device parameter sets: reference defaults, validation semantics and folding into kernel constants.

Mirrors (reference paths relative to spin_torque_gym/):
  - DeviceFactory.get_default_parameters          devices/device_factory.py:118-194
  - SpinTorqueEnv._get_default_device_params      envs/spin_torque_env.py:156-182
  - validate_parameters(params, 'stt_mram')       utils/validation.py:176-234, :491 (the solver always validates as STT)
  - device constructors' required keys            devices/stt_mram.py:33-55, sot_mram.py:46-59, vcma_mram.py:46-59
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _lib

DEVICE_KINDS = {"stt_mram": _lib.DEV_STT, "sot_mram": _lib.DEV_SOT, "vcma_mram": _lib.DEV_VCMA}


def default_device_parameters(device_type: str) -> Dict[str, Any]:
    """DeviceFactory.get_default_parameters (devices/device_factory.py:118-194)."""
    device_type = device_type.lower()
    if device_type == "stt_mram":
        return {
            "volume": 50e-9 * 100e-9 * 2e-9, "area": 50e-9 * 100e-9, "thickness": 2e-9, "aspect_ratio": 2.0,
            "saturation_magnetization": 800e3, "damping": 0.01, "uniaxial_anisotropy": 1.2e6,
            "exchange_constant": 20e-12, "polarization": 0.7, "resistance_parallel": 1e3,
            "resistance_antiparallel": 2e3, "easy_axis": np.array([0, 0, 1]),
            "reference_magnetization": np.array([0, 0, 1]),
        }
    if device_type == "sot_mram":
        return {
            "volume": 100e-9 * 100e-9 * 1e-9, "area": 100e-9 * 100e-9, "thickness": 1e-9,
            "saturation_magnetization": 800e3, "damping": 0.015, "uniaxial_anisotropy": 0.8e6,
            "exchange_constant": 20e-12, "spin_hall_angle": 0.2, "resistance_parallel": 500,
            "resistance_antiparallel": 1000, "easy_axis": np.array([0, 0, 1]),
        }
    if device_type == "vcma_mram":
        return {
            "volume": 80e-9 * 80e-9 * 1.5e-9, "area": 80e-9 * 80e-9, "thickness": 1.5e-9,
            "saturation_magnetization": 800e3, "damping": 0.008, "uniaxial_anisotropy": 1.5e6,
            "exchange_constant": 20e-12, "vcma_coefficient": 100e-6, "resistance_parallel": 2e3,
            "resistance_antiparallel": 4e3, "easy_axis": np.array([0, 0, 1]),
        }
    return {"volume": 1e-24, "saturation_magnetization": 800e3, "damping": 0.01, "uniaxial_anisotropy": 1e6,
            "exchange_constant": 20e-12, "polarization": 0.7}


def env_default_device_params(device_type: str) -> Dict[str, Any]:
    """SpinTorqueEnv._get_default_device_params (envs/spin_torque_env.py:156-182)."""
    if device_type == "stt_mram":
        return default_device_parameters("stt_mram")
    return {"volume": 1e-24, "saturation_magnetization": 800e3, "damping": 0.01, "uniaxial_anisotropy": 1e6,
            "polarization": 0.7}


_REQUIRED = {
    "stt_mram": ["volume", "saturation_magnetization", "damping", "uniaxial_anisotropy", "polarization"],
    "sot_mram": ["volume", "saturation_magnetization", "damping", "uniaxial_anisotropy", "easy_axis"],
    "vcma_mram": ["volume", "saturation_magnetization", "damping", "uniaxial_anisotropy", "easy_axis"],
}


def check_device_constructible(device_type: str, params: Dict[str, Any]) -> None:
    """Raise like DeviceFactory.create_device does when the device class' own validation fails
    (devices/device_factory.py:74-77 wraps the ValueError in RuntimeError)."""
    device_type = device_type.lower()
    if device_type not in _REQUIRED:
        raise ValueError(f"Unknown device type '{device_type}'. Available types: {list(_REQUIRED)}")
    for key in _REQUIRED[device_type]:
        if key not in params:
            raise RuntimeError(f"Failed to create {device_type} device: Missing required parameter: {key}")
    if device_type == "stt_mram":
        msg = None
        if params["volume"] <= 0:
            msg = "Volume must be positive"
        elif params["saturation_magnetization"] <= 0:
            msg = "Saturation magnetization must be positive"
        elif not 0 <= params["damping"] <= 1:
            msg = "Damping must be between 0 and 1"
        elif not 0 <= params["polarization"] <= 1:
            msg = "Polarization must be between 0 and 1"
        if msg:
            raise RuntimeError(f"Failed to create {device_type} device: {msg}")


def _finite_pos(x, lo) -> bool:
    try:
        x = float(x)
    except (TypeError, ValueError):
        return False
    return math.isfinite(x) and x > 0 and x >= lo


def _prob(x) -> bool:
    try:
        x = float(x)
    except (TypeError, ValueError):
        return False
    return math.isfinite(x) and 0.0 <= x <= 1.0


def solver_accepts(params: Dict[str, Any], temperature: float) -> bool:
    """True iff RobustLLGSSolver._validate_inputs passes (utils/robust_solver.py:152-190): the solver validates every
    device as 'stt_mram' (utils/validation.py:491), so a SOT/VCMA dict without `polarization` makes every solve fail and the
    env keeps its magnetisation."""
    if not (isinstance(temperature, (int, float)) and temperature > 0):
        return False
    if not _finite_pos(params.get("volume"), 1e-30):
        return False
    if not _finite_pos(params.get("saturation_magnetization"), 1e3):
        return False
    if not _prob(params.get("damping")):
        return False
    if not _finite_pos(params.get("uniaxial_anisotropy"), 1e3):
        return False
    if "easy_axis" not in params:
        return False
    e = np.asarray(params["easy_axis"], dtype=float)
    if e.shape != (3,) or not np.all(np.isfinite(e)) or np.linalg.norm(e) < 1e-12:
        return False
    if not _prob(params.get("polarization")):
        return False
    return True


def make_param_struct(device_type: str, params: Dict[str, Any], *, max_steps: int, max_current: float,
                      max_duration: float, temperature: float, thermal: bool, success_threshold: float,
                      energy_penalty_weight: float, applied_field: Sequence[float] = (0.0, 0.0, 0.0),
                      max_step: float = 1e-12) -> _lib.StgSttParams:
    kind = DEVICE_KINDS[device_type.lower()]
    p = _lib.StgSttParams()
    # same defaults as the solver's dict lookups (physics/simple_solver.py:310-315)
    p.damping = float(params.get("damping", 0.01))
    p.saturation_magnetization = float(params.get("saturation_magnetization", 800e3))
    p.uniaxial_anisotropy = float(params.get("uniaxial_anisotropy", 1e6))
    p.volume = float(params.get("volume", 1e-24))
    p.polarization = float(params.get("polarization", 0.7))
    e = np.asarray(params.get("easy_axis", [0, 0, 1]), dtype=float)
    r = np.asarray(params.get("reference_magnetization", [0, 0, 1]), dtype=float)
    p.easy_axis = _lib.c_double3(*e)
    p.reference_magnetization = _lib.c_double3(*r)
    p.resistance_parallel = float(params.get("resistance_parallel", 1e3))
    p.resistance_antiparallel = float(params.get("resistance_antiparallel", 2e3))
    p.area = float(params.get("area", 1e-14))                        # envs/spin_torque_env.py:476
    series = 0.0
    if kind == _lib.DEV_SOT:                                         # devices/sot_mram.py:61-76, 219-226
        t_hm = float(params.get("heavy_metal_thickness", 5e-9))
        rho = float(params.get("heavy_metal_resistivity", 2e-7))
        thickness = float(params.get("thickness", 1e-9))
        area = float(params.get("area", p.volume / thickness))
        series = 0.1 * ((rho / t_hm) / (area * 1e-12))
    p.series_resistance = series
    p.temperature = float(temperature)
    p.applied_field = _lib.c_double3(*[float(x) for x in applied_field])
    p.max_current = float(max_current)
    p.max_duration = float(max_duration)
    p.success_threshold = float(success_threshold)
    p.energy_penalty_weight = float(energy_penalty_weight)
    p.max_step = float(max_step)
    p.max_steps = int(max_steps)
    p.device_kind = kind
    p.thermal = 1 if thermal else 0
    p.solver_valid = 1 if solver_accepts(params, temperature) else 0
    return p


def fold(structs: List[_lib.StgSttParams]) -> np.ndarray:
    """stg_stt_fold on the host -> float64 array [n_sets, FOLDED_DOUBLES] ready to upload."""
    lib = _lib.load()
    n = len(structs)
    arr = (_lib.StgSttParams * n)(*structs)
    out = (_lib.StgSttFolded * n)()
    _lib.check(lib.stg_stt_fold(arr, n, out), "stg_stt_fold")
    return np.frombuffer(out, dtype=np.float64).reshape(n, _lib.FOLDED_DOUBLES).copy()


def all_axis_z(folded: np.ndarray) -> bool:
    lib = _lib.load()
    buf = np.ascontiguousarray(folded, dtype=np.float64)
    return bool(lib.stg_stt_all_axis_z(buf.ctypes.data_as(C.POINTER(_lib.StgSttFolded)), buf.shape[0]))


# ---- generalised LLGS right-hand side of the RK45 solver (K2) ---------------------------------------------------------------
GAMMA_LLGS = 2.21e5
MU0 = 4 * np.pi * 1e-7
KB_LLGS = 1.380649e-23     # physics/llgs_solver.py:49


def make_llg_struct(kind: str, params: Dict[str, Any], *, thermal: bool = False, temperature: float = 300.0,
                    current_direction: Optional[Sequence[float]] = None, gamma: float = GAMMA_LLGS) -> _lib.StgLlgParams:
    """StgLlgParams for one device parameter dict. kind 'stt_mram' is exactly LLGSSolver's RHS (physics/llgs_solver.py:
    92-126, 182-237); 'sot_mram' / 'vcma_mram' compose the device methods' terms into the same RHS."""
    kind = kind.lower()
    p = _lib.StgLlgParams()
    alpha = float(params.get("damping", 0.01))
    ms = float(params.get("saturation_magnetization", 800e3))
    vol = float(params.get("volume", 1e-24))
    p.gamma, p.mu0, p.alpha, p.saturation_magnetization, p.volume = gamma, MU0, alpha, ms, vol
    p.uniaxial_anisotropy = float(params.get("uniaxial_anisotropy", 1e6))
    p.easy_axis = _lib.c_double3(*np.asarray(params.get("easy_axis", [0, 0, 1]), dtype=float))
    p.h_th = math.sqrt(2 * alpha * KB_LLGS * temperature / (gamma * MU0 * ms * vol)) if thermal else 0.0
    p.p_hat = _lib.c_double3(0.0, 0.0, 1.0)
    if kind == "stt_mram":
        p.demag_n = _lib.c_double3(*np.asarray(params.get("demag_factors", [0, 0, 1]), dtype=float))
        a_ex = float(params.get("exchange_constant", 20e-12))
        p.exchange_coeff = (2 * a_ex / (MU0 * ms)) * 0.1 if a_ex > 0 else 0.0
        beta = float(params.get("polarization", 0.7)) * gamma / (2 * ms * vol)
        p.c_dl_p, p.c_fl_p = beta, 0.1 * beta
    elif kind in ("sot_mram", "vcma_mram"):
        ar = float(params.get("aspect_ratio", 1.0))
        nx, ny = (1.0 / (1.0 + ar), ar / (1.0 + ar)) if ar >= 1.0 else (ar / (1.0 + ar), 1.0 / (1.0 + ar))
        p.demag_n = _lib.c_double3(nx, ny, 1.0 - nx - ny)
        if kind == "sot_mram":
            t_hm = float(params.get("heavy_metal_thickness", 5e-9))
            js = float(params.get("spin_hall_angle", 0.1)) * float(params.get("interface_transparency", 0.5)) * \
                (t_hm / (t_hm + float(params.get("thickness", 1e-9))))
            p.c_dl_s = float(params.get("damping_like_efficiency", 0.2)) * js
            p.c_fl_s = float(params.get("field_like_efficiency", 0.1)) * js
            d = np.array([1.0, 0.0, 0.0]) if current_direction is None else np.asarray(current_direction, dtype=float)
            d = d / np.linalg.norm(d)
            p.sigma = _lib.c_double3(*np.cross(np.array([0.0, 0.0, 1.0]), d))
        else:
            p.use_vcma = 1
            p.vcma_coefficient = float(params.get("vcma_coefficient", 100e-6))
            p.dielectric_thickness = float(params.get("dielectric_thickness", 1e-9))
            p.breakdown_voltage = float(params.get("breakdown_voltage", 2.0))
    else:
        raise ValueError(f"Unknown device type '{kind}'")
    return p


def llg_table(structs: List[_lib.StgLlgParams]) -> np.ndarray:
    """Raw bytes of a StgLlgParams array as a uint8 NumPy array (uploaded as the device table)."""
    arr = (_lib.StgLlgParams * len(structs))(*structs)
    return np.frombuffer(arr, dtype=np.uint8).copy()
