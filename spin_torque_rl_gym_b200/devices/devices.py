from __future__ import annotations

import ctypes as C
import math
import warnings
from typing import Any, Dict, Optional, Tuple

import numpy as np

from .. import _lib, params as _params

MU0 = 4 * np.pi * 1e-7


def _to_dev(x, device, rows=None):
    """-> (contiguous float64 CUDA tensor [n,k] or [n], was_numpy, was_single)"""
    torch = _lib.require_cuda()
    was_numpy = not isinstance(x, torch.Tensor)
    t = torch.as_tensor(np.asarray(x, dtype=np.float64)) if was_numpy else x.to(torch.float64)
    t = t.to(device)
    return t.contiguous(), was_numpy


def _back(t, was_numpy, single):
    if single:
        t = t[0]
    return t.cpu().numpy() if was_numpy else t


class BaseSpintronicDevice:
    """devices/base_device.py:13-138."""

    KIND = _lib.DEV_STT

    def __init__(self, device_params: Dict[str, Any], device: Any = "cuda"):
        torch = _lib.require_cuda()
        self.device_params = dict(device_params)
        self.volume = device_params.get("volume", 1e-24)
        self.thickness = device_params.get("thickness", 1e-9)
        self.saturation_magnetization = device_params.get("saturation_magnetization", 800e3)
        self.mu0 = MU0
        self.kb = 1.380649e-23
        self.e = 1.602176634e-19
        self.hbar = 1.054571817e-34
        for key in ("volume", "saturation_magnetization"):
            if key not in self.device_params:
                raise ValueError(f"Missing required parameter: {key}")
        self._device = torch.device(device)
        self._lib = _lib.load()

    # ---- parameter access -------------------------------------------------------------------------------------------
    def get_parameter(self, key: str, default: Any = None) -> Any:
        return self.device_params.get(key, default)

    def set_parameter(self, key: str, value: Any) -> None:
        self.device_params[key] = value

    def validate_magnetization(self, magnetization):
        """devices/base_device.py:94-116 (batched: every row is normalised; a zero row raises)."""
        m = np.asarray(magnetization, dtype=float) if not hasattr(magnetization, "device") else magnetization
        if m.shape[-1] != 3 or m.ndim > 2:
            raise ValueError(f"Magnetization must be 3D vector, got shape {tuple(m.shape)}")
        if hasattr(m, "device"):
            mag = m.norm(dim=-1, keepdim=True)
            if bool((mag < 1e-12).any()):
                raise ValueError("Magnetization vector cannot be zero")
            return m / mag
        mag = np.linalg.norm(m, axis=-1, keepdims=True)
        if np.any(mag < 1e-12):
            raise ValueError("Magnetization vector cannot be zero")
        return m / mag

    # ---- C-ABI plumbing -----------------------------------------------------------------------------------------------
    def _demag_factors(self):
        ar = self.device_params.get("aspect_ratio", 1.0)
        if ar >= 1.0:
            nx, ny = 1.0 / (1.0 + ar), ar / (1.0 + ar)
        else:
            nx, ny = ar / (1.0 + ar), 1.0 / (1.0 + ar)
        return nx, ny, 1.0 - nx - ny

    def _struct(self) -> _lib.StgDeviceParams:
        p = _lib.StgDeviceParams()
        p.kind = self.KIND
        p.saturation_magnetization = float(self.saturation_magnetization)
        p.uniaxial_anisotropy = float(self.device_params.get("uniaxial_anisotropy", 1e6))
        p.mu0 = MU0
        p.easy_axis = _lib.c_double3(*np.asarray(self.device_params.get("easy_axis", [0, 0, 1]), dtype=float))
        if self.KIND != _lib.DEV_STT:
            p.demag_n = _lib.c_double3(*self._demag_factors())
        ref = np.asarray(self.device_params.get("reference_magnetization", [0, 0, 1]), dtype=float)
        p.reference_magnetization = _lib.c_double3(*(ref / np.linalg.norm(ref)))
        p.resistance_parallel = float(self.device_params.get("resistance_parallel", 1e3))
        p.resistance_antiparallel = float(self.device_params.get("resistance_antiparallel", 2e3))
        return p

    def _rows(self, magnetization) -> Tuple[Any, bool, bool]:
        m, was_numpy = _to_dev(magnetization, self._device)
        single = m.dim() == 1
        if single:
            m = m.reshape(1, 3)
        if m.dim() != 2 or m.shape[1] != 3:
            raise ValueError(f"Magnetization must be 3D vector, got shape {tuple(m.shape)}")
        return m.contiguous(), was_numpy, single

    def _stream(self):
        torch = _lib.require_cuda()
        return torch.cuda.current_stream(self._device).cuda_stream

    def _zero_counter(self):
        """Device int32 the kernels add the number of zero-norm STT rows to (devices/base_device.py:112-114 raises for them)."""
        torch = _lib.require_cuda()
        z = getattr(self, "_zero_rows", None)
        if z is None:
            z = self._zero_rows = torch.zeros(1, dtype=torch.int32, device=self._device)
        else:
            z.zero_()
        return z

    def _raise_if_zero_rows(self, z) -> None:
        if z is not None and int(z.item()) > 0:             # one 4-byte D2H instead of three passes over the input
            raise ValueError("Magnetization vector cannot be zero")

    def compute_effective_field(self, magnetization, applied_field, applied_voltage=None, check_zero: bool = True):
        """`check_zero=False` skips reading back the zero-row counter, which keeps a batched call fully asynchronous; the
        output rows of zero vectors are then unspecified (the reference raises for them)."""
        torch = _lib.require_cuda()
        m, was_numpy, single = self._rows(magnetization)
        z = self._zero_counter() if (check_zero and self.KIND == _lib.DEV_STT) else None
        h, _ = _to_dev(applied_field, self._device)
        h = h.reshape(-1, 3).contiguous()
        if h.shape[0] not in (1, m.shape[0]):
            raise ValueError("applied_field must be [3] or [N,3]")
        v = None
        if applied_voltage is not None:
            v, _ = _to_dev(applied_voltage, self._device)
            v = v.reshape(-1)
            if v.numel() == 1:
                v = v.expand(m.shape[0])
            v = v.contiguous()
        out = torch.empty_like(m)
        p = self._struct()
        with torch.cuda.device(self._device):
            _lib.check(self._lib.stg_device_field_f64(C.byref(p), m.data_ptr(), h.data_ptr(), h.shape[0], _lib.ptr(v),
                                                      out.data_ptr(), m.shape[0], _lib.ptr(z), self._stream()),
                       "stg_device_field_f64")
        self._raise_if_zero_rows(z)
        return _back(out, was_numpy, single)

    def compute_resistance(self, magnetization, check_zero: bool = True):
        torch = _lib.require_cuda()
        m, was_numpy, single = self._rows(magnetization)
        z = self._zero_counter() if (check_zero and self.KIND == _lib.DEV_STT) else None
        out = torch.empty(m.shape[0], dtype=torch.float64, device=self._device)
        p = self._struct()
        with torch.cuda.device(self._device):
            _lib.check(self._lib.stg_device_resistance_f64(C.byref(p), m.data_ptr(), out.data_ptr(), m.shape[0],
                                                           _lib.ptr(z), self._stream()), "stg_device_resistance_f64")
        self._raise_if_zero_rows(z)
        if single:
            return float(out[0])
        return out.cpu().numpy() if was_numpy else out

    def get_device_info(self) -> Dict[str, Any]:
        return {"device_type": self.__class__.__name__, "volume": self.volume, "thickness": self.thickness,
                "saturation_magnetization": self.saturation_magnetization, "parameters": self.device_params.copy()}

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(volume={self.volume:.2e}, Ms={self.saturation_magnetization:.0f})"


class STTMRAMDevice(BaseSpintronicDevice):
    """devices/stt_mram.py:13-98."""

    KIND = _lib.DEV_STT

    def __init__(self, device_params: Dict[str, Any], device: Any = "cuda"):
        for key in ("volume", "saturation_magnetization", "damping", "uniaxial_anisotropy", "polarization"):
            if key not in device_params:
                raise ValueError(f"Missing required parameter: {key}")
        super().__init__(device_params, device)
        if self.device_params["volume"] <= 0:
            raise ValueError("Volume must be positive")
        if self.device_params["saturation_magnetization"] <= 0:
            raise ValueError("Saturation magnetization must be positive")
        if not 0 <= self.device_params["damping"] <= 1:
            raise ValueError("Damping must be between 0 and 1")
        if not 0 <= self.device_params["polarization"] <= 1:
            raise ValueError("Polarization must be between 0 and 1")
        ref = np.asarray(self.get_parameter("reference_magnetization", np.array([0, 0, 1])), dtype=float)
        self.reference_magnetization = ref / np.linalg.norm(ref)


class SOTMRAMDevice(BaseSpintronicDevice):
    """devices/sot_mram.py:16-441."""

    KIND = _lib.DEV_SOT

    def __init__(self, device_params: Dict[str, Any], device: Any = "cuda"):
        super().__init__(device_params, device)
        for key in ("volume", "saturation_magnetization", "damping", "uniaxial_anisotropy", "easy_axis"):
            if key not in self.device_params:
                raise ValueError(f"Missing required parameter: {key}")
        if self.device_params.get("spin_hall_angle", 0.1) > 1.0:
            warnings.warn("Spin Hall angle > 1.0 is physically unrealistic")
        g = device_params.get
        self.spin_hall_angle = g("spin_hall_angle", 0.1)
        self.heavy_metal_thickness = g("heavy_metal_thickness", 5e-9)
        self.heavy_metal_resistivity = g("heavy_metal_resistivity", 2e-7)
        self.interface_transparency = g("interface_transparency", 0.5)
        self.field_like_efficiency = g("field_like_efficiency", 0.1)
        self.damping_like_efficiency = g("damping_like_efficiency", 0.2)
        self._update_cached_parameters()

    def _update_cached_parameters(self) -> None:
        """devices/sot_mram.py:61-76."""
        self.j_s_efficiency = (self.spin_hall_angle * self.interface_transparency *
                               (self.heavy_metal_thickness / (self.heavy_metal_thickness + self.thickness)))
        self.tau_dl_factor = self.damping_like_efficiency * self.j_s_efficiency
        self.tau_fl_factor = self.field_like_efficiency * self.j_s_efficiency
        self.sheet_resistance_hm = self.heavy_metal_resistivity / self.heavy_metal_thickness
        self.area = self.device_params.get("area", self.volume / self.thickness)

    def _struct(self):
        p = super()._struct()
        p.tau_dl_factor = float(self.tau_dl_factor)
        p.tau_fl_factor = float(self.tau_fl_factor)
        r_hm = self.sheet_resistance_hm / (self.area * 1e-12)
        p.series_resistance = float(r_hm * 0.1)
        return p

    def compute_spin_torque(self, current_density, magnetization, current_direction: Optional[np.ndarray] = None):
        """(tau_DL, tau_FL) for one state or a batch (devices/sot_mram.py:163-194)."""
        torch = _lib.require_cuda()
        m, was_numpy, single = self._rows(magnetization)
        j, _ = _to_dev(current_density, self._device)
        j = j.reshape(-1).contiguous()
        if j.numel() not in (1, m.shape[0]):
            raise ValueError("current_density must be a scalar or [N]")
        d = np.array([1.0, 0.0, 0.0]) if current_direction is None else np.asarray(current_direction, dtype=float)
        dl, fl = torch.empty_like(m), torch.empty_like(m)
        p = self._struct()
        darr = (C.c_double * 3)(*d)
        with torch.cuda.device(self._device):
            _lib.check(self._lib.stg_device_sot_torque_f64(C.byref(p), j.data_ptr(), j.numel(), m.data_ptr(), darr,
                                                           dl.data_ptr(), fl.data_ptr(), m.shape[0], self._stream()),
                       "stg_device_sot_torque_f64")
        return _back(dl, was_numpy, single), _back(fl, was_numpy, single)

    def compute_power_consumption(self, current_density: float, pulse_duration: float, magnetization=None) -> float:
        """devices/sot_mram.py:230-255 (scalar bookkeeping, not on the LLGS path)."""
        if abs(current_density) < 1e-12:
            return 0.0
        current = current_density * self.area
        voltage = current * self.sheet_resistance_hm / (self.area * 1e-12)
        return voltage * current * pulse_duration

    def get_switching_threshold(self) -> Dict[str, float]:
        alpha = self.device_params["damping"]
        h_k = 2 * self.device_params["uniaxial_anisotropy"] / (self.mu0 * self.saturation_magnetization)
        j_c = 5e6 * (1 + alpha) * (1 + h_k / 1e6) / (1 + self.tau_dl_factor)
        return {"critical_current_density": j_c, "critical_field": h_k, "damping_like_efficiency": self.tau_dl_factor,
                "field_like_efficiency": self.tau_fl_factor}


class VCMAMRAMDevice(BaseSpintronicDevice):
    """devices/vcma_mram.py:16-510."""

    KIND = _lib.DEV_VCMA

    def __init__(self, device_params: Dict[str, Any], device: Any = "cuda"):
        super().__init__(device_params, device)
        for key in ("volume", "saturation_magnetization", "damping", "uniaxial_anisotropy", "easy_axis"):
            if key not in self.device_params:
                raise ValueError(f"Missing required parameter: {key}")
        g = device_params.get
        if g("vcma_coefficient", 100e-6) < 0:
            warnings.warn("Negative VCMA coefficient indicates inverted VCMA effect")
        self.vcma_coefficient = g("vcma_coefficient", 100e-6)
        self.dielectric_thickness = g("dielectric_thickness", 1e-9)
        self.dielectric_constant = g("dielectric_constant", 25.0)
        self.breakdown_voltage = g("breakdown_voltage", 2.0)
        self.leakage_resistance = g("leakage_resistance", 1e12)
        self.capacitance_per_area = g("capacitance_per_area", None)
        self.area = g("area", self.volume / self.thickness)
        if self.capacitance_per_area is None:
            self.capacitance = 8.854e-12 * self.dielectric_constant * self.area / self.dielectric_thickness
        else:
            self.capacitance = self.capacitance_per_area * self.area
        self.base_anisotropy = self.device_params["uniaxial_anisotropy"]
        self.max_electric_field = self.breakdown_voltage / self.dielectric_thickness

    def _struct(self):
        p = super()._struct()
        p.vcma_coefficient = float(self.vcma_coefficient)
        p.dielectric_thickness = float(self.dielectric_thickness)
        p.breakdown_voltage = float(self.breakdown_voltage)
        return p

    def _compute_effective_anisotropy(self, voltage):
        """K_eff(V) for a scalar or a batch of voltages (devices/vcma_mram.py:122-147)."""
        torch = _lib.require_cuda()
        v, was_numpy = _to_dev(voltage, self._device)
        single = v.dim() == 0
        v = v.reshape(-1).contiguous()
        out = torch.empty_like(v)
        p = self._struct()
        with torch.cuda.device(self._device):
            _lib.check(self._lib.stg_vcma_anisotropy_f64(C.byref(p), v.data_ptr(), out.data_ptr(), v.numel(),
                                                         self._stream()), "stg_vcma_anisotropy_f64")
        if single:
            return float(out[0])
        return out.cpu().numpy() if was_numpy else out

    def compute_switching_probability(self, voltage: float, pulse_duration: float, temperature: float = 300.0,
                                      initial_state=None) -> float:
        """Arrhenius switching probability (devices/vcma_mram.py:187-234)."""
        k_eff = self._compute_effective_anisotropy(voltage)
        barrier = k_eff * self.volume
        thermal = 1.38e-23 * temperature
        if barrier <= 0:
            return 1.0
        if thermal <= 0:
            return 0.0
        rate = 1e9 * math.exp(-barrier / thermal)
        return min(1.0 - math.exp(-rate * pulse_duration), 1.0)

    def compute_power_consumption(self, voltage: float, pulse_duration: float, magnetization=None) -> float:
        """0.5 C V^2 + V^2 T / R_leak (devices/vcma_mram.py:259-287)."""
        if abs(voltage) < 1e-12:
            return 0.0
        return 0.5 * self.capacitance * voltage ** 2 + voltage ** 2 * pulse_duration / self.leakage_resistance


_TYPES = {"stt_mram": STTMRAMDevice, "sot_mram": SOTMRAMDevice, "vcma_mram": VCMAMRAMDevice}


class DeviceFactory:
    """devices/device_factory.py:17-265 (registry + default parameter sets)."""

    def __init__(self, device: Any = "cuda"):
        self._device_types = dict(_TYPES)
        self._device = device

    def create_device(self, device_type: str, device_params: Dict[str, Any]):
        device_type = device_type.lower()
        if device_type not in self._device_types:
            raise ValueError(f"Unknown device type '{device_type}'. Available types: {list(self._device_types)}")
        try:
            return self._device_types[device_type](device_params, device=self._device)
        except _lib.StgError:
            raise
        except Exception as e:  # noqa: BLE001 - same wrapping as the reference (devices/device_factory.py:74-77)
            raise RuntimeError(f"Failed to create {device_type} device: {e}")

    def get_available_devices(self) -> list:
        return list(self._device_types)

    def get_default_parameters(self, device_type: str) -> Dict[str, Any]:
        return _params.default_device_parameters(device_type)

    def create_default_device(self, device_type: str):
        return self.create_device(device_type, self.get_default_parameters(device_type))


def create_device(device_type: str, device_params: Optional[Dict[str, Any]] = None, device: Any = "cuda"):
    f = DeviceFactory(device)
    return f.create_device(device_type, device_params or f.get_default_parameters(device_type))
