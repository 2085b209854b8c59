"""Batched ThermalFluctuations (reference: physics/thermal_model.py:12-336). The white / Ornstein-Uhlenbeck field generator
runs on the GPU for `num_devices` independent devices (Philox stream). The Neel-Brown analytics keep the reference's scalar
methods and add a batched form: `batch_analytics` evaluates stability factor, switching probability, retention time and noise
strength on a (temperature x device) grid in ONE launch of the K4 kernel `thermal_analytics_kernel`, tensors in / tensors out;
`generate_temperature_sweep` and tensor arguments of the scalar methods run through it."""
from __future__ import annotations

import ctypes as C
import math
from typing import Any, Dict, Optional, Sequence, Tuple, Union

import numpy as np

from .. import _lib


class ThermalFluctuations:
    def __init__(self, temperature: float = 300.0, correlation_time: float = 1e-12, seed: Optional[int] = None,
                 num_devices: int = 1, device="cuda"):
        torch = _lib.require_cuda()
        self.temperature = temperature
        self.correlation_time = correlation_time
        self.k_b = 1.380649e-23
        self.mu_0 = 4 * np.pi * 1e-7
        self.seed = 0 if seed is None else int(seed)
        self.rng = np.random.default_rng(seed)        # host stream of sample_switching_time (thermal_model.py:36)
        self.num_devices = int(num_devices)
        self._device = torch.device(device)
        self._lib = _lib.load()
        self._previous_noise = torch.zeros(self.num_devices, 3, dtype=torch.float64, device=self._device)
        self._calls = 0

    def set_temperature(self, temperature: float) -> None:
        self.temperature = temperature

    def compute_noise_strength(self, damping: float, saturation_magnetization: float, volume: float,
                               gamma: float = 2.21e5) -> float:
        """physics/thermal_model.py:46-73."""
        if self.temperature <= 0:
            return 0.0
        variance = 2 * damping * self.k_b * self.temperature / (gamma * self.mu_0 * saturation_magnetization * volume)
        return math.sqrt(variance)

    def generate_thermal_field(self, damping: float, saturation_magnetization: float, volume: float, dt: float,
                               gamma: float = 2.21e5, correlated: bool = True):
        """Thermal field for every device: [num_devices, 3] float64 CUDA tensor ([3] NumPy when num_devices == 1)
        (physics/thermal_model.py:75-137)."""
        torch = _lib.require_cuda()
        s = self.compute_noise_strength(damping, saturation_magnetization, volume, gamma)
        out = torch.zeros(self.num_devices, 3, dtype=torch.float64, device=self._device)
        if s != 0:
            ou = correlated and self.correlation_time > 0
            decay = math.exp(-dt / self.correlation_time) if ou else 0.0
            with torch.cuda.device(self._device):
                _lib.check(self._lib.stg_thermal_field_f64(
                    s, decay, self._previous_noise.data_ptr() if ou else None, out.data_ptr(), self.seed, 0, self._calls,
                    self.num_devices, torch.cuda.current_stream(self._device).cuda_stream), "stg_thermal_field_f64")
            self._calls += 1
        return out[0].cpu().numpy() if self.num_devices == 1 else out

    # ---- Neel-Brown analytics (scalar; physics/thermal_model.py:139-258) ----------------------------------------------
    def compute_thermal_barrier(self, anisotropy_constant: float, volume: float) -> float:
        if self.temperature <= 0:
            return float("inf")
        return anisotropy_constant * volume / (self.k_b * self.temperature)

    def compute_switching_probability(self, energy_barrier: float, attempt_frequency: float = 1e9,
                                      measurement_time: float = 1e-9) -> float:
        if self.temperature <= 0:
            return 0.0
        rate = attempt_frequency * math.exp(-energy_barrier / (self.k_b * self.temperature))
        return min(1 - math.exp(-rate * measurement_time), 1.0)

    def sample_switching_time(self, energy_barrier: float, attempt_frequency: float = 1e9) -> float:
        """One exponential waiting time at the Neel-Brown rate (physics/thermal_model.py:185-207)."""
        if self.temperature <= 0:
            return float("inf")
        rate = attempt_frequency * np.exp(-energy_barrier / (self.k_b * self.temperature))
        if rate <= 0:
            return float("inf")
        return self.rng.exponential(1.0 / rate)

    def compute_retention_time(self, energy_barrier: float, failure_rate: float = 1e-9,
                               attempt_frequency: float = 1e9) -> float:
        if self.temperature <= 0 or failure_rate <= 0:
            return float("inf")
        thermal_factor = energy_barrier / (self.k_b * self.temperature)
        return -math.log(failure_rate) / (attempt_frequency * math.exp(-thermal_factor))

    def analyze_thermal_stability(self, device_params: dict, time_scale: float = 10.0) -> dict:
        volume = device_params.get("volume", 1e-24)
        k_u = device_params.get("uniaxial_anisotropy", 1e6)
        barrier = k_u * volume
        delta = self.compute_thermal_barrier(k_u, volume)
        secs = time_scale * 365.25 * 24 * 3600
        return {
            "thermal_stability_factor": delta, "energy_barrier_J": barrier,
            "energy_barrier_kT": barrier / (self.k_b * self.temperature),
            "switching_probability": self.compute_switching_probability(barrier, measurement_time=secs),
            "retention_time_years": self.compute_retention_time(barrier) / (365.25 * 24 * 3600),
            "is_thermally_stable": delta > 40, "temperature_K": self.temperature,
        }

    # ---- batched analytics (one K4 launch) -----------------------------------------------------------------------------------
    def batch_analytics(self, temperatures, device_params: Union[Dict[str, Any], Sequence[Dict[str, Any]]],
                        attempt_frequency: float = 1e9, measurement_time: float = 1e-9, failure_rate: float = 1e-9,
                        energy_barrier=None, gamma: float = 2.21e5) -> Dict[str, Any]:
        """Thermal properties on the grid temperatures [n_T] x devices [n_dev] in one kernel launch. `device_params`: one dict,
        a list of dicts, or a dict of equal-length arrays (keys volume, uniaxial_anisotropy, damping, saturation_magnetization;
        the reference's defaults apply, physics/thermal_model.py:300-303). `energy_barrier`: [n_dev] barriers in J (default
        K_u V). Returns float64 CUDA tensors [n_T, n_dev]: 'thermal_stability_factor', 'switching_probability' (over
        `measurement_time`), 'retention_time' (seconds, at `failure_rate`), 'noise_strength' - the formulas of
        compute_thermal_barrier / compute_switching_probability / compute_retention_time / compute_noise_strength."""
        torch = _lib.require_cuda()
        dev, f64 = self._device, torch.float64

        def vec(x):
            if isinstance(x, torch.Tensor):
                return x.to(device=dev, dtype=f64).reshape(-1).contiguous()
            return torch.as_tensor(np.asarray(x, dtype=np.float64).reshape(-1)).to(dev)

        defaults = (("volume", 1e-24), ("uniaxial_anisotropy", 1e6), ("damping", 0.01), ("saturation_magnetization", 800e3))
        if isinstance(device_params, dict):
            cols = {k: vec(device_params.get(k, d)) for k, d in defaults}
        else:
            cols = {k: vec([p.get(k, d) for p in device_params]) for k, d in defaults}
        n_dev = max(c.numel() for c in cols.values())
        cols = {k: (c.expand(n_dev).contiguous() if c.numel() == 1 else c) for k, c in cols.items()}
        if any(c.numel() != n_dev for c in cols.values()):
            raise ValueError("device parameter arrays must have the same length")
        temps = vec(temperatures)
        n_t = temps.numel()
        barrier = None
        if energy_barrier is not None:
            barrier = vec(energy_barrier)
            barrier = barrier.expand(n_dev).contiguous() if barrier.numel() == 1 else barrier
            if barrier.numel() != n_dev:
                raise ValueError("energy_barrier must have one entry per device")
        out = torch.empty(4, n_t, n_dev, dtype=f64, device=dev)
        a = _lib.StgThermalAnalyticsArgs()
        a.d_temperature, a.d_ku, a.d_volume = temps.data_ptr(), cols["uniaxial_anisotropy"].data_ptr(), cols["volume"].data_ptr()
        a.d_damping, a.d_ms, a.d_barrier = cols["damping"].data_ptr(), cols["saturation_magnetization"].data_ptr(), _lib.ptr(barrier)
        a.d_out, a.k_b, a.mu0, a.gamma = out.data_ptr(), self.k_b, self.mu_0, gamma
        a.attempt_frequency, a.measurement_time, a.failure_rate = attempt_frequency, measurement_time, failure_rate
        a.n_t, a.n_dev = n_t, n_dev
        with torch.cuda.device(dev):
            _lib.check(self._lib.stg_thermal_analytics_f64(C.byref(a), torch.cuda.current_stream(dev).cuda_stream),
                       "stg_thermal_analytics_f64")
        return {"thermal_stability_factor": out[0], "switching_probability": out[1], "retention_time": out[2],
                "noise_strength": out[3]}

    def generate_temperature_sweep(self, temp_range: Tuple[float, float], device_params: dict, n_points: int = 100) -> dict:
        """Stability factor, one-year switching probability, retention time (years) and noise strength on a temperature grid
        (physics/thermal_model.py:274-336) - the whole grid in one kernel launch; `device_params` may also be a list of dicts
        or a dict of arrays, which adds a device axis ([n_points, n_dev] arrays). The instance temperature is untouched."""
        year = 365.25 * 24 * 3600
        temperatures = np.linspace(temp_range[0], temp_range[1], n_points)
        r = self.batch_analytics(temperatures, device_params, measurement_time=year)
        single = isinstance(device_params, dict) and all(np.ndim(v) == 0 for v in device_params.values())
        out = {"temperature": temperatures}
        for k, v in r.items():
            x = v.cpu().numpy()
            if k == "retention_time":
                x = x / year
            out[k] = x[:, 0] if single else x
        return out
