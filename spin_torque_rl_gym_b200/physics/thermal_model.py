"""Batched ThermalFluctuations (reference: physics/thermal_model.py:12-336). The white / Ornstein-Uhlenbeck field generator
runs on the GPU for `num_devices` independent devices (Philox stream); the Neel-Brown analytics are scalar formulas."""
from __future__ import annotations

import math
from typing import Optional, Tuple

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

    def generate_temperature_sweep(self, temp_range: Tuple[float, float], device_params: dict, n_points: int = 100) -> dict:
        """Stability factor, one-year switching probability, retention time (years) and noise strength on a temperature
        grid; the instance temperature is restored afterwards (physics/thermal_model.py:274-336)."""
        year = 365.25 * 24 * 3600
        temperatures = np.linspace(temp_range[0], temp_range[1], n_points)
        volume = device_params.get("volume", 1e-24)
        k_u = device_params.get("uniaxial_anisotropy", 1e6)
        damping = device_params.get("damping", 0.01)
        ms = device_params.get("saturation_magnetization", 800e3)
        keep = self.temperature
        cols = {"thermal_stability_factor": [], "switching_probability": [], "retention_time": [], "noise_strength": []}
        for temp in temperatures:
            self.set_temperature(temp)
            barrier = k_u * volume
            cols["thermal_stability_factor"].append(self.compute_thermal_barrier(k_u, volume))
            cols["switching_probability"].append(self.compute_switching_probability(barrier, measurement_time=year))
            cols["retention_time"].append(self.compute_retention_time(barrier) / year)
            cols["noise_strength"].append(self.compute_noise_strength(damping, ms, volume))
        self.set_temperature(keep)
        return {"temperature": temperatures, **{k: np.array(v) for k, v in cols.items()}}
