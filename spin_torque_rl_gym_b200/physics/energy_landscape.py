"""Batched EnergyLandscape.compute_energy / compute_energy_gradient / generate_phase_diagram (reference:
physics/energy_landscape.py:16-104,282-359) on the GPU:
the step *before* the hot path (choosing targets / parameters, phase diagrams over (J, H) grids are embarrassingly parallel)."""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional, Tuple

import numpy as np

from .. import _lib


class EnergyLandscape:
    def __init__(self, device_params: Dict[str, Any], device: Any = "cuda"):
        torch = _lib.require_cuda()
        self.device_params = device_params
        self.mu_0 = 4 * np.pi * 1e-7
        self.ms = device_params.get("saturation_magnetization", 800e3)
        self.volume = device_params.get("volume", 1e-24)
        self.k_u = device_params.get("uniaxial_anisotropy", 1e6)
        self.easy_axis = np.asarray(device_params.get("easy_axis", np.array([0, 0, 1])), dtype=float)
        self.demag_factors = np.asarray(device_params.get("demag_factors", np.array([0, 0, 1])), dtype=float)
        self.k_b = 1.380649e-23
        self._device = torch.device(device)
        self._lib = _lib.load()

    def _call(self, magnetization, applied_field, want_energy, want_grad):
        torch = _lib.require_cuda()
        was_numpy = not isinstance(magnetization, torch.Tensor)
        m = torch.as_tensor(np.asarray(magnetization, dtype=np.float64)) if was_numpy else magnetization.to(torch.float64)
        m = m.to(self._device)
        single = m.dim() == 1
        m = m.reshape(-1, 3).contiguous()
        h = None
        if applied_field is not None:
            h = torch.as_tensor(np.asarray(applied_field, dtype=np.float64)) if not isinstance(applied_field, torch.Tensor) \
                else applied_field.to(torch.float64)
            h = h.to(self._device).reshape(-1, 3).contiguous()
        p = _lib.StgEnergyParams(self.mu_0, float(self.ms), float(self.volume), float(self.k_u),
                                 _lib.c_double3(*self.easy_axis), _lib.c_double3(*self.demag_factors))
        e = torch.empty(m.shape[0], dtype=torch.float64, device=self._device) if want_energy else None
        g = torch.empty_like(m) if want_grad else None
        with torch.cuda.device(self._device):
            _lib.check(self._lib.stg_energy_landscape_f64(
                C.byref(p), m.data_ptr(), _lib.ptr(h), 0 if h is None else h.shape[0], _lib.ptr(e), _lib.ptr(g), m.shape[0],
                torch.cuda.current_stream(self._device).cuda_stream), "stg_energy_landscape_f64")
        out = e if want_energy else g
        if single:
            return float(out[0]) if want_energy else out[0].cpu().numpy()
        return out.cpu().numpy() if was_numpy else out

    def compute_energy(self, magnetization, applied_field=None, current: float = 0.0):
        """Total energy (J) of one state [3] or a batch [N,3] (physics/energy_landscape.py:36-71)."""
        return self._call(magnetization, applied_field, True, False)

    def compute_energy_gradient(self, magnetization, applied_field=None):
        """Effective field of one state or a batch (physics/energy_landscape.py:73-104)."""
        return self._call(magnetization, applied_field, False, True)

    def compute_energy_barrier(self, initial_state, final_state, applied_field=None, n_intermediate: int = 50):
        """(barrier height, energy path) along the normalised straight-line path between two states, all points in one
        launch of the energy kernel (physics/energy_landscape.py:179-221)."""
        a, b = np.asarray(initial_state, dtype=float), np.asarray(final_state, dtype=float)
        t = np.linspace(0, 1, n_intermediate)[:, None]
        path = (1 - t) * a + t * b
        path = path / np.linalg.norm(path, axis=1, keepdims=True)
        happ = np.zeros(3) if applied_field is None else np.asarray(applied_field, dtype=float)
        energy_path = np.asarray(self._call(path, happ.reshape(1, 3), True, False))
        return np.max(energy_path) - min(energy_path[0], energy_path[-1]), energy_path

    def generate_phase_diagram(self, current_range: Tuple[float, float], field_range: Tuple[float, float],
                               resolution: int = 50, save_path: Optional[str] = None) -> Dict[str, np.ndarray]:
        """Switching map over a (current, field) grid: 1 where |H| exceeds h_k - |beta I| (physics/energy_landscape.py:282-340).
        `switching_probability[i, j]` belongs to `fields[i]`, `currents[j]`. The reference also draws a matplotlib figure;
        plotting is outside the path, so `save_path` is accepted and ignored."""
        torch = _lib.require_cuda()
        currents = np.linspace(current_range[0], current_range[1], resolution)
        fields = np.linspace(field_range[0], field_range[1], resolution)
        beta = self.device_params.get("polarization", 0.7) * 2.21e5 / (2 * self.ms * self.volume)
        h_k = 2 * self.k_u / (self.mu_0 * self.ms)
        d_i = torch.as_tensor(currents).to(self._device)
        d_h = torch.as_tensor(fields).to(self._device)
        out = torch.empty(resolution, resolution, dtype=torch.float64, device=self._device)
        with torch.cuda.device(self._device):
            _lib.check(self._lib.stg_phase_diagram_f64(
                d_i.data_ptr(), d_h.data_ptr(), resolution, resolution, float(beta), float(h_k), out.data_ptr(),
                torch.cuda.current_stream(self._device).cuda_stream), "stg_phase_diagram_f64")
        return {"currents": currents, "fields": fields, "switching_probability": out.cpu().numpy()}

    def compute_thermal_stability_factor(self, temperature: float = 300.0) -> float:
        """Delta = K_u V / (k_B T) (physics/energy_landscape.py:342-359)."""
        if temperature <= 0:
            return float("inf")
        return self.k_u * self.volume / (self.k_b * temperature)
