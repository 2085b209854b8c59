"""Batched EnergyLandscape.compute_energy / compute_energy_gradient (reference: physics/energy_landscape.py:16-104) on the GPU:
the step *before* the hot path (choosing targets / parameters, phase diagrams over (J, H) grids are embarrassingly parallel)."""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional

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
