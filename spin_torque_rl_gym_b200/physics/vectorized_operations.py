"""VectorizedSolver with the reference's batch API (utils/vectorized_operations.py:14-300) on the K1 solver kernel:
explicit Euler, no current, no thermal field, `n_steps = max(10, int(T/dt))`, one parameter dict per trajectory; and
VectorizedMagneticsOperations (utils/vectorized_operations.py:288-393) on the K4 row-wise kernel."""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Tuple

import numpy as np

from .. import _lib, params as _params


class VectorizedSolver:
    def __init__(self, max_batch_size: int = 1000, device: Any = "cuda", dtype: Any = None):
        torch = _lib.require_cuda()
        self.max_batch_size = max_batch_size
        self.gamma = 2.21e5
        self.mu_0 = 4 * np.pi * 1e-7
        self.vectorized_operations = 0
        self.batch_count = 0
        self._device = torch.device(device)
        self._dtype = torch.float64 if dtype is None else dtype
        self._lib = _lib.load()

    def solve_batch_tensors(self, m_initial_batch, t_span: Tuple[float, float], device_params_batch: List[Dict[str, Any]],
                            dt: float = 1e-12, return_trajectory: bool = True) -> Dict[str, Any]:
        """GPU-resident result: {'m_final' [N,3], 'traj' [N, n_steps+1, 3], 'n_steps'}. A single dict in
        device_params_batch is shared by every trajectory."""
        torch = _lib.require_cuda()
        dev, f64 = self._device, torch.float64
        m0 = torch.as_tensor(np.asarray(m_initial_batch, dtype=np.float64)) if not isinstance(m_initial_batch, torch.Tensor) \
            else m_initial_batch.to(f64)
        m0 = m0.to(dev).reshape(-1, 3).contiguous()
        n = m0.shape[0]
        dur = float(t_span[1]) - float(t_span[0])
        structs = [_params.make_param_struct("stt_mram", dict(p, polarization=p.get("polarization", 0.7)), max_steps=1,
                                             max_current=1.0, max_duration=1.0, temperature=300.0, thermal=False,
                                             success_threshold=0.9, energy_penalty_weight=0.1, max_step=dt)
                   for p in device_params_batch]
        folded = _params.fold(structs)
        table = torch.from_numpy(folded).to(dev)
        pidx = None
        if len(structs) > 1:
            if len(structs) != n:
                raise ValueError("device_params_batch must have one entry per trajectory (or a single shared entry)")
            pidx = torch.arange(n, dtype=torch.int32, device=dev)
        pulse = torch.zeros(n, 3, dtype=f64, device=dev)
        pulse[:, 1] = dur
        pulse[:, 2] = dur
        n_steps = max(10, int(dur / dt))
        out = {"m_final": torch.empty(n, 3, dtype=f64, device=dev), "n_steps": n_steps}
        a = _lib.StgSttSolveArgs()
        a.d_table, a.d_param_index, a.d_m0, a.d_pulse = table.data_ptr(), _lib.ptr(pidx), m0.data_ptr(), pulse.data_ptr()
        a.d_m_out = out["m_final"].data_ptr()
        if return_trajectory:
            out["traj"] = torch.zeros(n, n_steps + 1, 3, dtype=f64, device=dev)
            a.d_traj, a.traj_stride = out["traj"].data_ptr(), n_steps + 1
        flags = _lib.F_EULER | _lib.F_VECTORIZED_PLAN
        if _params.all_axis_z(folded):
            flags |= _lib.F_AXIS_Z
        a.n_envs, a.n_sets, a.flags = n, len(structs), flags
        fn = self._lib.stg_stt_solve_f64 if self._dtype == torch.float64 else self._lib.stg_stt_solve_f32
        with torch.cuda.device(dev):
            _lib.check(fn(C.byref(a), torch.cuda.current_stream(dev).cuda_stream), "stg_stt_solve")
        self._keep = (table, pidx, m0, pulse)
        self.batch_count += 1
        self.vectorized_operations += 1
        return out

    def solve_batch(self, m_initial_batch: np.ndarray, t_span: Tuple[float, float],
                    device_params_batch: List[Dict[str, Any]], dt: float = 1e-12) -> List[Dict[str, Any]]:
        """Reference return format: one dict per trajectory (utils/vectorized_operations.py:32-121)."""
        m_initial_batch = np.asarray(m_initial_batch, dtype=float)
        if m_initial_batch.shape[0] == 0:
            return []
        r = self.solve_batch_tensors(m_initial_batch, t_span, device_params_batch, dt)
        n_steps = r["n_steps"]
        t = np.linspace(t_span[0], t_span[1], n_steps + 1)
        traj = r["traj"].cpu().numpy()
        traj[:, 0, :] = m_initial_batch            # the reference stores the initial rows un-normalised
        return [{"t": t.copy(), "m": traj[j], "success": True, "message": "Vectorized integration completed",
                 "n_steps": n_steps, "vectorized": True} for j in range(traj.shape[0])]

    def solve_single(self, m_initial: np.ndarray, t_span: Tuple[float, float], device_params: Dict[str, Any], **kwargs):
        return self.solve_batch(np.asarray(m_initial, dtype=float).reshape(1, -1), t_span, [device_params])[0]


_VEC3 = {"cross": 0, "dot": 1, "normalize": 2, "anis_energy": 3, "tmr_resistance": 4}     # include/stg.h STG_VEC3_*


def _vec3_op(op: str, a, b=None, p0=None, p1=None, device: Any = "cuda"):
    """Run one STG_VEC3_* op. NumPy inputs come back as NumPy, CUDA tensors stay on the device."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    was_numpy = not isinstance(a, torch.Tensor)
    dev = torch.device(device) if was_numpy else a.device
    if dev.type != "cuda":
        raise _lib.StgError("VectorizedMagneticsOperations runs on CUDA tensors only (there is no CPU path)")

    def dev64(x, cols=None):
        if x is None:
            return None
        t = torch.as_tensor(np.asarray(x, dtype=np.float64)) if not isinstance(x, torch.Tensor) else x
        t = t.to(device=dev, dtype=torch.float64)
        return (t.reshape(-1, cols) if cols else t.reshape(-1)).contiguous()

    ta, tb = dev64(a, 3), dev64(b, 3)
    n = ta.shape[0]
    if tb is not None and tb.shape[0] not in (1, n):
        raise ValueError(f"second operand has {tb.shape[0]} rows, expected 1 or {n}")
    tp0, tp1 = dev64(p0), dev64(p1)
    for t in (tp0, tp1):
        if t is not None and t.numel() != n:
            raise ValueError(f"per-row parameter has {t.numel()} entries, expected {n}")
    out = torch.empty((n, 3) if op in ("cross", "normalize") else (n,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.stg_vec3_op_f64(_VEC3[op], ta.data_ptr(), _lib.ptr(tb), 0 if tb is None else tb.shape[0],
                                       _lib.ptr(tp0), _lib.ptr(tp1), out.data_ptr(), n,
                                       torch.cuda.current_stream(dev).cuda_stream), "stg_vec3_op_f64")
    return out.cpu().numpy() if was_numpy else out


class VectorizedMagneticsOperations:
    """Row-wise helpers over [N,3] batches with the reference's names and argument meaning
    (utils/vectorized_operations.py:288-393); results carry NumPy's roundings bit for bit."""

    @staticmethod
    def batch_cross_product(a_batch, b_batch):
        return _vec3_op("cross", a_batch, b_batch)

    @staticmethod
    def batch_dot_product(a_batch, b_batch):
        return _vec3_op("dot", a_batch, b_batch)

    @staticmethod
    def batch_normalize(vectors):
        return _vec3_op("normalize", vectors)

    @staticmethod
    def batch_energy_computation(m_batch, params_batch: Dict[str, Any]):
        """Anisotropy energy -K_u V (m.e)^2 per row; defaults K_u = 1e6, V = 1e-24, e = z (:332-364)."""
        n = int(np.prod(m_batch.shape[:-1]))
        k_u = params_batch.get("uniaxial_anisotropy", np.full(n, 1e6))
        volume = params_batch.get("volume", np.full(n, 1e-24))
        easy = params_batch.get("easy_axis", np.array([0.0, 0.0, 1.0]))
        return _vec3_op("anis_energy", m_batch, easy, k_u, volume)

    @staticmethod
    def batch_resistance_computation(m_batch, reference_magnetizations, r_p_batch, r_ap_batch):
        return _vec3_op("tmr_resistance", m_batch, reference_magnetizations, r_p_batch, r_ap_batch)
