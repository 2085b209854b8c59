"""Batched SimpleLLGSSolver / RobustLLGSSolver (reference: physics/simple_solver.py:21-399, utils/robust_solver.py:22-345) on
the K1 fixed-step kernels (stg_stt_solve_*). Reference signature for one trajectory; `solve_batch` for N trajectories.

Python callables cannot cross into a kernel: `current_func` is a number, a (J, t_pulse) tuple, or a callable. A callable that
samples as a rectangular pulse and a `field_func` that is constant run the fast kernels (the env passes exactly that,
envs/spin_torque_env.py:442-447); anything else is sampled by the host at the stage times the reference evaluates it at
(`stage_times`) and integrated by the FP64 grid kernel (include/stg.h, StgSttSolveArgs.d_current_grid / d_field_grid)."""
from __future__ import annotations

import ctypes as C
import time
import warnings
from typing import Any, Callable, Dict, Optional, Tuple, Union

import numpy as np

from .. import _lib, params as _params
from .llgs_solver import _pulse_from_callable


class SimpleLLGSSolver:
    def __init__(self, method: str = "euler", rtol: float = 1e-3, atol: float = 1e-6, max_step: float = 1e-12,
                 timeout: float = 2.0, device: Any = "cuda", dtype: Any = None):
        torch = _lib.require_cuda()
        self.method = method.lower()
        if self.method not in ("euler", "rk4"):
            warnings.warn(f"Unknown method '{method}', using 'euler'")
            self.method = "euler"
        self.rtol, self.atol, self.max_step = rtol, atol, max_step
        self.timeout = timeout            # accepted for API compatibility; there is no wall-clock fallback on the GPU
        self.gamma = 2.21e5
        self.mu_0 = 4 * np.pi * 1e-7
        self.solve_count = 0
        self.timeout_count = 0
        self.last_solve_time = 0.0
        self._device = torch.device(device)
        self._dtype = torch.float64 if dtype is None else dtype
        self._lib = _lib.load()

    def solve_batch(self, m_initial, t_end, device_params: Dict[str, Any], current=0.0, t_pulse=None, applied_field=None,
                    thermal_noise: bool = False, temperature: float = 300.0, return_trajectory: bool = False, noise=None,
                    seed: int = 0, env_offset: int = 0, device_type: str = "stt_mram", current_grid=None,
                    field_grid=None) -> Dict[str, Any]:
        """Batched fixed-step solve. `current_grid` [n_sub,3] or [N,n_sub,3] and `field_grid` [n_sub,3,3] or [N,n_sub,3,3]
        carry current_func / field_func sampled at the stage times (t_i, t_i+dt/2, t_i+dt), see `stage_times`; they select
        the FP64 general-geometry kernel (include/stg.h, StgSttSolveArgs)."""
        torch = _lib.require_cuda()
        dev, f64 = self._device, torch.float64

        def arr(x, shape):
            t = torch.as_tensor(np.asarray(x, dtype=np.float64)) if not isinstance(x, torch.Tensor) else x.to(f64)
            t = t.to(dev)
            return (t.expand(shape) if tuple(t.shape) != tuple(shape) else t).contiguous()

        m0 = torch.as_tensor(np.asarray(m_initial, dtype=np.float64)) if not isinstance(m_initial, torch.Tensor) \
            else m_initial.to(f64)
        m0 = m0.to(dev).reshape(-1, 3).contiguous()
        n = m0.shape[0]
        te = arr(t_end, (n,))
        j = arr(current, (n,))
        tp = te if t_pulse is None else arr(t_pulse, (n,))
        pulse = torch.stack([j, tp, te], dim=1).contiguous()
        st = _params.make_param_struct(device_type, device_params, max_steps=1, max_current=1.0, max_duration=1.0,
                                       temperature=temperature, thermal=thermal_noise, success_threshold=0.9,
                                       energy_penalty_weight=0.1,
                                       applied_field=(0, 0, 0) if applied_field is None else applied_field,
                                       max_step=self.max_step)
        folded = _params.fold([st])
        table = torch.from_numpy(folded).to(dev)
        a = _lib.StgSttSolveArgs()
        out = {"m": torch.empty(n, 3, dtype=f64, device=dev), "n_steps": torch.zeros(n, dtype=torch.int32, device=dev),
               "guard": torch.zeros(n, dtype=torch.int32, device=dev)}
        a.d_table, a.d_m0, a.d_pulse, a.d_m_out = table.data_ptr(), m0.data_ptr(), pulse.data_ptr(), out["m"].data_ptr()
        a.d_n_sub, a.d_guard = out["n_steps"].data_ptr(), out["guard"].data_ptr()
        flags = _lib.F_EULER if self.method == "euler" else 0
        if _params.all_axis_z(folded):
            flags |= _lib.F_AXIS_Z
        keep = [table, m0, pulse]
        if return_trajectory:
            rows = int(np.ceil(float(te.max()) / min(self.max_step, float(te.max()) / 100))) + 16
            out["traj"] = torch.zeros(n, rows, 3, dtype=f64, device=dev)
            a.d_traj, a.traj_stride = out["traj"].data_ptr(), rows
        if noise is not None:
            nz = arr(noise, tuple(noise.shape))
            keep.append(nz)
            a.d_noise, a.noise_stride = nz.data_ptr(), nz.shape[1]
            flags |= _lib.F_THERMAL_INJECT
        elif thermal_noise and temperature > 0:
            flags |= _lib.F_THERMAL_PHILOX
        for name, g, tail in (("d_current_grid", current_grid, (3,)), ("d_field_grid", field_grid, (3, 3))):
            if g is None:
                continue
            gt = torch.as_tensor(np.asarray(g, dtype=np.float64)) if not isinstance(g, torch.Tensor) else g.to(f64)
            gt = gt.to(dev)
            if gt.dim() == 1 + len(tail):
                gt = gt[None]
            if tuple(gt.shape[2:]) != tail or gt.shape[0] not in (1, n):
                raise ValueError(f"{name[2:]} must be [n_sub, {tail}] or [{n}, n_sub, {tail}], got {tuple(gt.shape)}")
            if a.grid_stride and (gt.shape[1] != a.grid_stride or gt.shape[0] != a.grid_envs):
                raise ValueError("current_grid and field_grid must cover the same trajectories and substeps")
            gt = gt.contiguous()
            keep.append(gt)
            setattr(a, name, gt.data_ptr())
            a.grid_stride, a.grid_envs = gt.shape[1], gt.shape[0]
        a.seed, a.env_offset, a.n_envs, a.n_sets, a.flags = seed & 0xFFFFFFFFFFFFFFFF, env_offset, n, 1, flags
        fn = self._lib.stg_stt_solve_f64 if self._dtype == torch.float64 else self._lib.stg_stt_solve_f32
        with torch.cuda.device(dev):
            _lib.check(fn(C.byref(a), torch.cuda.current_stream(dev).cuda_stream), "stg_stt_solve")
        self._keep = keep
        self.solve_count += n
        return out

    def stage_times(self, t0: float, t1: float) -> np.ndarray:
        """[n_sub, 3] times the reference evaluates current_func / field_func at: t_i, t_i + dt/2, t_i + dt with
        dt = min(max_step, T/100), n = max(10, int(T/dt)), dt = T/n, t = linspace(t0, t1, n+1)
        (physics/simple_solver.py:137-145, 263-295; Euler uses column 0 only)."""
        dur = t1 - t0
        dt = min(self.max_step, dur / 100)
        n = max(10, int(dur / dt))
        dt = dur / n
        t = np.linspace(t0, t1, n + 1)[:n]
        return np.stack([t, t + dt / 2, t + dt], axis=1)

    def solve(self, m_initial: np.ndarray, time_span: Tuple[float, float], device_params: Dict[str, Any],
              current_func: Union[Callable[[float], float], float, Tuple[float, float], None] = None,
              field_func: Optional[Callable[[float], np.ndarray]] = None, thermal_noise: bool = False,
              temperature: float = 300.0) -> Dict[str, Any]:
        """One trajectory, reference return dict (physics/simple_solver.py:184-191)."""
        t_wall = time.time()
        m_initial = np.asarray(m_initial, dtype=float)
        if m_initial.shape != (3,):
            raise ValueError("Magnetization must be a 3D numpy array")
        t0, t1 = float(time_span[0]), float(time_span[1])
        if t1 <= t0:                                    # trivial solution (:122-123, 231-240)
            mag = np.linalg.norm(m_initial)
            m = m_initial / mag if np.isfinite(m_initial).all() and mag >= 1e-12 else np.array([0.0, 0.0, 1.0])
            return {"t": np.array([t0, t1]), "m": np.array([m, m]), "success": True,
                    "message": "Trivial solution (zero time span)", "solve_time": 0.0, "n_steps": 1}
        dur = t1 - t0
        times = None
        jgrid = hgrid = None
        if current_func is None:
            j, tp = 0.0, dur
        elif callable(current_func):
            try:
                j, tp = _pulse_from_callable(lambda t: current_func(t + t0), dur)
            except ValueError:      # not a rectangular pulse: sample it where the reference evaluates it
                times = self.stage_times(t0, t1)
                jgrid = np.array([[float(current_func(float(x))) for x in row] for row in times])
                j, tp = 0.0, dur
        elif isinstance(current_func, (tuple, list)):
            j, tp = float(current_func[0]), float(current_func[1])
        else:
            j, tp = float(current_func), dur
        happ = (0.0, 0.0, 0.0)
        if field_func is not None:
            times = self.stage_times(t0, t1) if times is None else times
            hs = np.array([[np.asarray(field_func(float(x)), dtype=float) for x in row] for row in times])
            if hs.shape[2:] != (3,):
                raise ValueError("field_func must return a 3-vector")
            if (hs == hs[0, 0]).all():
                happ = hs[0, 0]                         # constant on every evaluated time: folded into the table
            else:
                hgrid = hs
        r = self.solve_batch(m_initial[None], np.array([dur]), device_params, current=np.array([j]),
                             t_pulse=np.array([min(tp, dur * 4)]), applied_field=happ, thermal_noise=thermal_noise,
                             temperature=temperature, return_trajectory=True, current_grid=jgrid, field_grid=hgrid)
        n = int(r["n_steps"][0])
        m = r["traj"][0, : n + 1].cpu().numpy()
        self.last_solve_time = time.time() - t_wall
        return {"t": np.linspace(t0, t1, n + 1), "m": m, "success": True, "message": "Integration completed successfully",
                "solve_time": self.last_solve_time, "n_steps": n, "guard": bool(r["guard"][0])}

    def get_solver_info(self) -> Dict[str, Any]:
        return {"method": self.method, "solve_count": self.solve_count, "timeout_count": self.timeout_count,
                "last_solve_time": self.last_solve_time, "timeout_rate": 0.0, "avg_solve_time": self.last_solve_time}


class RobustLLGSSolver(SimpleLLGSSolver):
    """utils/robust_solver.py:22-345: same integrator plus input validation; a failed validation returns the fallback result
    (success=False, trajectory = initial state) instead of raising."""

    def __init__(self, method: str = "euler", rtol: float = 1e-3, atol: float = 1e-6, max_step: float = 1e-12,
                 timeout: float = 2.0, max_retries: int = 3, fallback_method: str = "euler", enable_monitoring: bool = True,
                 enable_validation: bool = True, **kw):
        super().__init__(method, rtol, atol, max_step, timeout, **kw)
        self.max_retries, self.fallback_method = max_retries, fallback_method
        self.enable_validation = enable_validation
        self.stats = {"total_solves": 0, "successful_solves": 0, "failed_solves": 0, "validation_errors": 0}

    def solve(self, m_initial, t_span, device_params, current_func=None, field_func=None, thermal_noise=False,
              temperature=300.0, **kwargs):
        self.stats["total_solves"] += 1
        m_initial = np.asarray(m_initial, dtype=float)
        ok = m_initial.shape == (3,) and np.isfinite(m_initial).all() and np.linalg.norm(m_initial) >= 1e-12 \
            and t_span[1] > t_span[0] and _params.solver_accepts(device_params, temperature)
        if self.enable_validation and not ok:
            self.stats["failed_solves"] += 1
            self.stats["validation_errors"] += 1
            n_points = max(2, int((t_span[1] - t_span[0]) / self.max_step))
            return {"t": np.linspace(t_span[0], t_span[1], n_points), "m": np.tile(m_initial, (n_points, 1)),
                    "success": False, "message": "Fallback result: input validation failed", "solve_time": 0.0,
                    "is_fallback": True}
        res = super().solve(m_initial, t_span, device_params, current_func, field_func, thermal_noise, temperature)
        if res.get("guard"):
            # a trajectory row failed validation in the reference => whole solve discarded (utils/robust_solver.py:192-205)
            res["success"] = False
        self.stats["successful_solves" if res["success"] else "failed_solves"] += 1
        return res
