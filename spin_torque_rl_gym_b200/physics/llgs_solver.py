"""Batched LLGSSolver (reference: physics/llgs_solver.py:20-305) on the K2 adaptive-RK45 CUDA kernel.

`solve()` keeps the reference signature for one trajectory and returns the same dict ('t', 'm', 'energy', 'torques', 'success');
`solve_batch()` integrates N trajectories in one launch. Python callables cannot cross into a kernel: current_func must be a
rectangular pulse, given as a number, a (J, t_pulse) tuple, or a callable that is sampled to recover (J, t_pulse) and verified
to be rectangular; field_func must be constant in time."""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Dict, Optional, Sequence, Tuple, Union

import numpy as np

from .. import _lib, params as _params


def _pulse_from_callable(fn: Callable[[float], float], t_end: float) -> Tuple[float, float]:
    """Recover (J, t_pulse) of current_func(t) = J if t <= t_pulse else 0 by bisection; reject anything else."""
    ts = np.linspace(0.0, t_end, 257)
    vals = np.array([float(fn(float(t))) for t in ts])
    j = vals[0]
    on = vals == j
    if on.all():
        return j, float("inf") if j != 0.0 else 0.0
    k = int(np.argmin(on))
    if not on[:k].all() or np.any(vals[k:] != 0.0):
        raise ValueError("current_func must be a rectangular pulse (J while t <= t_pulse, 0 afterwards) for the CUDA solver")
    lo, hi = float(ts[k - 1]), float(ts[k])
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if mid == lo or mid == hi:
            break
        if float(fn(mid)) == j:
            lo = mid
        else:
            hi = mid
    return j, lo


class LLGSSolver:
    def __init__(self, method: str = "RK45", rtol: float = 1e-6, atol: float = 1e-9, max_step: float = 1e-12,
                 gamma: float = 2.21e5, device: Any = "cuda", device_type: str = "stt_mram", sort_trajectories: bool = True):
        """`sort_trajectories`: batches of >= 64 trajectories are launched through a permutation sorted by (parameter set,
        t_end descending) so the lanes of a warp integrate trajectories of similar length (StgRk45Args.d_perm; results are
        identical, inputs and outputs are not moved)."""
        if method != "RK45":
            raise ValueError("only method='RK45' (SciPy's default Dormand-Prince pair) is implemented on the GPU")
        torch = _lib.require_cuda()
        self.method, self.rtol, self.atol, self.max_step, self.gamma = method, rtol, atol, max_step, gamma
        self.mu_0 = 4 * np.pi * 1e-7
        self.k_b = 1.380649e-23
        self.device_type = device_type
        self._device = torch.device(device)
        self._lib = _lib.load()
        self.sort_trajectories = bool(sort_trajectories)

    # -----------------------------------------------------------------------------------------------------------------
    def solve_batch(self, m_initial, t_end, device_params: Union[Dict[str, Any], Sequence[Dict[str, Any]]], current=0.0,
                    t_pulse=None, applied_field=None, voltage=None, thermal_noise: bool = False,
                    temperature: float = 300.0, param_index=None, device_type=None, current_direction=None,
                    return_trajectory: bool = False, max_traj_rows: Optional[int] = None, noise=None,
                    seed: int = 0, env_offset: int = 0) -> Dict[str, Any]:
        """N trajectories of LLGSSolver.solve in one launch. Arrays may be NumPy or torch; outputs are CUDA tensors."""
        torch = _lib.require_cuda()
        dev, f64 = self._device, torch.float64

        def arr(x, shape, fill=None):
            if x is None:
                return None if fill is None else torch.full(shape, fill, dtype=f64, device=dev)
            t = torch.as_tensor(np.asarray(x, dtype=np.float64)) if not isinstance(x, torch.Tensor) else x.to(f64)
            t = t.to(dev)
            return t.expand(shape).contiguous() if t.dim() < len(shape) or tuple(t.shape) != tuple(shape) else t.contiguous()

        m0 = torch.as_tensor(np.asarray(m_initial, dtype=np.float64)) if not isinstance(m_initial, torch.Tensor) \
            else m_initial.to(f64)
        m0 = m0.to(dev).reshape(-1, 3).contiguous()
        n = m0.shape[0]
        plist = [device_params] if isinstance(device_params, dict) else list(device_params)
        types = device_type if device_type is not None else self.device_type
        types = [types] * len(plist) if isinstance(types, str) else list(types)
        structs = [_params.make_llg_struct(t, p, thermal=thermal_noise, temperature=temperature,
                                           current_direction=current_direction, gamma=self.gamma)
                   for t, p in zip(types, plist)]
        table = torch.from_numpy(_params.llg_table(structs)).to(dev)
        pidx = None
        if param_index is not None:
            pidx = torch.as_tensor(param_index, dtype=torch.int32).to(dev).contiguous()
        elif len(structs) != 1:
            raise ValueError("several parameter sets need a param_index")
        a = _lib.StgRk45Args()
        t_end_t = arr(t_end, (n,))
        keep = [table, pidx, m0, t_end_t]
        if self.sort_trajectories and n >= 64:
            # lanes whose trajectory has ended idle until the slowest lane of the warp is done: launch in an order that puts
            # trajectories of the same parameter set and similar length into the same warp (longest first)
            perm = torch.argsort(t_end_t, descending=True, stable=True)
            if pidx is not None and len(structs) > 1:
                perm = perm[torch.argsort(pidx[perm], stable=True)]
            perm = perm.to(torch.int32).contiguous()
            keep.append(perm)
            a.d_perm = perm.data_ptr()
        a.d_table, a.d_param_index, a.d_m0, a.d_t_end = table.data_ptr(), _lib.ptr(pidx), m0.data_ptr(), t_end_t.data_ptr()
        for name, val, shape in (("d_current", current, (n,)), ("d_t_pulse", t_pulse, (n,)),
                                 ("d_happ", applied_field, (n, 3)), ("d_voltage", voltage, (n,))):
            t = arr(val, shape)
            keep.append(t)
            setattr(a, name, _lib.ptr(t))
        out = {
            "y": torch.empty(n, 3, dtype=f64, device=dev),
            "n_accepted": torch.zeros(n, dtype=torch.int32, device=dev),
            "n_rejected": torch.zeros(n, dtype=torch.int32, device=dev),
            "n_rhs": torch.zeros(n, dtype=torch.int32, device=dev),
            "status": torch.zeros(n, dtype=torch.int32, device=dev),
            "t_reached": torch.zeros(n, dtype=f64, device=dev),
        }
        a.d_y_out, a.d_n_accepted, a.d_n_rejected = out["y"].data_ptr(), out["n_accepted"].data_ptr(), out["n_rejected"].data_ptr()
        a.d_n_rhs, a.d_status, a.d_t_reached = out["n_rhs"].data_ptr(), out["status"].data_ptr(), out["t_reached"].data_ptr()
        if return_trajectory:
            if max_traj_rows is None:
                tmax = float(t_end_t.max())
                max_traj_rows = int(2.5 * tmax / self.max_step) + 64
            out["traj"] = torch.zeros(n, max_traj_rows, 6, dtype=f64, device=dev)
            a.d_traj, a.traj_stride = out["traj"].data_ptr(), max_traj_rows
        flags = 0
        if noise is not None:
            nz = arr(noise, tuple(np.shape(noise)) if not isinstance(noise, torch.Tensor) else tuple(noise.shape))
            if nz.dim() != 3 or nz.shape[0] != n or nz.shape[2] != 3:
                raise ValueError("noise must have shape [N, max_rhs_evaluations, 3]")
            keep.append(nz)
            a.d_noise, a.noise_stride = nz.data_ptr(), nz.shape[1]
            flags |= _lib.F_THERMAL_INJECT
        elif thermal_noise:
            flags |= _lib.F_THERMAL_PHILOX
        a.rtol, a.atol, a.max_step = self.rtol, self.atol, self.max_step
        a.seed, a.env_offset, a.n_envs, a.n_sets, a.flags = seed & 0xFFFFFFFFFFFFFFFF, env_offset, n, len(structs), flags
        with torch.cuda.device(dev):
            _lib.check(self._lib.stg_llgs_rk45_f64(C.byref(a), torch.cuda.current_stream(dev).cuda_stream),
                       "stg_llgs_rk45_f64")
        out["m"] = out["y"] / out["y"].norm(dim=1, keepdim=True)
        out["success"] = out["status"] == 0
        self._keep = keep
        return out

    # -----------------------------------------------------------------------------------------------------------------
    def solve(self, m_initial: np.ndarray, time_span: Tuple[float, float], device_params: Dict[str, Any],
              current_func: Union[Callable[[float], float], float, Tuple[float, float]],
              field_func: Optional[Callable[[float], np.ndarray]] = None, thermal_noise: bool = True,
              temperature: float = 300.0, seed: int = 0) -> Dict[str, np.ndarray]:
        """Reference signature, one trajectory (physics/llgs_solver.py:51-180)."""
        t0, t1 = float(time_span[0]), float(time_span[1])
        if t0 != 0.0:
            raise ValueError("time_span must start at 0 (the pulse and field are defined relative to the step start)")
        if callable(current_func):
            j, tp = _pulse_from_callable(current_func, t1)
        elif isinstance(current_func, (tuple, list)):
            j, tp = float(current_func[0]), float(current_func[1])
        else:
            j, tp = float(current_func), float("inf")
        happ = np.zeros(3)
        if field_func is not None:
            happ = np.asarray(field_func(0.0), dtype=float)
            if not np.array_equal(happ, np.asarray(field_func(t1), dtype=float)):
                raise ValueError("field_func must be constant in time for the CUDA solver")
        tp = min(tp, 1.0e300)
        r = self.solve_batch(np.asarray(m_initial, dtype=float)[None], np.array([t1]), device_params, current=np.array([j]),
                             t_pulse=np.array([tp]), applied_field=happ[None], thermal_noise=thermal_noise,
                             temperature=temperature, return_trajectory=True, seed=seed)
        rows = int(r["n_accepted"][0]) + 1
        traj = r["traj"][0, :rows].cpu().numpy()
        status = int(r["status"][0])
        if status & 2:
            raise RuntimeError("LLGS integration failed: trajectory buffer overflow")
        return {"t": traj[:, 0], "m": traj[:, 1:4], "energy": traj[:, 4], "torques": traj[:, 5], "success": status == 0,
                "n_accepted": rows - 1, "n_rejected": int(r["n_rejected"][0])}

    def find_stable_states(self, device_params: Dict[str, Any], n_trials: int = 100, threshold: float = 1e-6,
                           relax_time: float = 10e-9, seed: Optional[int] = None) -> np.ndarray:
        """Relax `n_trials` random starts for 10 ns without current / field / noise in ONE launch and return the distinct end
        states (physics/llgs_solver.py:264-305)."""
        rng = np.random.default_rng(seed)
        m0 = rng.normal(0, 1, (n_trials, 3))
        m0 /= np.linalg.norm(m0, axis=1, keepdims=True)
        r = self.solve_batch(m0, relax_time, device_params, current=0.0, thermal_noise=False)
        ok = r["success"].cpu().numpy()
        finals = r["m"].cpu().numpy()[ok]
        stable = []
        for m in finals:
            if all(np.linalg.norm(m - s) >= threshold for s in stable):
                stable.append(m)
        return np.array(stable) if stable else np.array([[0, 0, 1]])
