"""Batched LLGSSolver (reference: physics/llgs_solver.py:20-305) on the K2 adaptive-RK45 CUDA kernel.

`solve()` keeps the reference signature for one trajectory and returns the same dict ('t', 'm', 'energy', 'torques', 'success');
`solve_batch()` integrates N trajectories in one launch. Python callables cannot cross into a kernel and an adaptive controller
evaluates them at data-dependent times, so current_func / field_func must be PIECEWISE CONSTANT in time: a number / vector, a
(J, t_pulse) tuple, a `PiecewiseConstant` table, or any callable that turns out to be piecewise constant on the time span (its
break points are located by bisection to adjacent doubles and the table is evaluated inside the kernel). Anything else - a
ramp, a sine - raises ValueError instead of being approximated."""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Dict, Optional, Sequence, Tuple, Union

import numpy as np

from .. import _lib, params as _params

_MAX_SEGMENTS = 64


class PiecewiseConstant:
    """f(t) = values[k] for breaks[k-1] < t <= breaks[k] (right-closed like the env's `J if t <= t_pulse else 0`), values[-1]
    after the last break. Callable, so the same object drives the reference's LLGSSolver.solve."""

    def __init__(self, breaks: Sequence[float], values: Sequence[Any]):
        self.breaks = np.asarray(breaks, dtype=np.float64).reshape(-1)
        self.values = np.asarray(values, dtype=np.float64)
        if self.values.shape[0] != self.breaks.size + 1:
            raise ValueError("PiecewiseConstant needs len(values) == len(breaks) + 1")
        if np.any(np.diff(self.breaks) <= 0):
            raise ValueError("PiecewiseConstant breaks must be strictly ascending")

    def __call__(self, t: float):
        k = int(np.searchsorted(self.breaks, t, side="left"))       # first break >= t
        v = self.values[k]
        return float(v) if v.ndim == 0 else v.copy()


def _table_from_callable(fn: Callable[[float], Any], t0: float, t1: float, what: str) -> PiecewiseConstant:
    """Recover the break points of a piecewise-constant callable on [t0, t1]; ValueError if it is not one."""
    ts = np.linspace(t0, t1, 1025)
    vals = [np.asarray(fn(float(t)), dtype=np.float64) for t in ts]
    change = [i for i in range(1, len(ts)) if not np.array_equal(vals[i], vals[i - 1])]
    if len(change) > _MAX_SEGMENTS:
        raise ValueError(f"{what} is not piecewise constant in time (more than {_MAX_SEGMENTS} changes on the time span): an "
                         "adaptive solver evaluates it at data-dependent times, which a CUDA kernel cannot do for a Python callable")
    breaks, values = [], [vals[0]]
    for i in change:
        lo, hi = float(ts[i - 1]), float(ts[i])
        v_lo = vals[i - 1]
        while True:                                   # last time with the left value, to adjacent doubles
            mid = 0.5 * (lo + hi)
            if mid <= lo or mid >= hi:
                break
            if np.array_equal(np.asarray(fn(mid), dtype=np.float64), v_lo):
                lo = mid
            else:
                hi = mid
        if not np.array_equal(np.asarray(fn(hi), dtype=np.float64), vals[i]):
            raise ValueError(f"{what} changes more than once between two sample times: not representable as a table")
        breaks.append(lo)
        values.append(vals[i])
    table = PiecewiseConstant(breaks, np.array(values))
    rng = np.random.default_rng(0)
    edges = [t0] + breaks + [t1]
    for k in range(len(edges) - 1):                   # constant inside every segment?
        for t in rng.uniform(edges[k], edges[k + 1], 4):
            if np.nextafter(edges[k], np.inf) < t <= edges[k + 1] and not np.array_equal(
                    np.asarray(fn(float(t)), dtype=np.float64), np.asarray(table(float(t)), dtype=np.float64)):
                raise ValueError(f"{what} varies inside a segment: not piecewise constant in time")
    return table


def _as_table(x, t0: float, t1: float, what: str, vector: bool) -> PiecewiseConstant:
    if x is None:
        return PiecewiseConstant([], np.zeros((1, 3)) if vector else np.zeros(1))
    if isinstance(x, PiecewiseConstant):
        return x
    if callable(x):
        return _table_from_callable(x, t0, t1, what)
    if not vector and isinstance(x, (tuple, list)) and len(x) == 2:
        return PiecewiseConstant([float(x[1])], [float(x[0]), 0.0])          # (J, t_pulse)
    v = np.asarray(x, dtype=np.float64)
    return PiecewiseConstant([], v.reshape(1, 3) if vector else v.reshape(1))


def _pulse_from_callable(fn: Callable[[float], float], t_end: float) -> Tuple[float, float]:
    """(J, t_pulse) of current_func(t) = J if t <= t_pulse else 0 (physics/simple_solver.py mirror); ValueError for anything else."""
    tab = _table_from_callable(fn, 0.0, t_end, "current_func")
    if tab.breaks.size == 0:
        j = float(tab.values[0])
        return j, float("inf") if j != 0.0 else 0.0
    if tab.breaks.size != 1 or float(tab.values[1]) != 0.0:
        raise ValueError("current_func must be a rectangular pulse (J while t <= t_pulse, 0 afterwards) for the CUDA solver")
    return float(tab.values[0]), float(tab.breaks[0])


def _merge_tables(cur: PiecewiseConstant, fld: PiecewiseConstant):
    """Common break points of the two tables -> (ends [K], current [K+1], field [K+1, 3])."""
    ends = np.union1d(cur.breaks, fld.breaks)
    probe = np.concatenate([ends, [np.inf]])          # a time inside every segment: its (closed) right end
    j = np.array([float(cur(t)) if np.isfinite(t) else float(cur.values[-1]) for t in probe])
    h = np.array([np.asarray(fld(t)) if np.isfinite(t) else fld.values[-1] for t in probe]).reshape(-1, 3)
    return ends, j, h


class LLGSSolver:
    def __init__(self, method: str = "RK45", rtol: float = 1e-6, atol: float = 1e-9, max_step: float = 1e-12,
                 gamma: float = 2.21e5, device: Any = "cuda", device_type: str = "stt_mram", sort_trajectories: bool = True):
        """`sort_trajectories`: batches of >= 64 trajectories are launched through a permutation sorted by (parameter set,
        t_end descending) so the lanes of a warp integrate trajectories of similar length (StgRk45Args.d_perm; results are
        identical, inputs and outputs are not moved)."""
        if method != "RK45":
            raise ValueError("only method='RK45' (SciPy's default Dormand-Prince pair, the reference's default) is implemented on "
                             "the GPU; RK23 / DOP853 / Radau / BDF / LSODA are not")
        torch = _lib.require_cuda()
        self.method, self.rtol, self.atol, self.max_step, self.gamma = method, rtol, atol, max_step, gamma
        self.mu_0 = 4 * np.pi * 1e-7
        self.k_b = 1.380649e-23
        self.device_type = device_type
        self._device = torch.device(device)
        self._lib = _lib.load()
        self.sort_trajectories = bool(sort_trajectories)
        self._tables: Dict[bytes, Any] = {}

    # -----------------------------------------------------------------------------------------------------------------
    def solve_batch(self, m_initial, t_end, device_params: Union[Dict[str, Any], Sequence[Dict[str, Any]]], current=0.0,
                    t_pulse=None, applied_field=None, voltage=None, thermal_noise: bool = False,
                    temperature: float = 300.0, param_index=None, device_type=None, current_direction=None,
                    return_trajectory: bool = False, max_traj_rows: Optional[int] = None, noise=None,
                    seed: int = 0, env_offset: int = 0, t_start=None, segments=None, host_outputs: bool = False) -> Dict[str, Any]:
        """N trajectories of LLGSSolver.solve in one launch. Arrays may be NumPy or torch; outputs are CUDA tensors.
        `t_start`: [N] or scalar start times (default 0; t_end is absolute). `segments`: piecewise-constant controls
        (ends [K] or [N,K], current [K+1] or [N,K+1], field [K+1,3] or [N,K+1,3] or None) replacing current / t_pulse / applied_field
        (include/stg.h StgRk45Args.d_seg_*). `host_outputs`: end states, counters and status are written by the kernel straight
        into pinned host memory (CPU tensors are returned, the stream is synchronised before returning)."""
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
        # the device copy of the parameter table is kept across calls with the same parameters (a pageable H2D copy per call
        # otherwise: the host path is a visible share of a 2 ms solve)
        raw = _params.llg_table(structs)
        key = raw.tobytes()
        table = self._tables.get(key)
        if table is None:
            if len(self._tables) >= 16:
                self._tables.clear()
            table = self._tables[key] = torch.from_numpy(raw).to(dev)
        pidx = None
        if param_index is not None:
            pidx = torch.as_tensor(param_index, dtype=torch.int32).to(dev).contiguous()
        elif len(structs) != 1:
            raise ValueError("several parameter sets need a param_index")
        a = _lib.StgRk45Args()
        t_end_t = arr(t_end, (n,))
        keep = [table, pidx, m0, t_end_t]
        a.d_table, a.d_param_index, a.d_m0, a.d_t_end = table.data_ptr(), _lib.ptr(pidx), m0.data_ptr(), t_end_t.data_ptr()
        for name, val, shape in (("d_current", current, (n,)), ("d_t_pulse", t_pulse, (n,)),
                                 ("d_happ", applied_field, (n, 3)), ("d_voltage", voltage, (n,))):
            t = arr(val, shape)
            keep.append(t)
            setattr(a, name, _lib.ptr(t))
        t_start_t = arr(t_start, (n,))
        keep.append(t_start_t)
        a.d_t_start = _lib.ptr(t_start_t)
        if segments is not None:
            ends, seg_j, seg_h = segments
            ends = torch.as_tensor(np.asarray(ends, dtype=np.float64)).to(dev)
            per_env = ends.dim() == 2
            k = ends.shape[-1]
            if k > 0:
                rows = n if per_env else 1
                ends = ends.reshape(rows, k).contiguous()
                seg_j = torch.as_tensor(np.asarray(seg_j, dtype=np.float64)).to(dev).reshape(rows, k + 1).contiguous()
                keep += [ends, seg_j]
                a.d_seg_t, a.d_seg_current, a.n_seg, a.seg_rows = ends.data_ptr(), seg_j.data_ptr(), k, rows
                if seg_h is not None:
                    seg_h = torch.as_tensor(np.asarray(seg_h, dtype=np.float64)).to(dev).reshape(rows, k + 1, 3).contiguous()
                    keep.append(seg_h)
                    a.d_seg_field = seg_h.data_ptr()
        a.rtol, a.atol, a.max_step = self.rtol, self.atol, self.max_step
        a.n_envs, a.n_sets = n, len(structs)
        if self.sort_trajectories and n >= 64:
            # lanes whose trajectory has ended idle until the slowest lane of the warp is done: launch in an order that puts
            # trajectories of the same parameter set and similar estimated cost into the same warp (most expensive first).
            # The estimate is span length x max(1/max_step, 8 (gamma |H_eff(m0)| + torque rate)); counting sort on the device
            # (stg_llgs_rk45_sort_f64: three small launches, no host synchronisation).
            perm = torch.empty(n, dtype=torch.int32, device=dev)
            work = torch.empty(_lib.SORT_WORK_INTS, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(self._lib.stg_llgs_rk45_sort_f64(C.byref(a), perm.data_ptr(), work.data_ptr(),
                                                            torch.cuda.current_stream(dev).cuda_stream), "stg_llgs_rk45_sort_f64")
            keep.append(work)
            keep.append(perm)
            a.d_perm = perm.data_ptr()
        def out_buf(shape, dtype):                                          # every entry is written by the kernel
            return torch.empty(shape, dtype=dtype).pin_memory() if host_outputs else torch.empty(shape, dtype=dtype, device=dev)
        counts = out_buf((4, n), torch.int32)
        out = {
            "y": out_buf((n, 3), f64),
            "n_accepted": counts[0], "n_rejected": counts[1], "n_rhs": counts[2], "status": counts[3],
            "t_reached": out_buf((n,), f64),
        }
        a.d_y_out, a.d_t_reached = _lib.ptr(out["y"]), _lib.ptr(out["t_reached"])
        cbase = _lib.ptr(counts)
        a.d_n_accepted, a.d_n_rejected, a.d_n_rhs, a.d_status = cbase, cbase + 4 * n, cbase + 8 * n, cbase + 12 * n
        if return_trajectory:
            if max_traj_rows is None:
                tmax = float((t_end_t - t_start_t).max()) if t_start_t is not None else float(t_end_t.max())
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
        if host_outputs:
            torch.cuda.current_stream(dev).synchronize()
        out["m"] = out["y"] / out["y"].norm(dim=1, keepdim=True)
        out["success"] = out["status"] == 0
        self._keep = keep
        return out

    # -----------------------------------------------------------------------------------------------------------------
    def solve(self, m_initial: np.ndarray, time_span: Tuple[float, float], device_params: Dict[str, Any],
              current_func: Union[Callable[[float], float], float, Tuple[float, float]],
              field_func: Optional[Callable[[float], np.ndarray]] = None, thermal_noise: bool = True,
              temperature: float = 300.0, seed: int = 0) -> Dict[str, np.ndarray]:
        """Reference signature, one trajectory (physics/llgs_solver.py:51-180). current_func / field_func: see the module
        docstring (piecewise constant in time, or ValueError)."""
        t0, t1 = float(time_span[0]), float(time_span[1])
        if not t1 > t0:
            raise ValueError("time_span must run forward (t_end > t_start)")
        cur = _as_table(current_func, t0, t1, "current_func", vector=False)
        fld = _as_table(field_func, t0, t1, "field_func", vector=True)
        if fld.values.ndim != 2 or fld.values.shape[1] != 3:
            raise ValueError("field_func must return a 3-vector")
        ends, seg_j, seg_h = _merge_tables(cur, fld)
        kw: Dict[str, Any] = {}
        if ends.size == 0:
            kw = dict(current=np.array([seg_j[0]]), applied_field=seg_h[:1])
        elif ends.size == 1 and seg_j[1] == 0.0 and np.array_equal(seg_h[0], seg_h[1]):
            kw = dict(current=np.array([seg_j[0]]), t_pulse=np.array([min(ends[0], 1.0e300)]), applied_field=seg_h[:1])
        else:
            kw = dict(segments=(ends, seg_j, seg_h))
        r = self.solve_batch(np.asarray(m_initial, dtype=float)[None], np.array([t1]), device_params, t_start=np.array([t0]),
                             thermal_noise=thermal_noise, temperature=temperature, return_trajectory=True, seed=seed, **kw)
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
        states (physics/llgs_solver.py:264-305). The starts come from NumPy's legacy stream like the reference's
        (`np.random.normal(0, 1, 3)` per trial: the global stream when `seed` is None, RandomState(seed) otherwise)."""
        rs = np.random if seed is None else np.random.RandomState(seed)
        m0 = np.array([rs.normal(0, 1, 3) for _ in range(n_trials)])
        m0 /= np.linalg.norm(m0, axis=1, keepdims=True)
        r = self.solve_batch(m0, relax_time, device_params, current=0.0, thermal_noise=False, return_trajectory=False)
        ok = r["success"].cpu().numpy()
        finals = r["m"].cpu().numpy()[ok]
        stable = []
        for m in finals:
            if all(np.linalg.norm(m - s) >= threshold for s in stable):
                stable.append(m)
        return np.array(stable) if stable else np.array([[0, 0, 1]])
