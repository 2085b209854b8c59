"""Round-2 additions pinned to the live reference (tests/golden/round2.npz, oracle/gen_golden.py::gen_round2):
ThermalFluctuations analytics batched on a (temperature x device) grid, LLGSSolver.solve with a shifted time span and with
piecewise-constant current_func / field_func (tables evaluated inside the K2 kernel at the controller's stage times), and
LLGSSolver.find_stable_states. CPU tests run the kernel bodies through tests/hostsim; GPU tests run the CUDA kernels."""
import ctypes as C
import os

import numpy as np
import pytest

from spin_torque_rl_gym_b200 import _lib, params as P
from spin_torque_rl_gym_b200.physics.llgs_solver import PiecewiseConstant, _as_table, _merge_tables, _table_from_callable
from tests.helpers import GOLDEN

G = np.load(os.path.join(GOLDEN, "round2.npz"))
SOLVE_CASES = ["shift", "pwc", "shift_pwc"]


def _solve_params():
    p = P.default_device_parameters("stt_mram")
    p["volume"] = float(G["solve/volume"])
    p["demag_factors"] = G["solve/demag"]
    return p


def _tables(name):
    cur = PiecewiseConstant(G[f"solve/{name}/cur_breaks"], G[f"solve/{name}/cur_values"])
    fld = PiecewiseConstant(G[f"solve/{name}/fld_breaks"], G[f"solve/{name}/fld_values"])
    return cur, fld


# ---- host logic: tables -------------------------------------------------------------------------------------------------------
def test_piecewise_constant_tables_and_detection():
    cur = PiecewiseConstant([1e-11, 3e-11], [5.0, -2.0, 0.0])
    assert cur(0.0) == 5.0 and cur(1e-11) == 5.0 and cur(np.nextafter(1e-11, 1)) == -2.0 and cur(3e-11) == -2.0 and cur(1.0) == 0.0
    # a plain Python callable with the same steps is recovered to adjacent doubles, right-closed and left-closed edges alike
    tab = _table_from_callable(lambda t: 5.0 if t <= 1e-11 else (-2.0 if t < 3e-11 else 0.0), 0.0, 5e-11, "current_func")
    assert np.array_equal(tab.values, [5.0, -2.0, 0.0])
    assert tab.breaks[0] == 1e-11 and tab.breaks[1] == np.nextafter(3e-11, 0)
    vec = _table_from_callable(lambda t: np.array([1.0, 0, 0]) if t < 2e-11 else np.array([0, 2.0, 0]), 0.0, 5e-11, "field_func")
    assert vec.values.shape == (2, 3) and vec.breaks[0] == np.nextafter(2e-11, 0)
    ends, j, h = _merge_tables(cur, vec)
    assert np.array_equal(ends, [1e-11, np.nextafter(2e-11, 0), 3e-11]) and np.array_equal(j, [5.0, -2.0, -2.0, 0.0])
    assert np.array_equal(h[:, 0], [1.0, 1.0, 0.0, 0.0])
    # not representable: raise, never coerce
    for bad in (lambda t: 1e11 * t, lambda t: np.sin(1e12 * t)):
        with pytest.raises(ValueError):
            _table_from_callable(bad, 0.0, 5e-11, "current_func")
    assert _as_table((3.0, 2e-11), 0, 1, "c", False).values.tolist() == [3.0, 0.0]
    with pytest.raises(ValueError):
        PiecewiseConstant([2.0, 1.0], [0, 1, 2])


# ---- K2 body on the host: shifted span + tables vs the live reference ----------------------------------------------------------
def _host_rk45_tables(name):
    from tests.hostsim.harness import lib
    p = _solve_params()
    st = P.make_llg_struct("stt_mram", p)
    cur, fld = _tables(name)
    ends, seg_j, seg_h = _merge_tables(cur, fld)
    t0, t1 = G[f"solve/{name}/span"]
    table = (_lib.StgLlgParams * 1)(st)
    a = _lib.StgRk45Args()
    rows = 4096
    arrs = dict(m0=np.ascontiguousarray(G["solve/m0"][None]), t_end=np.array([t1]), t_start=np.array([t0]),
                ends=np.ascontiguousarray(ends), j=np.ascontiguousarray(seg_j), h=np.ascontiguousarray(seg_h),
                y=np.zeros((1, 3)), acc=np.zeros(1, np.int32), rej=np.zeros(1, np.int32), status=np.zeros(1, np.int32),
                traj=np.zeros((1, rows, 6)))
    a.d_table = C.addressof(table)
    a.d_m0, a.d_t_end, a.d_t_start = arrs["m0"].ctypes.data, arrs["t_end"].ctypes.data, arrs["t_start"].ctypes.data
    a.d_seg_t, a.d_seg_current, a.d_seg_field = arrs["ends"].ctypes.data, arrs["j"].ctypes.data, arrs["h"].ctypes.data
    a.n_seg, a.seg_rows = len(ends), 1
    a.d_y_out, a.d_n_accepted, a.d_n_rejected, a.d_status = (arrs[k].ctypes.data for k in ("y", "acc", "rej", "status"))
    a.d_traj, a.traj_stride = arrs["traj"].ctypes.data, rows
    a.rtol, a.atol, a.max_step, a.n_envs, a.n_sets = 1e-6, 1e-9, 1e-12, 1, 1
    lib().hostsim_llgs_rk45(C.byref(a))
    return arrs


def _check_against_golden(name, t, m, energy, torques):
    g = {k: G[f"solve/{name}/{k}"] for k in ("t", "m", "energy", "torques")}
    assert len(t) == len(g["t"])
    assert np.allclose(t, g["t"], rtol=1e-9, atol=1e-9 * (g["t"][-1] - g["t"][0]))      # spans may cross t = 0
    assert np.abs(m - g["m"]).max() < 1e-6
    assert np.allclose(energy, g["energy"], rtol=1e-6, atol=1e-6 * np.abs(g["energy"]).max())
    assert np.allclose(torques, g["torques"], rtol=1e-6, atol=1e-6 * (np.abs(g["torques"]).max() + 1e-300))


@pytest.mark.parametrize("name", SOLVE_CASES)
def test_kernel_body_shifted_span_and_tables_vs_live_reference(name):
    o = _host_rk45_tables(name)
    rows = len(G[f"solve/{name}/t"])
    assert o["status"][0] == 0 and o["acc"][0] == rows - 1
    tr = o["traj"][0, :rows]
    _check_against_golden(name, tr[:, 0], tr[:, 1:4], tr[:, 4], tr[:, 5])


@pytest.mark.gpu
@pytest.mark.parametrize("name", SOLVE_CASES)
@pytest.mark.parametrize("as_callable", [False, True])
def test_cuda_llgssolver_shifted_span_and_tables(name, as_callable, cuda_device):
    """LLGSSolver.solve with the reference's signature: tables given as PiecewiseConstant objects, or hidden in plain Python
    callables whose break points the host recovers."""
    from spin_torque_rl_gym_b200.physics import LLGSSolver
    cur, fld = _tables(name)
    if as_callable:
        cur_f, fld_f = (lambda t, c=cur: c(t)), (lambda t, f=fld: f(t))
    else:
        cur_f, fld_f = cur, fld
    solver = LLGSSolver(device=cuda_device)
    r = solver.solve(G["solve/m0"], tuple(G[f"solve/{name}/span"]), _solve_params(), cur_f, fld_f, thermal_noise=False)
    assert r["success"]
    _check_against_golden(name, r["t"], r["m"], r["energy"], r["torques"])


@pytest.mark.gpu
def test_cuda_llgssolver_rejects_what_it_cannot_represent(cuda_device):
    from spin_torque_rl_gym_b200.physics import LLGSSolver
    solver = LLGSSolver(device=cuda_device)
    m0, p = G["solve/m0"], _solve_params()
    with pytest.raises(ValueError):
        solver.solve(m0, (0, 4e-11), p, lambda t: 1e12 * t, None, thermal_noise=False)            # ramp
    with pytest.raises(ValueError):
        solver.solve(m0, (0, 4e-11), p, 5.0, lambda t: np.array([np.cos(1e12 * t), 0, 0]), thermal_noise=False)
    with pytest.raises(ValueError):
        solver.solve(m0, (4e-11, 0), p, 5.0, None, thermal_noise=False)                            # backward span
    with pytest.raises(ValueError):
        LLGSSolver(method="BDF", device=cuda_device)


@pytest.mark.gpu
def test_cuda_find_stable_states_vs_live_reference(cuda_device):
    """physics/llgs_solver.py:264-305 re-run by the generator with a short relaxation: same starts (NumPy's legacy stream), same
    end states, same distinct-state list."""
    from spin_torque_rl_gym_b200.physics import LLGSSolver
    p = P.default_device_parameters("stt_mram")
    p["volume"], p["damping"] = float(G["fss/volume"]), float(G["fss/damping"])
    solver = LLGSSolver(device=cuda_device)
    got = solver.find_stable_states(p, n_trials=int(G["fss/n_trials"]), threshold=float(G["fss/threshold"]),
                                    relax_time=float(G["fss/relax"]), seed=int(G["fss/seed"]))
    assert got.shape == G["fss/stable"].shape and np.abs(got - G["fss/stable"]).max() < 1e-6
    # the global stream is the default, like the reference
    np.random.seed(int(G["fss/seed"]))
    got2 = solver.find_stable_states(p, n_trials=int(G["fss/n_trials"]), threshold=float(G["fss/threshold"]),
                                     relax_time=float(G["fss/relax"]))
    assert np.array_equal(got, got2)


# ---- thermal analytics on a grid ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_thermal_analytics_grid_vs_live_reference(cuda_device):
    import torch
    from spin_torque_rl_gym_b200.physics import ThermalFluctuations
    th = ThermalFluctuations(temperature=300.0, device=cuda_device)
    cols = {"uniaxial_anisotropy": G["thg/ku"], "volume": G["thg/vol"], "damping": G["thg/damping"],
            "saturation_magnetization": G["thg/ms"]}
    r = th.batch_analytics(G["thg/temps"], cols, attempt_frequency=float(G["thg/f0"]), measurement_time=float(G["thg/tm"]),
                           failure_rate=float(G["thg/fr"]))
    want = G["thg/grid"]
    for k, name in enumerate(("thermal_stability_factor", "switching_probability", "retention_time", "noise_strength")):
        got = r[name].cpu().numpy()
        assert got.shape == want[k].shape
        fin = np.isfinite(want[k])
        assert np.array_equal(np.isfinite(got), fin) and np.array_equal(got[~fin], want[k][~fin])      # T = 0 row: inf / 0
        assert np.allclose(got[fin], want[k][fin], rtol=1e-12, atol=1e-15), name
    # list-of-dicts form and tensor inputs give the same grid; the sweep of the reference is one column of it
    plist = [{k: float(v[j]) for k, v in cols.items()} for j in range(len(G["thg/ku"]))]
    r2 = th.batch_analytics(torch.as_tensor(G["thg/temps"], device=cuda_device), plist, attempt_frequency=float(G["thg/f0"]),
                            measurement_time=float(G["thg/tm"]), failure_rate=float(G["thg/fr"]))
    assert all(torch.equal(r[k], r2[k]) for k in r)
    N = np.load(os.path.join(GOLDEN, "next.npz"))
    tp = dict(volume=1.5e-25, uniaxial_anisotropy=1.1e6, damping=0.02, saturation_magnetization=7.5e5)
    th2 = ThermalFluctuations(temperature=320.0, seed=9, device=cuda_device)
    sweep = th2.generate_temperature_sweep((50.0, 450.0), tp, n_points=23)
    for k in ("temperature", "thermal_stability_factor", "switching_probability", "retention_time", "noise_strength"):
        assert np.allclose(sweep[k], N[f"th/sweep/{k}"], rtol=1e-12, atol=1e-15), k
    assert th2.temperature == 320.0
    # a device axis on the sweep
    sw2 = th2.generate_temperature_sweep((50.0, 450.0), [tp, dict(tp, volume=3e-25)], n_points=23)
    assert sw2["retention_time"].shape == (23, 2) and np.allclose(sw2["retention_time"][:, 0], sweep["retention_time"], rtol=1e-12)


# ---- host_outputs for K2 and K3: the kernels write their results straight into pinned host memory --------------------------------
@pytest.mark.gpu
def test_cuda_host_outputs_of_rk45_and_array_env_equal_device_outputs(cuda_device):
    import torch
    from spin_torque_rl_gym_b200 import SpinTorqueArrayVectorEnv
    from spin_torque_rl_gym_b200.physics import LLGSSolver
    rng = np.random.default_rng(3)
    # K2
    n = 1000
    p = _solve_params()
    m0 = rng.normal(size=(n, 3))
    cur = rng.uniform(-15.0, 15.0, n)
    solver = LLGSSolver(device=cuda_device)
    rd = solver.solve_batch(m0, 4e-11, p, current=cur)
    rh = solver.solve_batch(m0, 4e-11, p, current=cur, host_outputs=True)
    for k in ("y", "n_accepted", "n_rejected", "n_rhs", "status", "t_reached"):
        assert rh[k].device.type == "cpu" and torch.equal(rh[k], rd[k].cpu()), k
    assert torch.allclose(rh["m"], rd["m"].cpu(), rtol=0, atol=1e-15)        # y / |y| by torch on the host vs on the device
    # K3
    na, size = 37, (8, 8)
    kw = dict(num_envs=na, array_size=size, action_mode="row", device=cuda_device, rng_seed=4, max_steps=3)
    ed, eh = SpinTorqueArrayVectorEnv(**kw), SpinTorqueArrayVectorEnv(host_outputs=True, **kw)
    od, _ = ed.reset(seed=4)
    oh, _ = eh.reset(seed=4)
    assert oh.is_pinned() and torch.equal(oh, od.cpu())
    for s in range(5):
        act = np.stack([rng.uniform(0, 7.49, na), rng.uniform(-2e6, 2e6, na), rng.uniform(0, 5e-9, na)], 1).astype(np.float32)
        od, rd_, ted, trd, idv = ed.step(act)
        oh, rh_, teh, trh, ih = eh.step(act)
        assert torch.equal(oh, od.cpu()) and torch.equal(rh_, rd_.cpu()) and torch.equal(teh, ted.cpu()) and torch.equal(trh, trd.cpu())
        done = (ted | trd).cpu()
        assert torch.equal(ih["final_observation"][done], idv["final_observation"].cpu()[done])
    assert eh.episode_stats()["steps"] == 5 * na and ed.episode_stats()["truncated"] > 0


# ---- torch.library registration ---------------------------------------------------------------------------------------
def test_torch_library_ops_registered_with_fake_impls_and_no_cpu_kernel():
    import torch
    from spin_torque_rl_gym_b200 import torch_ops  # noqa: F401  (defines the stg:: namespace)
    for name in ("vec3_cross", "vec3_dot", "vec3_normalize", "stt_env_step"):
        assert hasattr(torch.ops.stg, name), name
    schema = str(torch.ops.stg.stt_env_step.default._schema)
    for arg in ("m", "target", "total_energy", "last_action", "step_count", "episode"):
        assert f"Tensor(a" in schema and f") {arg}" in schema, schema           # declared as mutated
    a = torch.empty(7, 3, dtype=torch.float64, device="meta")
    assert torch.ops.stg.vec3_cross(a, a).shape == (7, 3) and torch.ops.stg.vec3_dot(a, a).shape == (7,)
    assert torch.ops.stg.vec3_normalize(a).shape == (7, 3)
    act = torch.empty(5, 2, dtype=torch.float32, device="meta")
    st = [torch.empty(3, 5, device="meta", dtype=torch.float64)] * 2 + [torch.empty(5, device="meta", dtype=torch.float64),
          torch.empty(2, 5, device="meta", dtype=torch.float64)] + [torch.empty(5, device="meta", dtype=torch.int32)] * 2
    obs, r, te, tr = torch.ops.stg.stt_env_step(1, act, *st)
    assert obs.shape == (5, 12) and obs.dtype == torch.float32 and r.dtype == torch.float64 and te.dtype == tr.dtype == torch.bool
    with pytest.raises((NotImplementedError, RuntimeError)):                      # no CPU kernel behind the op: loud, not a fallback
        torch.ops.stg.vec3_cross(torch.zeros(2, 3, dtype=torch.float64), torch.zeros(2, 3, dtype=torch.float64))


@pytest.mark.gpu
def test_cuda_torch_library_ops_match_the_classes_and_trace_without_graph_breaks(cuda_device):
    import torch
    from spin_torque_rl_gym_b200 import SpinTorqueVectorEnv, torch_ops
    from spin_torque_rl_gym_b200.physics import VectorizedMagneticsOperations as V
    rng = np.random.default_rng(5)
    a, b = (torch.from_numpy(rng.normal(size=(1000, 3))).to(cuda_device) for _ in range(2))
    assert torch.equal(torch.ops.stg.vec3_cross(a, b), V.batch_cross_product(a, b))
    assert torch.equal(torch.ops.stg.vec3_dot(a, b), V.batch_dot_product(a, b))
    assert torch.equal(torch.ops.stg.vec3_normalize(a), V.batch_normalize(a))
    kw = dict(num_envs=512, device=cuda_device, max_current=1.1e-6, rng_seed=11, dtype=torch.float64, max_steps=4)
    e_ref, e_op = SpinTorqueVectorEnv(**kw), SpinTorqueVectorEnv(**kw)
    e_ref.reset(seed=3)
    e_op.reset(seed=3)
    h = torch_ops.register_env(e_op)
    state = torch_ops.state_tensors(e_op)

    def rollout(acts):
        tot = torch.zeros((), dtype=torch.float64, device=acts.device)
        for k in range(acts.shape[0]):
            obs, r, te, tr = torch.ops.stg.stt_env_step(h, acts[k], *state)
            tot = tot + r.sum() + obs.sum().double()
        return tot, obs, te | tr

    acts = torch.from_numpy(np.stack([np.stack([rng.uniform(-1.1e-6, 1.1e-6, 512), rng.uniform(1e-11, 1e-9, 512)], 1)
                                      for _ in range(3)]).astype(np.float32)).to(cuda_device)
    compiled = torch.compile(rollout, backend="aot_eager", fullgraph=True)        # fullgraph: a graph break would raise
    tot, obs, done = compiled(acts)
    want = torch.zeros((), dtype=torch.float64, device=cuda_device)
    for k in range(3):
        o, r, te, tr, _ = e_ref.step(acts[k])
        want = want + r.sum() + o.sum().double()
    assert torch.equal(obs, o) and torch.equal(done, te | tr) and torch.equal(tot, want)
    assert torch.equal(e_op.magnetization, e_ref.magnetization)
    with pytest.raises(RuntimeError):
        torch.ops.stg.stt_env_step(h, acts[0], state[0][:, :100].contiguous(), *state[1:])     # wrong layout is refused
