"""K2 adaptive RK45: oracle (SciPy + restated RHS) pinned to LLGSSolver.solve goldens; the kernel body (host build on CPU, CUDA
kernel on the GPU) must reproduce SciPy's accepted/rejected step sequence and the trajectory within 1e-6 relative."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import rk45_oracle as RO
from spin_torque_rl_gym_b200 import _lib, params as P
from tests.helpers import GOLDEN

G = np.load(os.path.join(GOLDEN, "rk45.npz"))
CASES = ["stt_on", "stt_pulse", "stt_nocurrent", "stt_tilted", "stt_short"]


def _case(name):
    c = {k.split("/", 1)[1]: G[k] for k in G.files if k.startswith(name + "/")}
    p = P.default_device_parameters("stt_mram")
    for k, v in c.items():
        if k.startswith("param/"):
            p[k[6:]] = v if v.ndim else float(v)
    return c, p


@pytest.mark.parametrize("name", CASES)
def test_oracle_pinned_to_llgssolver(name):
    c, p = _case(name)
    J, tp, happ = float(c["J"]), float(c["t_pulse"]), c["happ"]
    r = RO.solve(c["m0"], float(c["t_end"]), p, lambda t: J if t <= tp else 0.0, lambda t: happ)
    # same accepted-step sequence; last-bit differences of the restated RHS move the adaptive step times by ~1e-13 relative
    assert len(r["t"]) == len(c["t"]) and np.allclose(r["t"], c["t"], rtol=1e-10, atol=0)
    assert np.abs(r["m"] - c["m"]).max() < 1e-10
    assert np.allclose(r["energy"], c["energy"], rtol=1e-9, atol=1e-9 * np.abs(c["energy"]).max())
    assert np.allclose(r["torques"], c["torques"], rtol=1e-9, atol=1e-300)


def test_oracle_thermal_pinned():
    c, p = _case("stt_thermal")
    np.random.seed(int(c["seed"]))
    r = RO.solve(c["m0"], float(c["t_end"]), p, lambda t: float(c["J"]), None, thermal_noise=True)
    # with noise inside the RHS the error estimate (hence every step size) is rough: last-bit differences move t by ~1e-9
    assert len(r["t"]) == len(c["t"]) and np.allclose(r["t"], c["t"], rtol=1e-7, atol=0)
    assert np.abs(r["m"] - c["m"]).max() < 1e-8


def _host_rk45(m0, t_end, structs, pidx=None, current=None, t_pulse=None, happ=None, voltage=None, noise=None,
               rtol=1e-6, atol=1e-9, max_step=1e-12, traj_rows=4096):
    from tests.hostsim.harness import lib
    n = len(m0)
    table = (_lib.StgLlgParams * len(structs))(*structs)
    a = _lib.StgRk45Args()
    keep = []

    def ptr(x, dt=np.float64):
        if x is None:
            return None
        x = np.ascontiguousarray(x, dt)
        keep.append(x)
        return x.ctypes.data
    a.d_table = C.addressof(table)
    a.d_param_index = ptr(pidx, np.int32)
    a.d_m0, a.d_t_end = ptr(m0), ptr(np.broadcast_to(t_end, (n,)))
    a.d_current, a.d_t_pulse, a.d_happ, a.d_voltage = ptr(current), ptr(t_pulse), ptr(happ), ptr(voltage)
    out = dict(y=np.zeros((n, 3)), acc=np.zeros(n, np.int32), rej=np.zeros(n, np.int32), rhs=np.zeros(n, np.int32),
               status=np.zeros(n, np.int32), t=np.zeros(n), traj=np.zeros((n, traj_rows, 6)))
    a.d_y_out, a.d_n_accepted, a.d_n_rejected = out["y"].ctypes.data, out["acc"].ctypes.data, out["rej"].ctypes.data
    a.d_n_rhs, a.d_status, a.d_t_reached = out["rhs"].ctypes.data, out["status"].ctypes.data, out["t"].ctypes.data
    a.d_traj, a.traj_stride = out["traj"].ctypes.data, traj_rows
    if noise is not None:
        a.d_noise, a.noise_stride, a.flags = ptr(noise), noise.shape[1], _lib.F_THERMAL_INJECT
    a.rtol, a.atol, a.max_step, a.n_envs, a.n_sets = rtol, atol, max_step, n, len(structs)
    lib().hostsim_llgs_rk45(C.byref(a))
    return out


@pytest.mark.parametrize("name", CASES)
def test_kernel_body_reproduces_scipy_step_sequence(name):
    c, p = _case(name)
    st = P.make_llg_struct("stt_mram", p)
    o = _host_rk45(c["m0"][None], float(c["t_end"]), [st], current=[float(c["J"])], t_pulse=[min(float(c["t_pulse"]), 1e300)],
                   happ=c["happ"][None])
    rows = len(c["t"])
    assert o["status"][0] == 0 and o["acc"][0] == rows - 1
    tr = o["traj"][0, :rows]
    assert np.allclose(tr[:, 0], c["t"], rtol=1e-9, atol=0)
    assert np.abs(tr[:, 1:4] - c["m"]).max() < 1e-6
    assert np.allclose(tr[:, 4], c["energy"], rtol=1e-6, atol=1e-6 * np.abs(c["energy"]).max())
    assert np.allclose(tr[:, 5], c["torques"], rtol=1e-6, atol=1e-6 * (np.abs(c["torques"]).max() + 1e-300))
    J, tp, happ = float(c["J"]), float(c["t_pulse"]), c["happ"]
    ref = RO.solve(c["m0"], float(c["t_end"]), p, lambda t: J if t <= tp else 0.0, lambda t: happ)
    assert o["rej"][0] == ref["n_rejected"] and o["rhs"][0] == ref["nfev"]
    assert np.abs(o["y"][0] - ref["y_end"]).max() < 1e-9


def test_kernel_body_thermal_injected_noise():
    c, p = _case("stt_thermal")
    np.random.seed(int(c["seed"]))
    noise = np.random.normal(0, 1, (1, 2000, 3))
    st = P.make_llg_struct("stt_mram", p, thermal=True, temperature=300.0)
    o = _host_rk45(c["m0"][None], float(c["t_end"]), [st], current=[float(c["J"])], noise=noise)
    rows = len(c["t"])
    assert o["acc"][0] == rows - 1
    assert np.abs(o["traj"][0, :rows, 1:4] - c["m"]).max() < 1e-6


def _mix_setup(n, seed=3):
    rng = np.random.default_rng(seed)
    sot = P.default_device_parameters("sot_mram")
    sot.update(aspect_ratio=2.0, spin_hall_angle=0.3)
    vcma = P.default_device_parameters("vcma_mram")
    vcma.update(aspect_ratio=1.5)
    m0 = rng.normal(size=(n, 3))
    pidx = (np.arange(n) % 2).astype(np.int32)
    cur = np.where(pidx == 0, rng.uniform(-3e11, 3e11, n), 0.0)      # SOT rate 0.2*j_s*J ~ 1e10 1/s at this scale
    volt = np.where(pidx == 1, rng.uniform(-2.5, 2.5, n), 0.0)
    happ = rng.normal(size=(n, 3)) * 2e4
    t_end = rng.uniform(2e-12, 6e-11, n)
    return sot, vcma, m0, pidx, cur, volt, happ, t_end


def test_kernel_body_sot_vcma_mix_vs_composed_oracle():
    """Config C3 physics (SOT + VCMA device mix) at oracle-sized N: end state and step counts against SciPy driving the RHS
    composed from the reference's device methods."""
    n = 24
    sot, vcma, m0, pidx, cur, volt, happ, t_end = _mix_setup(n)
    structs = [P.make_llg_struct("sot_mram", sot), P.make_llg_struct("vcma_mram", vcma)]
    o = _host_rk45(m0, t_end, structs, pidx=pidx, current=cur, happ=happ, voltage=volt, traj_rows=8)
    for i in range(n):
        kind, prm = ("sot_mram", sot) if pidx[i] == 0 else ("vcma_mram", vcma)
        J, h = cur[i], happ[i]
        ref = RO.solve(m0[i], t_end[i], prm, lambda t: J, lambda t: h, kind=kind, voltage=volt[i])
        assert o["acc"][i] == ref["n_accepted"] and o["rej"][i] == ref["n_rejected"], i
        assert np.abs(o["y"][i] - ref["y_end"]).max() < 1e-6 * max(1.0, np.abs(ref["y_end"]).max())
    assert (o["status"] & ~2 == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES + ["stt_thermal"])
def test_cuda_llgssolver_matches_reference(name, cuda_device):
    from spin_torque_rl_gym_b200.physics import LLGSSolver
    c, p = _case(name)
    solver = LLGSSolver(method="RK45", rtol=1e-6, atol=1e-9, max_step=1e-12, device=cuda_device)
    if name == "stt_thermal":
        np.random.seed(int(c["seed"]))
        noise = np.random.normal(0, 1, (1, 2000, 3))
        r = solver.solve_batch(c["m0"][None], float(c["t_end"]), p, current=float(c["J"]), thermal_noise=True,
                               noise=noise, return_trajectory=True)
        rows = len(c["t"])
        assert int(r["n_accepted"][0]) == rows - 1
        assert np.abs(r["traj"][0, :rows, 1:4].cpu().numpy() - c["m"]).max() < 1e-6
        return
    J, tp, happ = float(c["J"]), float(c["t_pulse"]), c["happ"]
    r = solver.solve(c["m0"], (0, float(c["t_end"])), p, lambda t: J if t <= tp else 0.0, lambda t: happ,
                     thermal_noise=False)
    assert r["success"] and len(r["t"]) == len(c["t"])
    assert np.allclose(r["t"], c["t"], rtol=1e-9, atol=0)
    assert np.abs(r["m"] - c["m"]).max() < 1e-6
    assert np.allclose(r["energy"], c["energy"], rtol=1e-6, atol=1e-6 * np.abs(c["energy"]).max())
    assert np.allclose(r["torques"], c["torques"], rtol=1e-6, atol=1e-6 * (np.abs(c["torques"]).max() + 1e-300))


@pytest.mark.gpu
def test_cuda_sot_vcma_mix_batch(cuda_device):
    """C3 at a GPU-sized batch: 256 envs checked against the composed SciPy oracle, 262,144 envs for invariants."""
    import torch
    from spin_torque_rl_gym_b200.physics import LLGSSolver
    solver = LLGSSolver(device=cuda_device)
    n = 256
    sot, vcma, m0, pidx, cur, volt, happ, t_end = _mix_setup(n, seed=8)
    r = solver.solve_batch(m0, t_end, [sot, vcma], current=cur, applied_field=happ, voltage=volt, param_index=pidx,
                           device_type=["sot_mram", "vcma_mram"])
    y = r["y"].cpu().numpy()
    acc, rej = r["n_accepted"].cpu().numpy(), r["n_rejected"].cpu().numpy()
    for i in range(0, n, 4):
        kind, prm = ("sot_mram", sot) if pidx[i] == 0 else ("vcma_mram", vcma)
        J, h = cur[i], happ[i]
        ref = RO.solve(m0[i], t_end[i], prm, lambda t: J, lambda t: h, kind=kind, voltage=volt[i])
        assert acc[i] == ref["n_accepted"] and rej[i] == ref["n_rejected"], i
        assert np.abs(y[i] - ref["y_end"]).max() < 1e-6
    assert bool(r["success"].all())
    # full C3 size: every trajectory reaches t_end, stays near the unit sphere, and is independent of its batch neighbours
    N = 262144
    reps = N // n
    big = solver.solve_batch(np.tile(m0, (reps, 1)), np.tile(t_end, reps), [sot, vcma], current=np.tile(cur, reps),
                             applied_field=np.tile(happ, (reps, 1)), voltage=np.tile(volt, reps),
                             param_index=np.tile(pidx, reps), device_type=["sot_mram", "vcma_mram"])
    assert bool(big["success"].all())
    assert torch.equal(big["y"][:n], r["y"]) and torch.equal(big["y"][-n:], r["y"])
    assert torch.equal(big["t_reached"], torch.as_tensor(np.tile(t_end, reps)).to(cuda_device))
    # the raw end state is not renormalised (sol.y[:, -1]); the SOT field-like torque is not tangential, so |y| drifts (up to ~20 % here)
    assert float((big["y"].norm(dim=1) - 1).abs().max()) < 0.5 and bool(torch.isfinite(big["y"]).all())


@pytest.mark.gpu
def test_cuda_rk45_sorted_launch_identical_to_caller_order(cuda_device):
    """Large ragged batches are launched through a permutation sorted by (parameter set, t_end) so warps hold trajectories of
    similar length (StgRk45Args.d_perm). Which thread integrates a trajectory must not matter: every output equals the
    caller-order launch bit for bit, also with Philox noise (counters come from the trajectory id) and recorded rows."""
    import torch
    from spin_torque_rl_gym_b200.physics import LLGSSolver
    n = 256
    sot, vcma, m0, pidx, cur, volt, happ, t_end = _mix_setup(n, seed=21)
    N = 262144
    reps = N // n
    rng = np.random.default_rng(21)
    kw = dict(current=np.tile(cur, reps) * rng.uniform(0.5, 1.5, N), applied_field=np.tile(happ, (reps, 1)),
              voltage=np.tile(volt, reps) * rng.uniform(0.5, 1.0, N), param_index=np.tile(pidx, reps),
              device_type=["sot_mram", "vcma_mram"])
    M0 = rng.normal(size=(N, 3))
    T = np.tile(t_end, reps) * rng.uniform(0.3, 1.0, N)    # ragged trajectory lengths
    a = LLGSSolver(device=cuda_device).solve_batch(M0, T, [sot, vcma], **kw)
    b = LLGSSolver(device=cuda_device, sort_trajectories=False).solve_batch(M0, T, [sot, vcma], **kw)
    for k in ("y", "n_accepted", "n_rejected", "n_rhs", "status", "t_reached"):
        assert torch.equal(a[k], b[k]), k
    spread = a["n_accepted"].float()
    assert float(spread.max()) > 1.5 * float(spread.min()) and bool(a["success"].all())       # the batch really is ragged
    # thermal noise from the in-kernel stream + recorded rows, 90,112 short trajectories
    N2 = 90112
    sel = slice(0, N2)
    kw2 = {k: (v[sel] if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
    T2 = T[sel] * 0.15
    c = LLGSSolver(device=cuda_device).solve_batch(M0[sel], T2, [sot, vcma], thermal_noise=True, temperature=300.0, seed=5,
                                                   return_trajectory=True, **kw2)
    d = LLGSSolver(device=cuda_device, sort_trajectories=False).solve_batch(M0[sel], T2, [sot, vcma], thermal_noise=True,
                                                                            temperature=300.0, seed=5, return_trajectory=True,
                                                                            **kw2)
    for k in ("y", "n_accepted", "n_rejected", "n_rhs", "status", "traj"):
        assert torch.equal(c[k], d[k]), k
    assert not torch.equal(c["y"], LLGSSolver(device=cuda_device).solve_batch(M0[sel], T2, [sot, vcma], **kw2)["y"])
