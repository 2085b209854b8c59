"""Parity tests proper: the CUDA path (SpinTorqueVectorEnv -> C-ABI -> sm_100a kernels) against the golden vectors of the
live reference and against the C oracle on identical inputs. Tolerances (north star): 1e-6 relative with FP64 stage
arithmetic, 1e-4 with FP32 stage arithmetic; binomial / F-test confidence intervals with the in-kernel Philox stream."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN, load_case, noise_for_step, rel_err, stt_params_for, transverse_rel_err

pytestmark = pytest.mark.gpu
TOL = {"f64": 1e-6, "f32": 1e-4}


def _torch():
    import torch
    return torch


def _dtype(name):
    torch = _torch()
    return torch.float64 if name == "f64" else torch.float32


def _make(n, prec, cuda_device, **kw):
    """prec 'f32-pair': FP32 stages through the two-envs-per-thread kernels at every batch size (the thermal one is otherwise
    dispatched only from 262,144 envs - it is the bench's headline kernel)."""
    from spin_torque_rl_gym_b200 import SpinTorqueVectorEnv
    kw.setdefault("autoreset", False)
    if prec == "f32-pair":
        prec = "f32"
        kw.setdefault("pair_kernel", "always")
    if prec == "f32-pair-philox":           # ... and with every word of the stream from Philox4x32-10
        prec = "f32"
        kw.setdefault("pair_kernel", "always")
        kw.setdefault("thermal_stream", "philox")
    return SpinTorqueVectorEnv(num_envs=n, device=cuda_device, dtype=_dtype(prec), **kw)


def _case_env(c, prec, cuda_device):
    return _make(1, prec, cuda_device, device_params=stt_params_for(c), max_current=float(c["max_current"]),
                 include_thermal_fluctuations=bool(c["thermal"]), integrator=str(c["method"]),
                 max_steps=int(c.get("max_steps", 100)))


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ["bigvol_det", "tilted_rk4", "thermal_injected"])
def test_golden_episode(name, prec, cuda_device):
    c = load_case("stt_env.npz", name)
    env = _case_env(c, prec, cuda_device)
    obs, _ = env.reset(options={"initial_state": c["m0"], "target_state": c["target"]})
    assert np.array_equal(obs.cpu().numpy()[0], c["obs"][0])
    tol = TOL[prec]
    for k, a in enumerate(c["actions"]):
        noise = noise_for_step(c["seeds"][k], a, float(c["max_current"]))[None] if "seeds" in c else None
        o, r, te, tr, info = env.step(a[None].copy(), noise=noise)
        m = env.magnetization.cpu().numpy()[0]
        assert np.abs(m - c["m"][k + 1]).max() < tol, k
        assert np.abs(o.cpu().numpy()[0] - c["obs"][k + 1]).max() < tol
        assert abs(float(r[0]) - c["reward"][k]) < tol * max(1.0, abs(c["reward"][k]))
        e_ref = c["energy"][k]
        assert abs(float(info["step_energy"][0]) - e_ref) <= tol * e_ref
        assert bool(te[0]) == c["terminated"][k] and bool(tr[0]) == c["truncated"][k]
        assert int(info["status"][0]) == 0


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_c1_single_env_100_pulses(prec, cuda_device):
    """BASELINE config[0]: N=1, thermal off, 100 random pulses. Free-running while the golden is well-conditioned, then every
    step from the golden's own pre-step state (the episode collapses onto a pole down to denormals; tests/test_oracle_golden)."""
    torch = _torch()
    c = load_case("stt_env.npz", "c1_det")
    env = _case_env(c, prec, cuda_device)
    env.reset(options={"initial_state": c["m0"], "target_state": c["target"]})
    tol = TOL[prec]
    for k, a in enumerate(c["actions"]):
        if np.hypot(*c["m"][k + 1][:2]) < 1e-300:
            break
        o, r, te, tr, info = env.step(a[None].copy())
        m = env.magnetization.cpu().numpy()[0]
        assert np.abs(m - c["m"][k + 1]).max() < tol                                  # the contract
        # beyond the contract: the transverse pair (1e-73 ... 1e-256 here) agrees as a vector relative to its own magnitude;
        # free-running, the per-step relative errors accumulate in the exponent of the decay / regrowth
        assert transverse_rel_err(m, c["m"][k + 1]) < (1e-6 if prec == "f64" else 2e-3), k
        assert abs(float(r[0]) - c["reward"][k]) < tol * max(1.0, abs(c["reward"][k]))
        assert bool(te[0]) == c["terminated"][k]
    assert k >= 8
    # teacher-forced over all 100 steps
    for k, a in enumerate(c["actions"]):
        t_in = np.hypot(*c["m"][k][:2])
        env._m.copy_(torch.from_numpy(c["m"][k]).to(cuda_device).reshape(3, 1))
        env._total_energy.fill_(float(c["total_energy"][k - 1]) if k else 0.0)
        env._step_count.fill_(k)
        o, r, te, tr, info = env.step(a[None].copy())
        m = env.magnetization.cpu().numpy()[0]
        if t_in > 1e-200 and np.hypot(*c["m"][k + 1][:2]) > 1e-300:
            # beyond the contract: one step from the golden's state keeps the exponentially small transverse pair within the
            # mode's tolerance relative to its own magnitude (FP32: block-scaled state)
            assert transverse_rel_err(m, c["m"][k + 1]) < tol, k
            if prec == "f64":
                assert rel_err(m, c["m"][k + 1]) < 1e-6, k
        assert np.abs(m - c["m"][k + 1]).max() < tol, k
        assert np.abs(o.cpu().numpy()[0] - c["obs"][k + 1]).max() < tol
        assert abs(float(r[0]) - c["reward"][k]) < tol * max(1.0, abs(c["reward"][k]))
        assert bool(te[0]) == c["terminated"][k] and bool(tr[0]) == c["truncated"][k]


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_multi_episode_golden(prec, cuda_device):
    g = np.load(os.path.join(GOLDEN, "stt_multi.npz"))
    n = len(g["actions"])
    env = _make(n, prec, cuda_device, max_current=float(g["max_current"]), include_thermal_fluctuations=False)
    obs0, _ = env.reset(options={"initial_state": g["m0"], "target_state": g["target"]})
    assert np.array_equal(obs0.cpu().numpy(), g["obs0"])
    o, r, te, tr, info = env.step(g["actions"].copy())
    tol = TOL[prec]
    assert np.abs(env.magnetization.cpu().numpy() - g["m"]).max() < tol
    assert np.abs(o.cpu().numpy() - g["obs"]).max() < tol
    assert np.allclose(r.cpu().numpy(), g["reward"], rtol=tol, atol=tol)
    assert np.array_equal(te.cpu().numpy(), g["terminated"]) and np.array_equal(tr.cpu().numpy(), g["truncated"])
    assert np.allclose(info["step_energy"].cpu().numpy(), g["energy"], rtol=1e-9, atol=0)


def _random_setup(n, seed, jm=1.1e-6, tmax=1.5e-9):
    rng = np.random.default_rng(seed)
    m0 = rng.normal(size=(n, 3))
    tgt = np.where(rng.integers(2, size=(n, 1)) == 0, 1.0, -1.0) * np.array([[0, 0, 1.0]])
    acts = [np.stack([rng.uniform(-jm, jm, n), rng.uniform(0, tmax, n)], 1).astype(np.float32) for _ in range(2)]
    acts[0][0] = [np.nan, 1e-9]
    acts[0][1] = [jm, np.inf]
    acts[0][2] = [0.0, 5e-10]
    acts[0][3] = [-1.0, 1e-13]
    acts[0][4] = [1e-13, 1e-9]
    return m0, tgt, acts


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("variant", ["default", "tilted", "euler"])
def test_random_batch_vs_c_oracle(prec, variant, cuda_device):
    """4096 envs with ragged substep counts (10..5000: pulses up to max_duration), edge-case actions, two consecutive steps.
    Every env must meet the mode's tolerance: FP32 stages repeat the few ill-conditioned trajectories with FP64 stages (status
    bit 2, llgs_core.cuh CondTrack); Euler always runs FP64 stages."""
    from oracle.c_oracle import COracleEnv
    from oracle.stt_oracle import default_stt_params
    n, jm = 4096, 1.1e-6
    m0, tgt, acts = _random_setup(n, 21, tmax=5e-9)
    p = default_stt_params()
    method = "rk4"
    if variant == "tilted":
        p.update(easy_axis=np.array([0.2, -0.1, 1.0]), reference_magnetization=np.array([0.0, 0.3, 1.0]), damping=0.03)
    if variant == "euler":
        method = "euler"
        acts = [np.stack([a[:, 0], np.minimum(a[:, 1], 2e-10)], 1) for a in acts]   # the Euler map amplifies rounding
    env = _make(n, prec, cuda_device, device_params=p, max_current=jm, include_thermal_fluctuations=False,
                integrator=method)
    ora = COracleEnv(n, device_params=p, max_current=jm, include_thermal=False, method=method,
                     nthreads=os.cpu_count() or 1)
    obs0, _ = env.reset(options={"initial_state": m0, "target_state": tgt})
    assert np.array_equal(obs0.cpu().numpy(), ora.reset(m0, tgt))
    tol = TOL[prec]
    for k, a in enumerate(acts):
        o, r, te, tr, info = env.step(a.copy())
        oo, orr, ote, otr = ora.step(a)
        assert np.array_equal(info["n_sub"].cpu().numpy(), ora.n_sub)
        err = np.abs(env.magnetization.cpu().numpy() - ora.m).max(1)
        assert err.max() < tol
        redone = (info["status"].cpu().numpy() & 4) != 0
        if prec == "f32" and variant == "default":
            assert 0 < redone.mean() < 0.03
            if k == 0:      # from identical start states: a repeated env is FP64-exact, the others stay 5x inside the contract
                assert err[redone].max() < 1e-9 and err[~redone].max() < 2e-5
        else:
            assert not redone.any()              # FP64 stages (also: tilted geometry and Euler through the f32 entry point)
        # an env sitting within tol of the success threshold may legitimately flip its flag (and with it the +100 of its reward)
        mism = (te.cpu().numpy() != ote)
        align = (ora.m * tgt).sum(1)
        assert np.all(np.abs(align[mism] - 0.9) < tol)
        assert np.abs(o.cpu().numpy() - oo).max() < tol
        assert np.allclose(r.cpu().numpy()[~mism], orr[~mism], rtol=tol, atol=tol)
        assert np.array_equal(tr.cpu().numpy(), otr)
        assert np.allclose(info["step_energy"].cpu().numpy(), ora.step_energy, rtol=tol, atol=0)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_injected_noise_batch_vs_c_oracle(prec, cuda_device):
    """Thermal on, identical N(0,1) tensor fed to both sides (order: substep, stage, xyz)."""
    from oracle.c_oracle import COracleEnv
    n, jm = 128, 1.1e-6
    rng = np.random.default_rng(5)
    m0 = rng.normal(size=(n, 3))
    m0[:16] = [1.0, 0.0, 0.0]      # unstable equilibrium: the only place the noise matters dynamically
    tgt = np.tile([0.0, 0.0, 1.0], (n, 1))
    act = np.stack([rng.uniform(-jm, jm, n), rng.uniform(1e-11, 3e-10, n)], 1).astype(np.float32)
    act[:16, 0] = 0.0
    noise = rng.normal(size=(n, 300, 4, 3))
    env = _make(n, prec, cuda_device, max_current=jm, include_thermal_fluctuations=True, temperature=300.0)
    ora = COracleEnv(n, max_current=jm, include_thermal=True, temperature=300.0, nthreads=4)
    env.reset(options={"initial_state": m0, "target_state": tgt})
    ora.reset(m0, tgt)
    o, r, te, tr, info = env.step(act.copy(), noise=noise)
    oo, orr, ote, otr = ora.step(act, noise)
    m = env.magnetization.cpu().numpy()
    assert np.abs(m - ora.m).max() < TOL[prec]
    # at the equator m_z is O(1e-8): compare it relatively (never through 1-|m.z|)
    mz, mz_ref = m[:16, 2], ora.m[:16, 2]
    assert np.all(np.abs(mz_ref) < 1e-6)
    assert np.abs(mz - mz_ref).max() < TOL[prec] * np.abs(mz_ref).max()
    assert np.allclose(r.cpu().numpy(), orr, rtol=TOL[prec], atol=TOL[prec])


@pytest.mark.parametrize("pair", [True, False])
def test_fp32_second_pass_equals_fp64_mode(pair, cuda_device):
    """stg_stt_step_f32 = FP32 kernel + compacted FP64 pass over the envs it declined (include/stg.h d_redo). The redone envs
    must carry status bit 2 and the values of a plain FP64-mode env; statistics, auto-reset rows, sorted launches and pinned host
    outputs must see them exactly once."""
    torch = _torch()
    n, jm = 16384, 1.1e-6
    m0, tgt, acts = _random_setup(n, 77, tmax=5e-9)
    for a in acts:
        a[: n // 4, 0] *= 0.1                     # slow regime: trajectories that linger where the polar rate vanishes
    kw = dict(max_current=jm, include_thermal_fluctuations=False, autoreset=True, max_steps=2, rng_seed=5)
    e32 = _make(n, "f32", cuda_device, pair_kernel=pair, sort_by_substeps=True, **kw)
    e64 = _make(n, "f64", cuda_device, sort_by_substeps=False, **kw)
    eh = _make(n, "f32", cuda_device, pair_kernel=pair, sort_by_substeps=False, host_outputs=True, **kw)
    for e in (e32, e64, eh):
        e.reset(options={"initial_state": m0, "target_state": tgt})
    seen = 0
    for a in acts:
        o32, r32, te32, tr32, i32 = e32.step(a.copy())
        o64, r64, te64, tr64, i64 = e64.step(a.copy())
        oh, rh, teh, trh, ih = eh.step(a.copy())
        redone = (i32["status"] & 4) != 0
        seen += int(redone.sum())
        assert not bool(((i64["status"] & 4) != 0).any())
        # same FP64 source in two kernels: the compiler may contract differently, so FP64 values agree to rounding, FP32 rows bitwise
        assert torch.equal(o32[redone], o64[redone]) and float((r32[redone] - r64[redone]).abs().max()) < 1e-12
        ended = redone & (te32 | tr32)
        assert torch.equal(i32["final_observation"][ended], i64["final_observation"][ended])
        assert float((e32.magnetization[redone] - e64.magnetization[redone]).abs().max()) < 1e-13
        assert float((e32.magnetization - e64.magnetization).abs().max()) < 1e-4
        # host-output env (unsorted launch): same bits as the sorted device-output env, redone rows included
        assert torch.equal(oh, o32.cpu()) and torch.equal(teh, te32.cpu()) and torch.equal(ih["final_observation"], i32["final_observation"].cpu())
        assert torch.equal(ih["status"], i32["status"])
        # envs whose flags differ between the modes sit within tol of the success threshold
        mism = te32 != te64
        align = (e64.magnetization * torch.as_tensor(tgt, device=cuda_device)).sum(1)
        assert bool(((align[mism] - 0.9).abs() < 1e-4).all())
        # every step is compared from identical start states (the statistics buffers are left alone)
        for e in (e32, eh):
            for name in ("_m", "_target", "_total_energy", "_last_action", "_step_count", "_episode"):
                getattr(e, name).copy_(getattr(e64, name))
    assert 0 < seen < 0.05 * n * len(acts)
    st32, st64 = e32.episode_stats(), e64.episode_stats()
    assert st32["steps"] == st64["steps"] == n * len(acts) and st32["substeps"] == st64["substeps"]


@pytest.mark.parametrize("prec", ["f32", "f32-pair", "f32-pair-philox", "f64"])
def test_philox_switching_statistics(prec, cuda_device):
    """In-kernel RNG: start on the equator (m_z = 0), no current, 100 substeps. The sign of m_z is decided by the noise:
    P(m_z > 0) must sit inside the binomial 95 % CI, and Var(m_z) must match the oracle's (same experiment with NumPy noise)
    within the F-test interval — the sign checks the symmetry of the stream, the variance checks the amplitude h_th."""
    from oracle.c_oracle import COracleEnv
    n_gpu, n_cpu = 65536, 4096
    act = np.tile(np.array([[0.0, 1e-10]], np.float32), (n_gpu, 1))
    env = _make(n_gpu, prec, cuda_device, include_thermal_fluctuations=True, temperature=300.0, rng_seed=1234)
    env.reset(options={"initial_state": np.array([1.0, 0.0, 0.0]), "target_state": np.array([0.0, 0.0, 1.0])})
    env.step(act)
    mz = env.magnetization.cpu().numpy()[:, 2]
    rng = np.random.default_rng(99)
    ora = COracleEnv(n_cpu, include_thermal=True, temperature=300.0, nthreads=os.cpu_count() or 1)
    ora.reset(np.array([1.0, 0.0, 0.0]), np.array([0.0, 0.0, 1.0]))
    ora.step(act[:n_cpu], rng.normal(size=(n_cpu, 100, 4, 3)))
    mz_ref = ora.m[:, 2]
    p_gpu = (mz > 0).mean()
    half = 1.96 * np.sqrt(0.25 / n_gpu)
    assert abs(p_gpu - 0.5) < half + 1.96 * np.sqrt(0.25 / n_cpu) * 0 + 1e-12, p_gpu
    p_ref = (mz_ref > 0).mean()
    assert abs(p_gpu - p_ref) < 1.96 * np.sqrt(0.25 / n_gpu + 0.25 / n_cpu)
    ratio = mz.var() / mz_ref.var()
    ci = 1.96 * np.sqrt(2.0 / n_gpu + 2.0 / n_cpu)        # normal approx. of the F interval (both samples Gaussian)
    assert abs(ratio - 1.0) < ci, ratio
    assert abs(mz.mean()) < 4 * mz.std() / np.sqrt(n_gpu)
    assert 1e-9 < mz.std() < 1e-7


def test_full_size_properties(cuda_device):
    """BASELINE config[1] size (65,536 envs, T=300 K, RK4, 999 substeps): size-independent properties."""
    torch = _torch()
    n = 65536
    act = torch.zeros(n, 2, dtype=torch.float32, device=cuda_device)
    act[:, 0] = torch.linspace(-2e6, 2e6, n, device=cuda_device) * 5e-13   # well-conditioned current scale
    act[:, 1] = 1e-9
    kw = dict(max_current=1.1e-6, include_thermal_fluctuations=True, temperature=300.0, rng_seed=7)
    env = _make(n, "f32", cuda_device, **kw)
    env.reset(seed=7)
    m_before = env.magnetization.clone()
    o, r, te, tr, info = env.step(act)
    m = env.magnetization
    assert torch.all(info["n_sub"] == 999)
    assert float((m.norm(dim=1) - 1).abs().max()) < 1e-12            # renormalised in FP64
    assert torch.isfinite(o).all() and torch.isfinite(r).all()
    assert int(info["status"].max()) == 0
    assert float((m - m_before).abs().max()) > 1e-3
    # determinism: same seed, same inputs -> bit-identical
    env2 = _make(n, "f32", cuda_device, **kw)
    env2.reset(seed=7)
    o2, r2, *_ = env2.step(act)
    assert torch.equal(o, o2) and torch.equal(r, r2)
    # sharding invariance: two half-size envs with env_offset reproduce the full batch (Philox counters use global ids)
    h = n // 2
    parts = []
    for k in range(2):
        e = _make(h, "f32", cuda_device, env_offset=k * h, **kw)
        e.reset(seed=7)
        ok, *_ = e.step(act[k * h:(k + 1) * h].contiguous())
        parts.append(ok.clone())
    assert torch.equal(torch.cat(parts), o)
    # different seed -> different thermal stream
    env3 = _make(n, "f32", cuda_device, **dict(kw, rng_seed=8))
    env3.reset(seed=8)                                  # reset(seed) re-keys the Philox stream
    env3._m.copy_(m_before.t().contiguous())
    env3._target.copy_(env2._target)
    o3, *_ = env3.step(act)
    assert not torch.equal(o3[:, :3], o[:, :3])


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_sorted_launch_is_identical(prec, cuda_device):
    """Sorting envs by substep count only changes the thread mapping, never the result."""
    n, jm = 8192, 1.1e-6
    m0, tgt, acts = _random_setup(n, 3, tmax=5e-9)
    outs = []
    for sort in (False, True):
        env = _make(n, prec, cuda_device, max_current=jm, include_thermal_fluctuations=True, rng_seed=3,
                    sort_by_substeps=sort)
        env.reset(options={"initial_state": m0, "target_state": tgt})
        o, r, te, tr, info = env.step(acts[0].copy())
        outs.append((o.clone(), r.clone(), te.clone(), info["n_sub"].clone(), env.magnetization.clone()))
        if sort:
            perm = env._perm.cpu().numpy()
            assert np.array_equal(np.sort(perm), np.arange(n))
            ns = info["n_sub"].cpu().numpy()[perm]
            assert np.all(np.diff(ns) <= 0)
    for a, b in zip(outs[0], outs[1]):
        assert _torch().equal(a, b)


def test_autoreset_final_observation_and_stats(cuda_device):
    torch = _torch()
    n = 4096
    env = _make(n, "f64", cuda_device, max_steps=3, include_thermal_fluctuations=False, autoreset=True, rng_seed=11,
                max_current=1.1e-6)
    obs, _ = env.reset(seed=11)
    m0 = env.magnetization.cpu().numpy()
    assert np.allclose(np.linalg.norm(m0, axis=1), 1.0, atol=1e-15)
    assert abs(m0.mean()) < 0.03 and abs(m0[:, 2].var() - 1 / 3) < 0.03           # uniform on the sphere
    tz = env.target.cpu().numpy()[:, 2]
    assert set(np.unique(tz)) == {-1.0, 1.0} and abs((tz > 0).mean() - 0.5) < 0.05
    act = torch.zeros(n, 2, dtype=torch.float32, device=cuda_device)
    act[:, 1] = 1e-11
    ended = 0
    for s in range(3):
        fin_before = env._final_obs.clone()
        o, r, te, tr, info = env.step(act)
        done = te | tr
        ended += int(done.sum())
        assert torch.all(info["step_count"][done] == 0)
        assert torch.all(o[done, 8] == 1.0)
        fin = info["final_observation"]
        assert torch.equal(fin[~done], fin_before[~done])      # only the rows of ended episodes are written (include/stg.h)
        if done.any():
            assert torch.all(fin[done][:, :3].norm(dim=1) > 0.99)
    st = env.episode_stats()
    assert st["steps"] == 3 * n
    assert st["terminated"] + st["truncated"] == ended
    assert st["substeps"] == 3 * n * 100
    assert ended >= n                     # max_steps=3 truncates every env at the latest in step 3
    sd = env.state_dict()
    env2 = _make(n, "f64", cuda_device, max_steps=3, include_thermal_fluctuations=False, autoreset=True, rng_seed=11,
                 max_current=1.1e-6)
    env2.load_state_dict(sd)
    o1, *_ = env.step(act)
    o1 = o1.clone()
    o2, *_ = env2.step(act)
    assert torch.equal(o1, o2)


def test_invalid_solver_params_keep_magnetisation(cuda_device):
    """SOT dict without `polarization`: the reference's solver validation fails and m never moves."""
    from spin_torque_rl_gym_b200 import params as P
    p = P.default_device_parameters("sot_mram")
    env = _make(16, "f64", cuda_device, device_type="sot_mram", device_params=p, include_thermal_fluctuations=False)
    env.reset(seed=1)
    m0 = env.magnetization.clone()
    o, r, te, tr, info = env.step(np.tile(np.array([[1e6, 1e-9]], np.float32), (16, 1)))
    assert _torch().equal(env.magnetization, m0)
    assert int(info["status"].min()) == 2 and np.isfinite(r.cpu().numpy()).all()
    assert float(info["step_energy"].min()) > 0


@pytest.mark.parametrize("n", [1, 31, 65, 1000])
def test_odd_batch_sizes_and_extreme_actions(n, cuda_device):
    """Batch sizes that are not multiples of the 64-thread CTA, durations at both clamps (10 and 5000 substeps), currents at
    both clamps, against the C oracle."""
    from oracle.c_oracle import COracleEnv
    jm = 1.1e-6
    rng = np.random.default_rng(n)
    m0 = rng.normal(size=(n, 3))
    tgt = np.tile([0.0, 0.0, -1.0], (n, 1))
    act = np.stack([rng.uniform(-jm, jm, n), rng.uniform(1e-12, 6e-10, n)], 1).astype(np.float32)
    act[0] = [5 * jm, 1.0]              # both clamped: J -> max_current, T -> max_duration (5000 substeps, capped below)
    act[-1] = [-5 * jm, 0.0]            # T -> 1e-12 (100 substeps of 1e-14)
    env = _make(n, "f64", cuda_device, max_current=jm, max_duration=2e-9, include_thermal_fluctuations=False)
    ora = COracleEnv(n, max_current=jm, max_duration=2e-9, include_thermal=False, nthreads=2)
    env.reset(options={"initial_state": m0, "target_state": tgt})
    ora.reset(m0, tgt)
    o, r, te, tr, info = env.step(act.copy())
    oo, orr, ote, otr = ora.step(act)
    assert np.array_equal(info["n_sub"].cpu().numpy(), ora.n_sub)
    assert int(info["n_sub"][-1]) == 100 and (n == 1 or int(info["n_sub"][0]) == 2000)
    assert np.abs(env.magnetization.cpu().numpy() - ora.m).max() < 1e-6
    assert np.abs(o.cpu().numpy() - oo).max() < 1e-6 and np.allclose(r.cpu().numpy(), orr, rtol=1e-6, atol=1e-6)
    assert o.shape == (n, 12) and float(o[-1, 10]) == -1.0 and (n == 1 or float(o[0, 10]) == 1.0)


def test_device_mix_param_index(cuda_device):
    """Three parameter sets (STT / SOT / VCMA resistance forms, different damping / anisotropy / volume) selected per env through
    param_index: each env must behave exactly like a single-set env of its own kind, and the resistance-dependent outputs
    (Joule energy, obs[6]) must follow the device formulas of the oracle (devices/*:compute_resistance)."""
    import torch
    from oracle import devices_oracle as DO
    from spin_torque_rl_gym_b200 import SpinTorqueVectorEnv, params as P
    jm, n = 1.1e-6, 96
    stt = P.default_device_parameters("stt_mram")
    sot = dict(P.default_device_parameters("sot_mram"), polarization=0.6, reference_magnetization=np.array([0.0, 0.0, 1.0]))
    vcma = dict(P.default_device_parameters("vcma_mram"), polarization=0.5, volume=2e-23)
    types, plist = ["stt_mram", "sot_mram", "vcma_mram"], [stt, sot, vcma]
    rng = np.random.default_rng(1)
    m0 = rng.normal(size=(n, 3))
    tgt = np.tile([0.0, 0.0, 1.0], (n, 1))
    act = np.stack([rng.uniform(-jm, jm, n), rng.uniform(1e-11, 4e-10, n)], 1).astype(np.float32)
    pidx = np.arange(n) % 3
    kw = dict(max_current=jm, include_thermal_fluctuations=False, autoreset=False, dtype=torch.float64, device=cuda_device)
    mix = SpinTorqueVectorEnv(num_envs=n, device_type=types, device_params=plist, param_index=pidx, **kw)
    mix.reset(options={"initial_state": m0, "target_state": tgt})
    o, r, te, tr, info = mix.step(act.copy())
    m0n = m0 / np.linalg.norm(m0, axis=1, keepdims=True)
    for k, (t, p) in enumerate(zip(types, plist)):
        sel = pidx == k
        one = SpinTorqueVectorEnv(num_envs=int(sel.sum()), device_type=t, device_params=p, **kw)
        one.reset(options={"initial_state": m0[sel], "target_state": tgt[sel]})
        o1, r1, *_ , i1 = one.step(act[sel].copy())
        assert torch.equal(o[torch.as_tensor(sel)], o1) and torch.equal(r[torch.as_tensor(sel)], r1)
        res_before = DO.resistance(t, p, m0n[sel])
        J = np.clip(act[sel, 0].astype(np.float64), -jm, jm)
        T = np.clip(act[sel, 1].astype(np.float64), 1e-12, 5e-9)
        e_ref = (J * res_before * p.get("area", 1e-14)) ** 2 / res_before * T
        assert np.allclose(i1["step_energy"].cpu().numpy(), e_ref, rtol=1e-12)
        m_after = one.magnetization.cpu().numpy()
        assert np.allclose(o1.cpu().numpy()[:, 6], DO.resistance(t, p, m_after) / p["resistance_parallel"], rtol=1e-6)
        assert int(i1["status"].max()) == 0 and float((m_after - m0n[sel]).__abs__().max()) > 1e-6


@pytest.mark.parametrize("thermal", [False, True, "philox"])
def test_packed_pair_kernel_is_bit_identical(thermal, cuda_device):
    """FP32 / e=z / RK4 has two-envs-per-thread FFMA2 kernels (stt_env_step_pair_kernel<0> without noise, <1> with the in-kernel
    thermal stream - the bench's headline kernel, forced here at a small batch with pair_kernel='always'). Every output must
    equal the one-env-per-thread kernel's bit for bit - FP32 and FP64 alike (the FP64 bookkeeping is contraction-proof) - for
    ragged substep counts, an odd batch size, auto-reset, sorted and unsorted launches."""
    torch = _torch()
    n, jm = 4099, 1.1e-6
    m0, tgt, acts = _random_setup(n, 13, tmax=2e-9)
    res = {}
    for pair in ("always", False):
        for sort in (False, True):
            env = _make(n, "f32", cuda_device, max_current=jm, include_thermal_fluctuations=bool(thermal), autoreset=True,
                        max_steps=2, rng_seed=4, pair_kernel=pair, sort_by_substeps=sort,
                        thermal_stream="philox" if thermal == "philox" else "xoshiro")
            env.reset(options={"initial_state": m0, "target_state": tgt})
            out = []
            for a in acts + acts[:1]:
                o, r, te, tr, info = env.step(a.copy())
                out += [o.clone(), r.clone(), te.clone(), tr.clone(), info["final_observation"].clone(),
                        env.magnetization.clone(), info["n_sub"].clone()]
            res[(pair, sort)] = out + [env.stats_tensor().clone()]
    ref = res[(False, False)]
    for key, val in res.items():
        for x, y in zip(val[:-1], ref[:-1]):
            assert torch.equal(x, y), key
        assert torch.allclose(val[-1], ref[-1], rtol=1e-12)          # statistics: same values, different summation order


def test_headline_thermal_kernel_at_dispatch_size_equals_one_env_per_thread(cuda_device):
    """262,144 envs x 999 substeps with the thermal stream: the default dispatch takes stt_env_step_pair_kernel<1> (the bench's
    kernel), pair_kernel=False the one-env-per-thread kernel; same bits per env, and the state stays on the unit sphere."""
    torch = _torch()
    n, jm = 1 << 18, 1.1e-6
    rng = np.random.default_rng(31)
    act = torch.from_numpy(np.stack([rng.uniform(-jm, jm, n), np.full(n, 1e-9)], 1).astype(np.float32)).to(cuda_device)
    outs = []
    for pair in (True, False):
        env = _make(n, "f32", cuda_device, max_current=jm, include_thermal_fluctuations=True, rng_seed=77, pair_kernel=pair,
                    autoreset=True)
        env.reset(seed=5)
        for _ in range(2):
            o, r, te, tr, info = env.step(act)
        assert int(info["n_sub"].min()) == 999 == int(info["n_sub"].max())
        m = env.magnetization
        assert float((m.norm(dim=1) - 1).abs().max()) < 1e-12
        outs.append((o.clone(), r.clone(), te.clone(), m.clone(), info["final_observation"].clone()))
    for x, y in zip(*outs):
        assert torch.equal(x, y)


def test_reset_seed_reproducibility_and_masked_reset(cuda_device):
    """reset(seed) is reproducible (tests/integration/test_environment.py:77-93 of the reference), differs across seeds, and a
    masked reset touches only the selected envs."""
    torch = _torch()
    n = 512
    a = _make(n, "f32", cuda_device, rng_seed=1)
    b = _make(n, "f32", cuda_device, rng_seed=99)
    oa, _ = a.reset(seed=42)
    ob, _ = b.reset(seed=42)
    assert torch.equal(oa, ob)
    oc, _ = b.reset(seed=43)
    assert not torch.equal(oa, oc)
    oa = oa.clone()
    mask = torch.zeros(n, dtype=torch.bool, device=cuda_device)
    mask[::3] = True
    a.step(torch.zeros(n, 2, device=cuda_device))
    before = a.magnetization.clone()
    o2, _ = a.reset(mask=mask)
    after = a.magnetization
    assert torch.equal(after[~mask], before[~mask]) and not torch.equal(after[mask], before[mask])
    assert torch.all(a._step_count[mask] == 0) and torch.all(a._step_count[~mask] == 1)
    # list / numpy / host-tensor actions are all accepted and equivalent
    e1 = _make(4, "f64", cuda_device, include_thermal_fluctuations=False, max_current=1.1e-6)
    outs = []
    for conv in (lambda x: x.tolist(), lambda x: x, lambda x: torch.from_numpy(x), lambda x: torch.from_numpy(x).to(cuda_device)):
        e1.reset(options={"initial_state": np.array([0.3, 0.2, 0.9]), "target_state": np.array([0, 0, 1.0])})
        act = np.array([[5e-7, 1e-10]] * 4, dtype=np.float32)
        o, *_ = e1.step(conv(act))
        outs.append(o.clone())
    assert all(torch.equal(outs[0], x) for x in outs[1:])
    with pytest.raises(ValueError):
        e1.step(act, noise=np.zeros((4, 10, 3, 3)))
    with pytest.raises(RuntimeError):
        _make(2, "f32", cuda_device).step(act[:2])            # step before reset


@pytest.mark.parametrize("prec", ["f32", "f32-pair", "f32-pair-philox", "f64"])
def test_thermal_switching_probability(prec, cuda_device):
    """North-star statistical parity: start exactly on the +z pole, drive with a destabilising current; whether (and when) the
    magnetisation switches is decided by the thermal kicks that seed the transverse component (the switching time goes with the
    log of the noise amplitude: a 3 % amplitude error moves these probabilities by ~0.015). P(crossed the equator) after 310
    substeps and P(success) after 330 substeps must agree with the oracle (its own Gaussian stream) within the binomial 95 % CI."""
    from oracle.c_oracle import COracleEnv
    n_gpu, n_cpu, jm = 65536, 16384, 1.1e-6
    for T, observable in ((3.1e-10, "crossed"), (3.3e-10, "success")):
        act = np.tile(np.array([[8e-7, T]], np.float32), (n_gpu, 1))
        env = _make(n_gpu, prec, cuda_device, max_current=jm, include_thermal_fluctuations=True, temperature=300.0,
                    rng_seed=2024)
        env.reset(options={"initial_state": np.array([0.0, 0.0, 1.0]), "target_state": np.array([0.0, 0.0, -1.0])})
        o, r, te, tr, info = env.step(act)
        ora = COracleEnv(n_cpu, max_current=jm, include_thermal=True, temperature=300.0, nthreads=os.cpu_count() or 1)
        ora.rng_seed = 777
        ora.reset(np.array([0.0, 0.0, 1.0]), np.array([0.0, 0.0, -1.0]))
        _, _, ote, _ = ora.step(act[:n_cpu])
        assert int(info["n_sub"][0]) == ora.n_sub[0]
        if observable == "crossed":
            p_gpu = float((env.magnetization[:, 2] < 0).double().mean())
            p_ref = float((ora.m[:, 2] < 0).mean())
        else:
            p_gpu, p_ref = float(te.double().mean()), float(ote.mean())
        assert 0.2 < p_ref < 0.8                                   # the experiment sits on the steep part of the S-curve
        half = 1.96 * np.sqrt(p_ref * (1 - p_ref) * (1.0 / n_gpu + 1.0 / n_cpu))
        assert abs(p_gpu - p_ref) < half, (observable, p_gpu, p_ref, half)


def test_standalone_stats_reduce_matches_fused_epilogue(cuda_device):
    """K5 standalone (stg_stats_reduce_f64) over stored step results == the statistics the step kernel fuses into its epilogue
    == plain NumPy sums of the same arrays (SURVEY §8b lists stg_stats_reduce in the boundary)."""
    torch = _torch()
    from spin_torque_rl_gym_b200 import _lib
    from spin_torque_rl_gym_b200.parallel import all_reduce_stats, reduce_step_stats
    n, steps = 10007, 6                                  # odd size: partial warps and a partial last CTA
    env = _make(n, "f32", cuda_device, max_steps=4, include_thermal_fluctuations=True, autoreset=False, rng_seed=5,
                max_current=1.1e-6)
    env.reset(seed=5)
    rng = np.random.default_rng(5)
    keep = {k: [] for k in ("reward", "step_energy", "terminated", "truncated", "n_sub", "status", "step_count")}
    for _ in range(steps):
        a = np.stack([rng.uniform(-1.1e-6, 1.1e-6, n), rng.uniform(1e-11, 4e-10, n)], 1).astype(np.float32)
        _, r, te, tr, info = env.step(torch.from_numpy(a).to(cuda_device))
        for k, v in (("reward", r), ("terminated", te), ("truncated", tr), ("step_energy", info["step_energy"]),
                     ("n_sub", info["n_sub"]), ("status", info["status"]), ("step_count", info["step_count"])):
            keep[k].append(v.clone())
    stored = {k: torch.stack(v) for k, v in keep.items()}                    # rollout-buffer layout [T, N]
    fused = env.stats_tensor().cpu().numpy()
    mine = reduce_step_stats(torch.zeros(_lib.NSTATS, dtype=torch.float64, device=cuda_device), **stored).cpu().numpy()
    te, tr = stored["terminated"].cpu().numpy(), stored["truncated"].cpu().numpy()
    want = np.array([n * steps, stored["n_sub"].sum().item(), te.sum(), (~te & tr).sum(),
                     stored["step_energy"].cpu().numpy().sum(), stored["reward"].double().cpu().numpy().sum(),
                     (stored["status"].cpu().numpy() & 1).sum(), stored["step_count"].cpu().numpy()[te | tr].sum()], dtype=float)
    assert want[2] > 0 and want[3] > 0 and want[7] > 0                       # both kinds of episode end occur
    assert np.allclose(mine, want, rtol=1e-12, atol=0) and np.allclose(fused, want, rtol=1e-12, atol=0)
    assert np.array_equal(mine[[0, 1, 2, 3, 6, 7]], want[[0, 1, 2, 3, 6, 7]])        # counts are exact
    # accumulates (does not overwrite); optional inputs leave their statistic untouched
    acc = torch.zeros(_lib.NSTATS, dtype=torch.float64, device=cuda_device)
    reduce_step_stats(acc, reward=stored["reward"][:3])
    reduce_step_stats(acc, reward=stored["reward"][3:])
    got = acc.cpu().numpy()
    assert got[0] == n * steps and got[5] == pytest.approx(want[5], rel=1e-12) and not got[[1, 2, 3, 4, 6, 7]].any()
    assert all_reduce_stats(torch.from_numpy(mine))["success_rate"] == pytest.approx(want[2] / (want[2] + want[3]))
    with pytest.raises(ValueError):
        reduce_step_stats(acc, reward=stored["reward"], n_sub=stored["n_sub"][:2])
    with pytest.raises(ValueError):
        reduce_step_stats(torch.zeros(_lib.NSTATS, dtype=torch.float64), reward=stored["reward"])      # CPU vector


def test_replicated_statistics_fold_and_state_dict(cuda_device):
    """include/stg.h STG_STAT_REPLICAS: the step kernels spread their atomics over 256 copies of the statistics vector
    (one L2 line would serialise every warp of a launch); stg_stats_fold_f64 gives the column sums."""
    import ctypes as C
    torch = _torch()
    from spin_torque_rl_gym_b200 import _lib
    lib = _lib.load()
    rep = torch.rand(_lib.STAT_REPLICAS, _lib.NSTATS, dtype=torch.float64, device=cuda_device) * 1e3
    out = torch.full((_lib.NSTATS,), 7.0, dtype=torch.float64, device=cuda_device)
    stream = torch.cuda.current_stream(cuda_device).cuda_stream
    _lib.check(lib.stg_stats_fold_f64(rep.data_ptr(), out.data_ptr(), 0, stream), "fold")
    want = rep.sum(0)
    assert torch.allclose(out, want, rtol=1e-13, atol=0)
    _lib.check(lib.stg_stats_fold_f64(rep.data_ptr(), out.data_ptr(), 1, stream), "fold")
    assert torch.allclose(out, 2 * want, rtol=1e-13, atol=0)
    assert lib.stg_stats_fold_f64(None, out.data_ptr(), 0, stream) == -1          # STG_E_NULL
    # the env really uses more than one copy, reports their sum, and carries it through a state dict
    n = 65536
    env = _make(n, "f32", cuda_device, max_steps=3, include_thermal_fluctuations=False, autoreset=True, rng_seed=2,
                max_current=1.1e-6)
    env.reset(seed=2)
    act = torch.zeros(n, 2, dtype=torch.float32, device=cuda_device)
    act[:, 0] = 6e-7
    act[:, 1] = 5e-11
    for _ in range(4):
        env.step(act)
    assert int((env._stats[:, 0] != 0).sum()) == _lib.STAT_REPLICAS               # every copy received env-steps
    st = env.episode_stats()
    assert st["steps"] == 4 * n and st["truncated"] + st["terminated"] >= n
    sd = env.state_dict()
    other = _make(n, "f32", cuda_device, max_steps=3, include_thermal_fluctuations=False, autoreset=True, rng_seed=2,
                  max_current=1.1e-6)
    other.reset(seed=2)
    other.load_state_dict(sd)
    assert other.episode_stats() == st
    other.reset_stats()
    assert other.episode_stats()["steps"] == 0 and torch.equal(other.stats_tensor(), torch.zeros_like(other.stats_tensor()))


@pytest.mark.parametrize("thermal", [False, True])
def test_cuda_graph_replay_equals_eager_steps(thermal, cuda_device):
    """capture_step(): a CUDA-graph replay of step() (counting sort + step kernel) gives the same bits as eager steps,
    draws fresh Philox noise on every replay, and keeps the statistics and launch counters right."""
    torch = _torch()
    n = 8192                                               # >= 4096: the captured step includes the counting sort
    kw = dict(max_steps=4, include_thermal_fluctuations=thermal, autoreset=True, rng_seed=9, max_current=1.1e-6)
    eager, graphed = _make(n, "f32", cuda_device, **kw), _make(n, "f32", cuda_device, **kw)
    eager.reset(seed=9)
    graphed.reset(seed=9)
    rng = np.random.default_rng(9)
    acts = [np.stack([rng.uniform(-1.1e-6, 1.1e-6, n), rng.uniform(1e-12, 3e-10, n)], 1).astype(np.float32) for _ in range(6)]
    static = torch.from_numpy(acts[0]).to(cuda_device)
    g = graphed.capture_step(static)
    per_replay = 4 if thermal else 5       # counting sort (3) + step kernel (+ the FP64 second pass of the deterministic FP32 mode)
    assert g.launches_per_replay == per_replay and graphed.gpu_launches == 1        # reset only: capture executed nothing
    m_before = graphed.magnetization.clone()
    assert torch.equal(m_before, eager.magnetization)
    prev_obs = None
    for a in acts:
        o0, r0, te0, tr0, i0 = eager.step(torch.from_numpy(a).to(cuda_device))
        static.copy_(torch.from_numpy(a))
        o1, r1, te1, tr1, i1 = g.replay()
        assert torch.equal(o0, o1) and torch.equal(te0, te1) and torch.equal(tr0, tr1)
        # FP64 bookkeeping of the packed kernel under the counting sort is reproducible to an ulp, not bitwise
        assert torch.equal(r0, r1) if thermal else float((r0 - r1).abs().max()) < 1e-12
        assert torch.equal(i0["n_sub"], i1["n_sub"]) and torch.equal(i0["final_observation"], i1["final_observation"])
        assert torch.equal(eager.magnetization, graphed.magnetization)
        if prev_obs is not None:
            assert not torch.equal(prev_obs, o1)
        prev_obs = o1.clone()
    se, sg = eager.episode_stats(), graphed.episode_stats()
    for k in ("steps", "substeps", "terminated", "truncated", "guard", "episode_length"):
        assert se[k] == sg[k], k
    for k in ("energy", "reward"):                          # FP64 atomics: same terms, unordered sum
        assert se[k] == pytest.approx(sg[k], rel=1e-12), k
    assert se["steps"] == 6 * n and graphed.gpu_launches == 1 + 6 * per_replay
    with pytest.raises(ValueError):
        graphed.capture_step(acts[0])                       # a NumPy array would need staging copies


@pytest.mark.parametrize("thermal", [False, True])
def test_host_outputs_equal_device_outputs(thermal, cuda_device):
    """host_outputs=True: obs / final_obs / reward / flags are written by the kernels straight into pinned host memory
    (no device-to-host copy after the launch). Same bits as the default CUDA outputs; step() returns after the stream drained."""
    torch = _torch()
    n = 50001                                              # odd: partial last CTA, also for the two-envs-per-thread kernel
    kw = dict(max_steps=3, include_thermal_fluctuations=thermal, autoreset=True, rng_seed=13, max_current=1.1e-6)
    dev_env, host_env = _make(n, "f32", cuda_device, **kw), _make(n, "f32", cuda_device, host_outputs=True, **kw)
    o_d, _ = dev_env.reset(seed=13)
    o_h, _ = host_env.reset(seed=13)
    assert o_h.device.type == "cpu" and o_h.is_pinned() and o_d.is_cuda and torch.equal(o_h, o_d.cpu())
    rng = np.random.default_rng(13)
    for _ in range(5):
        a = np.stack([rng.uniform(-1.1e-6, 1.1e-6, n), rng.uniform(1e-12, 4e-10, n)], 1).astype(np.float32)
        od, rd, ted, trd, idv = dev_env.step(a)
        oh, rh, teh, trh, ih = host_env.step(a)
        for x in (oh, rh, teh, trh, ih["final_observation"]):
            assert x.device.type == "cpu"
        assert torch.equal(oh, od.cpu()) and torch.equal(teh, ted.cpu()) and torch.equal(trh, trd.cpu())
        # two independent runs: FP64 rewards of the packed kernel under the counting sort agree to an ulp
        assert torch.equal(rh, rd.cpu()) if thermal else float((rh - rd.cpu()).abs().max()) < 1e-12
        assert torch.equal(ih["final_observation"], idv["final_observation"].cpu())
        assert ih["n_sub"].is_cuda and torch.equal(ih["n_sub"], idv["n_sub"])          # diagnostics stay on the device
    assert host_env.episode_stats()["steps"] == 5 * n
    # masked reset writes the selected observation rows into the host buffer too
    mask = np.zeros(n, bool)
    mask[::3] = True
    oh, _ = host_env.reset(mask=mask)
    od, _ = dev_env.reset(mask=mask)
    assert torch.equal(oh, od.cpu())
