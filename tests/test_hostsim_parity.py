"""CPU-side check of the exact arithmetic the CUDA kernels run: the per-env bodies of csrc/stt_env_core.cuh compiled for the
host (tests/hostsim) against the golden vectors of the live reference and against the C oracle.

Tolerances are the north-star ones: 1e-6 relative (FP64 stage arithmetic), 1e-4 (FP32 stage arithmetic, FP64 state).
The GPU tests (-m gpu) repeat the same comparisons through libstg.so's CUDA kernels."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN, load_case, noise_for_step, rel_err, stt_params_for, transverse_rel_err
from tests.hostsim.harness import HostSimEnv
from oracle.c_oracle import COracleEnv

TOL = {True: 1e-6, False: 1e-4}


def _env(c, f64, **kw):
    return HostSimEnv(1, device_params=stt_params_for(c), max_current=float(c["max_current"]),
                      include_thermal_fluctuations=bool(c["thermal"]), integrator=str(c["method"]),
                      max_steps=int(c.get("max_steps", 100)), f64=f64, **kw)


@pytest.mark.parametrize("f64", [True, False])
@pytest.mark.parametrize("name", ["bigvol_det", "tilted_rk4", "thermal_injected"])
def test_free_running_episode(name, f64):
    c = load_case("stt_env.npz", name)
    env = _env(c, f64)
    obs = env.reset(c["m0"], c["target"])
    assert np.array_equal(obs[0], c["obs"][0])
    for k, a in enumerate(c["actions"]):
        noise = noise_for_step(c["seeds"][k], a, float(c["max_current"]))[None] if "seeds" in c else None
        o, r, te, tr = env.step(a[None], noise)
        assert np.abs(env.m[:, 0] - c["m"][k + 1]).max() < TOL[f64], k
        assert np.abs(o[0] - c["obs"][k + 1]).max() < TOL[f64]
        assert abs(r[0] - c["reward"][k]) < TOL[f64] * max(1.0, abs(c["reward"][k]))
        e_ref = c["energy"][k]
        assert abs(env.step_energy[0] - e_ref) <= TOL[f64] * e_ref
        assert te[0] == c["terminated"][k] and tr[0] == c["truncated"][k]


@pytest.mark.parametrize("f64,general", [(True, False), (True, True), (False, False), (False, True)])
def test_c1_teacher_forced(f64, general):
    """Config C1: every step from the golden's own pre-step state (see test_oracle_golden for why)."""
    c = load_case("stt_env.npz", "c1_det")
    env = _env(c, f64, force_general=general)
    env.reset(c["m0"], c["target"])
    for k, a in enumerate(c["actions"]):
        t_in = np.hypot(*c["m"][k][:2])
        if not f64 and general and t_in < 1e-30:
            continue   # only the axis-z FP32 path block-scales the transverse pair (llgs_core.cuh: ScaledState)
        env.m[:, 0] = c["m"][k]
        env.total_energy[0] = c["total_energy"][k - 1] if k else 0.0
        env.step_count[0] = k
        o, r, te, tr = env.step(a[None])
        if t_in > 1e-200 and np.hypot(*c["m"][k + 1][:2]) > 1e-300 and (f64 or not general):
            # beyond the contract: the exponentially small transverse pair (down to 1e-257) agrees as a vector, relative to
            # its own magnitude, within the mode's tolerance (FP32: block-scaled state, llgs_core.cuh ScaledState)
            assert transverse_rel_err(env.m[:, 0], c["m"][k + 1]) < TOL[f64], k
            if f64:
                assert rel_err(env.m[:, 0], c["m"][k + 1]) < 1e-6, k
        assert np.abs(env.m[:, 0] - c["m"][k + 1]).max() < TOL[f64], k
        assert abs(r[0] - c["reward"][k]) < TOL[f64] * max(1.0, abs(c["reward"][k]))
        assert te[0] == c["terminated"][k] and tr[0] == c["truncated"][k]


@pytest.mark.parametrize("f64", [True, False])
def test_c1_free_running_prefix(f64):
    """C1 free-running for as long as the golden itself is well-conditioned (transverse above the denormal range)."""
    c = load_case("stt_env.npz", "c1_det")
    env = _env(c, f64)
    env.reset(c["m0"], c["target"])
    for k, a in enumerate(c["actions"]):
        o, r, te, tr = env.step(a[None])
        if np.hypot(*c["m"][k + 1][:2]) < 1e-300:
            break
        assert np.abs(env.m[:, 0] - c["m"][k + 1]).max() < TOL[f64]                 # the contract
        # beyond the contract: free-running over nine steps the transverse pair decays to 1e-256 and regrows; the per-step
        # relative errors (test_c1_teacher_forced) accumulate in the exponent, so the free-running bound is looser
        assert transverse_rel_err(env.m[:, 0], c["m"][k + 1]) < (1e-6 if f64 else 2e-3), k
        assert abs(r[0] - c["reward"][k]) < TOL[f64] * max(1.0, abs(c["reward"][k]))
        assert te[0] == c["terminated"][k] and tr[0] == c["truncated"][k]
    assert k >= 8


@pytest.mark.parametrize("f64", [True, False])
def test_multi_episode_golden(f64):
    g = np.load(os.path.join(GOLDEN, "stt_multi.npz"))
    n = len(g["actions"])
    env = HostSimEnv(n, max_current=float(g["max_current"]), include_thermal_fluctuations=False, f64=f64)
    obs0 = env.reset(g["m0"], g["target"])
    assert np.array_equal(obs0, g["obs0"])
    o, r, te, tr = env.step(g["actions"])
    assert np.abs(env.m.T - g["m"]).max() < TOL[f64]
    assert np.abs(o - g["obs"]).max() < TOL[f64]
    assert np.allclose(r, g["reward"], rtol=TOL[f64], atol=TOL[f64])
    assert np.array_equal(te, g["terminated"]) and np.array_equal(tr, g["truncated"])


@pytest.mark.parametrize("f64", [True, False])
def test_euler_free_running(f64):
    """The explicit-Euler map with 0.35 rad/substep amplifies rounding differences chaotically, so Euler runs FP64 stages
    through both entry points (stt_kernels.cu: launch_step) and the whole free-running episode agrees to 1e-6."""
    c = load_case("stt_env.npz", "tilted_euler")
    env = _env(c, f64)
    env.reset(c["m0"], c["target"])
    for k, a in enumerate(c["actions"]):
        env.step(a[None])
        assert np.abs(env.m[:, 0] - c["m"][k + 1]).max() < 1e-6


@pytest.mark.parametrize("f64", [True, False])
def test_random_batch_vs_c_oracle(f64):
    """256 envs, random states/targets/actions incl. edge actions, 3 steps, against the C oracle."""
    rng = np.random.default_rng(11)
    n, jm = 256, 1.1e-6
    m0 = rng.normal(size=(n, 3))
    tgt = np.where(rng.integers(2, size=(n, 1)) == 0, 1.0, -1.0) * np.array([[0, 0, 1.0]])
    h = HostSimEnv(n, max_current=jm, include_thermal_fluctuations=False, f64=f64)
    o = COracleEnv(n, max_current=jm, include_thermal=False, nthreads=4)
    assert np.array_equal(h.reset(m0, tgt), o.reset(m0, tgt))
    for s in range(3):
        act = np.stack([rng.uniform(-jm, jm, n), rng.uniform(0, 1.5e-9, n)], 1).astype(np.float32)
        act[0] = [np.nan, 1e-9]; act[1] = [jm, np.inf]; act[2] = [0.0, 5e-10]; act[3] = [-1.0, 1e-13]
        act[4] = [1e-13, 1e-9]
        ho, hr, hte, htr = h.step(act)
        oo, orr, ote, otr = o.step(act)
        assert np.array_equal(h.n_sub, o.n_sub)
        assert np.abs(h.m.T - o.m).max() < TOL[f64]
        assert np.abs(ho - oo).max() < TOL[f64]
        assert np.allclose(hr, orr, rtol=TOL[f64], atol=TOL[f64])
        assert np.array_equal(hte, ote) and np.array_equal(htr, otr)
        assert np.allclose(h.step_energy, o.step_energy, rtol=TOL[f64], atol=0)


@pytest.mark.parametrize("pair", [False, True])
def test_fp32_conditioning_flag(pair):
    """FP32 stages + conditioning bound (llgs_core.cuh CondTrack): over uniformly random (state, current, pulse <= 5 ns) samples
    EVERY env ends within the FP32 contract of the FP64 oracle - the few trajectories held near an unstable polar angle for
    thousands of substeps are detected in flight and repeated with FP64 stages (status bit 2) - and the worst error of the envs
    that stayed on FP32 stages keeps a wide margin to 1e-4."""
    rng = np.random.default_rng(31 + pair)
    n, jm = 8192, 1.1e-6
    m0 = rng.normal(size=(n, 3))
    tgt = np.where(rng.integers(2, size=(n, 1)) == 0, 1.0, -1.0) * np.array([[0, 0, 1.0]])
    act = np.stack([rng.uniform(-jm, jm, n), rng.uniform(0, 5e-9, n)], 1).astype(np.float32)
    act[: n // 4, 0] *= 0.1                      # a quarter of the batch in the slow regime where W ~ 0 is reachable
    h = HostSimEnv(n, max_current=jm, include_thermal_fluctuations=False, f64=False, pair=pair)
    o = COracleEnv(n, max_current=jm, include_thermal=False, nthreads=os.cpu_count() or 1)
    h.reset(m0, tgt)
    o.reset(m0, tgt)
    ho, hr, hte, htr = h.step(act)
    oo, orr, ote, otr = o.step(act)
    err = np.abs(h.m.T - o.m).max(1)
    redone = (h.status & 4) != 0
    assert err.max() < 1e-4
    assert err[~redone].max() < 2e-5 and err[redone].max() < 1e-9
    assert 0 < redone.mean() < 0.03
    assert np.abs(ho - oo).max() < 1e-4 and np.allclose(hr, orr, rtol=1e-4, atol=1e-4)
    align = (o.m * tgt).sum(1)
    assert np.all(np.abs(align[hte != ote] - 0.9) < 1e-4)          # a flag may only differ within tol of the threshold


def test_substep_plan_matches_numpy_quirks():
    """float32 durations floor to 999/4999/... substeps and tiny durations give 99 or 100."""
    from oracle.stt_oracle import substep_plan
    durs = np.array([1e-9, 5e-9, 2e-9, 5e-10, 1e-10, 7.3e-11, 1e-12, 3.3e-12, 9.99e-11], dtype=np.float32)
    h = HostSimEnv(len(durs), include_thermal_fluctuations=False, f64=True)
    h.reset(np.array([0.3, 0.2, 0.9]), [0, 0, 1.0])
    h.step(np.stack([np.zeros(len(durs), np.float32), durs], 1))
    assert list(h.n_sub) == [substep_plan(float(d))[0] for d in durs]
    assert h.n_sub[0] == 999 and h.n_sub[1] == 4999


def test_autoreset_and_truncation_semantics():
    h = HostSimEnv(8, max_steps=3, include_thermal_fluctuations=False, f64=True, autoreset=True, rng_seed=5)
    h.reset()
    m_start = h.m.copy()
    assert np.allclose(np.linalg.norm(m_start, axis=0), 1.0)
    act = np.tile(np.array([[0.0, 1e-11]], np.float32), (8, 1))
    ends = 0
    for s in range(3):
        fin_before = h.final_obs.copy()
        o, r, te, tr = h.step(act)
        done = te | tr
        ends += done.sum()
        # envs that ended were reset in the same call: fresh counters, obs of the new episode, final_obs = their last obs;
        # the final_obs rows of the running envs are not written (include/stg.h StgSttStepOut.final_obs)
        assert np.all(h.step_count[done] == 0) and np.all(h.total_energy[done] == 0)
        assert np.all(o[done, 8] == 1.0)
        assert np.array_equal(h.final_obs[~done], fin_before[~done])
        if done.any():
            assert np.all(np.abs(h.final_obs[done][:, :3]).sum(1) > 0)
            assert np.all(h.final_obs[done][:, 8] == 0.0) or not np.all(tr[done])     # truncated: no steps remaining
    assert np.all(h.episode >= 1 + 1)   # reset() + at least one auto-reset within 3 steps (max_steps=3)
    assert ends >= 8


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    import ctypes as C
    from tests.hostsim.harness import lib
    out = (C.c_uint32 * 4)()
    lib().hostsim_philox(0, 0, 0, 0, 0, 0, out)
    assert list(out) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    lib().hostsim_philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, out)
    assert list(out) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    lib().hostsim_philox(0xa4093822, 0x299f31d0, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, out)
    assert list(out) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _xoshiro128pp_py(state, n):
    """xoshiro128++ 1.0 restated from its published definition (Blackman & Vigna): result = rotl(s0 + s3, 7) + s0; t = s1 << 9;
    s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t; s3 = rotl(s3, 11)."""
    M = 0xffffffff
    rotl = lambda x, k: ((x << k) | (x >> (32 - k))) & M
    s0, s1, s2, s3 = state
    out = []
    for _ in range(n):
        out.append((rotl((s0 + s3) & M, 7) + s0) & M)
        t = (s1 << 9) & M
        s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t
        s3 = rotl(s3, 11)
    return out


def test_xoshiro_known_answers_and_stream_seeding():
    """The thermal stream of the RK4 paths: xoshiro128++ (checked against an independent restatement of the published algorithm,
    incl. the first output of state (1, 2, 3, 4): rotl(5, 7) + 1 = 641), seeded from block 0 of the env-step's Philox stream."""
    import ctypes as C
    from tests.hostsim.harness import lib
    out = (C.c_uint32 * 64)()
    for st in ((1, 2, 3, 4), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8), (0xffffffff, 0, 0x80000000, 7)):
        lib().hostsim_xoshiro(*st, 64, out)
        assert list(out) == _xoshiro128pp_py(st, 64)
    lib().hostsim_xoshiro(1, 2, 3, 4, 1, out)
    assert out[0] == 641
    seed4, ph4 = (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
    for seed, gid, ep, step in ((1234, 5, 2, 3), (0xdeadbeefcafe, (1 << 33) | 17, 4, 9)):
        lib().hostsim_stream_seed(C.c_uint64(seed), C.c_uint64(gid), ep, step, seed4)
        lib().hostsim_philox(seed & 0xffffffff, seed >> 32, gid & 0xffffffff, ep, step, (gid >> 32) << 20, ph4)
        assert list(seed4) == list(ph4)
    # distinct (env, episode, step) -> unrelated states
    seen = set()
    for gid in range(64):
        for step in range(4):
            lib().hostsim_stream_seed(C.c_uint64(7), C.c_uint64(gid), 0, step, seed4)
            seen.add(tuple(seed4))
    assert len(seen) == 256


def test_thermal_stream_normals_moments():
    import ctypes as C
    from tests.hostsim.harness import lib
    buf = (C.c_float * 12)()
    xs = []
    for g in range(6000):                       # substeps 7 / 8 of 6000 different env-steps
        lib().hostsim_normals12(C.c_uint64(1234), C.c_uint64(g), 2, 3, 7 + (g & 1), buf)
        xs.append(np.array(buf[:]))
    x = np.concatenate(xs)
    assert abs(x.mean()) < 4 / np.sqrt(x.size)
    assert abs(x.var() - 1) < 0.03
    assert abs((x ** 4).mean() - 3) < 0.15
    assert abs((x ** 6).mean() - 15) < 1.5
    assert np.abs(x).max() < 4.86                      # 16-bit radius uniform: tail to sqrt(2 ln 2^17)
    pairs = np.stack(xs)
    c = np.corrcoef(pairs.T)
    assert np.abs(c - np.eye(12)).max() < 0.07
    # along ONE stream: consecutive substeps are uncorrelated, moments hold, and the stream is a pure function of its identity
    ys = []
    for sub in range(1500):
        lib().hostsim_normals12(C.c_uint64(99), C.c_uint64(1 << 33 | 17), 4, 9, sub, buf)
        ys.append(np.array(buf[:]))
    y = np.stack(ys)
    assert abs(y.mean()) < 4 / np.sqrt(y.size) and abs(y.var() - 1) < 0.05
    assert np.abs(np.corrcoef(y[:-1].T, y[1:].T)[:12, 12:]).max() < 0.12          # lag-1 cross-correlation of the 12 components
    lib().hostsim_normals12(C.c_uint64(99), C.c_uint64(1 << 33 | 17), 4, 9, 1499, buf)
    assert np.array_equal(np.array(buf[:]), ys[-1])
    lib().hostsim_normals12(C.c_uint64(99), C.c_uint64(1 << 33 | 18), 4, 9, 1499, buf)
    assert not np.array_equal(np.array(buf[:]), ys[-1])


def test_all_philox_stream_normals_and_block_layout():
    """thermal_stream='philox' (STG_F_STREAM_PHILOX10): substep pair (2g, 2g+1) consumes Philox blocks 3g, 3g+1, 3g+2 of the
    env-step's counter space, one 16 + 16 bit word per Box-Muller pair; moments as for the default stream."""
    import ctypes as C
    from tests.hostsim.harness import lib
    buf = (C.c_float * 12)()
    xs = []
    for g in range(4000):
        lib().hostsim_normals12_philox(C.c_uint64(1234), C.c_uint64(g), 2, 3, 7 + (g & 1), buf)
        xs.append(np.array(buf[:]))
    x = np.concatenate(xs)
    assert abs(x.mean()) < 4 / np.sqrt(x.size) and abs(x.var() - 1) < 0.03 and abs((x ** 4).mean() - 3) < 0.2
    assert np.abs(x).max() < 4.86
    # words -> samples: the first pair of substep 0 comes from word 0 of block 0
    seed, gid, ep, step = 99, (1 << 33) | 17, 4, 9
    ph = (C.c_uint32 * 4)()
    lib().hostsim_philox(seed & 0xffffffff, seed >> 32, gid & 0xffffffff, ep, step, ((gid >> 32) << 20) + 0, ph)
    lib().hostsim_normals12_philox(C.c_uint64(seed), C.c_uint64(gid), ep, step, 0, buf)
    w = ph[0]
    u = ((w >> 16) + 0.5) / 65536.0
    r, ang = np.sqrt(-2.0 * np.log(u)), 2 * np.pi * (w & 0xffff) / 65536.0
    assert abs(buf[0] - r * np.cos(ang)) < 2e-3 and abs(buf[1] - r * np.sin(ang)) < 2e-3
    # substep 1 starts with the two words carried from block 1 (words 2, 3)
    lib().hostsim_philox(seed & 0xffffffff, seed >> 32, gid & 0xffffffff, ep, step, ((gid >> 32) << 20) + 1, ph)
    lib().hostsim_normals12_philox(C.c_uint64(seed), C.c_uint64(gid), ep, step, 1, buf)
    w = ph[2]
    u = ((w >> 16) + 0.5) / 65536.0
    r, ang = np.sqrt(-2.0 * np.log(u)), 2 * np.pi * (w & 0xffff) / 65536.0
    assert abs(buf[0] - r * np.cos(ang)) < 2e-3 and abs(buf[1] - r * np.sin(ang)) < 2e-3
    # and it is a different stream from the default one
    b2 = (C.c_float * 12)()
    lib().hostsim_normals12(C.c_uint64(seed), C.c_uint64(gid), ep, step, 1, b2)
    assert not np.array_equal(np.array(buf[:]), np.array(b2[:]))


@pytest.mark.parametrize("thermal", [False, True, "philox"])
def test_pair_path_is_bit_identical_to_scalar_path(thermal):
    """Two envs per thread (FP32x2 pack) vs one env per thread: same IEEE operations per component, so bit-identical results,
    for ragged substep counts (partners finish at different substeps), odd batch sizes and a permuted launch."""
    rng = np.random.default_rng(17)
    n, jm = 257, 1.1e-6
    m0 = rng.normal(size=(n, 3))
    m0[5] = [1e-20, -2e-21, 1.0]           # deep-pole state: exercises the block scaling
    tgt = np.where(rng.integers(2, size=(n, 1)) == 0, 1.0, -1.0) * np.array([[0, 0, 1.0]])
    act = np.stack([rng.uniform(-jm, jm, n), rng.uniform(1e-12, 1.2e-9, n)], 1).astype(np.float32)
    perm = rng.permutation(n).astype(np.int32)
    outs = []
    for pair in (True, False):
        h = HostSimEnv(n, max_current=jm, include_thermal_fluctuations=bool(thermal), f64=False, rng_seed=9, pair=pair,
                       autoreset=True, max_steps=2, thermal_stream="philox" if thermal == "philox" else "xoshiro")
        h.reset(m0, tgt)
        res = []
        for s in range(3):
            o, r, te, tr = h.step(act, perm=perm if s == 1 else None)
            res.append((o, r, te, tr, h.m.copy(), h.step_energy.copy(), h.final_obs.copy(), h.episode.copy()))
        outs.append(res)
    for a, b in zip(outs[0], outs[1]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_sampled_callables_grid_solver_vs_live_reference():
    """SimpleLLGSSolver with a time-dependent current_func / field_func: the host samples the callables at the
    reference's stage times and the FP64 grid integrator (integrate_grid, the same source the CUDA kernel compiles) must
    reproduce the live reference's trajectories (tests/golden/next.npz, grid/*) to 1e-6."""
    import types
    from tests.hostsim.harness import host_solve
    from spin_torque_rl_gym_b200.params import default_device_parameters
    from spin_torque_rl_gym_b200.physics.simple_solver import SimpleLLGSSolver
    G = np.load(os.path.join(GOLDEN, "next.npz"))
    p = default_device_parameters("stt_mram")
    stage_times = lambda t0, t1: SimpleLLGSSolver.stage_times(types.SimpleNamespace(max_step=1e-12), t0, t1)   # noqa: E731
    w = 2 * np.pi / 1.7e-10
    cur_sin = lambda t: 1.0e-6 * np.sin(w * t) + 2.0e-7                                  # noqa: E731
    cur_steps = lambda t: 9e-7 if t < 0.8e-10 else (-6e-7 if t < 2.1e-10 else 3e-7)     # noqa: E731
    cur_rect = lambda t: 8e-7 if t <= 1.3e-10 else 0.0                                   # noqa: E731
    fld_rot = lambda t: 2.0e5 * np.array([np.cos(w * t), np.sin(w * t), 0.3])           # noqa: E731
    fld_const = lambda t: np.array([1.0e5, -5.0e4, 2.0e4])                               # noqa: E731
    cases = {
        "sin_rk4": ("rk4", (0.0, 3.0e-10), cur_sin, None, 300),
        "sin_constfield_rk4": ("rk4", (0.0, 3.0e-10), cur_sin, fld_const, 300),
        "rect_rotfield_rk4": ("rk4", (0.0, 3.0e-10), cur_rect, fld_rot, 300),
        "steps_rotfield_rk4": ("rk4", (0.0, 3.0e-10), cur_steps, fld_rot, 300),
        "steps_rotfield_euler": ("euler", (0.0, 3.0e-10), cur_steps, fld_rot, 300),
        "sin_rotfield_offset_rk4": ("rk4", (1.0e-10, 3.5e-10), cur_sin, fld_rot, 249),
        "sin_short_rk4": ("rk4", (0.0, 4.0e-12), cur_sin, fld_rot, 100),
    }
    for name, (method, (t0, t1), cf, ff, n) in cases.items():
        times = stage_times(t0, t1)
        assert times.shape == (n, 3), name
        jg = np.array([[cf(x) for x in row] for row in times])
        hg = None if ff is None else np.array([[ff(x) for x in row] for row in times])
        traj, nsub = host_solve(G["grid/m0"], t1 - t0, p, method=method, current_grid=jg, field_grid=hg)
        want = G[f"grid/{name}/m"]
        assert nsub[0] == n and np.abs(traj[0, : n + 1] - want).max() < 1e-6, (name, np.abs(traj[0, : n + 1] - want).max())


def test_stage_times_agree_with_the_kernel_plan():
    """The host samples callables on `stage_times`; the kernel integrates `substep_plan(T)` substeps. Both must pick the same
    n for every duration (float32 pulse lengths included), and the times must be the floats the reference passes to
    current_func: t[i], t[i] + dt/2, t[i] + dt with t = linspace(t0, t1, n + 1) (physics/simple_solver.py:137-145, 278-295)."""
    import types
    from oracle.stt_oracle import substep_plan
    from tests.hostsim.harness import host_solve
    from spin_torque_rl_gym_b200.params import default_device_parameters
    from spin_torque_rl_gym_b200.physics.simple_solver import SimpleLLGSSolver
    ns = types.SimpleNamespace(max_step=1e-12)
    rng = np.random.default_rng(12)
    durs = np.concatenate([rng.uniform(1e-12, 5e-9, 200).astype(np.float32).astype(np.float64),
                           [1e-9, 5e-9, 2.5e-10, 9.99e-11, 1e-10, 1.0000001e-10, 3.3e-12, 1e-12, 7e-13]])
    for T in durs:
        times = SimpleLLGSSolver.stage_times(ns, 0.0, float(T))
        n, dt = substep_plan(float(T))[:2]
        assert times.shape == (n, 3), T
        t = np.linspace(0.0, float(T), n + 1)[:n]
        assert np.array_equal(times[:, 0], t) and np.array_equal(times[:, 1], t + dt / 2) and np.array_equal(times[:, 2], t + dt)
    # and the kernel body really runs that many substeps when it is handed the grid
    p = default_device_parameters("stt_mram")
    for T in (3.3e-12, 9.99e-11, float(np.float32(1e-9)), 2.5000000000000003e-10):
        times = SimpleLLGSSolver.stage_times(ns, 0.0, T)
        _, nsub = host_solve([0.3, 0.2, 0.9], T, p, current_grid=np.full(times.shape, 5e-7))
        assert nsub[0] == times.shape[0], T
