"""Drop-in API: single-env façade with the reference's step/reset/info contract, make(), batched solvers with the reference's
signatures, the SB3 VecEnv protocol and the GPU-resident rollout collector."""
import numpy as np
import pytest

from tests.helpers import load_case, stt_params_for

pytestmark = pytest.mark.gpu


def test_single_env_facade_reproduces_reference_episode(cuda_device):
    import spin_torque_rl_gym_b200 as stg
    c = load_case("stt_env.npz", "bigvol_det")
    env = stg.make("SpinTorque-v0", device_type="stt_mram", device_params=stt_params_for(c),
                   include_thermal_fluctuations=False, max_steps=int(c["max_steps"]), device=cuda_device)
    assert env.action_space.shape == (2,) and env.observation_space.shape == (12,)
    obs, info = env.reset(options={"initial_state": c["m0"], "target_state": c["target"]})
    assert obs.dtype == np.float32 and np.array_equal(obs, c["obs"][0])
    for k in ("step_count", "total_energy", "current_alignment", "is_success", "target_reached", "magnetization_magnitude",
              "device_type"):
        assert k in info
    for k, a in enumerate(c["actions"]):
        obs, r, te, tr, info = env.step(a.copy())
        assert isinstance(r, float) and isinstance(te, bool) and isinstance(tr, bool)
        assert np.abs(obs - c["obs"][k + 1]).max() < 1e-6
        assert r == pytest.approx(c["reward"][k], rel=1e-6, abs=1e-6)
        assert te == c["terminated"][k] and tr == c["truncated"][k]
        assert info["energy_consumed"] == pytest.approx(c["energy"][k], rel=1e-9)
        assert info["simulation_success"] is True and info["step_count"] == k + 1
        for key in ("final_magnetization", "pulse_duration", "current_density", "alignment_improvement"):
            assert key in info
    rep = env.analyze_episode()
    assert rep["episode_length"] == len(c["actions"]) and rep["total_energy"] == pytest.approx(c["total_energy"][-1], rel=1e-9)
    with pytest.raises(RuntimeError):
        stg.make("SpinTorque-v0", device_type="sot_mram", device=cuda_device)      # the reference ctor fails the same way


def test_simple_solver_signature_and_trajectory(cuda_device):
    """SimpleLLGSSolver.solve(m, t_span, params, current_func, field_func, ...) against the NumPy oracle trajectory."""
    from oracle.stt_oracle import integrate, prepare_params
    from spin_torque_rl_gym_b200.physics import RobustLLGSSolver, SimpleLLGSSolver
    from spin_torque_rl_gym_b200.params import default_device_parameters
    p = default_device_parameters("stt_mram")
    m0 = np.array([0.5, -0.3, 0.8])
    J, T = 6e-7, 3.3e-10
    for method in ("rk4", "euler"):
        s = SimpleLLGSSolver(method=method, device=cuda_device)
        r = s.solve(m0, (0, T), p, lambda t: J if t <= T else 0.0, lambda t: np.zeros(3), thermal_noise=False)
        traj, n, _ = integrate(m0, J, T, prepare_params(p, False, 300.0), method, return_traj=True)
        assert r["success"] and r["n_steps"] == n and r["m"].shape == (n + 1, 3) and len(r["t"]) == n + 1
        assert np.abs(r["m"] - traj).max() < (1e-9 if method == "rk4" else 1e-6)
        assert set(r) >= {"t", "m", "success", "message", "solve_time", "n_steps"}
    # pulse shorter than the span: current_func is sampled and recognised as a rectangular pulse
    s = SimpleLLGSSolver(method="rk4", device=cuda_device)
    r = s.solve(m0, (0, T), p, lambda t: J if t <= 1.2e-10 else 0.0)
    traj, n, _ = integrate(m0, J, 1.2e-10, prepare_params(p, False, 300.0), "rk4", return_traj=True, t_end=T)
    assert np.abs(r["m"] - traj).max() < 1e-9
    ramp = s.solve(m0, (0, T), p, lambda t: J * t / T)       # not rectangular: sampled at the stage times (grid kernel)
    assert ramp["success"] and ramp["m"].shape == (n + 1, 3) and np.abs(np.linalg.norm(ramp["m"], axis=1) - 1).max() < 1e-12
    assert np.abs(ramp["m"][-1] - r["m"][-1]).max() > 1e-3   # and it is not the rectangular-pulse answer
    assert s.solve(m0, (1e-9, 1e-9), p)["message"].startswith("Trivial")
    rb = RobustLLGSSolver(method="rk4", device=cuda_device)
    bad = rb.solve(m0, (0, T), default_device_parameters("sot_mram"), lambda t: J)
    assert bad["success"] is False and bad["is_fallback"] and np.array_equal(bad["m"][0], m0)


def test_simple_solver_arbitrary_callables_vs_live_reference(cuda_device):
    """current_func / field_func that are not a rectangular pulse / a constant are sampled by the host at the
    reference's stage times (t_i, t_i+dt/2, t_i+dt) and integrated by the FP64 grid kernel. Goldens: the live reference's
    SimpleLLGSSolver.solve with the same callables (oracle/gen_golden.py gen_next), 1e-6 as north_star asks for FP64."""
    import os
    import torch
    from tests.helpers import GOLDEN
    from spin_torque_rl_gym_b200.physics import SimpleLLGSSolver
    from spin_torque_rl_gym_b200.params import default_device_parameters
    G = np.load(os.path.join(GOLDEN, "next.npz"))
    p = default_device_parameters("stt_mram")
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
    m0 = G["grid/m0"]
    for name, (method, span, cf, ff, n) in cases.items():
        for dtype in (torch.float64, torch.float32):          # both entry points run the FP64 grid stages
            s = SimpleLLGSSolver(method=method, device=cuda_device, dtype=dtype)
            r = s.solve(m0, span, p, cf, ff)
            want = G[f"grid/{name}/m"]
            assert r["success"] and r["n_steps"] == n and r["m"].shape == want.shape, name
            assert np.array_equal(r["t"], G[f"grid/{name}/t"]), name
            assert np.abs(r["m"] - want).max() < 1e-6, (name, np.abs(r["m"] - want).max())
    # the sampled grids can also be passed for a whole batch: every trajectory shares one grid here
    s = SimpleLLGSSolver(method="rk4", device=cuda_device)
    times = s.stage_times(0.0, 3.0e-10)
    jg = np.array([[cur_sin(x) for x in row] for row in times])
    hg = np.array([[fld_rot(x) for x in row] for row in times])
    rb = s.solve_batch(np.tile(m0, (1000, 1)), np.full(1000, 3.0e-10), p, current_grid=jg, field_grid=hg)
    want = s.solve(m0, (0.0, 3.0e-10), p, cur_sin, fld_rot)["m"][-1]
    assert np.array_equal(rb["m"].cpu().numpy(), np.tile(want, (1000, 1)))
    with pytest.raises(ValueError):
        s.solve_batch(m0[None], np.array([3.0e-10]), p, current_grid=jg[:, :2])


def test_sb3_vecenv_protocol_and_rollout_collector(cuda_device):
    import torch
    import spin_torque_rl_gym_b200 as stg
    n = 512
    env = stg.make("SpinTorque-v0", num_envs=n, device=cuda_device, max_current=1.1e-6, max_steps=5, rng_seed=3)
    venv = stg.SB3VecEnvAdapter(env)
    venv.seed(3)
    obs = venv.reset()
    assert obs.shape == (n, 12) and obs.dtype == np.float32 and venv.num_envs == n
    rng = np.random.default_rng(0)
    n_done = 0
    for _ in range(6):
        act = np.stack([rng.uniform(-1.1e-6, 1.1e-6, n), rng.uniform(1e-11, 2e-10, n)], 1).astype(np.float32)
        venv.step_async(act)
        obs, rew, dones, infos = venv.step_wait()
        assert obs.shape == (n, 12) and rew.dtype == np.float32 and dones.dtype == bool and len(infos) == n
        for i in np.nonzero(dones)[0]:
            assert infos[i]["terminal_observation"].shape == (12,) and "TimeLimit.truncated" in infos[i]
            assert obs[i, 8] == 1.0                       # already the first observation of the next episode
        n_done += int(dones.sum())
    assert n_done >= n                                    # max_steps=5 ends every episode within 6 steps
    assert venv.get_attr("max_steps")[0] == 5 and venv.env_is_wrapped(object) == [False] * n
    # GPU-resident collector with a torch policy
    policy_w = torch.randn(12, 2, device=cuda_device) * 1e-7

    def policy(o):
        a = o @ policy_w
        a[:, 1] = 1e-10
        return a
    col = stg.RolloutCollector(env, n_steps=16)
    stats = col.collect(policy)
    assert col.observations.shape == (16, n, 12) and col.rewards.shape == (16, n) and col.dones.dtype == torch.bool
    assert stats["steps"] >= 16 * n and 0.0 <= stats["success_rate"] <= 1.0
    assert torch.isfinite(col.rewards).all()
