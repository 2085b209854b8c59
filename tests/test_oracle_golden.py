"""The oracle is pinned: NumPy restatement == live reference (bit-exact), C restatement within 1e-12, on the golden vectors
generated from /root/reference by oracle/gen_golden.py."""
import numpy as np
import pytest

from tests.helpers import GOLDEN, load_case, noise_for_step, rel_err, stt_params_for
from oracle.c_oracle import COracleEnv
from oracle.stt_oracle import SttOracleEnv, substep_plan

CASES = ["bigvol_det", "tilted_rk4", "tilted_euler", "thermal_injected"]


def _np_env(c):
    return SttOracleEnv(device_params=stt_params_for(c), max_current=float(c["max_current"]),
                        include_thermal=bool(c["thermal"]), method=str(c["method"]),
                        max_steps=int(c.get("max_steps", 100)))


@pytest.mark.parametrize("name,nsteps", [("bigvol_det", 10), ("tilted_rk4", 6), ("tilted_euler", 16),
                                         ("thermal_injected", 5), ("c1_det", 2)])
def test_numpy_oracle_bit_exact(name, nsteps):
    c = load_case("stt_env.npz", name)
    env = _np_env(c)
    obs = env.reset(c["m0"], c["target"])
    assert np.array_equal(obs, c["obs"][0])
    for k in range(nsteps):
        a = c["actions"][k]
        noise = noise_for_step(c["seeds"][k], a, env.max_current) if "seeds" in c else None
        o, r, te, tr, info = env.step(a.copy(), noise)
        assert np.array_equal(env.m, c["m"][k + 1]), k
        assert np.array_equal(o, c["obs"][k + 1]), k
        assert r == c["reward"][k] and info["energy"] == c["energy"][k]
        assert te == c["terminated"][k] and tr == c["truncated"][k]


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_golden(name):
    c = load_case("stt_env.npz", name)
    env = COracleEnv(1, device_params=stt_params_for(c), max_current=float(c["max_current"]),
                     include_thermal=bool(c["thermal"]), method=str(c["method"]), max_steps=int(c.get("max_steps", 100)))
    obs = env.reset(c["m0"], c["target"])
    assert np.array_equal(obs[0], c["obs"][0])
    tol = 1e-9 if name == "tilted_euler" else 1e-12      # the Euler map amplifies last-bit differences
    for k, a in enumerate(c["actions"]):
        noise = noise_for_step(c["seeds"][k], a, env.p.max_current)[None] if "seeds" in c else None
        o, r, te, tr = env.step(a[None], noise)
        assert rel_err(env.m[0], c["m"][k + 1]) < tol, k
        assert np.abs(o[0] - c["obs"][k + 1]).max() <= 1e-6
        assert abs(r[0] - c["reward"][k]) <= 1e-9 * max(1.0, abs(c["reward"][k]))
        assert te[0] == c["terminated"][k] and tr[0] == c["truncated"][k]
        assert env.n_sub[0] == substep_plan(float(env.last_action[0, 1]))[0]


def test_c_oracle_c1_teacher_forced():
    """C1 (100 random pulses, no reset) collapses onto a pole with transverse components down to denormals, where any
    last-bit difference decides later switch times. Each step is therefore checked from the golden's own pre-step state."""
    c = load_case("stt_env.npz", "c1_det")
    env = COracleEnv(1, device_params=stt_params_for(c), max_current=float(c["max_current"]), include_thermal=False)
    env.reset(c["m0"], c["target"])
    for k, a in enumerate(c["actions"]):
        env.m[0] = c["m"][k]
        env.total_energy[0] = c["total_energy"][k - 1] if k else 0.0
        env.step_count[0] = k
        o, r, te, tr = env.step(a[None])
        if np.hypot(*c["m"][k][:2]) > 1e-200:
            assert rel_err(env.m[0], c["m"][k + 1]) < 1e-11, k
        assert np.abs(env.m[0] - c["m"][k + 1]).max() < 1e-12
        assert r[0] == pytest.approx(c["reward"][k], rel=1e-12, abs=1e-12)
        assert te[0] == c["terminated"][k] and tr[0] == c["truncated"][k]


def test_c_oracle_multi_episode_golden():
    import os
    g = np.load(os.path.join(GOLDEN, "stt_multi.npz"))
    n = len(g["actions"])
    env = COracleEnv(n, max_current=float(g["max_current"]), include_thermal=False, nthreads=4)
    obs0 = env.reset(g["m0"], g["target"])
    assert np.array_equal(obs0, g["obs0"])
    o, r, te, tr = env.step(g["actions"])
    assert rel_err(env.m, g["m"], floor=1e-30) < 1e-10
    assert np.abs(o - g["obs"]).max() <= 1e-6
    assert np.allclose(r, g["reward"], rtol=1e-11, atol=1e-11)
    assert np.array_equal(te, g["terminated"]) and np.array_equal(tr, g["truncated"])
    assert np.allclose(env.step_energy, g["energy"], rtol=1e-12, atol=0)


def test_thermal_equator_golden_is_noise_driven():
    """Reference behaviour pinned at the unstable equilibrium: m_z after 100 substeps is O(1e-8), sign set by the noise."""
    c = load_case("stt_env.npz", "thermal_equator")
    mz = c["m_final"][:, 2]
    assert np.all(np.abs(mz) < 1e-6) and np.all(np.abs(mz) > 1e-12)
    env = SttOracleEnv(max_current=float(c["max_current"]), include_thermal=True, temperature=300.0)
    for s, ref in zip(c["seeds"], c["m_final"]):
        env.reset([1.0, 0.0, 0.0], [0.0, 0.0, 1.0])
        env.step(c["actions"][0].copy(), noise_for_step(s, c["actions"][0], env.max_current))
        assert np.array_equal(env.m, ref)
