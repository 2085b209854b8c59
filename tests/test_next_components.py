"""SURVEY §8(f) adjacent components on the GPU kernels, against goldens from the live reference (tests/golden/next.npz):
VectorizedSolver.solve_batch, EnergyLandscape.compute_energy / compute_energy_gradient, LLGSSolver.find_stable_states."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(GOLDEN, "next.npz"))


def _plist():
    from spin_torque_rl_gym_b200.params import default_device_parameters
    out = []
    for i in range(len(G["m0"])):
        p = default_device_parameters("stt_mram")
        p.update(damping=float(G["vec/damping"][i]), uniaxial_anisotropy=float(G["vec/uniaxial_anisotropy"][i]),
                 saturation_magnetization=float(G["vec/saturation_magnetization"][i]), easy_axis=G["vec/easy_axis"][i])
        out.append(p)
    return out


def test_vectorized_solver_batch_api(cuda_device):
    from spin_torque_rl_gym_b200.physics import VectorizedSolver
    plist = _plist()
    res = VectorizedSolver(device=cuda_device).solve_batch(G["m0"], (0, 1.5e-10), plist, dt=1e-12)
    assert len(res) == len(plist) and res[0]["n_steps"] == 150 and res[0]["vectorized"]
    assert np.allclose(res[0]["t"], G["vec/t"], rtol=1e-15)
    got = np.array([r["m"] for r in res])
    assert np.abs(got - G["vec/m"]).max() < 1e-6          # 150 Euler steps: rounding amplified by the Euler map
    res2 = VectorizedSolver(device=cuda_device).solve_batch(G["m0"], (0, 3e-12), plist[:1] * len(plist), dt=1e-12)
    assert res2[0]["n_steps"] == 10                        # max(10, int(T/dt)): not SimpleLLGSSolver's >= 100
    assert np.abs(np.array([r["m"] for r in res2]) - G["vec_short/m"]).max() < 1e-9
    big = VectorizedSolver(device=cuda_device).solve_batch_tensors(np.tile(G["m0"], (4096, 1)), (0, 1e-10), plist[:1],
                                                                   return_trajectory=False)
    assert big["m_final"].shape == (4096 * len(plist), 3)
    assert float((big["m_final"].norm(dim=1) - 1).abs().max()) < 1e-12


def test_energy_landscape_batch(cuda_device):
    from spin_torque_rl_gym_b200.physics import EnergyLandscape
    from spin_torque_rl_gym_b200.params import default_device_parameters
    p = default_device_parameters("stt_mram")
    p.update(demag_factors=np.array([0.1, 0.3, 0.6]), easy_axis=np.array([0.0, 0.6, 0.8]))
    land = EnergyLandscape(p, device=cuda_device)
    e = land.compute_energy(G["land/m"], G["land/happ"])
    assert np.allclose(e, G["land/energy"], rtol=1e-12, atol=1e-12 * np.abs(G["land/energy"]).max())
    g = land.compute_energy_gradient(G["land/m"], G["land/happ"])
    assert np.allclose(g, G["land/grad"], rtol=1e-12, atol=1e-6)
    assert land.compute_energy(G["land/m"][0]) == pytest.approx(G["land/energy0"][0], rel=1e-12)
    assert land.compute_energy_gradient(G["land/m"][1], G["land/happ"][1]).shape == (3,)


def test_find_stable_states(cuda_device):
    """Relaxation without current ends on the easy axis: the default device has exactly the two states +-z."""
    from spin_torque_rl_gym_b200.physics import LLGSSolver
    from spin_torque_rl_gym_b200.params import default_device_parameters
    s = LLGSSolver(device=cuda_device)
    states = s.find_stable_states(default_device_parameters("stt_mram"), n_trials=64, threshold=1e-3, relax_time=4e-9, seed=0)
    assert states.shape == (2, 3)
    assert np.allclose(np.sort(states[:, 2]), [-1.0, 1.0], atol=1e-4) and np.abs(states[:, :2]).max() < 1e-2
