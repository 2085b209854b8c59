"""SURVEY §8(f) adjacent components on the GPU kernels, against goldens from the live reference (tests/golden/next.npz):
VectorizedSolver.solve_batch, VectorizedMagneticsOperations, EnergyLandscape.compute_energy / compute_energy_gradient /
generate_phase_diagram, ThermalFluctuations analytics, LLGSSolver.find_stable_states."""
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
    # compute_energy_barrier (physics/energy_landscape.py:179-221): 41 path points in one launch
    bh, path = land.compute_energy_barrier(G["barrier/s0"], G["barrier/s1"], G["barrier/happ"], n_intermediate=41)
    scale = np.abs(G["barrier/path"]).max()
    assert path.shape == (41,) and np.allclose(path, G["barrier/path"], rtol=0, atol=1e-12 * scale)
    assert float(G["barrier/height"]) > 0 and bh == pytest.approx(float(G["barrier/height"]), rel=1e-10)
    assert land.compute_energy_barrier(G["barrier/s0"], G["barrier/s1"])[0] == pytest.approx(
        float(G["barrier/height_nofield"]), rel=1e-10)


def test_find_stable_states(cuda_device):
    """Relaxation without current ends on the easy axis: the default device has exactly the two states +-z."""
    from spin_torque_rl_gym_b200.physics import LLGSSolver
    from spin_torque_rl_gym_b200.params import default_device_parameters
    s = LLGSSolver(device=cuda_device)
    states = s.find_stable_states(default_device_parameters("stt_mram"), n_trials=64, threshold=1e-3, relax_time=4e-9, seed=0)
    assert states.shape == (2, 3)
    assert np.allclose(np.sort(states[:, 2]), [-1.0, 1.0], atol=1e-4) and np.abs(states[:, :2]).max() < 1e-2


def test_vectorized_magnetics_operations_bit_exact(cuda_device):
    """utils/vectorized_operations.py:288-393 of the reference: every helper reproduces NumPy's roundings exactly, including
    the 1e-12 norm floor (zero row) and the R_P/2 resistance floor (rows 0-5)."""
    import torch
    from spin_torque_rl_gym_b200.physics import VectorizedMagneticsOperations as VMO
    a, b = G["vmo/a"], G["vmo/b"]
    assert np.array_equal(VMO.batch_cross_product(a, b), G["vmo/cross"])
    assert np.array_equal(VMO.batch_dot_product(a, b), G["vmo/dot"])
    nrm = VMO.batch_normalize(a)
    assert np.array_equal(nrm, G["vmo/normalize"]) and np.array_equal(nrm[5], np.zeros(3))
    pb = {"uniaxial_anisotropy": G["vmo/k_u"], "volume": G["vmo/volume"], "easy_axis": G["vmo/easy"]}
    assert np.array_equal(VMO.batch_energy_computation(b, pb), G["vmo/energy"])
    assert np.array_equal(VMO.batch_energy_computation(b, {}), G["vmo/energy_default"])
    pb["easy_axis"] = np.array([0.0, 0.6, 0.8])
    assert np.array_equal(VMO.batch_energy_computation(b, pb), G["vmo/energy_one_axis"])
    r = VMO.batch_resistance_computation(G["vmo/bn"], G["vmo/ref"], G["vmo/r_p"], G["vmo/r_ap"])
    assert np.array_equal(r, G["vmo/resistance"]) and np.array_equal(r[:6], 0.5 * G["vmo/r_p"][:6])
    # CUDA tensors stay on the device; a large batch agrees with NumPy on the host
    rng = np.random.default_rng(3)
    x, y = rng.normal(size=(1 << 20, 3)), rng.normal(size=(1 << 20, 3))
    tx, ty = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)
    c = VMO.batch_cross_product(tx, ty)
    assert c.is_cuda and np.array_equal(c.cpu().numpy(), np.cross(x, y, axis=1))
    assert np.array_equal(VMO.batch_dot_product(tx, ty).cpu().numpy(), np.sum(x * y, axis=1))
    with pytest.raises(ValueError):
        VMO.batch_cross_product(a, b[:7])
    with pytest.raises(Exception):
        VMO.batch_dot_product(torch.from_numpy(a), torch.from_numpy(b))      # CPU tensors: no CPU path


def test_phase_diagram_and_stability_factor(cuda_device):
    """physics/energy_landscape.py:282-359: identical 0/1 map on a grid that straddles the h_k - |beta I| boundary."""
    from spin_torque_rl_gym_b200.physics import EnergyLandscape
    from spin_torque_rl_gym_b200.params import default_device_parameters
    p = default_device_parameters("stt_mram")
    p.update(volume=float(G["phase/volume"]))
    land = EnergyLandscape(p, device=cuda_device)
    i_max, h_max = float(G["phase/i_max"]), float(G["phase/h_max"])
    pd = land.generate_phase_diagram((-i_max, i_max), (-h_max, h_max), resolution=37)
    assert np.array_equal(pd["currents"], G["phase/currents"]) and np.array_equal(pd["fields"], G["phase/fields"])
    want = G["phase/switching_probability"]
    assert 0.1 < want.mean() < 0.9                                  # both phases present
    assert np.array_equal(pd["switching_probability"], want)
    assert land.compute_thermal_stability_factor(300.0) == float(G["phase/stability_300"])
    assert land.compute_thermal_stability_factor(77.0) == float(G["phase/stability_77"])
    assert land.compute_thermal_stability_factor(0.0) == float("inf")
    big = land.generate_phase_diagram((-i_max, i_max), (-h_max, h_max), resolution=1024)["switching_probability"]
    assert big.shape == (1024, 1024) and abs(big.mean() - want.mean()) < 0.02


def test_thermal_analytics(cuda_device):
    """physics/thermal_model.py:139-336: sample_switching_time (same host stream for the same seed), the temperature sweep
    through the 0 -> 1 switching transition, and analyze_thermal_stability."""
    from spin_torque_rl_gym_b200.physics import ThermalFluctuations
    tp = dict(volume=1.5e-25, uniaxial_anisotropy=1.1e6, damping=0.02, saturation_magnetization=7.5e5)
    th = ThermalFluctuations(temperature=320.0, seed=9, device=cuda_device)
    barrier = float(G["th/barrier_J"])
    assert np.array_equal([th.sample_switching_time(barrier) for _ in range(5)], G["th/switch_times"])
    sweep = th.generate_temperature_sweep((50.0, 450.0), tp, n_points=23)
    assert th.temperature == float(G["th/temperature_after_sweep"]) == 320.0
    assert np.array_equal(sweep["temperature"], G["th/sweep/temperature"])
    ps = G["th/sweep/switching_probability"]
    assert ps.min() == 0.0 and ps.max() == 1.0 and ((ps > 1e-3) & (ps < 0.99)).sum() >= 3
    for k in ("thermal_stability_factor", "switching_probability", "retention_time", "noise_strength"):
        assert np.allclose(sweep[k], G[f"th/sweep/{k}"], rtol=1e-12, atol=0), k
    st = th.analyze_thermal_stability(tp, time_scale=3.0)
    for k, v in st.items():
        want = G[f"th/stability/{k}"]
        assert (bool(v) == bool(want)) if isinstance(v, (bool, np.bool_)) else v == pytest.approx(float(want), rel=1e-12), k
    cold = ThermalFluctuations(temperature=0.0, device=cuda_device)
    assert cold.sample_switching_time(barrier) == float("inf")
