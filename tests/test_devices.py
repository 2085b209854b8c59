"""K4 device-class operations: the NumPy oracle is pinned to the live-reference goldens (CPU); the CUDA kernels are checked
against both on the GPU, together with the reference's own known answers (tests/unit/test_devices.py of the reference)."""
import os

import numpy as np
import pytest

from oracle import devices_oracle as DO
from tests.helpers import GOLDEN

G = np.load(os.path.join(GOLDEN, "devices.npz"))
VARIANTS = {"stt": "stt_mram", "sot": "sot_mram", "sot_ar": "sot_mram", "vcma": "vcma_mram"}


def _params(key):
    from spin_torque_rl_gym_b200.params import default_device_parameters
    p = default_device_parameters(VARIANTS[key])
    for k in G.files:
        if k.startswith(f"{key}/param/"):
            v = G[k]
            p[k.split("/")[-1]] = v if v.ndim else float(v)
    return p


@pytest.mark.parametrize("key", list(VARIANTS))
def test_oracle_pinned_to_reference(key):
    p, kind = _params(key), VARIANTS[key]
    m, happ, volt, cur = G["m"], G["happ"], G["volt"], G["cur"]
    f = DO.effective_field(kind, p, m, happ, volt if kind == "vcma_mram" else None)
    assert np.allclose(f, G[f"{key}/field"], rtol=1e-14, atol=1e-9)
    assert np.allclose(DO.resistance(kind, p, m), G[f"{key}/resistance"], rtol=1e-14)
    if kind == "sot_mram":
        dl, fl = DO.sot_torque(p, cur, m, [1.0, 2.0, 0.0])
        assert np.allclose(dl, G[f"{key}/tau_dl"], rtol=1e-13, atol=1e-9) and np.allclose(fl, G[f"{key}/tau_fl"], rtol=1e-13)
    if kind == "vcma_mram":
        assert np.allclose(DO.vcma_keff(p, volt), G[f"{key}/keff"], rtol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("key", list(VARIANTS))
def test_cuda_device_ops_match_reference(key, cuda_device):
    import torch
    from spin_torque_rl_gym_b200.devices import DeviceFactory
    p, kind = _params(key), VARIANTS[key]
    dev = DeviceFactory(cuda_device).create_device(kind, p)
    m, happ, volt, cur = G["m"], G["happ"], G["volt"], G["cur"]
    args = (m, happ, volt) if kind == "vcma_mram" else (m, happ)
    f = dev.compute_effective_field(*args)
    assert isinstance(f, np.ndarray) and np.allclose(f, G[f"{key}/field"], rtol=1e-13, atol=1e-8)
    # torch in -> torch out (device resident), one row in -> one row out
    ft = dev.compute_effective_field(torch.from_numpy(m).to(cuda_device), torch.from_numpy(happ).to(cuda_device),
                                     *( [torch.from_numpy(volt).to(cuda_device)] if kind == "vcma_mram" else []))
    assert ft.is_cuda and np.allclose(ft.cpu().numpy(), G[f"{key}/field"], rtol=1e-13, atol=1e-8)
    f0 = dev.compute_effective_field(m[0], happ[0], *([volt[0]] if kind == "vcma_mram" else []))
    assert f0.shape == (3,) and np.allclose(f0, G[f"{key}/field"][0], rtol=1e-13, atol=1e-8)
    assert np.allclose(dev.compute_resistance(m), G[f"{key}/resistance"], rtol=1e-13)
    assert dev.compute_resistance(m[3]) == pytest.approx(G[f"{key}/resistance"][3], rel=1e-13)
    if kind == "sot_mram":
        dl, fl = dev.compute_spin_torque(cur, m, np.array([1.0, 2.0, 0.0]))
        assert np.allclose(dl, G[f"{key}/tau_dl"], rtol=1e-12, atol=1e-9) and np.allclose(fl, G[f"{key}/tau_fl"], rtol=1e-12)
        dl0, _ = dev.compute_spin_torque(cur, m)
        assert np.allclose(dl0, G[f"{key}/tau_dl_default"], rtol=1e-12, atol=1e-9)
        assert np.abs((dl0 * m).sum(1)).max() < 1e-10 * np.abs(dl0).max()            # tau_DL is perpendicular to m
        assert dev.compute_power_consumption(cur[0], 1e-9, m[0]) == pytest.approx(G[f"{key}/power"][0], rel=1e-13)
    if kind == "vcma_mram":
        assert np.allclose(dev._compute_effective_anisotropy(volt), G[f"{key}/keff"], rtol=1e-14)
        assert dev._compute_effective_anisotropy(0.0) == p["uniaxial_anisotropy"]
        assert dev._compute_effective_anisotropy(10.0) == dev._compute_effective_anisotropy(p["breakdown_voltage"])
        assert dev.compute_power_consumption(volt[0], 1e-9) == pytest.approx(G[f"{key}/power"][0], rel=1e-13)
        assert dev.compute_switching_probability(volt[1], 1e-9) == pytest.approx(G[f"{key}/pswitch"][1], rel=1e-9, abs=1e-300)


@pytest.mark.gpu
def test_reference_known_answers(cuda_device):
    """tests/unit/test_devices.py:79-94 of the reference: R(+z)=R_P, R(-z)=R_AP, R_P < R(x) < R_AP."""
    from spin_torque_rl_gym_b200.devices import create_device
    from spin_torque_rl_gym_b200 import _lib
    dev = create_device("stt_mram", device=cuda_device)
    assert dev.compute_resistance(np.array([0.0, 0.0, 1.0])) == pytest.approx(1e3)
    assert dev.compute_resistance(np.array([0.0, 0.0, -1.0])) == pytest.approx(2e3)
    assert 1e3 < dev.compute_resistance(np.array([1.0, 0.0, 0.0])) < 2e3
    f = dev.compute_effective_field(np.array([0.0, 0.0, 2.0]), np.zeros(3))
    assert f.shape == (3,) and f[2] > 0
    with pytest.raises(ValueError):
        dev.compute_resistance(np.zeros(3))
    with pytest.raises(RuntimeError, match="Missing required parameter"):
        from spin_torque_rl_gym_b200.devices import DeviceFactory
        DeviceFactory(cuda_device).create_device("stt_mram", {"volume": 1e-24})


@pytest.mark.gpu
def test_thermal_field_statistics(cuda_device):
    """tests/test_comprehensive_suite.py:447-476 of the reference: std within 20 % of compute_noise_strength, |mean| < 0.1 std;
    OU process: stationary variance 1 and lag-1 autocorrelation exp(-dt/tau_c)."""
    from spin_torque_rl_gym_b200.physics import ThermalFluctuations
    th = ThermalFluctuations(temperature=350.0, correlation_time=2e-12, seed=1, num_devices=20000, device=cuda_device)
    s = th.compute_noise_strength(0.01, 800e3, 1e-23)
    assert s == pytest.approx(float(G["thermal/strength"]), rel=1e-14)
    assert th.compute_thermal_barrier(1.2e6, 1e-23) == pytest.approx(float(G["thermal/barrier"]), rel=1e-14)
    assert th.compute_switching_probability(1.2e6 * 1e-23 * 0.05) == pytest.approx(float(G["thermal/pswitch"]), rel=1e-12)
    assert th.compute_retention_time(1.2e6 * 1e-23 * 0.05) == pytest.approx(float(G["thermal/retention"]), rel=1e-12)
    w = th.generate_thermal_field(0.01, 800e3, 1e-23, 1e-12, correlated=False)
    assert abs(float(w.std()) / s - 1) < 0.02 and abs(float(w.mean())) < 0.02 * s
    prev = None
    for _ in range(40):       # burn-in towards the stationary OU state
        prev = th.generate_thermal_field(0.01, 800e3, 1e-23, 1e-12, correlated=True).clone()
    cur = th.generate_thermal_field(0.01, 800e3, 1e-23, 1e-12, correlated=True)
    assert abs(float(cur.std()) / s - 1) < 0.03
    rho = float(((cur * prev).mean() / (cur.std() * prev.std())))
    assert abs(rho - np.exp(-0.5)) < 0.03


@pytest.mark.gpu
def test_zero_magnetisation_raises_like_the_reference(cuda_device):
    """devices/base_device.py:112-114: a zero vector is a ValueError for the STT device (validate_magnetization normalises).
    The check runs inside the kernel (a 4-byte row counter), also for one bad row inside a large batch."""
    import torch
    from spin_torque_rl_gym_b200.devices import create_device
    from spin_torque_rl_gym_b200.params import default_device_parameters
    stt = create_device("stt_mram", default_device_parameters("stt_mram"), device=cuda_device)
    for bad in (np.zeros(3), np.array([0.0, 1e-13, 0.0])):
        with pytest.raises(ValueError, match="cannot be zero"):
            stt.compute_resistance(bad)
        with pytest.raises(ValueError, match="cannot be zero"):
            stt.compute_effective_field(bad, np.zeros(3))
    m = torch.randn(100003, 3, dtype=torch.float64, device=cuda_device)
    good_r = stt.compute_resistance(m)
    good_h = stt.compute_effective_field(m, np.zeros(3))
    assert torch.isfinite(good_r).all() and torch.isfinite(good_h).all()
    m[77777] = 0.0
    with pytest.raises(ValueError, match="cannot be zero"):
        stt.compute_resistance(m)
    with pytest.raises(ValueError, match="cannot be zero"):
        stt.compute_effective_field(m, np.zeros(3))
    r = stt.compute_resistance(m, check_zero=False)                 # asynchronous use: no read-back, no exception;
    assert torch.equal(r[:77777], good_r[:77777]) and torch.equal(r[77778:], good_r[77778:])   # the bad row is unspecified
    m[77777] = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    assert torch.equal(stt.compute_resistance(m)[:77777], good_r[:77777])          # the counter was reset
    p = default_device_parameters("sot_mram")
    sot = create_device("sot_mram", p, device=cuda_device)
    assert np.isfinite(sot.compute_resistance(np.zeros(3)))         # SOT/VCMA do not normalise and do not raise
