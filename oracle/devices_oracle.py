"""CPU oracle for the device-class operations (NumPy restatement, vectorised over rows). TEST INFRASTRUCTURE ONLY.

Parity status: PINNED against tests/golden/devices.npz (generated from the live reference by oracle/gen_golden.py devices).
Reference lines: devices/stt_mram.py:56-94, devices/sot_mram.py:61-132,163-255, devices/vcma_mram.py:86-166,187-287.
"""
from __future__ import annotations

import numpy as np

MU0 = 4 * np.pi * 1e-7


def demag_factors(aspect_ratio):
    """devices/sot_mram.py:120-129."""
    if aspect_ratio >= 1.0:
        nx, ny = 1.0 / (1.0 + aspect_ratio), aspect_ratio / (1.0 + aspect_ratio)
    else:
        nx, ny = aspect_ratio / (1.0 + aspect_ratio), 1.0 / (1.0 + aspect_ratio)
    return np.array([nx, ny, 1.0 - nx - ny])


def vcma_keff(p, v):
    """devices/vcma_mram.py:122-147."""
    vbd = p.get('breakdown_voltage', 2.0)
    td = p.get('dielectric_thickness', 1e-9)
    v = np.clip(v, -vbd, vbd)
    k = p['uniaxial_anisotropy'] + (-p.get('vcma_coefficient', 100e-6) * np.abs(v) / td ** 2)
    return np.maximum(k, -0.5 * p['uniaxial_anisotropy'])


def effective_field(kind, p, m, happ, volt=None):
    m = np.asarray(m, float)
    ms = p.get('saturation_magnetization', 800e3)
    e = np.asarray(p.get('easy_axis', [0, 0, 1]), float)
    ku = p.get('uniaxial_anisotropy', 1e6)
    if kind == 'stt_mram':
        m = m / np.linalg.norm(m, axis=-1, keepdims=True)
        return happ + (2 * ku / (MU0 * ms)) * (m @ e)[..., None] * e
    if kind == 'vcma_mram':
        ku = vcma_keff(p, 0.0 if volt is None else volt)
    hk = (2 * ku / (MU0 * ms)) * (m @ e)
    return happ + hk[..., None] * e + (-ms * demag_factors(p.get('aspect_ratio', 1.0)) * m)


def resistance(kind, p, m):
    m = np.asarray(m, float)
    ref = np.asarray(p.get('reference_magnetization', [0, 0, 1]), float)
    ref = ref / np.linalg.norm(ref)
    rp, rap = p.get('resistance_parallel', 1e3), p.get('resistance_antiparallel', 2e3)
    if kind == 'stt_mram':
        m = m / np.linalg.norm(m, axis=-1, keepdims=True)
        r = rp * (1 + ((rap - rp) / rp) * (1 - m @ ref) / 2)
        return np.maximum(r, rp * 0.5)
    r = rp + (rap - rp) * (1 - m @ ref) / 2
    if kind == 'sot_mram':
        thickness = p.get('thickness', 1e-9)
        area = p.get('area', p.get('volume', 1e-24) / thickness)
        sheet = p.get('heavy_metal_resistivity', 2e-7) / p.get('heavy_metal_thickness', 5e-9)
        r = r + (sheet / (area * 1e-12)) * 0.1
    return np.maximum(r, 1.0)


def sot_factors(p):
    """devices/sot_mram.py:61-72."""
    t_hm = p.get('heavy_metal_thickness', 5e-9)
    js = p.get('spin_hall_angle', 0.1) * p.get('interface_transparency', 0.5) * (t_hm / (t_hm + p.get('thickness', 1e-9)))
    return p.get('damping_like_efficiency', 0.2) * js, p.get('field_like_efficiency', 0.1) * js


def sot_torque(p, current, m, direction=None):
    d = np.array([1.0, 0.0, 0.0]) if direction is None else np.asarray(direction, float)
    d = d / np.linalg.norm(d)
    sigma = np.cross(np.array([0.0, 0.0, 1.0]), d)
    f_dl, f_fl = sot_factors(p)
    cur = np.asarray(current, float)[..., None]
    return f_dl * cur * np.cross(sigma, m), f_fl * cur * sigma
