"""CPU oracle for the adaptive-RK45 solver path (LLGSSolver.solve). TEST INFRASTRUCTURE ONLY.

The arithmetic of this path lives in a third-party dependency of the reference: scipy.integrate.solve_ivp(method='RK45')
(pyproject.toml pins scipy>=1.7.0; scipy 1.18.1 is installed in this image, on the GPU box too). The oracle therefore calls
solve_ivp exactly like the reference does (physics/llgs_solver.py:130-139: rtol, atol, max_step, dense_output=True) on a
NumPy restatement of llgs_rhs (physics/llgs_solver.py:92-126, 182-237).

Parity status: PINNED for the STT form against tests/golden/rk45.npz (LLGSSolver.solve of the live reference).
The SOT / VCMA forms compose the reference's own device methods (devices/sot_mram.py:163-194, vcma_mram.py:122-147) into the
same RHS; the reference itself never wires them into a solver, so that composition is "parity unpinned upstream" 
and is anchored on the device-method goldens (tests/golden/devices.npz) plus SciPy.
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import solve_ivp

from . import devices_oracle as DO

GAMMA = 2.21e5
MU0 = 4 * np.pi * 1e-7
KB = 1.380649e-23          # physics/llgs_solver.py:49


def make_rhs(params, current_func, field_func=None, thermal_noise=False, temperature=300.0, kind='stt_mram',
             voltage=0.0, current_direction=None, noise=None):
    """llgs_rhs closure. kind selects which torque / anisotropy terms are composed in. `noise`: optional [k,3] array replacing
    np.random.normal(0,1,3) call by call."""
    alpha = params.get('damping', 0.01)
    ms = params.get('saturation_magnetization', 800e3)
    volume = params.get('volume', 1e-24)
    pol = params.get('polarization', 0.7)
    strength = np.sqrt(2 * alpha * KB * temperature / (GAMMA * MU0 * ms * volume)) if thermal_noise else 0.0
    easy = np.asarray(params.get('easy_axis', np.array([0, 0, 1])), float)
    ku = params.get('uniaxial_anisotropy', 1e6)
    if kind == 'vcma_mram':
        ku = float(DO.vcma_keff(params, voltage))
    if kind == 'stt_mram':
        demag = np.asarray(params.get('demag_factors', np.array([0, 0, 1])), float)
        a_ex = params.get('exchange_constant', 20e-12)
        exch = (2 * a_ex / (MU0 * ms)) * 0.1 if a_ex > 0 else 0.0
    else:
        demag = DO.demag_factors(params.get('aspect_ratio', 1.0))
        exch = 0.0
    state = {'calls': 0}

    def field(m, t):
        h_app = np.asarray(field_func(t), float) if field_func else np.zeros(3)
        h = h_app.copy()
        h += (2 * ku / (MU0 * ms)) * np.dot(m, easy) * easy
        h += -ms * demag * m
        if exch:
            h += exch * m
        return h

    def torques(m, current):
        if abs(current) < 1e-12:
            return np.zeros(3), np.zeros(3)
        if kind == 'stt_mram':                                     # physics/llgs_solver.py:213-237
            p_hat = np.array([0, 0, 1])
            beta = pol * GAMMA / (2 * ms * volume)
            mxp = np.cross(m, p_hat)
            return beta * current * np.cross(m, mxp), 0.1 * beta * current * mxp
        if kind == 'sot_mram':
            return DO.sot_torque(params, current, m, current_direction)
        return np.zeros(3), np.zeros(3)

    def rhs(t, y):
        m = y[:3]
        nrm = np.linalg.norm(m)
        m = m / nrm if nrm > 1e-12 else np.array([0, 0, 1])
        h = field(m, t)
        if thermal_noise:
            k = state['calls']
            xi = np.random.normal(0, 1, 3) if noise is None else noise[min(k, len(noise) - 1)]
            h = h + strength * xi
        state['calls'] += 1
        tau_dl, tau_fl = torques(m, current_func(t))
        dm = -GAMMA * np.cross(m, h)
        dm = dm + alpha * np.cross(m, dm)
        return dm + tau_dl + tau_fl

    rhs.state = state
    rhs.field = field
    rhs.torques = torques
    rhs.ku = ku
    rhs.demag = demag
    return rhs


def solve(m_initial, t_end, params, current_func, field_func=None, thermal_noise=False, temperature=300.0,
          rtol=1e-6, atol=1e-9, max_step=1e-12, **kw):
    """LLGSSolver.solve (physics/llgs_solver.py:51-180). Returns the reference's dict + accepted/rejected bookkeeping."""
    m0 = np.asarray(m_initial, float)
    m0 = m0 / np.linalg.norm(m0)
    rhs = make_rhs(params, current_func, field_func, thermal_noise, temperature, **kw)
    sol = solve_ivp(rhs, (0, t_end), m0, method='RK45', rtol=rtol, atol=atol, max_step=max_step, dense_output=True)
    y_end = sol.y[:, -1].copy()
    m = sol.y.T.copy()
    m /= np.linalg.norm(m, axis=1, keepdims=True)
    ms, vol = params.get('saturation_magnetization', 800e3), params.get('volume', 1e-24)
    easy = np.asarray(params.get('easy_axis', np.array([0, 0, 1])), float)
    energy, torque = [], []
    for ti, mi in zip(sol.t, m):
        h_app = np.asarray(field_func(ti), float) if field_func else np.zeros(3)
        e = -MU0 * ms * vol * np.dot(mi, h_app) - rhs.ku * vol * np.dot(mi, easy) ** 2 \
            + 0.5 * MU0 * ms ** 2 * vol * np.sum(rhs.demag * mi ** 2)
        dl, fl = rhs.torques(mi, current_func(ti))
        energy.append(e)
        torque.append(np.linalg.norm(dl) + np.linalg.norm(fl))
    n_acc = len(sol.t) - 1
    # every attempted step costs 6 RHS evaluations (FSAL), plus 2 for f0 and select_initial_step
    n_rej = (sol.nfev - 2) // 6 - n_acc
    return {'t': sol.t, 'm': m, 'energy': np.array(energy), 'torques': np.array(torque), 'success': sol.success,
            'y_end': y_end, 'n_accepted': n_acc, 'n_rejected': n_rej, 'nfev': sol.nfev}
