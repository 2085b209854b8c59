"""Generate golden vectors by running the LIVE reference (/root/reference) in this container.

TEST INFRASTRUCTURE ONLY. Usage (build container only; /root/reference does not exist on the GPU box):

    python oracle/gen_golden.py [stt] [array] [rk45] [devices]      # default: all

The reference is imported through oracle/shims (gymnasium / matplotlib are not installed) and sanitised as
SURVEY.md §8c prescribes: wall-clock timeout off, memo caches off, global NumPy RNG seeded immediately
before each thermal step. Outputs land in tests/golden/*.npz (small, committed).
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("STG_REFERENCE", "/root/reference")


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference checkout not found at {REF}; goldens can only be generated in the build container")
    sys.path.insert(0, os.path.join(HERE, "shims"))
    sys.path.insert(0, REF)
    import logging
    logging.disable(logging.CRITICAL)
    warnings.filterwarnings("ignore")
    import spin_torque_gym  # noqa: F401
    from spin_torque_gym.envs import SpinTorqueEnv, SpinTorqueArrayEnv
    return SpinTorqueEnv, SpinTorqueArrayEnv


def _sanitise(env):
    """no RK4->Euler wall-clock switch, no stale caches."""
    env.solver.timeout = 1e9
    env.optimizer.cache.ttl = -1
    env.cache_observations = False
    return env


def _stt_params(**over):
    from spin_torque_gym.devices import DeviceFactory
    p = DeviceFactory().get_default_parameters('stt_mram')
    p.update(over)
    return p


def _run_stt_episode(env, m0, target, actions, seeds=None):
    obs0, _ = env.reset(seed=0, options={'initial_state': m0, 'target_state': target})
    out = dict(obs=[obs0], reward=[], terminated=[], truncated=[], m=[env.current_magnetization.copy()],
               energy=[], total_energy=[], sim_ok=[])
    for k, a in enumerate(actions):
        if seeds is not None:
            np.random.seed(int(seeds[k]))
        obs, r, term, trunc, info = env.step(a.copy())
        out['obs'].append(obs)
        out['reward'].append(r)
        out['terminated'].append(term)
        out['truncated'].append(trunc)
        out['m'].append(env.current_magnetization.copy())
        out['energy'].append(info['energy_consumed'])
        out['total_energy'].append(env.total_energy)
        out['sim_ok'].append(info['simulation_success'])
    return {k: np.array(v) for k, v in out.items()}


def gen_stt():
    SpinTorqueEnv, _ = _import_reference()
    cases = {}

    # ---- C1: N=1, thermal off, 100 steps of random pulses, well-conditioned regime ----------
    rng = np.random.default_rng(0)
    jm = 1.1e-6
    m0 = rng.normal(0, 1, 3)
    m0 /= np.linalg.norm(m0)
    acts = np.stack([rng.uniform(-jm, jm, 100), rng.uniform(1e-12, 5e-9, 100)], axis=1).astype(np.float32)
    env = _sanitise(SpinTorqueEnv(device_type='stt_mram', device_params=_stt_params(), max_current=jm,
                                  include_thermal_fluctuations=False, seed=0))
    res = _run_stt_episode(env, m0, np.array([0.0, 0.0, 1.0]), acts)
    cases['c1_det'] = dict(res, m0=m0, target=np.array([0.0, 0.0, 1.0]), actions=acts, max_current=jm,
                           thermal=0, method='rk4')
    print("c1_det done", res['m'][-1], flush=True)

    # ---- same regime reached through a scaled volume and action-scale currents ---------------------------
    rng = np.random.default_rng(1)
    m0 = rng.normal(0, 1, 3)
    m0 /= np.linalg.norm(m0)
    vol = 1e-23 * 1e12
    acts = np.stack([rng.uniform(-2e6, 2e6, 24), rng.uniform(1e-12, 1.2e-9, 24)], axis=1).astype(np.float32)
    acts[3, 0] = 0.0          # zero-current step
    acts[5] = [3e6, 7e-9]     # clipped by max_current / max_duration
    acts[7, 1] = 0.0          # duration below the 1e-12 floor
    env = _sanitise(SpinTorqueEnv(device_type='stt_mram', device_params=_stt_params(volume=vol),
                                  include_thermal_fluctuations=False, max_steps=20, seed=0))
    res = _run_stt_episode(env, m0, np.array([0.0, 0.0, -1.0]), acts)
    cases['bigvol_det'] = dict(res, m0=m0, target=np.array([0.0, 0.0, -1.0]), actions=acts, max_current=2e6,
                               volume=vol, thermal=0, method='rk4', max_steps=20)
    print("bigvol_det done", flush=True)

    # ---- tilted easy axis / reference layer, Euler integrator ---------------------------------------------
    rng = np.random.default_rng(2)
    m0 = rng.normal(0, 1, 3)
    m0 /= np.linalg.norm(m0)
    e = np.array([1.0, 2.0, 2.0])
    refm = np.array([0.0, 1.0, 1.0])
    tgt = np.array([1.0, 2.0, 2.0])
    acts = np.stack([rng.uniform(-jm, jm, 16), rng.uniform(1e-12, 8e-10, 16)], axis=1).astype(np.float32)
    for method in ('rk4', 'euler'):
        env = _sanitise(SpinTorqueEnv(device_type='stt_mram',
                                      device_params=_stt_params(easy_axis=e, reference_magnetization=refm,
                                                                damping=0.05, polarization=0.55,
                                                                resistance_parallel=1500.0,
                                                                resistance_antiparallel=4000.0),
                                      max_current=jm, include_thermal_fluctuations=False, seed=0))
        env.solver.method = method
        res = _run_stt_episode(env, m0, tgt, acts)
        cases[f'tilted_{method}'] = dict(res, m0=m0, target=tgt, actions=acts, max_current=jm, thermal=0,
                                         method=method, easy_axis=e, reference_magnetization=refm, damping=0.05,
                                         polarization=0.55, resistance_parallel=1500.0,
                                         resistance_antiparallel=4000.0)
        print(f"tilted_{method} done", flush=True)

    # ---- thermal on, noise stream pinned by seeding the global NumPy RNG before every step -----------------
    rng = np.random.default_rng(3)
    m0 = rng.normal(0, 1, 3)
    m0 /= np.linalg.norm(m0)
    acts = np.stack([rng.uniform(-jm, jm, 12), rng.uniform(1e-12, 6e-10, 12)], axis=1).astype(np.float32)
    seeds = np.arange(100, 112)
    env = _sanitise(SpinTorqueEnv(device_type='stt_mram', device_params=_stt_params(), max_current=jm,
                                  temperature=300.0, include_thermal_fluctuations=True, seed=0))
    res = _run_stt_episode(env, m0, np.array([0.0, 0.0, 1.0]), acts, seeds)
    cases['thermal_injected'] = dict(res, m0=m0, target=np.array([0.0, 0.0, 1.0]), actions=acts, max_current=jm,
                                     thermal=1, method='rk4', seeds=seeds, temperature=300.0)
    print("thermal_injected done", flush=True)

    # ---- thermal at the unstable equilibrium (m_z = 0): the only place the noise is dynamically visible ----
    acts = np.tile(np.array([[0.0, 1e-10]], dtype=np.float32), (6, 1))
    seeds = np.arange(200, 206)
    outs = []
    for s in seeds:
        env = _sanitise(SpinTorqueEnv(device_type='stt_mram', device_params=_stt_params(), max_current=jm,
                                      temperature=300.0, include_thermal_fluctuations=True, seed=0))
        r = _run_stt_episode(env, np.array([1.0, 0.0, 0.0]), np.array([0.0, 0.0, 1.0]), acts[:1], [s])
        outs.append(r['m'][-1])
    cases['thermal_equator'] = dict(m_final=np.array(outs), seeds=seeds, actions=acts[:1], max_current=jm,
                                    temperature=300.0)
    print("thermal_equator done", flush=True)

    flat = {}
    for cname, c in cases.items():
        for k, v in c.items():
            flat[f"{cname}/{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(GOLD, "stt_env.npz"), **flat)
    print("wrote stt_env.npz")


def gen_multi():
    """64 independent well-conditioned single steps: fresh random m0 / target / action each (what an RL rollout with
    resets looks like), thermal off. Avoids the deep-pole regime a long reset-free episode collapses into."""
    SpinTorqueEnv, _ = _import_reference()
    rng = np.random.default_rng(7)
    jm = 1.1e-6
    n = 64
    m0 = rng.normal(0, 1, (n, 3))
    m0 /= np.linalg.norm(m0, axis=1, keepdims=True)
    tgt = np.where(rng.integers(2, size=(n, 1)) == 0, 1.0, -1.0) * np.array([[0.0, 0.0, 1.0]])
    acts = np.stack([rng.uniform(-jm, jm, n), rng.uniform(1e-12, 2e-9, n)], axis=1).astype(np.float32)
    env = _sanitise(SpinTorqueEnv(device_type='stt_mram', device_params=_stt_params(), max_current=jm,
                                  include_thermal_fluctuations=False, seed=0))
    out = dict(obs0=[], obs=[], reward=[], terminated=[], truncated=[], m=[], energy=[])
    for k in range(n):
        o0, _ = env.reset(seed=0, options={'initial_state': m0[k], 'target_state': tgt[k]})
        o, r, te, tr, info = env.step(acts[k].copy())
        out['obs0'].append(o0); out['obs'].append(o); out['reward'].append(r); out['terminated'].append(te)
        out['truncated'].append(tr); out['m'].append(env.current_magnetization.copy())
        out['energy'].append(info['energy_consumed'])
    np.savez_compressed(os.path.join(GOLD, "stt_multi.npz"), m0=m0, target=tgt, actions=acts, max_current=jm,
                        **{k: np.array(v) for k, v in out.items()})
    print("wrote stt_multi.npz")


def gen_devices():
    """Device-class methods of the reference on random inputs (devices/stt_mram.py, sot_mram.py, vcma_mram.py)."""
    _import_reference()
    from spin_torque_gym.devices import DeviceFactory
    from spin_torque_gym.physics import ThermalFluctuations
    fac = DeviceFactory()
    rng = np.random.default_rng(11)
    n = 48
    m = rng.normal(size=(n, 3))
    m[: n // 2] /= np.linalg.norm(m[: n // 2], axis=1, keepdims=True)      # half normalised, half not
    happ = rng.normal(size=(n, 3)) * 1e4
    volt = rng.uniform(-3.0, 3.0, n)
    cur = rng.uniform(-2e6, 2e6, n)
    out = dict(m=m, happ=happ, volt=volt, cur=cur)
    variants = {
        'stt': ('stt_mram', dict(easy_axis=np.array([0.0, 0.6, 0.8]), reference_magnetization=np.array([0.0, 1.0, 1.0]))),
        'sot': ('sot_mram', dict(aspect_ratio=2.5, spin_hall_angle=0.3, reference_magnetization=np.array([1.0, 0.0, 1.0]))),
        'sot_ar': ('sot_mram', dict(aspect_ratio=0.4)),
        'vcma': ('vcma_mram', dict(aspect_ratio=1.7, vcma_coefficient=80e-6, breakdown_voltage=1.5)),
    }
    for key, (dtype, over) in variants.items():
        p = fac.get_default_parameters(dtype)
        p.update(over)
        dev = fac.create_device(dtype, p)
        if dtype == 'vcma_mram':
            out[f'{key}/field'] = np.array([dev.compute_effective_field(m[i], happ[i], volt[i]) for i in range(n)])
            out[f'{key}/keff'] = np.array([dev._compute_effective_anisotropy(v) for v in volt])
            out[f'{key}/power'] = np.array([dev.compute_power_consumption(v, 1e-9) for v in volt])
            out[f'{key}/pswitch'] = np.array([dev.compute_switching_probability(v, 1e-9) for v in volt])
        else:
            out[f'{key}/field'] = np.array([dev.compute_effective_field(m[i], happ[i]) for i in range(n)])
        out[f'{key}/resistance'] = np.array([dev.compute_resistance(m[i]) for i in range(n)])
        if dtype == 'sot_mram':
            d = np.array([1.0, 2.0, 0.0])
            t = [dev.compute_spin_torque(cur[i], m[i], d) for i in range(n)]
            out[f'{key}/tau_dl'] = np.array([x[0] for x in t])
            out[f'{key}/tau_fl'] = np.array([x[1] for x in t])
            t = [dev.compute_spin_torque(cur[i], m[i]) for i in range(n)]
            out[f'{key}/tau_dl_default'] = np.array([x[0] for x in t])
            out[f'{key}/power'] = np.array([dev.compute_power_consumption(cur[i], 1e-9, m[i]) for i in range(n)])
        for k, v in over.items():
            out[f'{key}/param/{k}'] = np.asarray(v)
    th = ThermalFluctuations(temperature=350.0, correlation_time=2e-12, seed=1)
    out['thermal/strength'] = th.compute_noise_strength(0.01, 800e3, 1e-23)
    out['thermal/barrier'] = th.compute_thermal_barrier(1.2e6, 1e-23)
    out['thermal/pswitch'] = th.compute_switching_probability(1.2e6 * 1e-23 * 0.05)
    out['thermal/retention'] = th.compute_retention_time(1.2e6 * 1e-23 * 0.05)
    np.savez_compressed(os.path.join(GOLD, "devices.npz"), **out)
    print("wrote devices.npz")


def gen_rk45():
    """LLGSSolver.solve of the live reference (SciPy RK45, rtol 1e-6, atol 1e-9, max_step 1e-12)."""
    _import_reference()
    from spin_torque_gym.physics.llgs_solver import LLGSSolver
    solver = LLGSSolver(method='RK45', rtol=1e-6, atol=1e-9, max_step=1e-12)
    rng = np.random.default_rng(5)
    out = {}
    # LLGSSolver's torque prefactor is P*gamma/(2 Ms V) * J: with V = 1e-11 m^3 a current of ~10 A/m^2 gives a torque rate of
    # ~1e11 1/s, comparable to the precession rate; at action-scale currents SciPy's step size collapses
    base = _stt_params(volume=1e-23 * 1e12)
    cases = {
        'stt_on': dict(params=base, J=12.0, t_pulse=1e9, t_end=1.5e-10, happ=np.zeros(3)),
        'stt_pulse': dict(params=base, J=-15.0, t_pulse=6e-11, t_end=1.6e-10, happ=np.array([2e4, -1e4, 5e3])),
        'stt_nocurrent': dict(params=dict(base, demag_factors=np.array([0.1, 0.2, 0.7])), J=0.0, t_pulse=1e9, t_end=1e-10,
                              happ=np.array([0.0, 3e4, 0.0])),
        'stt_tilted': dict(params=dict(base, easy_axis=np.array([0.3, 0.0, 0.9]), exchange_constant=0.0, damping=0.05),
                           J=8.0, t_pulse=1e9, t_end=1.2e-10, happ=np.zeros(3)),
        'stt_short': dict(params=base, J=20.0, t_pulse=1e9, t_end=2.5e-12, happ=np.zeros(3)),
    }
    for name, c in cases.items():
        m0 = rng.normal(0, 1, 3)
        J, tp, happ = c['J'], c['t_pulse'], c['happ']
        res = solver.solve(m0, (0, c['t_end']), c['params'], lambda t: J if t <= tp else 0.0, lambda t: happ,
                           thermal_noise=False)
        for k in ('t', 'm', 'energy', 'torques'):
            out[f'{name}/{k}'] = np.asarray(res[k])
        out[f'{name}/success'] = res['success']
        out[f'{name}/m0'] = m0
        out[f'{name}/J'] = J; out[f'{name}/t_pulse'] = tp; out[f'{name}/t_end'] = c['t_end']; out[f'{name}/happ'] = happ
        for k, v in c['params'].items():
            if k in ('volume', 'easy_axis', 'demag_factors', 'exchange_constant', 'damping'):
                out[f'{name}/param/{k}'] = np.asarray(v)
        print(name, len(res['t']), res['success'], flush=True)
    # thermal: the global NumPy stream is consumed one normal(0,1,3) per RHS call
    m0 = rng.normal(0, 1, 3)
    np.random.seed(77)
    res = solver.solve(m0, (0, 4e-11), base, lambda t: 10.0, lambda t: np.zeros(3), thermal_noise=True, temperature=300.0)
    for k in ('t', 'm', 'energy', 'torques'):
        out[f'stt_thermal/{k}'] = np.asarray(res[k])
    out['stt_thermal/m0'] = m0; out['stt_thermal/seed'] = 77; out['stt_thermal/J'] = 10.0
    out['stt_thermal/t_end'] = 4e-11; out['stt_thermal/param/volume'] = np.asarray(base['volume'])
    print('stt_thermal', len(res['t']), flush=True)
    np.savez_compressed(os.path.join(GOLD, "rk45.npz"), **out)
    print("wrote rk45.npz")


def gen_array():
    """SpinTorqueArrayEnv.step of the live reference: 8x8 dipolar crossbar, all four action modes, a 3x5 stray-field array."""
    _, SpinTorqueArrayEnv = _import_reference()
    out = {}
    rng = np.random.default_rng(21)

    def run(name, size, mode, n_steps, **kw):
        env = SpinTorqueArrayEnv(array_size=size, action_mode=mode, seed=0, **kw)
        nd = size[0] * size[1]
        p0 = rng.normal(size=size + (3,))
        p0 /= np.linalg.norm(p0, axis=2, keepdims=True)
        obs0, _ = env.reset(seed=0, options={'initial_pattern': p0})
        hi = {'individual': nd - 1, 'row': size[0] - 1, 'column': size[1] - 1}.get(mode, 0)
        if mode == 'global':
            acts = np.stack([rng.uniform(-2e6, 2e6, n_steps), rng.uniform(0, 5e-9, n_steps)], 1).astype(np.float32)
        else:
            acts = np.stack([rng.uniform(-0.5, hi + 0.5, n_steps), rng.uniform(-2.2e6, 2.2e6, n_steps),
                             rng.uniform(-1e-10, 5.5e-9, n_steps)], 1).astype(np.float32)
            acts[1, 1] = 0.0
        rec = dict(obs=[obs0], reward=[], terminated=[], truncated=[], pattern=[env.current_pattern.copy()], energy=[],
                   similarity=[])
        for a in acts:
            o, r, te, tr, info = env.step(a.copy())
            rec['obs'].append(o); rec['reward'].append(r); rec['terminated'].append(te); rec['truncated'].append(tr)
            rec['pattern'].append(env.current_pattern.copy()); rec['energy'].append(info['energy_consumed'])
            rec['similarity'].append(info['pattern_similarity'])
        for k, v in rec.items():
            out[f'{name}/{k}'] = np.array(v)
        out[f'{name}/actions'] = acts
        out[f'{name}/p0'] = p0
        out[f'{name}/coupling'] = env.coupling_matrix
        print(name, 'done', flush=True)

    run('ind8', (8, 8), 'individual', 40)
    run('row8', (8, 8), 'row', 20)
    run('col8', (8, 8), 'column', 20)
    run('glob8', (8, 8), 'global', 6)
    run('stray35', (3, 5), 'row', 12, coupling_type='stray_field', coupling_strength=0.25, max_steps=10)
    np.savez_compressed(os.path.join(GOLD, "array_env.npz"), **out)
    print("wrote array_env.npz")


def gen_next():
    """Adjacent components of SURVEY §8(f): VectorizedSolver.solve_batch, EnergyLandscape, LLGSSolver.find_stable_states."""
    _import_reference()
    from spin_torque_gym.utils.vectorized_operations import VectorizedSolver
    from spin_torque_gym.physics.energy_landscape import EnergyLandscape
    rng = np.random.default_rng(31)
    n = 12
    m0 = rng.normal(size=(n, 3))
    m0 /= np.linalg.norm(m0, axis=1, keepdims=True)
    plist = []
    for i in range(n):
        p = _stt_params(damping=float(rng.uniform(0.005, 0.05)), uniaxial_anisotropy=float(rng.uniform(0.8e6, 1.6e6)),
                        saturation_magnetization=float(rng.uniform(6e5, 9e5)))
        if i % 3 == 0:
            p['easy_axis'] = np.array([0.2 * i / n, -0.1, 1.0])
        plist.append(p)
    out = dict(m0=m0)
    res = VectorizedSolver().solve_batch(m0, (0, 1.5e-10), plist, dt=1e-12)
    out['vec/m'] = np.array([r['m'] for r in res])
    out['vec/t'] = res[0]['t']
    res2 = VectorizedSolver().solve_batch(m0, (0, 3e-12), plist[:1] * n, dt=1e-12)       # T < 10 dt -> 10 steps
    out['vec_short/m'] = np.array([r['m'] for r in res2])
    for k in ('damping', 'uniaxial_anisotropy', 'saturation_magnetization'):
        out[f'vec/{k}'] = np.array([p[k] for p in plist])
    out['vec/easy_axis'] = np.array([np.asarray(p['easy_axis'], float) for p in plist])
    lp = _stt_params(demag_factors=np.array([0.1, 0.3, 0.6]), easy_axis=np.array([0.0, 0.6, 0.8]))
    land = EnergyLandscape(lp)
    happ = rng.normal(size=(n, 3)) * 1e4
    mm = rng.normal(size=(n, 3))
    out['land/m'] = mm; out['land/happ'] = happ
    out['land/energy'] = np.array([land.compute_energy(mm[i], happ[i]) for i in range(n)])
    out['land/grad'] = np.array([land.compute_energy_gradient(mm[i], happ[i]) for i in range(n)])
    out['land/energy0'] = np.array([land.compute_energy(mm[i]) for i in range(n)])
    # phase diagram over a (current, field) grid that straddles the h_k - |beta I| line; stability factor
    pp = _stt_params(volume=1e-11)
    land2 = EnergyLandscape(pp)
    h_k = 2 * pp['uniaxial_anisotropy'] / (4 * np.pi * 1e-7 * pp['saturation_magnetization'])
    beta = pp.get('polarization', 0.7) * 2.21e5 / (2 * pp['saturation_magnetization'] * pp['volume'])
    i_max = 1.5 * h_k / beta
    pd = land2.generate_phase_diagram((-i_max, i_max), (-1.2 * h_k, 1.2 * h_k), resolution=37)
    out['phase/volume'] = 1e-11; out['phase/i_max'] = i_max; out['phase/h_max'] = 1.2 * h_k
    for k, v in pd.items():
        out[f'phase/{k}'] = np.asarray(v)
    out['phase/stability_300'] = land2.compute_thermal_stability_factor(300.0)
    out['phase/stability_77'] = land2.compute_thermal_stability_factor(77.0)
    # energy barrier along the straight path between two tilted states of the anisotropic-demag landscape `land`
    s0, s1 = np.array([0.1, 0.55, 0.83]), np.array([0.2, -0.6, -0.7])
    hb = np.array([2e3, -1e3, 5e2])
    bh, ep = land.compute_energy_barrier(s0, s1, hb, n_intermediate=41)
    out['barrier/s0'] = s0; out['barrier/s1'] = s1; out['barrier/happ'] = hb
    out['barrier/height'] = bh; out['barrier/path'] = ep
    out['barrier/height_nofield'] = land.compute_energy_barrier(s0, s1)[0]
    # VectorizedMagneticsOperations on random rows (un-normalised, one zero row for the 1e-12 floor)
    from spin_torque_gym.utils.vectorized_operations import VectorizedMagneticsOperations as VMO
    nv = 64
    a = rng.normal(size=(nv, 3)) * rng.uniform(1e-3, 1e3, (nv, 1))
    b = rng.normal(size=(nv, 3))
    a[5] = 0.0
    k_u = rng.uniform(0.5e6, 2e6, nv); vol = rng.uniform(1e-24, 1e-22, nv)
    easy = rng.normal(size=(nv, 3)); easy /= np.linalg.norm(easy, axis=1, keepdims=True)
    r_p = rng.uniform(500.0, 2000.0, nv); r_ap = r_p * rng.uniform(0.2, 3.0, nv)     # some R_AP < R_P: the R_P/2 floor binds
    out.update({'vmo/a': a, 'vmo/b': b, 'vmo/k_u': k_u, 'vmo/volume': vol, 'vmo/easy': easy, 'vmo/r_p': r_p, 'vmo/r_ap': r_ap})
    out['vmo/cross'] = VMO.batch_cross_product(a, b)
    out['vmo/dot'] = VMO.batch_dot_product(a, b)
    out['vmo/normalize'] = VMO.batch_normalize(a)
    out['vmo/energy'] = VMO.batch_energy_computation(b, {'uniaxial_anisotropy': k_u, 'volume': vol, 'easy_axis': easy})
    out['vmo/energy_default'] = VMO.batch_energy_computation(b, {})
    out['vmo/energy_one_axis'] = VMO.batch_energy_computation(b, {'uniaxial_anisotropy': k_u, 'volume': vol,
                                                                  'easy_axis': np.array([0.0, 0.6, 0.8])})
    bn = b / np.linalg.norm(b, axis=1, keepdims=True)
    ref = easy.copy()
    ref[:6] = -bn[:6]; r_ap[:6] = 0.1 * r_p[:6]                  # antiparallel rows with R_AP << R_P sit on the floor
    out['vmo/r_ap'] = r_ap
    out['vmo/bn'] = bn; out['vmo/ref'] = ref
    out['vmo/resistance'] = VMO.batch_resistance_computation(bn, ref, r_p, r_ap)
    # ThermalFluctuations analytics (physics/thermal_model.py:139-336)
    from spin_torque_gym.physics import ThermalFluctuations
    tp = dict(volume=1.5e-25, uniaxial_anisotropy=1.1e6, damping=0.02, saturation_magnetization=7.5e5)   # Delta(300 K) ~ 40
    th = ThermalFluctuations(temperature=320.0, seed=9)
    barrier = 1.1e6 * 1.5e-25 * 0.5
    out['th/barrier_J'] = barrier
    out['th/switch_times'] = np.array([th.sample_switching_time(barrier) for _ in range(5)])
    sweep = th.generate_temperature_sweep((50.0, 450.0), tp, n_points=23)
    for k, v in sweep.items():
        out[f'th/sweep/{k}'] = np.asarray(v)
    out['th/temperature_after_sweep'] = th.temperature
    st = th.analyze_thermal_stability(tp, time_scale=3.0)
    for k, v in st.items():
        out[f'th/stability/{k}'] = np.asarray(v)
    # SimpleLLGSSolver.solve with callables that are NOT rectangular pulses / constant fields (sampled onto a
    # per-substep grid by the replacement). Well-conditioned STT regime (J ~ 1e-6 A/m^2 at the default volume).
    from spin_torque_gym.physics.simple_solver import SimpleLLGSSolver as RefSimple

    def ref_solver(method):
        sv = RefSimple(method=method)
        sv.timeout = 1e9                 # no wall-clock abort
        sv.optimizer.cache.ttl = -1      # the memo key omits current_func / field_func
        return sv

    gp = _stt_params()
    w = 2 * np.pi / 1.7e-10
    cur_sin = lambda t: 1.0e-6 * np.sin(w * t) + 2.0e-7                                  # noqa: E731
    cur_steps = lambda t: 9e-7 if t < 0.8e-10 else (-6e-7 if t < 2.1e-10 else 3e-7)     # noqa: E731  three levels
    cur_rect = lambda t: 8e-7 if t <= 1.3e-10 else 0.0                                   # noqa: E731
    fld_rot = lambda t: 2.0e5 * np.array([np.cos(w * t), np.sin(w * t), 0.3])           # noqa: E731
    fld_const = lambda t: np.array([1.0e5, -5.0e4, 2.0e4])                               # noqa: E731
    gm0 = np.array([0.45, -0.3, 0.84]); gm0 /= np.linalg.norm(gm0)
    gcases = {
        'sin_rk4': ('rk4', (0.0, 3.0e-10), cur_sin, None),
        'sin_constfield_rk4': ('rk4', (0.0, 3.0e-10), cur_sin, fld_const),
        'rect_rotfield_rk4': ('rk4', (0.0, 3.0e-10), cur_rect, fld_rot),
        'steps_rotfield_rk4': ('rk4', (0.0, 3.0e-10), cur_steps, fld_rot),
        'steps_rotfield_euler': ('euler', (0.0, 3.0e-10), cur_steps, fld_rot),
        'sin_rotfield_offset_rk4': ('rk4', (1.0e-10, 3.5e-10), cur_sin, fld_rot),        # t_start != 0: absolute times
        'sin_short_rk4': ('rk4', (0.0, 4.0e-12), cur_sin, fld_rot),                      # T < 10 ps: dt = T/100
    }
    out['grid/m0'] = gm0
    for name, (method, span, cf, ff) in gcases.items():
        res = ref_solver(method).solve(gm0.copy(), span, gp, cf, ff)
        assert res['success'], (name, res['message'])
        out[f'grid/{name}/m'] = res['m']; out[f'grid/{name}/t'] = res['t']
        print(name, res['n_steps'], res['m'][-1], flush=True)
    np.savez_compressed(os.path.join(GOLD, "next.npz"), **out)
    print("wrote next.npz")


def gen_round2():
    """Round-2 additions, all from the LIVE reference: (i) ThermalFluctuations analytics on a temperature x device grid
    (physics/thermal_model.py:46-73, 139-258), (ii) LLGSSolver.solve with t_span[0] != 0 and with piecewise-constant
    current_func / field_func (physics/llgs_solver.py:51-180), (iii) LLGSSolver.find_stable_states (:264-305) with its hard-coded
    10 ns relaxation shortened - the method is re-run here line for line with (0, relax_time) so that it finishes in seconds."""
    _import_reference()
    from spin_torque_gym.physics import ThermalFluctuations
    from spin_torque_gym.physics.llgs_solver import LLGSSolver
    rng = np.random.default_rng(77)
    out = {}
    # ---- (i) thermal grid
    temps = np.array([0.0, 4.2, 77.0, 150.0, 233.15, 300.0, 358.15, 398.15, 450.0, 600.0])
    n_dev = 9
    ku = rng.uniform(0.4e6, 1.6e6, n_dev); vol = rng.uniform(5e-26, 4e-25, n_dev)
    al = rng.uniform(0.005, 0.05, n_dev); ms = rng.uniform(5e5, 1.2e6, n_dev)
    f0, tm, fr = 2.5e9, 3.0e-3, 1e-6
    grid = np.zeros((4, len(temps), n_dev))
    for i, T in enumerate(temps):
        th = ThermalFluctuations(temperature=float(T), seed=0)
        for j in range(n_dev):
            e = ku[j] * vol[j]
            grid[0, i, j] = th.compute_thermal_barrier(ku[j], vol[j])
            grid[1, i, j] = th.compute_switching_probability(e, attempt_frequency=f0, measurement_time=tm)
            grid[2, i, j] = th.compute_retention_time(e, failure_rate=fr, attempt_frequency=f0)
            grid[3, i, j] = th.compute_noise_strength(al[j], ms[j], vol[j])
    out.update({'thg/temps': temps, 'thg/ku': ku, 'thg/vol': vol, 'thg/damping': al, 'thg/ms': ms, 'thg/f0': f0,
                'thg/tm': tm, 'thg/fr': fr, 'thg/grid': grid})
    # ---- (ii) LLGSSolver.solve: shifted time span; piecewise-constant controls
    base = _stt_params(volume=1e-11)                 # V = 1e-11 m^3: currents of ~10 A/m^2 give torque rates ~1e11 1/s (gen_rk45)
    base['demag_factors'] = np.array([0.05, 0.15, 0.8])
    solver = LLGSSolver()
    m0 = np.array([0.3, -0.2, 0.93])
    cases = {
        'shift': dict(span=(2.0e-11, 5.5e-11), cur=([3.1e-11], [14.0, 0.0]), fld=([], [[1.5e4, -0.5e4, 2e3]])),
        'pwc': dict(span=(0.0, 5.0e-11), cur=([1.2e-11, 2.6e-11, 4.1e-11], [15.0, -9.0, 0.0, 6.0]),
                    fld=([1.9e-11], [[2e4, 0.0, 0.0], [0.0, -1e4, 5e3]])),
        'shift_pwc': dict(span=(-1.0e-11, 3.0e-11), cur=([0.0, 1.5e-11], [0.0, 18.0, -18.0]),
                          fld=([0.7e-11], [[0.0, 0.0, 1e4], [1e4, 1e4, 0.0]])),
    }
    for name, c in cases.items():
        cb, cv = np.array(c['cur'][0]), np.array(c['cur'][1])
        fb, fv = np.array(c['fld'][0]), np.array(c['fld'][1], dtype=float)

        def cur(t, cb=cb, cv=cv):
            return float(cv[int(np.searchsorted(cb, t, side='left'))])

        def fld(t, fb=fb, fv=fv):
            return fv[int(np.searchsorted(fb, t, side='left'))].copy()

        res = solver.solve(m0, c['span'], base, cur, fld, thermal_noise=False)
        assert res['success']
        for k in ('t', 'm', 'energy', 'torques'):
            out[f'solve/{name}/{k}'] = np.asarray(res[k])
        out[f'solve/{name}/span'] = np.array(c['span'])
        out[f'solve/{name}/cur_breaks'] = cb; out[f'solve/{name}/cur_values'] = cv
        out[f'solve/{name}/fld_breaks'] = fb; out[f'solve/{name}/fld_values'] = fv
        print(name, len(res['t']), flush=True)
    out['solve/m0'] = m0
    out['solve/demag'] = base['demag_factors']
    # ---- (iii) find_stable_states with a short relaxation (the reference hard-codes (0, 10e-9): ~1e5 RHS calls per trial)
    fp = _stt_params(volume=1e-11, damping=0.3)
    relax, n_trials, threshold, seed = 3.0e-10, 8, 0.2, 5
    np.random.seed(seed)
    stable, finals = [], []
    for _ in range(n_trials):                                    # body of physics/llgs_solver.py:273-303
        m_init = np.random.normal(0, 1, 3)
        m_init = m_init / np.linalg.norm(m_init)
        result = solver.solve(m_init, (0, relax), fp, lambda t: 0.0, lambda t: np.zeros(3), thermal_noise=False)
        if result['success']:
            m_final = result['m'][-1]
            finals.append(m_final)
            if all(np.linalg.norm(m_final - s_) >= threshold for s_ in stable):
                stable.append(m_final)
    out.update({'fss/relax': relax, 'fss/n_trials': n_trials, 'fss/threshold': threshold, 'fss/seed': seed,
                'fss/damping': 0.3, 'fss/volume': 1e-11, 'solve/volume': 1e-11, 'fss/stable': np.array(stable), 'fss/finals': np.array(finals)})
    print('find_stable_states', len(stable), 'of', n_trials, flush=True)
    np.savez_compressed(os.path.join(GOLD, "round2.npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1:] or ["stt", "multi", "array", "rk45", "devices"]
    os.makedirs(GOLD, exist_ok=True)
    for w in what:
        fn = globals().get(f"gen_{w}")
        if fn is None:
            raise SystemExit(f"unknown golden set {w}")
        fn()
