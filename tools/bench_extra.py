#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs[2], [3]) and the ragged-duration variant of configs[1], device-timed.

    python tools/bench_extra.py            # prints one JSON line
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spin_torque_rl_gym_b200 as stg  # noqa: E402
from spin_torque_rl_gym_b200 import params as P  # noqa: E402
from spin_torque_rl_gym_b200.physics import LLGSSolver  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def collect(dev=None):
    dev = dev or torch.device("cuda", 0)
    out = {}
    rng = np.random.default_rng(0)
    # ---- configs[2]: SOT + VCMA mix, adaptive RK45 (atol 1e-9, rtol 1e-6), 262,144 envs ---------------------------------
    n = 262144
    sot = P.default_device_parameters("sot_mram"); sot.update(aspect_ratio=2.0, spin_hall_angle=0.3)
    vcma = P.default_device_parameters("vcma_mram"); vcma.update(aspect_ratio=1.5)
    m0 = torch.from_numpy(rng.normal(size=(n, 3))).to(dev)
    pidx = torch.arange(n, device=dev, dtype=torch.int32) % 2
    cur = torch.where(pidx == 0, torch.from_numpy(rng.uniform(-3e11, 3e11, n)).to(dev), torch.zeros(n, dtype=torch.float64, device=dev))
    volt = torch.where(pidx == 1, torch.from_numpy(rng.uniform(-2.5, 2.5, n)).to(dev), torch.zeros(n, dtype=torch.float64, device=dev))
    t_end = torch.full((n,), 1e-10, dtype=torch.float64, device=dev)
    solver = LLGSSolver(device=dev)
    res = {}

    def rk():
        res["r"] = solver.solve_batch(m0, t_end, [sot, vcma], current=cur, voltage=volt, param_index=pidx,
                                      device_type=["sot_mram", "vcma_mram"])
    ms = timed(rk, reps=5, warm=3)      # three warm-ups: the first calls after large frees pay allocator misses
    attempts = float((res["r"]["n_accepted"] + res["r"]["n_rejected"]).sum())
    out["rk45_mix_262144"] = {"ms_per_solve": ms, "attempted_steps_per_s": attempts / (ms * 1e-3),
                              "rhs_evals_per_s": float(res["r"]["n_rhs"].sum()) / (ms * 1e-3),
                              "accepted_per_env": attempts / n, "fp64_tflops_algorithmic": attempts * 766 / (ms * 1e-3) / 1e12}
    # end to end: host NumPy inputs (pageable -> device inside the call), results written by the kernel into pinned host memory
    h_m0, h_pidx, h_cur, h_volt = m0.cpu().numpy(), pidx.cpu().numpy(), cur.cpu().numpy(), volt.cpu().numpy()

    def rk_e2e():
        r = solver.solve_batch(h_m0, 1e-10, [sot, vcma], current=h_cur, voltage=h_volt, param_index=h_pidx,
                               device_type=["sot_mram", "vcma_mram"], host_outputs=True)
        res["e"] = float(r["y"][0, 0]) + int(r["n_accepted"][0])
    ms_e = timed(rk_e2e, reps=5, warm=2)
    out["rk45_mix_262144"]["e2e"] = {"ms_per_solve": ms_e, "attempted_steps_per_s": attempts / (ms_e * 1e-3),
                                     "h2d_bytes": n * (24 + 8 + 8 + 4), "d2h_bytes": n * (24 + 16 + 8)}
    # ---- configs[3]: 16,384 8x8 dipolar crossbars --------------------------------------------------------------------------
    A = 16384
    for mode in ("individual", "row", "global"):
        env = stg.SpinTorqueArrayVectorEnv(num_envs=A, array_size=(8, 8), action_mode=mode, device=dev, rng_seed=1)
        env.reset(seed=1)
        k = 2 if mode == "global" else 3
        act = torch.zeros(A, k, dtype=torch.float32, device=dev)
        if mode == "global":
            act[:, 0] = torch.from_numpy(rng.uniform(-2e6, 2e6, A)).to(dev); act[:, 1] = 2e-9
        else:
            hi = 63 if mode == "individual" else 7
            act[:, 0] = torch.from_numpy(rng.uniform(0, hi, A)).to(dev)
            act[:, 1] = torch.from_numpy(rng.uniform(-2e6, 2e6, A)).to(dev); act[:, 2] = 2e-9
        ms = timed(lambda: env.step(act), reps=10, warm=3)
        out[f"array_8x8_{mode}_16384"] = {"ms_per_step": ms, "array_steps_per_s": A / (ms * 1e-3),
                                          "hbm_gbs_algorithmic": A * 64 * 96 / (ms * 1e-3) / 1e9}
        if mode == "individual":
            # end to end: pinned host actions in, observations / rewards / flags written by the kernel into pinned host memory
            env_h = stg.SpinTorqueArrayVectorEnv(num_envs=A, array_size=(8, 8), action_mode=mode, device=dev, rng_seed=1,
                                                 host_outputs=True)
            env_h.reset(seed=1)
            act_h = act.cpu().pin_memory()

            def arr_e2e():
                o, r, te, tr, _ = env_h.step(act_h)
                return float(o[0, 0, 0, 0]) + float(r[0])
            ms_e = timed(arr_e2e, reps=10, warm=3)
            out[f"array_8x8_{mode}_16384"]["e2e"] = {"ms_per_step": ms_e, "array_steps_per_s": A / (ms_e * 1e-3),
                                                     "h2d_bytes_per_step": A * k * 4, "d2h_bytes_per_step": A * (64 * 24 + 10)}
    # ---- configs[1] with ragged durations T ~ U(1e-12, 5e-9): unsorted vs sorted launch -------------------------------------
    N = 1 << 20
    act = torch.zeros(N, 2, dtype=torch.float32, device=dev)
    act[:, 0] = torch.from_numpy(rng.uniform(-1.1e-6, 1.1e-6, N)).to(dev)
    act[:, 1] = torch.from_numpy(rng.uniform(1e-12, 5e-9, N)).to(dev)
    for sort in (False, True):
        env = stg.SpinTorqueVectorEnv(num_envs=N, device=dev, max_current=1.1e-6, include_thermal_fluctuations=True,
                                      rng_seed=2, sort_by_substeps=sort)
        env.reset(seed=2)
        ms = timed(lambda: env.step(act), reps=5, warm=2)
        sub = float(env._n_sub.sum())
        out[f"ragged_T_uniform_{'sorted' if sort else 'unsorted'}"] = {"ms_per_step": ms, "substeps_per_s": sub / (ms * 1e-3)}
    # ---- K4 elementwise FP64 ops and the standalone K5 reduction: HBM-bound, algorithmic bytes / device time ------------------
    from spin_torque_rl_gym_b200 import _lib
    from spin_torque_rl_gym_b200.devices import create_device
    from spin_torque_rl_gym_b200.parallel import reduce_step_stats
    from spin_torque_rl_gym_b200.physics import EnergyLandscape, VectorizedMagneticsOperations as VMO
    R = 1 << 24                                              # 16.8M rows: 403 MB per [R,3] f64 operand, far beyond the 126 MB L2
    a3 = torch.randn(R, 3, dtype=torch.float64, device=dev)
    b3 = torch.randn(R, 3, dtype=torch.float64, device=dev)
    k4 = {}
    for name, fn, nbytes in (
            ("vec3_cross", lambda: VMO.batch_cross_product(a3, b3), 72),          # 2 x 24 read + 24 written per row
            ("vec3_dot", lambda: VMO.batch_dot_product(a3, b3), 56),
            ("vec3_normalize", lambda: VMO.batch_normalize(a3), 48),
            ("stt_effective_field", None, 48),
            ("stt_resistance", None, 32),
            ("energy_and_gradient", None, 56)):
        if name == "stt_effective_field":
            dv = create_device("stt_mram", P.default_device_parameters("stt_mram"), device=dev)
            fn = lambda: dv.compute_effective_field(a3, np.zeros(3))                # noqa: E731
        elif name == "stt_resistance":
            fn = lambda: dv.compute_resistance(a3)                                   # noqa: E731
        elif name == "energy_and_gradient":
            land = EnergyLandscape(P.default_device_parameters("stt_mram"), device=dev)
            fn = lambda: (land.compute_energy(a3), land.compute_energy_gradient(a3))  # noqa: E731  (24+8) + (24+24) B per row
            nbytes = 80
        ms = timed(fn, reps=5, warm=2)
        k4[name] = {"ms": ms, "hbm_gbs_algorithmic": R * nbytes / (ms * 1e-3) / 1e9}
    out["k4_elementwise_16M_rows"] = k4
    T = 1 << 26                                              # 67M stored step results: 8+8+1+1+4+4+4 = 30 B each
    bufs = dict(reward=torch.randn(T, dtype=torch.float64, device=dev), step_energy=torch.rand(T, dtype=torch.float64, device=dev),
                terminated=(torch.rand(T, device=dev) < 0.1), truncated=(torch.rand(T, device=dev) < 0.1),
                n_sub=torch.randint(10, 5000, (T,), dtype=torch.int32, device=dev),
                status=torch.zeros(T, dtype=torch.int32, device=dev),
                step_count=torch.randint(1, 100, (T,), dtype=torch.int32, device=dev))
    bufs["terminated"], bufs["truncated"] = bufs["terminated"].to(torch.uint8), bufs["truncated"].to(torch.uint8)
    vec = torch.zeros(_lib.NSTATS, dtype=torch.float64, device=dev)
    ms = timed(lambda: reduce_step_stats(vec, **bufs), reps=5, warm=2)
    out["k5_stats_reduce_67M"] = {"ms": ms, "hbm_gbs_algorithmic": T * 30 / (ms * 1e-3) / 1e9}
    # ---- launch-bound regime: 1,024 envs, 100-substep pulses; eager step() vs CUDA-graph replay (capture_step) ------------------
    small = {}
    for nenv in (1024, 16384):
        env = stg.SpinTorqueVectorEnv(num_envs=nenv, device=dev, max_current=1.1e-6, include_thermal_fluctuations=True, rng_seed=4)
        env.reset(seed=4)
        a_s = torch.zeros(nenv, 2, dtype=torch.float32, device=dev)
        a_s[:, 0] = 5e-7
        a_s[:, 1] = 1e-10
        eager_ms = timed(lambda: env.step(a_s), reps=300, warm=20)
        g = env.capture_step(a_s)
        graph_ms = timed(g.replay, reps=300, warm=20)
        small[f"{nenv}_envs"] = {"eager_us_per_step": eager_ms * 1e3, "graph_us_per_step": graph_ms * 1e3,
                                 "launches_per_step": g.launches_per_replay}
    out["small_batch_100_substeps"] = small
    return out


if __name__ == "__main__":
    print(json.dumps(collect()))
