"""Time the LIVE reference's own CPU path (SURVEY §8d "CPU baseline timing") wherever a checkout of it is reachable.

    python tools/time_live_reference.py [--seconds S] [--procs N] [--json] [--no-vectorized]

Probed locations (first hit wins): $STG_REFERENCE, baseline/_ref, oracle/_ref, /root/reference - a directory that holds
spin_torque_gym/envs/spin_torque_env.py. gymnasium / matplotlib are not installed in this image, so the package is imported
through the stand-ins under oracle/shims. The reference is sanitised as the oracle requires (no wall-clock RK4->Euler switch,
no memo caches). Workload: the bench's (stt_mram defaults, well-conditioned max_current, 1 ns pulses = 999 RK4
substeps per env.step, T = 300 K), one process per host core, each stepping its own SpinTorqueEnv for S seconds; plus the
reference's own batched-NumPy entry point VectorizedSolver.solve_batch at N = 65,536 (Euler steps).

bench.py runs this file as a subprocess (`--json`) for its `cpu_baseline` / `--impl reference` legs when a checkout is found;
on a box without one (the GPU boxes: the reference is a pure-Python package whose own packaging installs only 3 of its 62
modules) the bench falls back to the C restatement and says so.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JM = 1.1e-6


def find_reference():
    if os.environ.get("STG_NO_LIVE_REFERENCE"):        # tests: exercise the port path on a box that has a checkout
        return None
    cands = [os.environ.get("STG_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), os.path.join(ROOT, "oracle", "_ref"),
             "/root/reference"]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "spin_torque_gym", "envs", "spin_torque_env.py")):
            return c
    return None


def _import(ref):
    import logging
    import warnings
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    sys.path.insert(0, ref)
    logging.disable(logging.CRITICAL)
    warnings.filterwarnings("ignore")
    import spin_torque_gym  # noqa: F401
    from spin_torque_gym.envs import SpinTorqueEnv
    return SpinTorqueEnv


def _worker(args):
    ref, seconds, thermal, seed = args
    import numpy as np
    SpinTorqueEnv = _import(ref)
    from spin_torque_gym.devices import DeviceFactory
    p = DeviceFactory().get_default_parameters("stt_mram")
    env = SpinTorqueEnv(device_type="stt_mram", device_params=p, max_current=JM, temperature=300.0,
                        include_thermal_fluctuations=thermal, max_steps=10 ** 9, seed=seed)
    env.solver.timeout = 1e9                 # SURVEY §8c sanitisation
    env.optimizer.cache.ttl = -1
    env.cache_observations = False
    rng = np.random.default_rng(seed)
    env.reset(seed=seed)
    env.step(np.array([0.5 * JM, 1e-9], dtype=np.float32))          # warm-up (imports, first-call paths)
    n, sub, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        a = np.array([rng.uniform(-JM, JM), 1e-9], dtype=np.float32)
        _, _, term, trunc, info = env.step(a)
        n += 1
        sub += 999
        if term or trunc:
            env.reset(seed=seed + n)
    return n, sub, time.perf_counter() - t0


def _vectorized(ref, n=65536, t_end=1e-10):
    """VectorizedSolver.solve_batch (utils/vectorized_operations.py:32-121): Euler substeps/s of the reference's NumPy batch path."""
    import numpy as np
    _import(ref)
    from spin_torque_gym.devices import DeviceFactory
    from spin_torque_gym.utils.vectorized_operations import VectorizedSolver
    p = DeviceFactory().get_default_parameters("stt_mram")
    rng = np.random.default_rng(0)
    m0 = rng.normal(size=(n, 3))
    m0 /= np.linalg.norm(m0, axis=1, keepdims=True)
    solver = VectorizedSolver()
    t0 = time.perf_counter()
    res = solver.solve_batch(m0, (0.0, t_end), [p] * n, dt=1e-12)
    dt = time.perf_counter() - t0
    steps = int(res[0].get("n_steps", 0)) if res and res[0].get("success") else 0
    return {"n": n, "euler_substeps_per_env": steps, "seconds": dt, "substeps_per_s": n * steps / dt if steps else None}


def measure(seconds=10.0, procs=None, thermal=True, vectorized=True):
    ref = find_reference()
    if ref is None:
        return {"available": False, "why": "no checkout of the reference under $STG_REFERENCE, baseline/_ref, oracle/_ref or "
                                           "/root/reference (pure-Python package, cannot travel to this box)"}
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_worker, [(ref, seconds, thermal, 1000 + i) for i in range(procs)])
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    sub_rate = sum(r[1] / r[2] for r in res)          # every process timed its own loop (start-up excluded)
    out = {"available": True, "path": ref, "procs": procs, "seconds_per_proc": seconds, "thermal": thermal, "env_steps": steps,
           "env_steps_per_s": sum(r[0] / r[2] for r in res), "substeps_per_s": sub_rate, "wall_s": wall}
    if vectorized:
        try:
            out["vectorized_solver"] = _vectorized(ref)
        except Exception as exc:  # noqa: BLE001
            out["vectorized_solver"] = {"error": repr(exc)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--no-thermal", action="store_true")
    ap.add_argument("--no-vectorized", action="store_true")
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args()
    out = measure(a.seconds, a.procs or None, not a.no_thermal, not a.no_vectorized)
    if a.json:
        print(json.dumps(out))
        return
    if not out["available"]:
        print("live reference not reachable:", out["why"])
        return
    print(f"live reference at {out['path']}: {out['procs']} processes x {out['seconds_per_proc']:.0f} s, thermal "
          f"{'on' if out['thermal'] else 'off'}: {out['env_steps']} env.step = {out['env_steps_per_s']:.3f} env-steps/s = "
          f"{out['substeps_per_s']:.4g} LLGS substeps/s (999 RK4 substeps per step)")
    if "vectorized_solver" in out:
        print("VectorizedSolver.solve_batch:", out["vectorized_solver"])


if __name__ == "__main__":
    main()
