"""Time the LIVE reference's own SpinTorqueEnv.step in the build container (SURVEY §8d "CPU baseline timing").

The reference is a Python package that cannot travel to the GPU box, so this figure is measured here, next to the checkout at
/root/reference, and recorded in DESIGN.md §6 as a build-container number; bench.py's cpu_baseline (the C and NumPy ports of
the same arithmetic) is what runs beside the GPU. Sanitised as the oracle requires (no wall-clock RK4->Euler switch, no memo
caches). Workload: the bench's (stt_mram, well-conditioned max_current, 1 ns pulses = 999 RK4 substeps), thermal off and on.

    timeout 900 python tools/time_live_reference.py [seconds_per_case]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import gen_golden as GG  # noqa: E402  (import helpers + sanitisation only)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
    SpinTorqueEnv, _ = GG._import_reference()
    jm = 1.1e-6
    rng = np.random.default_rng(0)
    for thermal in (False, True):
        env = GG._sanitise(SpinTorqueEnv(device_type="stt_mram", device_params=GG._stt_params(), max_current=jm,
                                         temperature=300.0, include_thermal_fluctuations=thermal, max_steps=10 ** 9, seed=0))
        env.reset(seed=0, options={"initial_state": np.array([0.3, 0.2, 0.9]), "target_state": np.array([0.0, 0.0, -1.0])})
        env.step(np.array([0.5 * jm, 1e-9], dtype=np.float32))                     # warm-up (imports, first-call paths)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < budget:
            a = np.array([rng.uniform(-jm, jm), 1e-9], dtype=np.float32)
            _, _, term, trunc, info = env.step(a)
            assert info["simulation_success"]
            n += 1
            if term or trunc:
                env.reset(seed=n)
        dt = time.perf_counter() - t0
        print(f"live reference, 1 process, thermal {'on' if thermal else 'off'}: {n} env.step in {dt:.1f} s = "
              f"{n / dt:.3f} env-steps/s = {n * 999 / dt:.4g} LLGS substeps/s (999 RK4 substeps per step)", flush=True)


if __name__ == "__main__":
    main()
