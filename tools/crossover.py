#!/usr/bin/env python
"""Two-envs-per-thread vs one-env-per-thread thermal step kernel by batch size (the scan behind STG_PAIR_THERMAL_MIN_ENVS,
profiles/README.md "Dispatch by batch size").

    python tools/crossover.py [--stream xoshiro|philox] [--pulse 1e-9]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spin_torque_rl_gym_b200 import SpinTorqueVectorEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--stream", default="xoshiro", choices=["xoshiro", "philox"])
ap.add_argument("--pulse", type=float, default=1e-9)
a = ap.parse_args()
dev = torch.device("cuda", 0)
for n in (1024, 4096, 16384, 32768, 65536, 98304, 131072, 196608, 262144, 524288, 1048576):
    row = [n]
    for pair in ("always", False):
        env = SpinTorqueVectorEnv(num_envs=n, device=dev, max_current=1.1e-6, include_thermal_fluctuations=True, rng_seed=1,
                                  pair_kernel=pair, sort_by_substeps=False, thermal_stream=a.stream)
        env.reset(seed=1)
        rng = np.random.default_rng(0)
        act = torch.from_numpy(np.stack([rng.uniform(-1.1e-6, 1.1e-6, n), np.full(n, a.pulse)], 1).astype(np.float32)).to(dev)
        for _ in range(3):
            env.step(act)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            env.step(act)
        e1.record()
        torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / reps)
        del env
    print(f"n={row[0]:8d}  two envs per thread {row[1]:8.3f} ms   one env per thread {row[2]:8.3f} ms   ratio {row[1] / row[2]:.3f}",
          flush=True)
