#!/usr/bin/env python
"""Tiny driver for ncu: a few launches of one K1 variant at bench size.

    python tools/profile_step.py --thermal 1 --envs 1048576 --steps 3 [--dtype f64] [--tilted]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spin_torque_rl_gym_b200 import SpinTorqueVectorEnv, params  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--thermal", type=int, default=1)
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--tilted", action="store_true")
ap.add_argument("--pulse", type=float, default=1e-9)
ap.add_argument("--no-pair", action="store_true")
ap.add_argument("--pair-always", action="store_true", help="two envs per thread at any batch size (STG_F_PAIR_ALWAYS)")
ap.add_argument("--stream", default="xoshiro", choices=["xoshiro", "philox"])
ap.add_argument("--no-sort", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
p = params.default_device_parameters("stt_mram")
if a.tilted:
    p["easy_axis"] = np.array([0.2, -0.1, 1.0])
env = SpinTorqueVectorEnv(num_envs=a.envs, device=dev, dtype=torch.float32 if a.dtype == "f32" else torch.float64,
                          device_params=p, max_current=1.1e-6, include_thermal_fluctuations=bool(a.thermal), rng_seed=1,
                          pair_kernel=("always" if a.pair_always else not a.no_pair), thermal_stream=a.stream, sort_by_substeps=False if a.no_sort else 'auto')
env.reset(seed=1)
rng = np.random.default_rng(0)
act = np.stack([rng.uniform(-1.1e-6, 1.1e-6, a.envs), np.full(a.envs, a.pulse)], 1).astype(np.float32)
act = torch.from_numpy(act).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
env.step(act)
torch.cuda.synchronize()
e0.record()
for _ in range(a.steps):
    env.step(act)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f"thermal={a.thermal} dtype={a.dtype} tilted={a.tilted} envs={a.envs}: {ms:.3f} ms/step, "
      f"{a.envs * env._n_sub[0].item() / ms / 1e6:.2f} G substeps/s")
