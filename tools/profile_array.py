#!/usr/bin/env python
"""Tiny driver for ncu / timing of the K3 crossbar kernel at BASELINE configs[3] size.

    python tools/profile_array.py --mode individual --arrays 16384 --steps 5
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spin_torque_rl_gym_b200 as stg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="individual")
ap.add_argument("--arrays", type=int, default=16384)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
env = stg.SpinTorqueArrayVectorEnv(num_envs=a.arrays, array_size=(8, 8), action_mode=a.mode, device=dev, rng_seed=1)
env.reset(seed=1)
rng = np.random.default_rng(0)
k = 2 if a.mode == "global" else 3
act = torch.zeros(a.arrays, k, dtype=torch.float32, device=dev)
if a.mode == "global":
    act[:, 0] = torch.from_numpy(rng.uniform(-2e6, 2e6, a.arrays)).to(dev); act[:, 1] = 2e-9
else:
    hi = 63 if a.mode == "individual" else 7
    act[:, 0] = torch.from_numpy(rng.uniform(0, hi, a.arrays)).to(dev)
    act[:, 1] = torch.from_numpy(rng.uniform(-2e6, 2e6, a.arrays)).to(dev); act[:, 2] = 2e-9
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
env.step(act)
torch.cuda.synchronize()
e0.record()
for _ in range(a.steps):
    env.step(act)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f"mode={a.mode} arrays={a.arrays}: {ms * 1e3:.1f} us/step, {a.arrays / ms / 1e3:.2f} M array-steps/s, "
      f"{a.arrays * 64 * 96 / ms / 1e6:.0f} GB/s algorithmic")
