#!/usr/bin/env python
"""Driver for ncu: the K2 RK45 kernel on the SOT/VCMA mix of BASELINE configs[2] (reduced trajectory length)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spin_torque_rl_gym_b200 import params as P  # noqa: E402
from spin_torque_rl_gym_b200.physics import LLGSSolver  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
t = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-10
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
sot = P.default_device_parameters("sot_mram"); sot.update(aspect_ratio=2.0, spin_hall_angle=0.3)
vcma = P.default_device_parameters("vcma_mram"); vcma.update(aspect_ratio=1.5)
m0 = torch.from_numpy(rng.normal(size=(n, 3))).to(dev)
pidx = (torch.arange(n, device=dev, dtype=torch.int32) % 2) if os.environ.get("RK_INTERLEAVE", "1") == "1" else (torch.arange(n, device=dev, dtype=torch.int32) >= n // 2).to(torch.int32)
z = torch.zeros(n, dtype=torch.float64, device=dev)
cur = torch.where(pidx == 0, torch.from_numpy(rng.uniform(-3e11, 3e11, n)).to(dev), z)
volt = torch.where(pidx == 1, torch.from_numpy(rng.uniform(-2.5, 2.5, n)).to(dev), z)
t_end = torch.full((n,), t, dtype=torch.float64, device=dev)
if os.environ.get("RK_RAGGED", "0") == "1":          # ragged trajectory lengths: t_end ~ U(0.05, 1) * t
    t_end = t_end * torch.from_numpy(rng.uniform(0.05, 1.0, n)).to(dev)
solver = LLGSSolver(device=dev, sort_trajectories=os.environ.get("RK_SORT", "1") == "1")
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = solver.solve_batch(m0, t_end, [sot, vcma], current=cur, voltage=volt, param_index=pidx,
                           device_type=["sot_mram", "vcma_mram"])
    e1.record()
    torch.cuda.synchronize()
att = float((r["n_accepted"] + r["n_rejected"]).sum())
amax, amin = int((r["n_accepted"] + r["n_rejected"]).max()), int((r["n_accepted"] + r["n_rejected"]).min())
print(f"sort={solver.sort_trajectories} ragged={os.environ.get('RK_RAGGED', '0')} attempts min/max {amin}/{amax}; "
      f"rk45 n={n}: {e0.elapsed_time(e1):.3f} ms, {att / n:.1f} attempts/env, {att / e0.elapsed_time(e1) / 1e6:.3f} G attempts/s")
