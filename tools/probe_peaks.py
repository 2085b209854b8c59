#!/usr/bin/env python
"""FMA-pipe peaks measured with stg_probe_fma: FFMA, packed FFMA2, DFMA."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spin_torque_rl_gym_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
blocks, iters = 148 * 16, 4096
buf = torch.empty(blocks * 256 * 2, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream
for mode, name, fma_per_iter in ((0, "FFMA  (fp32)", 64), (2, "FFMA2 (fp32x2)", 128), (1, "DFMA  (fp64)", 64)):
    best = 0.0
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.stg_probe_fma(buf.data_ptr(), blocks, iters, mode, stream))
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = max(best, blocks * 256 * iters * fma_per_iter * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    print(f"{name}: {best:.2f} TFLOP/s")
