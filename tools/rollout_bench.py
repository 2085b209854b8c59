#!/usr/bin/env python
"""BASELINE configs[4] on one rank's shard: 131,072 STT-MRAM envs (1,048,576 over 8 GPUs), a GPU-resident SB3-shaped rollout
of n_steps=2048 driven by a fixed random MLP policy, statistics all-reduced once per rollout.

    python tools/rollout_bench.py [--envs 131072] [--n-steps 2048]          # or under torchrun with N ranks
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spin_torque_rl_gym_b200 as stg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=131072)
ap.add_argument("--n-steps", type=int, default=2048)
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)
lo, hi = stg.shard_range(a.envs * world, rank, world)
env = stg.make("SpinTorque-v0", num_envs=hi - lo, device=dev, max_current=1.1e-6, rng_seed=7, env_offset=lo)
torch.manual_seed(0)
w1, w2 = torch.randn(12, 64, device=dev) * 0.3, torch.randn(64, 2, device=dev) * 0.3


def policy(obs):
    h = torch.tanh(obs @ w1) @ w2
    act = torch.empty(obs.shape[0], 2, device=dev)
    act[:, 0] = torch.tanh(h[:, 0]) * 1.1e-6
    act[:, 1] = torch.sigmoid(h[:, 1]) * 1e-9 + 1e-11
    return act


col = stg.RolloutCollector(env, n_steps=a.n_steps)
warm = stg.RolloutCollector(env, n_steps=2, store_observations=False)
warm.collect(policy)                       # warm-up: lazy NCCL communicator setup, first-launch overheads
env.reset(seed=7)
env.reset_stats()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
stats = col.collect(policy)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
if world > 1:                               # device time, maximum over ranks
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
if rank == 0:
    print(json.dumps({"n_gpus": world, "envs_total": a.envs * world, "n_steps": a.n_steps, "rollout_ms": ms,
                      "env_steps_per_s": a.envs * world * a.n_steps / (ms * 1e-3),
                      "substeps_per_s": stats["substeps"] / (ms * 1e-3),
                      "buffer_gb": (col.observations.numel() * 4 + col.actions.numel() * 4 + col.rewards.numel() * 4
                                    + col.dones.numel()) / 1e9,
                      "success_rate": stats["success_rate"], "mean_episode_length": stats["mean_episode_length"],
                      "episodes": stats["episodes"]}))
if world > 1:
    torch.distributed.destroy_process_group()
