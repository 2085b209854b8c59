"""Multi-GPU plumbing: the env batch shards by contiguous global env-id ranges (one process per GPU, no data-path
collective); the only collective is one all-reduce(SUM) of the STG_NSTATS-element episode-statistics vector per rollout
(SURVEY.md §8e — the reference analogue is the per-env EnvironmentMonitor, utils/monitoring.py:89-116)."""
from __future__ import annotations

from typing import Dict, Tuple

from . import _lib


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of the global env ids owned by `rank`. Philox counters use the GLOBAL id (env_offset=start), so results
    do not depend on the number of GPUs."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_reduce_stats(stats, group=None) -> Dict[str, float]:
    """SUM-reduce the per-rank statistics vector (NCCL for CUDA tensors, gloo for CPU tensors) and return it as a dict with
    the derived episode metrics. Without an initialised process group the local vector is returned."""
    import torch
    import torch.distributed as dist
    t = stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    vals = t.detach().to("cpu", torch.float64).tolist()
    out = dict(zip(_lib.STAT_NAMES, vals))
    episodes = out["terminated"] + out["truncated"]
    out["episodes"] = episodes
    out["success_rate"] = out["terminated"] / episodes if episodes else 0.0
    out["mean_episode_length"] = out["episode_length"] / episodes if episodes else 0.0
    out["mean_step_energy"] = out["energy"] / out["steps"] if out["steps"] else 0.0
    out["mean_reward"] = out["reward"] / out["steps"] if out["steps"] else 0.0
    return out
