"""Host-side multi-rank logic on CPU: env-id sharding and the one statistics all-reduce, world_size 2, gloo backend."""
import os
import subprocess
import sys

import pytest

from spin_torque_rl_gym_b200.parallel import shard_range
from tests.helpers import ROOT


def test_shard_ranges_partition_the_batch():
    for total, world in [(1 << 20, 8), (65536, 2), (10, 4), (7, 8)]:
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a1 >= a0
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(1 << 20, 3, 8) == (3 * 131072, 4 * 131072)
    with pytest.raises(ValueError):
        shard_range(8, 8, 8)


WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from spin_torque_rl_gym_b200.parallel import all_reduce_stats, shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(1000, rank, world)
# per-rank stats vector as the kernel accumulates it: steps, substeps, terminated, truncated, energy, reward, guard, eplen
s = torch.tensor([hi - lo, (hi - lo) * 999.0, 10.0 * (rank + 1), 5.0, 1e-12 * (rank + 1), 2.5, 0.0, 40.0 * (rank + 1)],
                 dtype=torch.float64)
out = all_reduce_stats(s)
if rank == 0:
    print("RESULT " + json.dumps(out))
dist.destroy_process_group()
'''


def test_stats_all_reduce_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][0]
    import json
    out = json.loads(line[7:])
    assert out["steps"] == 1000 and out["substeps"] == 999000
    assert out["terminated"] == 30 and out["truncated"] == 10 and out["episodes"] == 40
    assert out["success_rate"] == pytest.approx(0.75)
    assert out["mean_episode_length"] == pytest.approx(120 / 40)
    assert out["energy"] == pytest.approx(3e-12)
