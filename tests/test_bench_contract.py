"""bench.py driver contract: the reference arm runs on CPU here and prints one JSON line with the required keys."""
import json
import os
import subprocess
import sys

from tests.helpers import ROOT


import pytest


@pytest.mark.parametrize("live", [False, True])
def test_reference_arm_prints_one_contract_line(live):
    """`--impl reference`: the live reference (kind "reference") where a checkout is reachable - the build container - and the C
    restatement (kind "port", with the reason) elsewhere, e.g. on the GPU boxes."""
    from tools.time_live_reference import find_reference
    env = dict(os.environ)
    if live:
        if find_reference() is None:
            pytest.skip("no checkout of the reference on this box")
    else:
        env["STG_NO_LIVE_REFERENCE"] = "1"
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=400, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == ("reference" if live else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    if live:
        assert d["cpu_baseline"]["vectorized_solver"]["substeps_per_s"] > 0      # VectorizedSolver.solve_batch at N = 65,536
    else:
        assert "not reachable" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_graft_entry_has_build_and_smoke():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    assert callable(g.build) and callable(g.smoke)
