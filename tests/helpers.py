"""Shared helpers of the test-suite (golden loading, noise replay, error metrics)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_case(fname, name):
    g = np.load(os.path.join(GOLDEN, fname))
    return {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(name + "/")}


def stt_params_for(c):
    """Reference STT defaults overridden by whatever the golden case recorded."""
    from oracle.stt_oracle import default_stt_params
    p = default_stt_params()
    for k in ("volume", "easy_axis", "reference_magnetization", "damping", "polarization", "resistance_parallel",
              "resistance_antiparallel"):
        if k in c:
            p[k] = c[k] if c[k].ndim else float(c[k])
    return p


def noise_for_step(seed, action, max_current, max_duration=5e-9, stages=4):
    """The N(0,1) tensor the reference consumes in one thermal step after np.random.seed(seed)."""
    from oracle.stt_oracle import parse_action, substep_plan
    _, t = parse_action(np.array(action, dtype=np.float32), max_current, max_duration)
    n, _ = substep_plan(t)
    np.random.seed(int(seed))
    return np.random.normal(0, 1, (n, stages, 3))


def rel_err(a, ref, floor=1e-300):
    return float((np.abs(a - ref) / np.maximum(np.abs(ref), floor)).max())


def transverse_rel_err(m, ref):
    """|delta(m_x, m_y)| / |(m_x, m_y)_ref|: the transverse pair compared as a vector relative to its own magnitude (magnitude
    ratio and azimuth together), meaningful down to 1e-300 where the absolute error is zero for every practical purpose."""
    return float(np.hypot(m[0] - ref[0], m[1] - ref[1]) / max(np.hypot(ref[0], ref[1]), 1e-300))
