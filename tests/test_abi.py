"""The drop-in boundary: libstg.so loads, exports every symbol include/stg.h declares, the ctypes mirrors have the C
layout, argument validation returns the documented error codes (no GPU needed: validation precedes any launch)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from spin_torque_rl_gym_b200 import _lib, params as P
from tests.helpers import ROOT

HEADER = os.path.join(ROOT, "include", "stg.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stg_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in stg.h but not exported by libstg.so"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes binding table and stg.h disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.lib_path()], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", out), name
    assert lib.stg_abi_version() == _lib.ABI_VERSION == 7


def test_struct_layouts_match_the_c_header(tmp_path):
    """Compile a tiny C program against include/stg.h and compare sizeof/offsetof with the ctypes mirrors."""
    fields = {
        "StgSttParams": ["damping", "easy_axis", "reference_magnetization", "area", "temperature", "applied_field",
                         "max_step", "max_steps", "solver_valid"],
        "StgSttStepArgs": ["d_table", "state", "d_action", "out", "d_noise", "noise_stride", "d_perm",
                           "d_target_table", "seed", "env_offset", "n_envs", "n_sets", "n_targets", "flags"],
        "StgSttResetArgs": ["d_mask", "d_m0", "d_target0", "d_target_table", "d_obs", "seed", "n_envs", "n_targets"],
        "StgSttSolveArgs": ["d_m0", "d_pulse", "d_m_out", "d_traj", "traj_stride", "d_n_sub", "d_guard", "d_noise",
                            "noise_stride", "seed", "n_envs", "n_sets", "flags"],
        "StgSttStepOut": ["obs", "reward", "status", "final_obs", "stats"],
        "StgSttState": ["m", "episode"],
        "StgSttFolded": ["v"],
        "StgRk45Args": ["d_table", "d_t_end", "d_traj", "traj_stride", "rtol", "max_attempts", "n_envs", "flags", "d_perm",
                        "d_t_start", "d_seg_t", "d_seg_current", "d_seg_field", "n_seg", "seg_rows"],
        "StgThermalAnalyticsArgs": ["d_temperature", "d_barrier", "d_out", "k_b", "failure_rate", "n_t", "n_dev"],
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    for s, fs in fields.items():
        lines.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for f in fs:
            lines.append(f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    lines.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    got = dict(l.split() for l in out if l)
    for s, fs in fields.items():
        cls = getattr(_lib, s)
        assert int(got[s]) == C.sizeof(cls), s
        for f in fs:
            assert int(got[f"{s}.{f}"]) == getattr(cls, f).offset, f"{s}.{f}"


def test_fold_known_answers():
    """Folded constants against the reference formulas (physics/simple_solver.py:338,368,375-380)."""
    p = P.default_device_parameters("stt_mram")
    st = P.make_param_struct("stt_mram", p, max_steps=100, max_current=2e6, max_duration=5e-9, temperature=300.0,
                             thermal=True, success_threshold=0.9, energy_penalty_weight=0.1)
    f = P.fold([st])[0]
    mu0 = 4 * np.pi * 1e-7
    assert f[0] == 0.01
    assert f[1] == 2.21e5 / (1 + 0.01 ** 2)
    assert f[2] == (2 * 1.2e6) / (mu0 * 800e3)
    assert f[4] == pytest.approx(0.7 / (800e3 * p["volume"]), rel=1e-15)
    assert f[5] == pytest.approx(np.sqrt(2 * 0.01 * 1.38e-23 * 300.0 / (mu0 * 800e3 * p["volume"] * 2.21e5)), rel=1e-15)
    assert list(f[6:9]) == [0, 0, 1] and P.all_axis_z(f[None])
    st2 = P.make_param_struct("stt_mram", dict(p, easy_axis=np.array([1.0, 2.0, 2.0])), max_steps=100, max_current=2e6,
                              max_duration=5e-9, temperature=0.0, thermal=True, success_threshold=0.9,
                              energy_penalty_weight=0.1)
    f2 = P.fold([st2])[0]
    assert np.allclose(f2[6:9], np.array([1, 2, 2]) / 3.0, rtol=1e-16)
    assert f2[5] == 0.0 and not P.all_axis_z(f2[None])
    assert f2[28] == 0.0          # temperature <= 0 fails RobustLLGSSolver._validate_inputs => solver_valid = 0


def test_solver_validation_semantics():
    """SOT/VCMA dicts without `polarization` never integrate in the reference."""
    sot = P.default_device_parameters("sot_mram")
    assert not P.solver_accepts(sot, 300.0)
    assert P.solver_accepts(dict(sot, polarization=0.6), 300.0)
    assert P.solver_accepts(P.default_device_parameters("stt_mram"), 300.0)
    assert not P.solver_accepts(P.default_device_parameters("stt_mram"), 0.0)
    assert not P.solver_accepts(dict(P.default_device_parameters("stt_mram"), damping=1.5), 300.0)
    with pytest.raises(RuntimeError, match="Missing required parameter: easy_axis"):
        P.check_device_constructible("sot_mram", P.env_default_device_params("sot_mram"))
    with pytest.raises(ValueError, match="Unknown device type"):
        P.check_device_constructible("nope", {})


def test_argument_validation_error_codes():
    lib = _lib.load()
    assert lib.stg_stt_fold(None, 1, None) == -1
    a = _lib.StgSttStepArgs()
    assert lib.stg_stt_step_f32(None, None) == -1
    a.n_envs, a.n_sets = 4, 0
    assert lib.stg_stt_step_f32(C.byref(a), None) == -2            # STG_E_SIZE
    a.n_sets = 1
    assert lib.stg_stt_step_f64(C.byref(a), None) == -1            # STG_E_NULL: no buffers
    r = _lib.StgSttResetArgs()
    r.n_envs, r.n_sets = -1, 1
    assert lib.stg_stt_reset(C.byref(r), None) == -2
    s = _lib.StgSttSolveArgs()
    s.n_envs, s.n_sets = 1, 1
    assert lib.stg_stt_solve_f32(C.byref(s), None) == -1
    assert lib.stg_stt_sort_by_substeps(None, 1, None, None, None, None, 1, None) == -1
    assert lib.stg_error_string(-3).decode().startswith("unknown enum")
    bad = _lib.StgSttParams()
    bad.device_kind = 7
    out = _lib.StgSttFolded()
    assert lib.stg_stt_fold(C.byref(bad), 1, C.byref(out)) == -3   # STG_E_ENUM


def test_no_cpu_fallback_in_product_path():
    """The product package must not reach into oracle/ or tests/hostsim, and must refuse to run without CUDA."""
    pkg = os.path.join(ROOT, "spin_torque_rl_gym_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions oracle"
                assert "hostsim" not in src, f
    import torch
    if not torch.cuda.is_available():
        from spin_torque_rl_gym_b200 import SpinTorqueVectorEnv
        with pytest.raises(_lib.StgError):
            SpinTorqueVectorEnv(num_envs=4)


def test_integration_stub_names_the_current_abi_version():
    """INTEGRATION.md's binding stub asserts an ABI version: it must be the one the header and the package carry."""
    import re
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"stg_abi_version\(\) == (\d+)", text)
    hdr = re.search(r"#define STG_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "stg.h")).read())
    assert m and hdr and int(m.group(1)) == int(hdr.group(1)) == _lib.ABI_VERSION


def test_flag_constants_of_the_binding_equal_the_header():
    hdr = open(os.path.join(ROOT, "include", "stg.h")).read()
    flags = re.findall(r"#define (STG_F_\w+) (0x[0-9a-fA-F]+)u", hdr)
    assert len(flags) >= 11
    for name, val in flags:
        assert getattr(_lib, name[len("STG_"):]) == int(val, 16), name
    assert len({int(v, 16) for _, v in flags}) == len(flags)          # no two flags share a bit


def test_thermal_pair_kernel_dispatch_rule():
    """stg_stt_thermal_pair_dispatch on 148 SMs (B200): the packed kernel from 262,144 envs (all-Philox stream: 524,288), and in
    the one-wave window exactly where it halves the busiest scheduler's warp count (measured table in profiles/README.md)."""
    lib = _lib.load()
    f = _lib.F_THERMAL_PHILOX | _lib.F_AXIS_Z
    want = {1024: 0, 32768: 0, 49152: 0, 56832: 0, 57000: 1, 65536: 1, 75776: 1, 76000: 0, 90000: 0, 95000: 1, 113664: 1,
            114000: 0, 131072: 0, 196608: 0, 262143: 0, 262144: 1, 1 << 20: 1}
    for n, w in want.items():
        assert lib.stg_stt_thermal_pair_dispatch(n, f, 148) == w, n
        assert lib.stg_stt_thermal_pair_dispatch(n, f | _lib.F_NO_PAIR, 148) == 0
        assert lib.stg_stt_thermal_pair_dispatch(n, f | _lib.F_PAIR_ALWAYS, 148) == 1
        assert lib.stg_stt_thermal_pair_dispatch(n, f | _lib.F_STREAM_PHILOX10, 148) == (1 if n >= (1 << 19) else 0)
