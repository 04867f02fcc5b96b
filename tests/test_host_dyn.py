"""The kernel's recursion templates, instantiated in fp64 on the host (so100_host_forward), against the oracle's
independent formulation: quaternion FK + Jacobian mass matrix + world-frame RNE + Newton (oracle) versus link-local
RNEA + composite rigid bodies + projected Gauss-Seidel (kernels)."""
import ctypes

import numpy as np

from conftest import make_oracle


def _dp(x):
    return x.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def test_host_forward_matches_oracle(native_lib, spec):
    from so100_mujoco_rl_b200 import _native
    m = spec.to_ctypes()
    o = make_oracle(5, 1)
    rng = np.random.default_rng(1)
    n = 300
    lo, hi = spec.jnt_range[:, 0], spec.jnt_range[:, 1]
    q = rng.uniform(lo - 0.05, hi + 0.05, (n, 6))
    v = rng.normal(0, 1.5, (n, 6))
    u = q + rng.uniform(-1, 1, (n, 6)) * 0.3
    M, b, a, k = np.zeros((n, 21)), np.zeros((n, 6)), np.zeros((n, 6)), np.zeros((n, 18))
    _native.check(native_lib.so100_host_forward(ctypes.byref(m), n, _dp(q), _dp(v), _dp(u), _dp(M), _dp(b), _dp(a), _dp(k), 12, 0))
    for i in range(n):
        Mo = o.mass_matrix(q[i])
        assert np.abs(np.array([Mo[r, c] for r in range(6) for c in range(r + 1)]) - M[i]).max() < 1e-14
        assert np.abs(o.bias(q[i], v[i]) - b[i]).max() < 1e-12
        ao = o.forward(q[i], v[i], u[i])[0]
        assert np.abs(ao - a[i]).max() < 1e-10 * (1 + np.abs(ao).max())
        kk = o.fk(q[i])
        assert np.abs(np.concatenate([kk["end_pos"], kk["wrist_pos"], kk["cam_xpos"], kk["cam_xmat"]]) - k[i]).max() < 1e-13


def test_gauss_seidel_contraction(native_lib, spec):
    """5 sweeps from a cold start are within fp32 resolution of the exact minimiser; with fewer scheduled sweeps the
    solver keeps sweeping (per env) while the last sweep still moved qacc by more than 1e-3 relative."""
    from so100_mujoco_rl_b200 import _native
    m = spec.to_ctypes()
    o = make_oracle(1, 1)
    rng = np.random.default_rng(2)
    n = 200
    lo, hi = spec.jnt_range[:, 0], spec.jnt_range[:, 1]
    q = rng.uniform(lo, hi, (n, 6)); v = rng.normal(0, 1, (n, 6)); u = q + rng.uniform(-1, 1, (n, 6)) * 0.075
    errs = {}
    for sweeps in (2, 5):
        a = np.zeros((n, 6))
        _native.check(native_lib.so100_host_forward(ctypes.byref(m), n, _dp(q), _dp(v), _dp(u), None, None, _dp(a), None, sweeps, 0))
        errs[sweeps] = max(np.abs(o.forward(q[i], v[i], u[i])[0] - a[i]).max() / (1 + np.abs(a[i]).max()) for i in range(n))
    assert errs[5] < 1e-7 and errs[5] < errs[2] and errs[2] < 1e-5


def test_specialised_dynamics_match_oracle(native_lib, spec):
    """The generated, model-specialised straight-line code (csrc/so100_dyn_gen.cuh) in fp64 against the oracle.
    Its constants are the fp32-rounded ones the kernels use, so agreement is to fp32 rounding of the constants."""
    from so100_mujoco_rl_b200 import _native
    m = spec.to_ctypes()
    o = make_oracle(1, 1)
    rng = np.random.default_rng(7)
    n = 200
    lo, hi = spec.jnt_range[:, 0], spec.jnt_range[:, 1]
    q = rng.uniform(lo, hi, (n, 6)); v = rng.normal(0, 1.5, (n, 6)); u = q.copy()
    Mg, bg, Ms, bs = np.zeros((n, 21)), np.zeros((n, 6)), np.zeros((n, 21)), np.zeros((n, 6))
    _native.check(native_lib.so100_host_forward(ctypes.byref(m), n, _dp(q), _dp(v), _dp(u), _dp(Mg), _dp(bg), None, None, 1, 0))
    _native.check(native_lib.so100_host_forward(ctypes.byref(m), n, _dp(q), _dp(v), _dp(u), _dp(Ms), _dp(bs), None, None, 1, 1))
    assert np.abs(Ms - Mg).max() < 2e-8      # M entries ~0.1: relative 2e-7 = fp32 rounding of the baked constants
    assert np.abs(bs - bg).max() < 5e-7      # bias ~1 N m
    for i in range(0, n, 10):
        assert np.abs(o.bias(q[i], v[i]) - bs[i]).max() < 5e-7


def test_specialised_variant_rejects_other_models(native_lib, spec):
    import copy
    from so100_mujoco_rl_b200 import _native
    s2 = copy.deepcopy(spec)
    s2.body_mass = s2.body_mass * 1.01
    m = s2.to_ctypes()
    q = np.zeros((1, 6)); M = np.zeros((1, 21))
    assert native_lib.so100_host_forward(ctypes.byref(m), 1, _dp(q), _dp(q), _dp(q), _dp(M), None, None, None, 1, 1) == -3
    assert native_lib.so100_host_forward(ctypes.byref(m), 1, _dp(q), _dp(q), _dp(q), _dp(M), None, None, None, 1, 0) == 0


def test_generated_header_is_current():
    import subprocess, sys, os
    from conftest import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_so100_dyn.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
