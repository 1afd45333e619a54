"""Shared test helpers: synthetic problem generation for both the CUDA path
(mhpc_minimal_env_b200) and the oracle (oracle/oracle_py.py), from the same host-generated bits."""
from __future__ import annotations

import numpy as np

DEFAULT_X0 = np.array([0.0927, -0.1093, -0.1542, 1.0957, -2.2033, 0.9742, -1.7098,
                       0.9011, 0.2756, 0.7333, 0.0446, 0.0009, 1.3219, 2.7346])
SEED = 20260101
SCALE = np.array([0.01] * 3 + [0.02] * 4 + [0.05] * 7)


def perturbed_x0(i: int, scale: float = 1.0) -> np.ndarray:
    """x0_i = x0 + diag(s) z_i, z ~ N(0,1) clipped to +-2; counter-based stream = instance index."""
    if i == 0:
        return DEFAULT_X0.copy()
    g = np.random.Generator(np.random.Philox(key=SEED + i))
    z = np.clip(g.standard_normal(14), -2.0, 2.0)
    return DEFAULT_X0 + scale * SCALE * z


def make_specs(n: int, config: int = 2, scale: float = 1.0):
    """List of dicts with the per-instance problem fields for BASELINE.json configs 1-5."""
    specs = []
    for i in range(n):
        s = dict(n_wbphase=4, n_fbphase=4, dt_wb=0.001, dt_fb=0.001, cmode=1, vel=1.5, height=0.0,
                 gait_mode=[1, 2, 3, 4], gait_timing=[0.08, 0.1, 0.08, 0.1], x0=perturbed_x0(i, scale))
        if config == 1:
            s["x0"] = DEFAULT_X0.copy()
        elif config == 3:
            s["n_wbphase"], s["n_fbphase"] = 8, 0
        elif config == 4:
            g = np.random.Generator(np.random.Philox(key=SEED + 7919 * (i + 1)))
            nwb = int(g.integers(1, 9))
            s["n_wbphase"], s["n_fbphase"] = nwb, 8 - nwb
            s["gait_timing"] = [float(g.integers(60, 111)) / 1000.0 for _ in range(4)]
        specs.append(s)
    return specs


def to_gpu_problem(spec):
    import mhpc_minimal_env_b200 as m
    return m.default_problem(**spec)


def to_orc_problem(spec, **opt):
    from oracle.oracle_py import default_problem
    kw = dict(spec)
    kw.update(opt)
    return default_problem(**kw)


def relerr_rows(a, b, floor_frac=1e-6):
    """Worst PER-KNOT (per-matrix) norm-relative error of a [N, dim] field: max_k |a_k - b_k|_inf / |b_k|_inf.  A knot whose
    reference row is (numerically) empty is measured against floor_frac x the largest entry of the whole array instead, so
    that exact zeros / cancellation rows do not divide by nothing."""
    a = np.atleast_2d(np.asarray(a, dtype=np.float64))
    b = np.atleast_2d(np.asarray(b, dtype=np.float64))
    if a.size == 0:
        return 0.0
    glob = float(np.max(np.abs(b)))
    den = np.maximum(np.max(np.abs(b), axis=1), max(floor_frac * glob, 1e-300))
    return float(np.max(np.max(np.abs(a - b), axis=1) / den))


def relerr(a, b):
    """norm-relative error max|a-b| / max(|b|_inf, tiny) over the WHOLE array (one matrix / one vector)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(np.max(np.abs(b)) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b)) / den) if a.size else 0.0
