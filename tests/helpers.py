"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from mohid_b200.synthetic import make_case, default_params
from oracle.oracle import OracleAdvectionDiffusion, case_to_numpy

NULL_REAL = -9.9e15


def oracle_for(case, **kw):
    g, s, props, refs = case_to_numpy(case)
    o = OracleAdvectionDiffusion(case.I, case.J, case.K, case.ld, **kw)
    o.set_grid2d(g)
    o.set_step(s)
    return o, g, s, props, refs


def rel_err(a, b, mask):
    """max |a-b| / max(|b|, 1) over `mask` (property fields are O(1..40))."""
    d = np.abs(a[mask] - b[mask])
    return float((d / np.maximum(np.abs(b[mask]), 1.0)).max()) if d.size else 0.0


def water_mask(s):
    return s["WaterPoints3D"] == 1
