"""Regenerates tests/golden/transport_cases.npz.

The reference ships no golden vectors for this path and cannot be built here (parity unpinned),
so these vectors do NOT pin the oracle to the reference.  They pin the oracle -- and through tests/test_gpu_parity.py
the CUDA path -- to the state that was reviewed line by line against ModuleAdvectionDiffusion.F90 / ModuleFunctions.F90,
so that a later edit of either side cannot drift unnoticed.  Inputs are not stored: they come from the seeded generator
(mohid_b200/synthetic.py); a digest of the inputs is stored to catch generator drift.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mohid_b200.synthetic import make_case, default_params          # noqa: E402
from helpers import oracle_for                                        # noqa: E402

# name -> (I, J, K, nprop, params per property, steps)
CASES = {
    "upwind_implicit": (14, 11, 4, 1, [default_params(1, 4, 1, 4)], 3),
    "tvd_superbee_optimize": (14, 11, 4, 2, [default_params(4, 4, 4, 4, bc=4), default_params(4, 4, 4, 4, bc=4)], 3),
    "tvd_vanleer_explicit_v": (14, 11, 4, 1, [default_params(4, 2, 4, 2, impexp_advv=0.0, theta_difv=0.5)], 3),
    "quick_h_upwind_v": (14, 11, 4, 1, [default_params(2, 4, 1, 4, bc=1, decay_time=900.0)], 3),
    "quickest_explicit": (14, 11, 4, 1, [default_params(3, 4, 3, 4, impexp_advv=0.0)], 3),
    "central_massconsnullgrad": (14, 11, 4, 1, [default_params(5, 4, 5, 4, bc=7)], 3),
    "implicit_xx": (14, 11, 4, 1, [dict(default_params(4, 4, 4, 4), ImpExp_AdvXX=1.0)], 2),
    "implicit_yy": (14, 11, 4, 1, [dict(default_params(1, 4, 1, 4), ImpExp_AdvYY=1.0)], 2),
    "implicit_xx_2d_domain": (14, 11, 1, 2, [dict(default_params(4, 4, 4, 4, bc=4), ImpExp_AdvXX=1.0),
                                               dict(default_params(4, 4, 4, 4, bc=1), ImpExp_AdvYY=1.0)], 2),
    "cyclic_boundary": (14, 11, 4, 1, [default_params(4, 4, 4, 4, bc=8)], 3),
    "leapfrog_explicit_v": (14, 11, 4, 1, [default_params(6, 4, 6, 4, impexp_advv=0.0, theta_difv=0.3, bc=2)], 3),
    "tvd_pdm": (14, 11, 4, 2, [default_params(4, 5, 4, 5, bc=7), default_params(4, 5, 4, 5, decay_time=600.0, bc=5)], 3),
}


def digest(arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def run_case(name):
    I, J, K, N, prm, steps = CASES[name]
    case = make_case(I, J, K, nprop=N, stepped_bottom=K > 1, seed=20260101)
    o, g, s, props, refs = oracle_for(case)
    out = [p.copy() for p in props]
    for _ in range(steps):
        o.advect_batch(out, prm, refs)
    inputs = [g[k] for k in sorted(g)] + [s[k] for k in sorted(s)] + list(props) + list(refs)
    return np.stack(out), digest(inputs)


if __name__ == "__main__":
    data = {}
    for name in CASES:
        out, dg = run_case(name)
        data[name] = out
        data[name + "__inputs_sha256"] = np.frombuffer(bytes.fromhex(dg), dtype=np.uint8)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "transport_cases.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes")
