"""Small fused-kernel runs for compute-sanitizer (memcheck / racecheck / synccheck): TVD with an open boundary, a stepped
bottom and a ragged last strip; upwind; a forced narrow chunk."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["MOHID_ADT_LEAN_ALWAYS"] = "1"
os.environ["MOHID_ADT_CHUNK_COLS"] = "9"
from mohid_b200.synthetic import make_case, default_params
from mohid_b200.advection_diffusion import TransportStep
from helpers import oracle_for, rel_err, water_mask

for (I, J, K, n, m, bc) in [(40, 21, 6, 3, 4, 4), (33, 12, 5, 2, 1, 1), (64, 10, 9, 11, 4, 7)]:
    case = make_case(I, J, K, nprop=n, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    prm = [default_params(m, 4, m, 4, bc=bc, decay_time=900.0) for _ in range(n)]
    ts = TransportStep(I, J, K, case.ld)
    ts.set_grid2d(**g); ts.set_step(s)
    gpu = [p.copy() for p in props]
    cpu = [p.copy() for p in props]
    for _ in range(2):
        ts.advect_batch(gpu, prm, refs)
        o.advect_batch(cpu, prm, refs)
    w = water_mask(s)
    print(I, J, K, n, m, bc, "max rel err", max(rel_err(a, b, w) for a, b in zip(gpu, cpu)), flush=True)
    ts.close()
