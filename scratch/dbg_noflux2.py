import sys, numpy as np
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
from mohid_b200.synthetic import make_case, default_params
from helpers import oracle_for, rel_err, water_mask
from mohid_b200.advection_diffusion import TransportStep
case = make_case(52, 37, 9, nprop=2, stepped_bottom=True)
o, g, s, props, refs = oracle_for(case)
shape = s["OpenPoints3D"].shape
rng = np.random.default_rng(5)
nf = [np.ascontiguousarray((rng.random(shape) < 0.15).astype(np.int32)) for _ in range(3)]
w = water_mask(s)
for flags in ((1, 0), (0, 0)):
  for use_nf in (True, False):
    prm = [default_params(1, 4, 1, 4, impexp_advv=0.0, theta_difv=0.5), default_params(1, 4, 1, 4)]
    prm[0]["NoAdvFlux"], prm[0]["NoDifFlux"] = flags
    ts = TransportStep(case.I, case.J, case.K, case.ld); ts.set_grid2d(**g); ts.set_step(s)
    if use_nf:
        ts.set_noflux(*nf); o.set_noflux(*nf)
    else:
        if flags != (0, 0): continue
        o.set_noflux(None, None, None)
    a, b = [p.copy() for p in props], [p.copy() for p in props]
    for it in range(2):
        ts.advect_batch(a, prm); o.advect_batch(b, prm)
        for n in range(2):
            d = np.abs(a[n] - b[n]) * w
            print(flags, use_nf, "iter", it, "prop", n, "err", d.max(), "at", np.unravel_index(d.argmax(), d.shape), "nbad", int((d > 1e-9).sum()),
                  "nonwater equal", np.array_equal(a[n][~w], b[n][~w]))
    ts.close()
