import sys, numpy as np
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
from mohid_b200.synthetic import make_case, default_params
from helpers import oracle_for, rel_err, water_mask
from mohid_b200.advection_diffusion import TransportStep
case = make_case(52, 37, 9, nprop=2, stepped_bottom=True)
o, g, s, props, refs = oracle_for(case)
shape = s["OpenPoints3D"].shape
rng = np.random.default_rng(5)
full = [np.ascontiguousarray((rng.random(shape) < 0.15).astype(np.int32)) for _ in range(3)]
zero = np.zeros(shape, np.int32)
for name, nf in (("U", [full[0], zero, zero]), ("V", [zero, full[1], zero]), ("W", [zero, zero, full[2]]), ("none", [zero, zero, zero])):
    for mh, advv in ((1, 0.0), (1, 1.0), (4, 1.0)):
        for noadv, nodif in ((1, 0), (0, 1)):
            prm = [default_params(mh, 4, mh, 4, impexp_advv=advv, theta_difv=0.5)]
            prm[0]["NoAdvFlux"], prm[0]["NoDifFlux"] = noadv, nodif
            ts = TransportStep(case.I, case.J, case.K, case.ld); ts.set_grid2d(**g); ts.set_step(s)
            ts.set_noflux(*nf); o.set_noflux(*nf)
            a, b = [props[0].copy()], [props[0].copy()]
            ts.advect_batch(a, prm); o.advect_batch(b, prm)
            w = water_mask(s)
            d = np.abs(a[0] - b[0]) * w
            loc = np.unravel_index(d.argmax(), d.shape)
            print(name, mh, advv, noadv, nodif, "err", d.max(), "at k,j,i", loc, "nbad", int((d > 1e-9).sum()))
            ts.close()
