import sys, numpy as np
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
from mohid_b200.synthetic import make_case, default_params
from helpers import oracle_for, rel_err, water_mask
case = make_case(52, 37, 9, nprop=2, stepped_bottom=True)
o, g, s, props, refs = oracle_for(case)
shape = s["OpenPoints3D"].shape
rng = np.random.default_rng(5)
nf = [np.ascontiguousarray((rng.random(shape) < 0.15).astype(np.int32)) for _ in range(3)]
prm = [default_params(1, 4, 1, 4, impexp_advv=0.0, theta_difv=0.5), default_params(1, 4, 1, 4)]
prm[0]["NoAdvFlux"] = 1
o.set_noflux(*nf)
b = [p.copy() for p in props]
o.advect_batch(b, prm)
c = [props[1].copy()]
o.advect_batch(c, prm[1:])
print("oracle batch vs single prop1:", np.abs(b[1] - c[0]).max())
o2, *_ = oracle_for(case)
d = [props[1].copy()]
o2.advect_batch(d, prm[1:])
print("fresh single vs batch:", np.abs(b[1] - d[0]).max(), " fresh single vs later single:", np.abs(c[0] - d[0]).max())
e = [p.copy() for p in props]
o2.set_noflux(*nf)
o2.advect_batch(e, prm)
print("o2 batch vs o batch:", np.abs(e[1] - b[1]).max(), np.abs(e[0]-b[0]).max())
prm2 = [dict(prm[0]), dict(prm[1])]; prm2[0]["ImpExp_AdvV"] = 1.0
f = [p.copy() for p in props]; o2.advect_batch(f, prm2); print("prop0 implicit: batch prop1 vs fresh", np.abs(f[1]-d[0]).max())
zero = np.zeros(shape, np.int32)
for name, arrs in (("U", [nf[0], zero, zero]), ("V", [zero, nf[1], zero]), ("W", [zero, zero, nf[2]]), ("none", [zero]*3)):
    o2.set_noflux(*arrs)
    f = [p.copy() for p in props]; o2.advect_batch(f, prm)
    dd = np.abs(f[1]-d[0]); print(name, dd.max(), np.unravel_index(dd.argmax(), dd.shape), (dd>1e-9).sum())
