import sys, numpy as np
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
from mohid_b200.synthetic import make_case, default_params
from helpers import oracle_for, rel_err, water_mask
case = make_case(52, 37, 9, nprop=2, stepped_bottom=True)
o, g, s, props, refs = oracle_for(case)
shape = s["OpenPoints3D"].shape
rng = np.random.default_rng(5)
nf = [np.ascontiguousarray((rng.random(shape) < 0.15).astype(np.int32)) for _ in range(3)]
zero = np.zeros(shape, np.int32)
o.set_noflux(zero, nf[1], zero)
P = lambda a: dict(default_params(1, 4, 1, 4), NoAdvFlux=a)
def single(par, arr):
    x = [arr.copy()]; o.advect_batch(x, [par]); return x[0]
def second(par0, par1, arr0, arr1):
    x = [arr0.copy(), arr1.copy()]; o.advect_batch(x, [par0, par1]); return x
s1 = single(P(0), props[1]); s1f = single(P(1), props[1])
print("single flagged vs unflagged (V only, expect 0):", np.abs(s1 - s1f).max())
for a0 in (0, 1):
    for a1 in (0, 1):
        x = second(P(a0), P(a1), props[0], props[1])
        print("batch flags", a0, a1, "prop1 vs single:", np.abs(x[1] - (s1f if a1 else s1)).max())
x = second(P(1), P(0), props[1], props[1]); print("same array twice:", np.abs(x[0]-x[1]).max())
