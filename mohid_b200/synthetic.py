"""Synthetic sigma-grid cases for the MOHID property transport step.

Builds every input array of ``ModuleAdvectionDiffusion::AdvectionDiffusion`` with MOHID's
shapes and mask semantics:

* every 3-D array is dimensioned ``(0:I+1, 0:J+1, 0:K+1)`` column-major with ``i`` contiguous
  (reference: Geometry/Map allocation, ``MOHIDBase2/ModuleGeometry.F90:1624-1626``); here a
  torch tensor of shape ``(K+2, J+2, ld)`` whose C layout *is* that Fortran layout;
* metrics follow ``ModuleHorizontalGrid.F90:6283-6345`` (DZX/DZY) and
  ``ModuleGeometry.F90:3838-4282`` (DWZ, DZZ, VolumeZ, AreaU/V);
* masks follow ``ModuleMap.F90:985-1003`` (ComputeFacesU/V), ``:1611-1616`` (ComputeFacesW),
  ``:1787-1795`` (OpenPoints3D), ``:355-378`` (Water/LandPoints3D) and
  ``ModuleHorizontalMap.F90:939-942`` (no compute face between two boundary points).

Fluxes come from a discrete stream function (non-divergent per layer) plus a divergent
potential part; ``Wflux_Z`` closes continuity column by column, so a constant tracer is
preserved exactly in exact arithmetic.  Noise is a splitmix64 hash of the linear index, so
any sub-box of a case can be regenerated identically.

The generator runs on any torch device (CPU for the tests, CUDA for the full-size bench).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

TWO_PI = 2.0 * math.pi


def _splitmix64_uniform(idx: torch.Tensor, seed: int) -> torch.Tensor:
    """uniform [0,1) from a splitmix64 hash of (seed, idx); int64 arithmetic wraps mod 2**64."""
    def s64(v: int) -> int:                      # python int -> signed 64-bit constant
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v

    def lsr(x: torch.Tensor, n: int) -> torch.Tensor:   # logical shift right on int64
        return (x >> n) & ((1 << (64 - n)) - 1)

    z = idx.to(torch.int64) + s64(seed * 0x9E3779B97F4A7C15 + 0x9E3779B97F4A7C15)
    z = (z ^ lsr(z, 30)) * s64(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * s64(0x94D049BB133111EB)
    z = z ^ lsr(z, 31)
    return lsr(z, 11).to(torch.float64) * (1.0 / 9007199254740992.0)


@dataclass
class Case:
    """One synthetic transport case.  All tensors live on ``device``.

    3-D fields have shape ``(K+2, J+2, ld)``, 2-D fields ``(J+2, ld)``; element ``(i,j,k)`` of
    the Fortran array is ``t[k, j, i]``.
    """
    I: int
    J: int
    K: int
    ld: int
    dt: float
    grid2d: Dict[str, torch.Tensor] = field(default_factory=dict)   # DUX DVY DZX DZY KFloorZ BoundaryPoints2D
    step: Dict[str, torch.Tensor] = field(default_factory=dict)     # the 17 per-step 3-D inputs
    props: List[torch.Tensor] = field(default_factory=list)
    refs: List[torch.Tensor] = field(default_factory=list)
    j_offset: int = 0           # global j of local j=0 (for slabs of a decomposed domain)
    J_global: int = 0

    @property
    def shape3(self):
        return (self.K + 2, self.J + 2, self.ld)

    @property
    def cells(self) -> int:
        return self.I * self.J * self.K


GRID2D_ORDER = ["DUX", "DVY", "DZX", "DZY", "KFloorZ", "BoundaryPoints2D"]
STEP_ORDER = ["Wflux_X", "Wflux_Y", "Wflux_Z", "VolumeZOld", "VolumeZ", "Visc_H", "Diff_V", "DWZ", "DZZ",
              "AreaU", "AreaV", "OpenPoints3D", "LandPoints3D", "WaterPoints3D",
              "ComputeFacesU3D", "ComputeFacesV3D", "ComputeFacesW3D"]


def make_case(I: int, J: int, K: int, nprop: int = 1, *, dt: float = 30.0, seed: int = 20260101,
              ld: Optional[int] = None, device: str = "cpu", islands: bool = True,
              stepped_bottom: bool = False, courant_h: float = 0.4, courant_v: float = 0.3,
              volume_change: float = 2.0e-3, dx: float = 500.0, dy: float = 500.0,
              depth: float = 50.0, make_refs: bool = True, closed: bool = False,
              j_range: Optional[Tuple[int, int]] = None) -> Case:
    """Build a case on the global grid ``I x J x K``.

    ``j_range=(lo, hi)`` returns only the j-slab whose work columns are the global columns
    ``lo..hi`` (local ``J = hi-lo+1``, local column ``j`` is global ``lo-1+j``): every formula uses
    global coordinates and hashes, so a slab is bit-identical to the same columns of the global case.
    """
    dev = torch.device(device)
    f64 = dict(dtype=torch.float64, device=dev)
    lo, hi = (1, J) if j_range is None else (int(j_range[0]), int(j_range[1]))
    assert 1 <= lo <= hi <= J
    Jg, Jl = J, hi - lo + 1
    ni, nj, nk = I + 2, Jl + 2, K + 2
    ld = ni if ld is None else int(ld)
    assert ld >= ni
    c = Case(I=I, J=Jl, K=K, ld=ld, dt=float(dt), J_global=Jg, j_offset=lo - 1)

    # ---- 2-D stage on a j-range extended by a margin (the coastal taper below has an 8-cell reach) ----
    MARGIN = 12
    e_lo, e_hi = max(0, lo - 1 - MARGIN), min(Jg + 1, hi + 1 + MARGIN)      # global columns generated
    nje = e_hi - e_lo + 1
    crop = slice(lo - 1 - e_lo, lo - 1 - e_lo + nj)
    ii = torch.arange(ld, **f64).view(1, ld)
    jj = (torch.arange(nje, **f64) + e_lo).view(nje, 1)                     # GLOBAL j
    ji = (torch.arange(nje, device=dev) + e_lo).view(nje, 1)
    iw = torch.arange(ld, device=dev).view(1, ld)
    valid_i = (iw < ni)

    # horizontal metrics (HG:6283-6345); DZX/DZY use the next cell, replicated on the last one
    def dux_of(j):
        return dx * (1.0 + 0.1 * torch.sin(TWO_PI * j / Jg))

    def dvy_of(i):
        return dy * (1.0 + 0.1 * torch.cos(TWO_PI * i / I))
    DUX = dux_of(jj) + 0.0 * ii
    DVY = dvy_of(ii) + 0.0 * jj
    DZX = torch.where(jj < Jg + 1, 0.5 * (dux_of(jj) + dux_of(jj + 1.0)), dux_of(jj)) + 0.0 * ii
    DZY = torch.where(ii < I + 1, 0.5 * (dvy_of(ii) + dvy_of(ii + 1.0)), dvy_of(ii)) + 0.0 * jj

    # 2-D water mask, islands, boundary ring
    water2d = (ji >= 1) & (ji <= Jg) & (iw >= 1) & (iw <= I)
    if islands and I >= 16 and Jg >= 16:
        def rect(i0, i1, j0, j1):
            return (iw >= int(i0 * I)) & (iw <= int(i1 * I)) & (ji >= int(j0 * Jg)) & (ji <= int(j1 * Jg))
        land = rect(0.20, 0.30, 0.25, 0.40) | rect(0.55, 0.70, 0.60, 0.70) | rect(0.75, 0.80, 0.15, 0.30)
        # one enclosed lake cell: water point that never becomes an open point (AD:4003-4006)
        lake = (iw == int(0.25 * I)) & (ji == int(0.32 * Jg))
        water2d = water2d & (~land | lake)
    bnd2d = water2d & ((ji == 1) | (ji == Jg) | (iw == 1) | (iw == I))
    if closed:                                   # closed basin: no open-boundary ring
        bnd2d = torch.zeros_like(bnd2d)

    # bottom level
    if stepped_bottom and K >= 4:
        bump = 0.5 * (1.0 + torch.sin(TWO_PI * 2.0 * ii / I) * torch.sin(TWO_PI * 1.5 * jj / Jg))
        kfloor = (1 + torch.floor(bump * (K // 3))).to(torch.int32)
        kfloor = torch.where(bnd2d, torch.ones_like(kfloor), kfloor)
    else:
        kfloor = torch.ones((nje, ld), dtype=torch.int32, device=dev)
    kfloor = torch.where(water2d, kfloor, torch.ones_like(kfloor))

    # stream function at cell corners (SW corner of cell (i,j)): 24 x 20-cell eddies, tapered to zero on
    # every corner that touches a non-water cell, so the discrete flow stays exactly non-divergent per
    # layer and tangent to the coasts (no column-integrated convergence next to islands).
    psi = torch.sin(TWO_PI * (ii - 0.5) / 24.0) * torch.sin(TWO_PI * (jj - 0.5) / 20.0)
    cm = torch.zeros((nje, ld), dtype=torch.bool, device=dev)
    cm[1:, 1:] = water2d[1:, 1:] & water2d[:-1, 1:] & water2d[1:, :-1] & water2d[:-1, :-1]
    taper = cm.to(torch.float64).view(1, 1, nje, ld)
    for _ in range(8):
        taper = torch.nn.functional.avg_pool2d(taper, 3, stride=1, padding=1) * cm
    psi = psi * taper.view(nje, ld)
    qx2 = torch.zeros((nje, ld), **f64)
    qx2[:, :-1] = psi[:, 1:] - psi[:, :-1]            # Qx(i,j) = psi(i+1,j) - psi(i,j)
    qy2 = torch.zeros((nje, ld), **f64)
    qy2[:-1, :] = -(psi[1:, :] - psi[:-1, :])         # Qy(i,j) = -(psi(i,j+1) - psi(i,j))
    # divergent part from a potential whose sign changes over depth (column mean ~ 0)
    phi = torch.sin(TWO_PI * ii / 32.0) * torch.sin(TWO_PI * jj / 32.0)
    dxp = torch.zeros((nje, ld), **f64)
    dxp[1:, :] = phi[1:, :] - phi[:-1, :]             # across U face j
    dyp = torch.zeros((nje, ld), **f64)
    dyp[:, 1:] = phi[:, 1:] - phi[:, :-1]             # across V face i
    h = depth * (1.0 + 0.4 * torch.sin(TWO_PI * ii / I) * torch.cos(TWO_PI * jj / Jg))
    dvol2 = volume_change * torch.sin(TWO_PI * (ii / I + jj / Jg))
    blobs = []
    for n in range(nprop):
        ci = I * (0.35 + 0.3 * ((n * 0.37) % 1.0))
        cj = Jg * (0.35 + 0.3 * ((n * 0.61) % 1.0))
        sig_i, sig_j = 0.12 * I + 2.0, 0.12 * Jg + 2.0
        blobs.append(torch.exp(-0.5 * (((ii - ci) / sig_i) ** 2 + ((jj - cj) / sig_j) ** 2))[crop].contiguous())

    # ---- crop the 2-D fields to the slab (+1 halo column each side) ----
    DUX, DVY, DZX, DZY = (t[crop].contiguous() for t in (DUX, DVY, DZX, DZY))
    water2d, bnd2d, kfloor = water2d[crop].contiguous(), bnd2d[crop].contiguous(), kfloor[crop].contiguous()
    qx2, qy2, dxp, dyp, h, dvol2 = (t[crop].contiguous() for t in (qx2, qy2, dxp, dyp, h, dvol2))
    ji = ji[crop]                                       # global j of the local columns
    del psi, phi, taper, cm
    c.grid2d = dict(DUX=DUX, DVY=DVY, DZX=DZX, DZY=DZY, KFloorZ=kfloor,
                    BoundaryPoints2D=bnd2d.to(torch.int32).contiguous())

    # ---------------- vertical geometry: uniform sigma (GEO:5047-5170) ----------------
    DWZ = (h / K).unsqueeze(0).expand(nk, nj, ld).contiguous()
    DZZ = torch.empty_like(DWZ)
    DZZ[:-1] = 0.5 * (DWZ[1:] + DWZ[:-1])
    DZZ[-1] = DWZ[-1]
    VolumeZ = DWZ * (DUX * DVY).unsqueeze(0)
    kidx = torch.arange(nk, device=dev).view(nk, 1, 1)
    lin = (kidx * (Jg + 2) + ji.view(1, nj, 1)) * ni + iw.view(1, 1, ld)      # global linear index (hash key)
    kk = torch.arange(nk, **f64).view(nk, 1, 1)
    VolumeZOld = VolumeZ * (1.0 + dvol2.unsqueeze(0))

    # ---------------- 3-D masks ----------------
    inside_k = (kidx >= 1) & (kidx <= K)
    water3 = water2d.unsqueeze(0) & inside_k & (kidx >= kfloor.unsqueeze(0))
    inwork = ((ji >= 1) & (ji <= Jg) & (iw >= 1) & (iw <= I)).unsqueeze(0) & inside_k
    land3 = inwork & ~water3
    # 2-D compute faces: both sides water and not between two boundary points (HM:939-942)
    cf2u = torch.zeros((nj, ld), dtype=torch.bool, device=dev)
    cf2u[1:, :] = water2d[1:, :] & water2d[:-1, :] & ~(bnd2d[1:, :] & bnd2d[:-1, :])
    cf2v = torch.zeros((nj, ld), dtype=torch.bool, device=dev)
    cf2v[:, 1:] = water2d[:, 1:] & water2d[:, :-1] & ~(bnd2d[:, 1:] & bnd2d[:, :-1])
    kfu = kfloor.clone()
    kfu[1:, :] = torch.maximum(kfloor[1:, :], kfloor[:-1, :])
    kfv = kfloor.clone()
    kfv[:, 1:] = torch.maximum(kfloor[:, 1:], kfloor[:, :-1])
    CFU = cf2u.unsqueeze(0) & inside_k & (kidx >= kfu.unsqueeze(0))
    CFV = cf2v.unsqueeze(0) & inside_k & (kidx >= kfv.unsqueeze(0))
    # W faces: k = KFloorZ+1..KUB in columns whose surface cell has a horizontal compute face (MAP:1611-1616)
    surf = torch.zeros((nj, ld), dtype=torch.bool, device=dev)
    surf[:-1, :-1] = CFU[K, :-1, :-1] | CFU[K, 1:, :-1] | CFV[K, :-1, :-1] | CFV[K, :-1, 1:]
    CFW = surf.unsqueeze(0) & (kidx >= (kfloor.unsqueeze(0) + 1)) & (kidx <= K) & water2d.unsqueeze(0)
    openp = torch.zeros(c.shape3, dtype=torch.bool, device=dev)
    openp[:-1, :-1, :-1] = (CFU[:-1, :-1, :-1] | CFU[:-1, 1:, :-1] | CFV[:-1, :-1, :-1] | CFV[:-1, :-1, 1:] |
                            CFW[:-1, :-1, :-1] | CFW[1:, :-1, :-1])
    openp &= inwork

    # ---------------- water fluxes ----------------
    # amplitudes from analytic bounds (identical on every slab): |d psi| <= 2 sin(pi/24) + taper slope,
    # smallest cell volume = 0.9 dx * 0.9 dy * 0.6 depth / K
    vmin = 0.81 * dx * dy * 0.6 * depth / K
    amp = courant_h * vmin / dt / 0.40 * 0.85
    # peak of the potential-driven vertical flux: sum_k cos(.) <= K/pi, horizontal divergence <= 8 sin^2(pi/32)
    div_unit = 8.0 * math.sin(math.pi / 32.0) ** 2
    ampv = min(courant_v * vmin / dt / (div_unit * K / math.pi), 0.15 * courant_h * vmin / dt / (2.0 * math.sin(math.pi / 32.0)))
    g = (0.6 + 0.4 * kk / K)
    s = torch.cos(math.pi * (kk - 0.5) / K)
    Wflux_X = (qx2.unsqueeze(0) * (g * amp) + dxp.unsqueeze(0) * (s * ampv)) * CFU
    Wflux_Y = (qy2.unsqueeze(0) * (g * amp) + dyp.unsqueeze(0) * (s * ampv)) * CFV
    # continuity: Qz(k+1) = Qz(k) + Qx(j) - Qx(j+1) + Qy(i) - Qy(i+1) - (V - Vold)/dt   (AD:5718-5727 sign convention)
    colmask = (surf & ~bnd2d).unsqueeze(0) & water3        # interior columns with W faces
    conv = torch.zeros(c.shape3, **f64)
    conv[:, :-1, :-1] = (Wflux_X[:, :-1, :-1] - Wflux_X[:, 1:, :-1] + Wflux_Y[:, :-1, :-1] - Wflux_Y[:, :-1, 1:])
    conv = conv - (VolumeZ - VolumeZOld) / dt
    conv = conv * colmask
    Wflux_Z = torch.zeros(c.shape3, **f64)
    Wflux_Z[1:] = torch.cumsum(conv, dim=0)[:-1]            # Qz(k+1) = sum_{m<=k} conv(m)
    Wflux_Z = Wflux_Z * (colmask | torch.roll(colmask, 1, 0))
    Wflux_Z[0] = 0.0
    del conv, colmask

    # ---------------- turbulence and face areas ----------------
    Visc_H = 5.0 * (1.0 + 0.2 * _splitmix64_uniform(lin, seed + 101))
    Diff_V = 1.0e-3 * (1.0 + 0.5 * _splitmix64_uniform(lin, seed + 102))
    AreaU = torch.zeros(c.shape3, **f64)
    duxs = (DUX[1:, :] + DUX[:-1, :])
    AreaU[:, 1:, :] = ((DWZ[:, 1:, :] * DUX[:-1, :] + DWZ[:, :-1, :] * DUX[1:, :]) / duxs) * (0.5 * (DVY[1:, :] + DVY[:-1, :]))
    AreaV = torch.zeros(c.shape3, **f64)
    dvys = (DVY[:, 1:] + DVY[:, :-1])
    AreaV[:, :, 1:] = ((DWZ[:, :, 1:] * DVY[:, :-1] + DWZ[:, :, :-1] * DVY[:, 1:]) / dvys) * (0.5 * (DUX[:, 1:] + DUX[:, :-1]))

    pad = (~valid_i).view(1, 1, ld)

    def fin(t, dtype=None):
        t = t.to(dtype) if dtype is not None else t
        t = t.contiguous()
        if ld > ni:
            t = t.masked_fill(pad.expand_as(t), 0)
        return t

    c.step = dict(
        Wflux_X=fin(Wflux_X), Wflux_Y=fin(Wflux_Y), Wflux_Z=fin(Wflux_Z),
        VolumeZOld=fin(VolumeZOld), VolumeZ=fin(VolumeZ), Visc_H=fin(Visc_H), Diff_V=fin(Diff_V),
        DWZ=fin(DWZ), DZZ=fin(DZZ), AreaU=fin(AreaU), AreaV=fin(AreaV),
        OpenPoints3D=fin(openp, torch.int32), LandPoints3D=fin(land3, torch.int32),
        WaterPoints3D=fin(water3, torch.int32), ComputeFacesU3D=fin(CFU, torch.int32),
        ComputeFacesV3D=fin(CFV, torch.int32), ComputeFacesW3D=fin(CFW, torch.int32))
    del Wflux_X, Wflux_Y, Wflux_Z, VolumeZOld, Visc_H, Diff_V, AreaU, AreaV, openp, CFU, CFV, CFW
    if ld > ni:                                   # padded volumes / thicknesses stay finite divisors
        for name in ("VolumeZ", "VolumeZOld", "DWZ", "DZZ"):
            c.step[name].masked_fill_(pad.expand_as(c.step[name]), 1.0)

    # ---------------- tracers ----------------
    ranges = [(10.0, 20.0), (30.0, 36.0)]
    for n in range(nprop):
        lo_v, hi_v = ranges[n] if n < len(ranges) else (0.0, 1.0)
        vert = 0.5 + 0.5 * kk / K
        p = lo_v + (hi_v - lo_v) * (0.15 + 0.7 * blobs[n].unsqueeze(0) * vert +
                                    0.01 * _splitmix64_uniform(lin, seed + 1000 + n))
        p = torch.where(land3, torch.full_like(p, -9.9e15), p)      # land carries null_real (AD:1753)
        p = torch.where(inwork, p, torch.zeros_like(p))              # halos 0
        c.props.append(fin(p))
        if make_refs:
            r = lo_v + (hi_v - lo_v) * (0.5 + 0.1 * torch.sin(TWO_PI * kk / K)) + 0.0 * blobs[n].unsqueeze(0)
            c.refs.append(fin(r.expand(c.shape3)))
    return c


def default_params(method_h: int = 4, limiter_h: int = 4, method_v: int = 4, limiter_v: int = 4, *,
                   dt: float = 30.0, bc: int = 0, theta_difv: float = 1.0, impexp_advv: float = 1.0,
                   decay_time: float = 0.0, schmidt_h: float = 1.0):
    """Keyword defaults of the transport block (WP:9226-9733)."""
    return dict(Schmidt_H=schmidt_h, SchmidtCoef_V=1.0, SchmidtBackground_V=1.0e-8,
                AdvMethodH=method_h, TVDLimitationH=limiter_h, AdvMethodV=method_v, TVDLimitationV=limiter_v,
                Upwind2H=1, Upwind2V=1, VolumeRelMax=1.5, DTProp=dt,
                ImpExp_AdvV=impexp_advv, ImpExp_DifV=theta_difv, ImpExp_AdvXX=0.0, ImpExp_AdvYY=0.0,
                ImpExp_DifH=0.0, NullDif=0, BoundaryCondition=bc, DecayTime=decay_time,
                NoAdvFlux=0, NoDifFlux=0, CellFluxes=0)


def case_pieces(I: int, J: int, K: int, nprop: int, *, j_lo: int = 1, j_hi: Optional[int] = None, piece: int = 256,
                device: str = "cpu", make_refs: bool = False, **kw):
    """The columns ``j_lo-1 .. j_hi+1`` of a case (work columns ``j_lo..j_hi`` plus one halo column each side, as
    ``make_case(j_range=(j_lo, j_hi))`` returns them) in pieces of at most ``piece`` columns, so that a case that fills
    most of a GPU never exists twice.  Yields ``(j0, case)``: ``case`` holds the local columns ``j0 .. j0+ncols-1``
    (0-based inside the ``j_lo-1 .. j_hi+1`` window) in tensors of shape ``(K+2, ncols, ld)`` / ``(ncols, ld)``; every
    column is generated with two neighbour columns around it, which makes it identical to the same column of the
    undivided case.
    """
    j_hi = J if j_hi is None else j_hi
    first, last = j_lo - 1, j_hi + 1                      # global columns wanted (0 and J+1 are the domain halo)
    g = first
    while g <= last:
        e = min(g + piece - 1, last)
        lo, hi = max(1, g - 2), min(J, e + 2)             # work columns generated (margin 2)
        c = make_case(I, J, K, nprop, device=device, make_refs=make_refs, j_range=(lo, hi), **kw)
        a, b = g - (lo - 1), e - (lo - 1)                 # local columns of g .. e inside that piece
        cut3 = lambda t: t[:, a:b + 1, :].contiguous()
        cut2 = lambda t: t[a:b + 1, :].contiguous()
        out = Case(I=I, J=e - g + 1, K=K, ld=c.ld, dt=c.dt, J_global=J, j_offset=g)
        out.grid2d = {k: cut2(v) for k, v in c.grid2d.items()}
        out.step = {k: cut3(v) for k, v in c.step.items()}
        out.props = [cut3(p) for p in c.props]
        out.refs = [cut3(r) for r in c.refs]
        del c
        yield g - first, out
        g = e + 1
