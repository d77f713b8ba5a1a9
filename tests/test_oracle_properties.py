"""Self-consistency of the CPU oracle: the reference ships no golden
vectors for this path, so the restatement is pinned by properties the scheme must have."""
import numpy as np
import pytest

from mohid_b200.synthetic import make_case, default_params
from helpers import oracle_for, NULL_REAL, water_mask

CONFIGS = [  # (method_h, lim_h, method_v, lim_v, impexp_advv)
    (1, 4, 1, 4, 1.0), (2, 4, 1, 4, 1.0), (3, 4, 3, 4, 0.0), (4, 1, 4, 1, 1.0), (4, 2, 4, 2, 1.0),
    (4, 3, 4, 3, 0.0), (4, 4, 4, 4, 1.0), (4, 5, 4, 5, 1.0), (5, 4, 5, 4, 1.0),
]


def const_field(s, value):
    p = np.where(s["LandPoints3D"] == 1, NULL_REAL, value)
    p[0] = p[-1] = 0
    p[:, 0] = p[:, -1] = 0
    p[:, :, 0] = p[:, :, -1] = 0
    return np.ascontiguousarray(p)


@pytest.mark.parametrize("cfg", CONFIGS)
def test_constant_field_preserved(oracle_lib, cfg):
    """Continuity-consistent fluxes + MassConservation boundary with Ref = const keep a constant."""
    mh, lh, mv, lv, adv_v = cfg
    case = make_case(40, 36, 8, nprop=1, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    p = const_field(s, 7.0)
    ref = np.full_like(p, 7.0)
    for step in range(3):
        o.now += 30.0
        o.advection_diffusion(p, default_params(mh, lh, mv, lv, bc=1, impexp_advv=adv_v), ref)
    w = water_mask(s)
    assert np.abs(p[w] - 7.0).max() < 1e-12
    assert np.all(p[s["LandPoints3D"] == 1] == NULL_REAL)          # land: exactly null_real (AD:1753)
    assert np.all(p[-1] == 0.0)                                    # PROP(:,:,KUB+1) = G(KUB+1) = 0


@pytest.mark.parametrize("cfg", [(1, 4, 1, 4, 1.0), (4, 4, 4, 4, 1.0), (4, 4, 4, 4, 0.0), (2, 4, 1, 4, 1.0)])
@pytest.mark.parametrize("theta", [1.0, 0.5])
def test_mass_conserved_in_closed_basin(oracle_lib, cfg, theta):
    mh, lh, mv, lv, adv_v = cfg
    case = make_case(36, 32, 10, nprop=1, closed=True, volume_change=0.0)
    o, g, s, props, refs = oracle_for(case)
    p = props[0]
    w = s["OpenPoints3D"] == 1
    m0 = float((p[w] * s["VolumeZ"][w]).sum())
    for step in range(5):
        o.now += 30.0
        o.advection_diffusion(p, default_params(mh, lh, mv, lv, impexp_advv=adv_v, theta_difv=theta))
    m1 = float((p[w] * s["VolumeZ"][w]).sum())
    assert abs(m1 - m0) / abs(m0) < 1e-12


def test_upwind_is_monotone(oracle_lib):
    case = make_case(36, 32, 10, nprop=1, closed=True, volume_change=0.0)
    o, g, s, props, refs = oracle_for(case)
    p = props[0]
    w = s["OpenPoints3D"] == 1
    lo, hi = p[w].min(), p[w].max()
    for step in range(10):
        o.now += 30.0
        o.advection_diffusion(p, default_params(1, 4, 1, 4))
        assert p[w].min() >= lo - 1e-12 and p[w].max() <= hi + 1e-12


def test_optimize_path_matches_plain_path(oracle_lib):
    """The Optimize code path (AD:5241-5245, MF:10642) only changes rounding order (quirk A.4-5)."""
    case = make_case(40, 36, 12, nprop=3, stepped_bottom=True)
    o1, g, s, props1, refs = oracle_for(case)
    o2, _, _, props2, _ = oracle_for(case)
    params = [default_params(4, 4, 4, 4, bc=4) for _ in range(3)]
    for step in range(5):
        o1.advect_batch(props1, params, refs, force_optimize=0)
        o2.advect_batch(props2, params, refs, force_optimize=1)
    w = water_mask(s)
    for a, b in zip(props1, props2):
        assert np.abs(a[w] - b[w]).max() / np.abs(a[w]).max() < 1e-12
        assert np.array_equal(a == NULL_REAL, b == NULL_REAL)


def test_batch_decides_optimize_like_the_caller(oracle_lib):
    """WP:14580-14598: >= 2 properties, all P2_TVD + SuperBee, equal Schmidt_H -> Optimize."""
    case = make_case(24, 20, 6, nprop=2)
    o1, g, s, p1, refs = oracle_for(case)
    o2, _, _, p2, _ = oracle_for(case)
    params = [default_params(4, 4, 4, 4) for _ in range(2)]
    o1.advect_batch(p1, params)                       # decides Optimize = True
    o2.advect_batch(p2, params, force_optimize=1)
    assert all(np.array_equal(a, b) for a, b in zip(p1, p2))


def test_dry_columns_and_closed_cells_untouched(oracle_lib):
    case = make_case(40, 36, 8, nprop=1)
    o, g, s, props, refs = oracle_for(case)
    p = props[0]
    p0 = p.copy()
    o.now += 30.0
    o.advection_diffusion(p, default_params(4, 4, 4, 4))
    water_col = s["WaterPoints3D"][case.K] == 1               # (nj, ld)
    # columns with WaterPoints(i,j,KUB) /= 1 are not touched by the solver (MF:4086)
    assert np.array_equal(p[:, ~water_col], p0[:, ~water_col])
    # closed water cells (the lake cell) keep their concentration (AD:4003-4006)
    closed = (s["WaterPoints3D"] == 1) & (s["OpenPoints3D"] == 0)
    assert closed.any()
    assert np.array_equal(p[closed], p0[closed])


def test_reference_stop_conditions(oracle_lib):
    case = make_case(16, 16, 4, nprop=1)
    o, g, s, props, refs = oracle_for(case)
    with pytest.raises(RuntimeError, match="ERR200"):          # implicit vertical + QUICK (AD:1234-1237)
        o.advection_diffusion(props[0], default_params(1, 4, 2, 4, impexp_advv=1.0))
    bad = default_params(1, 4, 1, 4)
    bad["ImpExp_DifH"] = 1.0
    with pytest.raises(RuntimeError, match="ERR02"):           # AD:1340-1343
        o.advection_diffusion(props[0], bad)
    bad = default_params(1, 4, 1, 4)
    bad["ImpExp_AdvV"] = 0.5
    with pytest.raises(RuntimeError, match="VerticalAdvection"):   # AD:3124
        o.advection_diffusion(props[0], bad)


def test_openmp_threads_do_not_change_results(oracle_lib):
    case = make_case(40, 36, 8, nprop=2)
    o1, g, s, p1, refs = oracle_for(case, nthreads=1)
    o4, _, _, p4, _ = oracle_for(case, nthreads=4)
    params = [default_params(4, 4, 4, 4, bc=7) for _ in range(2)]
    for _ in range(3):
        o1.advect_batch(p1, params, refs)
        o4.advect_batch(p4, params, refs)
    assert all(np.array_equal(a, b) for a, b in zip(p1, p4))


@pytest.mark.parametrize("optimize", [False, True])
def test_cell_fluxes_close_the_cell_budget(oracle_lib, optimize):
    """CellFluxes (AD:3356-3954): V (Pnew - Pold Vold/V) / dt equals the net face flux of every open cell."""
    case = make_case(30, 26, 6, nprop=1, closed=True)
    o, g, s, props, refs = oracle_for(case)
    p = props[0]
    p0 = p.copy()
    prm = default_params(4, 4, 4, 4)
    prm["CellFluxes"] = 1
    o.now += 30.0
    o.advection_diffusion(p, prm, optimize=optimize, first_property=True)
    fl = o.get_cell_fluxes()
    K, J, I = case.K, case.J, case.I
    tot = {d: fl["AdvFlux" + d] + fl["DifFlux" + d] for d in "XYZ"}
    c = (slice(1, K + 1), slice(1, J + 1), slice(1, I + 1))
    net = (tot["X"][c] - tot["X"][1:K + 1, 2:J + 2, 1:I + 1] + tot["Y"][c] - tot["Y"][1:K + 1, 1:J + 1, 2:I + 2] +
           tot["Z"][c] - tot["Z"][2:K + 2, 1:J + 1, 1:I + 1])
    V = s["VolumeZ"]
    lhs = (V * (p - p0 * s["VolumeZOld"] / V) / 30.0)[c]
    w = s["OpenPoints3D"][c] == 1
    w[K - 1] = False
    assert np.abs(lhs - net)[w].max() / np.abs(tot["X"]).max() < 1e-12


@pytest.mark.parametrize("direction", ["XX", "YY"])
def test_cell_fluxes_close_the_budget_of_a_2d_implicit_step(oracle_lib, direction):
    """K = 1, horizontally implicit: one system, so the implicit direction's flux with the final field (AD:1895-1899) closes
    the cell budget exactly like the explicit ones."""
    case = make_case(30, 26, 1, nprop=1, closed=True)
    o, g, s, props, refs = oracle_for(case)
    p = props[0]
    p0 = p.copy()
    prm = dict(default_params(4, 4, 4, 4), CellFluxes=1, **{"ImpExp_Adv" + direction: 1.0})
    o.now += 30.0
    o.advection_diffusion(p, prm, optimize=False, first_property=True)
    fl = o.get_cell_fluxes()
    J, I = case.J, case.I
    tot = {d: fl["AdvFlux" + d] + fl["DifFlux" + d] for d in "XY"}
    c = (slice(1, 2), slice(1, J + 1), slice(1, I + 1))
    net = tot["X"][c] - tot["X"][1:2, 2:J + 2, 1:I + 1] + tot["Y"][c] - tot["Y"][1:2, 1:J + 1, 2:I + 2]
    V = s["VolumeZ"]
    # the only layer is the surface layer: its row carries the water flux through the top face implicitly (AD:3990-4005)
    lhs = (V * (p - p0 * s["VolumeZOld"] / V) / 30.0 + s["Wflux_Z"][2:3] * p)[c]
    w = s["OpenPoints3D"][c] == 1
    assert not np.array_equal(p, p0)
    assert np.abs(lhs - net)[w].max() / np.abs(tot["X"]).max() < 1e-12


def test_noflux_cells_block_fluxes_and_carry_over(oracle_lib):
    """NoAdvFlux / NoDifFlux (AD:4434-4443, 4804-4813, 3004-3013, 2497-2501): with every face listed and closed
    boundaries a flagged property only feels the volume change; NoFluxV alone changes nothing for the flagged
    property itself (its zeroing hits the XX arrays after their use) but is inherited by the next upwind property."""
    case = make_case(30, 26, 6, nprop=2, closed=True, volume_change=False)
    o, g, s, props, refs = oracle_for(case)
    ones = np.ones(s["OpenPoints3D"].shape, np.int32)
    zero = np.zeros_like(ones)
    up = lambda **kw: dict(default_params(1, 4, 1, 4), **kw)
    w = water_mask(s)
    o.set_noflux(ones, ones, ones)
    a = [props[0].copy()]
    o.advect_batch(a, [up(NoAdvFlux=1, NoDifFlux=1)])
    assert np.abs(a[0] - props[0])[w].max() > 1e-3           # YY advection is never switched off (quirk A.4-7)
    s0 = dict(s); s0["Wflux_Y"] = np.zeros_like(s["Wflux_Y"])
    o.set_step(s0)
    a = [props[0].copy()]
    o.advect_batch(a, [up(NoAdvFlux=1, NoDifFlux=1)])
    o.set_step(s)
    want = props[0] * s["VolumeZOld"] / np.where(w, s["VolumeZ"], 1.0)       # VolumeVariation only (AD:3966-4021)
    assert np.abs(a[0] - want)[w].max() < 1e-12
    o.set_noflux(zero, ones, zero)
    b = [props[0].copy()]; c = [props[0].copy()]
    o.advect_batch(b, [up(NoAdvFlux=1)])
    o.advect_batch(c, [up()])
    assert np.array_equal(b[0], c[0])                         # NoFluxV never reaches the YY coefficients
    d = [props[0].copy(), props[0].copy()]
    o.advect_batch(d, [up(NoAdvFlux=1), up()])
    assert np.array_equal(d[0], c[0]) and not np.array_equal(d[1], c[0])   # ... but the next property inherits it


@pytest.mark.parametrize("mh", [1, 4, 5])
@pytest.mark.parametrize("direction", ["xx", "yy"])
def test_horizontally_implicit_conserves_mass(oracle_lib, mh, direction):
    """Direction splitting (AD:4132-4265, THOMAS_3D MF:3751-3875): closed basin, mass is conserved to round-off and
    the result stays close to the explicit one at the generator's small Courant numbers."""
    case = make_case(40, 36, 8, nprop=1, closed=True, volume_change=False)
    o, g, s, props, refs = oracle_for(case)
    w = water_mask(s)
    V = s["VolumeZ"]
    mv = 1 if mh == 5 else mh
    imp = dict(default_params(mh, 4, mv, 4), **({"ImpExp_AdvXX": 1.0} if direction == "xx" else {"ImpExp_AdvYY": 1.0}))
    exp = default_params(mh, 4, mv, 4)
    a, b = [props[0].copy()], [props[0].copy()]
    m0 = (a[0] * V)[w].sum()
    for _ in range(3):
        o.advect_batch(a, [imp])
        o.advect_batch(b, [exp])
    assert abs((a[0] * V)[w].sum() - m0) <= 1e-12 * abs(m0)
    assert np.abs(a[0] - b[0])[w].max() < 0.5 and not np.array_equal(a[0], b[0])    # fields span ~7 units
    assert np.array_equal(a[0] == NULL_REAL, props[0] == NULL_REAL)


@pytest.mark.parametrize("direction", ["xx", "yy"])
def test_horizontally_implicit_2d_domain_conserves_mass(oracle_lib, direction):
    """K = 1 (AD:1758-1841): the line solve is the whole step.  Closed basin: mass conserved to round-off, land stays
    land, the result stays close to the explicit one."""
    case = make_case(40, 36, 1, nprop=1, closed=True, volume_change=False)
    o, g, s, props, refs = oracle_for(case)
    w = water_mask(s)
    V = s["VolumeZ"]
    imp = dict(default_params(4, 4, 4, 4), **({"ImpExp_AdvXX": 1.0} if direction == "xx" else {"ImpExp_AdvYY": 1.0}))
    exp = default_params(4, 4, 4, 4)
    a, b = [props[0].copy()], [props[0].copy()]
    m0 = (a[0] * V)[w].sum()
    for _ in range(3):
        o.advect_batch(a, [imp])
        o.advect_batch(b, [exp])
    assert abs((a[0] * V)[w].sum() - m0) <= 1e-12 * abs(m0)
    assert np.abs(a[0] - b[0])[w].max() < 0.5 and not np.array_equal(a[0], b[0])
    assert np.array_equal(a[0] == NULL_REAL, props[0] == NULL_REAL)


def test_caller_premix_conserves_mass_and_flags_shallow_columns(oracle_lib):
    """FreeConvection / SmallDepthsMixing_Processes (WP:13017-13074, 12939-13012) replace part of a column by its
    volume-weighted mean: the column mass is unchanged, mixed cells are uniform, the ON flag marks thin open columns."""
    case = make_case(30, 24, 8, nprop=1, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    rng = np.random.default_rng(1)
    shape3 = s["OpenPoints3D"].shape
    density = np.ascontiguousarray(1025.0 + rng.standard_normal(shape3) * 0.05)
    wcol = np.ascontiguousarray(5.0 + 40.0 * rng.random(shape3[1:]))
    p = props[0].copy()
    V = s["VolumeZ"]
    K = case.K
    cols = (s["OpenPoints3D"][K] == 1)
    m0 = (p * V)[1:K + 1][:, cols].sum(axis=0)
    on = o.caller_premix(p, density, wcol, 15.0, 0.0)
    m1 = (p * V)[1:K + 1][:, cols].sum(axis=0)
    assert np.allclose(m0, m1, rtol=1e-13)
    assert not np.array_equal(p, props[0])
    want = (s["OpenPoints3D"][K] == 1) & (wcol < 15.0)
    want[0] = want[-1] = False; want[:, 0] = False; want[:, case.I + 1:] = False
    assert np.array_equal(on == 1, want)
    jj, ii = np.argwhere(on == 1)[0]
    kb = g["KFloorZ"][jj, ii]
    assert np.ptp(p[kb:K + 1, jj, ii]) == 0.0
    q = props[0].copy()
    o.caller_premix(q, offset=2.5)
    w = water_mask(s)
    work = np.zeros_like(w); work[1:K + 1, 1:case.J + 1, 1:case.I + 1] = True
    assert np.array_equal(q[w & work], props[0][w & work] + 2.5) and np.array_equal(q[~(w & work)], props[0][~(w & work)])


def test_orlanski_keeps_a_constant_and_writes_the_exterior_cells(oracle_lib):
    """BC 6 (AD:5504-5570, MF:4129-4490): with the exterior cells and the reference at the same constant the field
    stays constant; the exterior (halo) cells next to open boundary cells are rewritten by the routine."""
    case = make_case(36, 30, 6, nprop=1, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    p = np.where(s["LandPoints3D"] == 1, NULL_REAL, 7.0)
    p[0] = p[-1] = 0
    p = np.ascontiguousarray(p)
    ref = np.full_like(p, 7.0)
    a = [p.copy()]
    for _ in range(3):
        o.advect_batch(a, [default_params(4, 4, 4, 4, bc=6)], [ref])
    w = water_mask(s)
    assert np.abs(a[0][w] - 7.0).max() < 1e-12
    b = [props[0].copy()]
    o.advect_batch(b, [default_params(1, 4, 1, 4, bc=6)], [refs[0]])
    changed = (b[0] != props[0]) & ~w
    K, J, I = case.K, case.J, case.I
    ring = np.zeros_like(w); ring[1:K + 1, 0, :] = ring[1:K + 1, J + 1, :] = ring[1:K + 1, :, 0] = ring[1:K + 1, :, I + 1] = True
    assert changed.sum() > 0 and not (changed & ~ring).any()
