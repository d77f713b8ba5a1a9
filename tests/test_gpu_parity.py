"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): masks / indices bit-exact (land = null_real exactly, dry columns
and closed cells untouched, KUB+1 row zero), property fields within 1e-10 relative after 100 steps
(fp64).  Per step the reformulated arithmetic (shared reciprocals) must stay below 1e-12.
"""
import numpy as np
import pytest

from mohid_b200.synthetic import make_case, default_params
from helpers import oracle_for, rel_err, water_mask, NULL_REAL

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-12      # relative, one step
TOL_100 = 1e-10       # relative, 100 steps (north star)


def gpu_for(case, g, s, **kw):
    from mohid_b200.advection_diffusion import TransportStep
    ts = TransportStep(case.I, case.J, case.K, case.ld, **kw)
    ts.set_grid2d(**g)
    ts.set_step(s)
    return ts


def compare(gpu, cpu, s, tol):
    w = water_mask(s)
    worst = 0.0
    for a, b in zip(gpu, cpu):
        assert np.array_equal(a == NULL_REAL, b == NULL_REAL)            # land pattern bit-exact
        assert np.array_equal(a[~w], b[~w])                              # everything that is not water: bit-exact
        assert np.array_equal(a[-1], b[-1])                              # KUB+1 row
        worst = max(worst, rel_err(a, b, w))
    assert worst <= tol, worst
    return worst


CONFIGS = [  # (method_h, lim_h, method_v, lim_v, impexp_advv, theta)
    (1, 4, 1, 4, 1.0, 1.0), (1, 4, 1, 4, 0.0, 0.5), (2, 4, 1, 4, 1.0, 1.0), (3, 4, 3, 4, 0.0, 1.0),
    (2, 4, 2, 4, 0.0, 0.0), (4, 1, 4, 1, 1.0, 1.0), (4, 2, 4, 2, 1.0, 1.0), (4, 3, 4, 3, 0.0, 1.0),
    (4, 4, 4, 4, 1.0, 1.0), (4, 4, 4, 4, 0.0, 0.3), (4, 5, 4, 5, 1.0, 1.0), (5, 4, 5, 4, 1.0, 1.0),
    (6, 4, 1, 4, 1.0, 1.0),
]


@pytest.mark.parametrize("cfg", CONFIGS)
def test_single_property_all_schemes(oracle_lib, cfg):
    mh, lh, mv, lv, adv_v, theta = cfg
    case = make_case(70, 45, 9, nprop=1, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    ts = gpu_for(case, g, s)
    gpu, cpu = [props[0].copy()], [props[0].copy()]
    prm = [default_params(mh, lh, mv, lv, impexp_advv=adv_v, theta_difv=theta)]
    for _ in range(2):
        ts.advect_batch(gpu, prm)
        o.advect_batch(cpu, prm)
    compare(gpu, cpu, s, 2 * TOL_STEP)
    assert ts.counters()["zero_pivots"] == 0
    ts.close()


@pytest.mark.parametrize("bc", [0, 1, 2, 4, 5, 7, 8])
def test_batched_tvd_with_boundary_conditions(oracle_lib, bc):
    case = make_case(66, 40, 8, nprop=3)
    o, g, s, props, refs = oracle_for(case)
    ts = gpu_for(case, g, s)
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    prm = [default_params(4, 4, 4, 4, bc=bc, decay_time=900.0) for _ in range(3)]
    for _ in range(2):
        ts.advect_batch(gpu, prm, refs)
        o.advect_batch(cpu, prm, refs)       # Optimize path, decided like WP:14580-14598
    compare(gpu, cpu, s, 2 * TOL_STEP)
    ts.close()


def test_mixed_boundary_conditions_and_schmidt_groups(oracle_lib):
    """Coastal3D-like: T,S with NullGradient, a tracer with MassConservNullGrad and another Schmidt number."""
    case = make_case(50, 38, 7, nprop=3)
    o, g, s, props, refs = oracle_for(case)
    ts = gpu_for(case, g, s)
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    prm = [default_params(4, 4, 4, 4, bc=4), default_params(4, 4, 4, 4, bc=4),
           default_params(4, 4, 4, 4, bc=7, schmidt_h=0.7)]
    prm[2]["SchmidtCoef_V"] = 0.5
    ts.advect_batch(gpu, prm, refs)
    o.advect_batch(cpu, prm, refs)
    compare(gpu, cpu, s, TOL_STEP)
    ts.close()


def test_hundred_steps_within_1e10(oracle_lib):
    """North-star tolerance: <= 1e-10 relative after 100 steps on identical inputs."""
    case = make_case(64, 48, 10, nprop=2)
    o, g, s, props, refs = oracle_for(case)
    ts = gpu_for(case, g, s)
    prm = [default_params(4, 4, 4, 4, bc=1, decay_time=3600.0) for _ in range(2)]
    cpu = [p.copy() for p in props]
    ts.upload(props, refs)
    for _ in range(100):
        o.advect_batch(cpu, prm, refs)
    ts.advect_device(prm, nsteps=100)
    gpu = [np.empty_like(p) for p in props]
    ts.download(gpu)
    worst = compare(gpu, cpu, s, TOL_100)
    print("100-step max relative difference:", worst)
    ts.close()


def test_upwind_100_steps(oracle_lib):
    """Config C2 numerics (upwind + implicit vertical, 1 tracer)."""
    case = make_case(64, 64, 10, nprop=1)
    o, g, s, props, refs = oracle_for(case)
    ts = gpu_for(case, g, s)
    prm = [default_params(1, 4, 1, 4)]
    cpu = [props[0].copy()]
    ts.upload(props)
    for _ in range(100):
        o.advect_batch(cpu, prm)
    ts.advect_device(prm, nsteps=100)
    gpu = [np.empty_like(props[0])]
    ts.download(gpu)
    compare(gpu, cpu, s, TOL_100)
    ts.close()


def test_padded_leading_dimension_and_ragged_tiles(oracle_lib):
    """_PAD_MATRICES-style ld > I+2, and I not a multiple of the 31-cell strip."""
    for I, ld in [(31, 40), (32, 48), (61, 64), (95, 97)]:
        case = make_case(I, 20, 5, nprop=2, ld=ld)
        o, g, s, props, refs = oracle_for(case)
        ts = gpu_for(case, g, s)
        gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
        prm = [default_params(4, 4, 4, 4) for _ in range(2)]
        ts.advect_batch(gpu, prm)
        o.advect_batch(cpu, prm)
        compare(gpu, cpu, s, TOL_STEP)
        ts.close()


def test_two_dimensional_case(oracle_lib):
    """K = 1: no vertical processes, the column solve degenerates to a division by E."""
    case = make_case(40, 33, 1, nprop=1)
    o, g, s, props, refs = oracle_for(case)
    ts = gpu_for(case, g, s)
    gpu, cpu = [props[0].copy()], [props[0].copy()]
    prm = [default_params(4, 4, 1, 4)]
    ts.advect_batch(gpu, prm)
    o.advect_batch(cpu, prm)
    compare(gpu, cpu, s, TOL_STEP)
    ts.close()


def test_small_depths_and_flags(oracle_lib):
    case = make_case(40, 30, 6, nprop=1)
    g_small = None
    for kw in (dict(xzflow=True), dict(vertical1d=True), dict()):
        o, g, s, props, refs = oracle_for(case, **kw)
        ts = gpu_for(case, g, s, **kw)
        small = np.zeros_like(g["KFloorZ"])
        small[5:15, 5:20] = 1
        o.set_step(s, small)
        ts.set_step(s, small)
        gpu, cpu = [props[0].copy()], [props[0].copy()]
        prm = [default_params(1, 4, 1, 4)]
        ts.advect_batch(gpu, prm)
        o.advect_batch(cpu, prm)
        compare(gpu, cpu, s, TOL_STEP)
        ts.close()


def test_device_resident_equals_host_path(oracle_lib):
    case = make_case(45, 37, 6, nprop=2)
    o, g, s, props, refs = oracle_for(case)
    ts = gpu_for(case, g, s)
    prm = [default_params(4, 4, 4, 4, bc=4) for _ in range(2)]
    a = [p.copy() for p in props]
    for _ in range(3):
        ts.advect_batch(a, prm, refs)
    ts.upload(props, refs)
    ts.advect_device(prm, nsteps=3)
    b = [np.empty_like(p) for p in props]
    ts.download(b)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    ts.close()


def test_reference_stop_conditions_become_errors(oracle_lib):
    from mohid_b200.capi import AdtError
    case = make_case(16, 16, 4, nprop=1)
    o, g, s, props, refs = oracle_for(case)
    ts = gpu_for(case, g, s)
    p0 = props[0].copy()
    with pytest.raises(AdtError, match="ERR200"):
        ts.advect_batch([p0], [default_params(1, 4, 2, 4, impexp_advv=1.0)])
    bad = default_params(1, 4, 1, 4); bad["ImpExp_DifH"] = 1.0
    with pytest.raises(AdtError, match="ERR02"):
        ts.advect_batch([p0], [bad])
    bad = default_params(1, 4, 1, 4); bad["ImpExp_AdvV"] = 0.5
    with pytest.raises(AdtError, match="VerticalAdvection"):
        ts.advect_batch([p0], [bad])
    bad = default_params(1, 4, 1, 4); bad["ImpExp_AdvXX"] = 1.0; bad["ImpExp_AdvYY"] = 1.0
    with pytest.raises(AdtError, match="ERR03"):
        ts.advect_batch([p0], [bad])
    bad = default_params(2, 4, 1, 4); bad["ImpExp_AdvXX"] = 1.0
    with pytest.raises(AdtError, match="ERR100"):
        ts.advect_batch([p0], [bad])
    bad = default_params(1, 4, 1, 4, bc=3)
    with pytest.raises(AdtError, match="ERR01"):       # not a boundary condition of the reference (AD:5816-5830)
        ts.advect_batch([p0], [bad], [props[0].copy()])
    assert np.array_equal(p0, props[0])                # nothing was touched
    ts.close()


def test_torch_cuda_inputs(oracle_lib):
    """The same entry points accept device pointers (UVA): inputs generated on the GPU."""
    import torch
    from mohid_b200.advection_diffusion import TransportStep
    case = make_case(48, 40, 6, nprop=2, device="cuda")
    ts = TransportStep(case.I, case.J, case.K)
    ts.set_grid2d(**case.grid2d)
    ts.set_step(case.step)
    ts.upload(case.props)
    prm = [default_params(4, 4, 4, 4) for _ in range(2)]
    ts.advect_device(prm, nsteps=2)
    out = [torch.empty_like(p) for p in case.props]
    ts.download(out)
    torch.cuda.synchronize()
    cpu_case = make_case(48, 40, 6, nprop=2)
    o, g, s, props, refs = oracle_for(cpu_case)
    # the generator is device independent up to libm rounding of sin/cos: feed the oracle the GPU-made inputs
    from oracle.oracle import case_to_numpy
    g, s, props, refs = case_to_numpy(case)
    o.set_grid2d(g); o.set_step(s)
    cpu = [p.copy() for p in props]
    for _ in range(2):
        o.advect_batch(cpu, prm)
    compare([t.cpu().numpy() for t in out], cpu, s, 2 * TOL_STEP)
    ts.close()


def _discharge_set(case, s, conc, concmf):
    """Three discharges: a bottom point source, a uniform-over-the-column source (kmin/kmax = FillValueInt), an
    ignored one, and a withdrawal (negative flow) that by-passes nothing."""
    K = case.K
    FILL = -9999999
    water = np.argwhere((s["OpenPoints3D"][K] == 1) & (s["OpenPoints3D"][1] == 1))
    (j1, i1), (j2, i2), (j3, i3) = water[len(water) // 5], water[len(water) // 2], water[4 * len(water) // 5]
    return dict(DischFlow=[40.0, 25.0, -30.0], DischConc=conc, DischConcMF=concmf,
                DischI=[i1, i2, i3], DischJ=[j1, j2, j3], DischK=[1, K, 2], DischKmin=[FILL, FILL, FILL],
                DischKmax=[FILL, FILL, FILL], DischVert=[1, 5, 9, 1], IgnoreDisch=[0, 0, 1, 0],
                DischnCells=[1, 1, 7, 1], ByPass=[0, 1, 0, 0])


def test_discharges_per_property(oracle_lib):
    """SetDischarges with property-specific concentrations (WP:14761-14773, AD:4025-4128)."""
    case = make_case(50, 40, 7, nprop=2)
    o, g, s, props, refs = oracle_for(case)
    ts = gpu_for(case, g, s)
    prm = [default_params(1, 4, 1, 4), default_params(1, 4, 1, 4)]
    sets = [_discharge_set(case, s, [35.0, 2.0, 0.0], [1.0, 1.0, 0.5]),
            _discharge_set(case, s, [0.5, 7.5, 0.0], [0.0, 1.0, 1.0])]
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    for n in range(2):
        ts.set_discharges(n, sets[n])
    for _ in range(3):
        ts.advect_batch(gpu, prm)
        for n in range(2):                      # the reference's per-property call sequence
            o.set_discharges(sets[n])
            o.now += 30.0 if n == 0 else 0.0
            o.advection_diffusion(cpu[n], prm[n], optimize=False, first_property=(n == 0))
            o.unset_discharges()
    compare(gpu, cpu, s, 3 * TOL_STEP)
    # discharges really acted
    ref_o, _, _, p0, _ = oracle_for(case)
    ref_o.advect_batch(p0, prm)
    assert not np.array_equal(p0[0], cpu[0])
    ts.unset_discharges()
    ts.close()


@pytest.mark.parametrize("cfg", [(4, 4, 4, 4, 1.0, 1.0, 2), (4, 4, 4, 4, 0.0, 0.4, 1), (1, 4, 1, 4, 1.0, 1.0, 1),
                                 (2, 4, 1, 4, 1.0, 0.0, 1), (5, 4, 5, 4, 1.0, 1.0, 1)])
def test_cell_fluxes_match_oracle_and_close_the_budget(oracle_lib, cfg):
    """CellFluxes outputs (AD:3356-3954, GetAdvFlux / GetDifFlux AD:697-851)."""
    mh, lh, mv, lv, adv_v, theta, nprop = cfg
    case = make_case(44, 38, 7, nprop=nprop, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    ts = gpu_for(case, g, s)
    prm = [default_params(mh, lh, mv, lv, impexp_advv=adv_v, theta_difv=theta) for _ in range(nprop)]
    for p in prm:
        p["CellFluxes"] = 1
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    old = [p.copy() for p in props]
    ts.advect_batch(gpu, prm)
    o.advect_batch(cpu, prm)
    compare(gpu, cpu, s, TOL_STEP)
    n = nprop - 1                                      # the oracle keeps the fluxes of its last call
    fg, fc = ts.get_cell_fluxes(n), o.get_cell_fluxes()
    for name in fc:
        scale = max(np.abs(fc[name]).max(), 1e-30)
        assert np.abs(fg[name] - fc[name]).max() / scale < 1e-12, name
        assert np.array_equal(fg[name] == 0, fc[name] == 0), name          # same faces carry a flux
    # budget of every open cell below the surface layer: V (Pnew - Pold Vold/V) / dt = sum of face fluxes
    K, J, I = case.K, case.J, case.I
    tot = {d: fg["AdvFlux" + d] + fg["DifFlux" + d] for d in "XYZ"}
    c = (slice(1, K + 1), slice(1, J + 1), slice(1, I + 1))
    net = (tot["X"][c] - tot["X"][1:K + 1, 2:J + 2, 1:I + 1] + tot["Y"][c] - tot["Y"][1:K + 1, 1:J + 1, 2:I + 2] +
           tot["Z"][c] - tot["Z"][2:K + 2, 1:J + 1, 1:I + 1])
    V = s["VolumeZ"]
    lhs = (V * (gpu[n] - old[n] * s["VolumeZOld"] / V) / case.dt)[c]
    w = (s["OpenPoints3D"][c] == 1) & (g["BoundaryPoints2D"][1:J + 1, 1:I + 1] == 0)[None]
    w[K - 1] = False                                   # the surface layer also exchanges through its free surface
    assert np.abs(lhs - net)[w].max() / np.abs(tot["X"]).max() < 1e-11
    ts.close()


@pytest.mark.parametrize("noadv,nodif", [(1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("mh", [1, 4])
def test_noflux_cells(oracle_lib, noadv, nodif, mh):
    """NoAdvFlux / NoDifFlux with the NoFluxU/V/W cell lists (AD:4438-4447, 4804-4813, 3003-3012, 2497-2501,
    2524-2528, 2737-2741): one property with the flags and one without, in the same batch."""
    case = make_case(52, 37, 9, nprop=2, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    rng = np.random.default_rng(5)
    nf = [np.asfortranarray((rng.random(s["OpenPoints3D"].shape) < 0.15).astype(np.int32)).reshape(s["OpenPoints3D"].shape)
          for _ in range(3)]
    nf = [np.ascontiguousarray(a) for a in nf]
    prm = [default_params(mh, 4, mh, 4, impexp_advv=0.0, theta_difv=0.5), default_params(mh, 4, mh, 4)]
    prm[0]["NoAdvFlux"], prm[0]["NoDifFlux"] = noadv, nodif
    ts = gpu_for(case, g, s)
    ts.set_noflux(*nf)
    o.set_noflux(*nf)
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    for _ in range(2):
        ts.advect_batch(gpu, prm)
        o.advect_batch(cpu, prm)
    compare(gpu, cpu, s, 2 * TOL_STEP)
    # the cell lists did change property 0 and left property 1 alone
    ts2 = gpu_for(case, g, s)
    ref = [p.copy() for p in props]
    prm0 = [dict(prm[0], NoAdvFlux=0, NoDifFlux=0), prm[1]]
    for _ in range(2):
        ts2.advect_batch(ref, prm0)
    assert np.abs(ref[0] - gpu[0]).max() > 0
    # P2_TVD coefficients are rebuilt for every property; the upwind ones are not (Set_Internal_State, AD:5768-5785),
    # so there the unflagged second property inherits the zeroed coefficients of the first; DifX / DifY are only
    # rebuilt when Schmidt_H changes, so their NoDifFlux zeroing is inherited with every method
    inherits = (mh != 4 and noadv) or nodif
    assert np.array_equal(ref[1], gpu[1]) != bool(inherits)
    ts.close(); ts2.close()


@pytest.mark.parametrize("case_id", ["nulldif_first", "nulldif_second", "optimize_schmidt_v", "schmidt_h_aba",
                                     "nodif_first", "noadv_second_upwind", "noadv_quick_vertical"])
def test_state_carried_between_properties(oracle_lib, case_id):
    """Coefficient arrays that Set_Internal_State (AD:5746-5835) does not rebuild keep what the previous property of
    the time step left in them (NullDif / NoDifFlux / NoAdvFlux zeroing, Optimize: first property's DifZ)."""
    case = make_case(48, 33, 8, nprop=3, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    # faces without flow, so that NullDif has something to zero
    s = dict(s)
    rng = np.random.default_rng(11)
    for name in ("Wflux_X", "Wflux_Y", "Wflux_Z"):
        a = s[name].copy(); a[rng.random(a.shape) < 0.2] = 0.0; s[name] = a
    o.set_step(s)
    nf = [np.ascontiguousarray((rng.random(s["OpenPoints3D"].shape) < 0.15).astype(np.int32)) for _ in range(3)]
    up = lambda **kw: dict(default_params(1, 4, 1, 4), **kw)
    tvd = lambda **kw: dict(default_params(4, 4, 4, 4), **kw)
    prm = {
        "nulldif_first": [up(NullDif=1), up(), up(SchmidtCoef_V=2.0)],
        "nulldif_second": [up(), up(NullDif=1), up()],
        "optimize_schmidt_v": [tvd(), tvd(SchmidtCoef_V=3.0), tvd(SchmidtBackground_V=1e-3)],
        "schmidt_h_aba": [tvd(Schmidt_H=1.0), tvd(Schmidt_H=2.0), tvd(Schmidt_H=1.0)],
        "nodif_first": [up(NoDifFlux=1), up(), up(Schmidt_H=0.5)],
        "noadv_second_upwind": [up(), up(NoAdvFlux=1), up()],
        "noadv_quick_vertical": [dict(default_params(4, 4, 2, 4, impexp_advv=0.0), NoAdvFlux=1),
                                 dict(default_params(4, 4, 2, 4, impexp_advv=0.0)),
                                 dict(default_params(4, 4, 2, 4, impexp_advv=0.0), NoAdvFlux=1)],
    }[case_id]
    ts = gpu_for(case, g, s)
    ts.set_noflux(*nf)
    o.set_noflux(*nf)
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    for _ in range(2):
        ts.advect_batch(gpu, prm)
        o.advect_batch(cpu, prm)
    compare(gpu, cpu, s, 2 * TOL_STEP)
    ts.close()


def test_noflux_flags_need_the_arrays():
    from mohid_b200.capi import AdtError
    case = make_case(20, 20, 5, nprop=1)
    o, g, s, props, refs = oracle_for(case)
    ts = gpu_for(case, g, s)
    prm = [default_params(1, 4, 1, 4)]
    prm[0]["NoAdvFlux"] = 1
    with pytest.raises(AdtError) as e:
        ts.advect_batch([props[0].copy()], prm)
    assert "NoFlux" in str(e.value)
    ts.close()


@pytest.mark.parametrize("shape", [(70, 45, 9, 4), (130, 37, 12, 10), (31, 20, 5, 7), (64, 33, 40, 10)])
@pytest.mark.parametrize("method", [4, 1])
@pytest.mark.parametrize("bc", [0, 4, 1])
def test_ring_kernel_matches_plain_kernel_and_oracle(oracle_lib, shape, method, bc, monkeypatch):
    """adt_transport_ring_kernel (inputs staged through cp.async / mbarrier rings, one consumer warp per property;
    opt-in with MOHID_ADT_RING=1) against adt_transport_kernel (bitwise: same arithmetic) and against the oracle."""
    I, J, K, N = shape
    case = make_case(I, J, K, nprop=N, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    prm = [default_params(method, 4, method, 4, bc=bc, decay_time=600.0) for _ in range(N)]
    out = {}
    for mode in ("ring", "ring2", "hsplit", "hsplit16", "plain"):
        monkeypatch.delenv("MOHID_ADT_RING", raising=False)
        monkeypatch.delenv("MOHID_ADT_HSPLIT", raising=False)
        if mode.startswith("ring"):
            monkeypatch.setenv("MOHID_ADT_RING", "2" if mode == "ring2" else "1")     # 2: two properties per warp
        elif mode.startswith("hsplit"):
            monkeypatch.setenv("MOHID_ADT_HSPLIT", "16" if mode == "hsplit16" else "12")   # warps of the column kernel
        ts = gpu_for(case, g, s)
        a = [p.copy() for p in props]
        for _ in range(3):
            ts.advect_batch(a, prm, refs)
        out[mode] = a
        assert ts.counters()["zero_pivots"] == 0
        ts.close()
    for mode in ("ring", "ring2", "hsplit", "hsplit16"):
        for a, b in zip(out[mode], out["plain"]):
            assert np.array_equal(a, b), mode
    cpu = [p.copy() for p in props]
    for _ in range(3):
        o.advect_batch(cpu, prm, refs)
    compare(out["ring"], cpu, s, 3 * TOL_STEP)


@pytest.mark.parametrize("method", [(1, 4), (4, 4), (4, 2), (5, 4)])
@pytest.mark.parametrize("bc", [0, 4, 1])
def test_horizontally_implicit_advection(oracle_lib, method, bc):
    """ImpExp_AdvXX / ImpExp_AdvYY = 1 (AD:4132-4265): implicit D/E fluxes of one horizontal direction, THOMAS_3D along
    it, then the vertical half of the step from the intermediate field.  One batch mixes an XX-implicit, a
    YY-implicit and an explicit property, as the alternating ImplicitH_Direction of WP:14676-14700 does."""
    mh, lim = method
    case = make_case(45, 38, 7, nprop=3, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    mv = 1 if mh == 5 else mh
    base = default_params(mh, lim, mv, lim, bc=bc, decay_time=600.0)
    prm = [dict(base, ImpExp_AdvXX=1.0), dict(base, ImpExp_AdvYY=1.0), dict(base)]
    ts = gpu_for(case, g, s)
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    for step in range(3):
        ts.advect_batch(gpu, prm, refs)
        o.advect_batch(cpu, prm, refs)
        prm[0], prm[1] = dict(prm[1]), dict(prm[0])          # the direction alternates from step to step
    compare(gpu, cpu, s, 3 * TOL_STEP)
    assert ts.counters()["zero_pivots"] == 0
    ts.close()


@pytest.mark.parametrize("method", [(1, 4), (4, 4)])
@pytest.mark.parametrize("bc", [0, 4, 1, 5])
def test_horizontally_implicit_2d_domain(oracle_lib, method, bc):
    """K = 1 (AD:1758-1841): no vertical system; explicit terms, the implicit direction, open-boundary rows and land fill
    are one tridiagonal system per line (THOMAS_3D).  XX-implicit, YY-implicit and explicit properties in one batch."""
    mh, lim = method
    case = make_case(45, 38, 1, nprop=3)
    o, g, s, props, refs = oracle_for(case)
    base = default_params(mh, lim, mh, lim, bc=bc, decay_time=600.0)
    prm = [dict(base, ImpExp_AdvXX=1.0), dict(base, ImpExp_AdvYY=1.0), dict(base)]
    ts = gpu_for(case, g, s)
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    for step in range(3):
        ts.advect_batch(gpu, prm, refs)
        o.advect_batch(cpu, prm, refs)
        prm[0], prm[1] = dict(prm[1]), dict(prm[0])
    compare(gpu, cpu, s, 3 * TOL_STEP)
    assert ts.counters()["zero_pivots"] == 0
    ts.close()


@pytest.mark.parametrize("K", [7, 1])
@pytest.mark.parametrize("direction", ["XX", "YY"])
@pytest.mark.parametrize("method", [1, 4])
def test_cell_fluxes_with_horizontally_implicit_advection(oracle_lib, K, direction, method):
    """AD:1895-1899 / 1908-1912: the implicit direction's cell flux is its coefficients (field at time n) times the final
    field; after a split step the vertical shares start from the line solve's result."""
    case = make_case(44, 38, K, nprop=2, stepped_bottom=K > 1)
    o, g, s, props, refs = oracle_for(case)
    ts = gpu_for(case, g, s)
    prm = [dict(default_params(method, 4, method, 4, impexp_advv=1.0, theta_difv=0.6, bc=4), CellFluxes=1,
                **{"ImpExp_Adv" + direction: 1.0}) for _ in range(2)]
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    for _ in range(2):
        ts.advect_batch(gpu, prm, refs)
        o.advect_batch(cpu, prm, refs)
    compare(gpu, cpu, s, 2 * TOL_STEP)
    fg, fc = ts.get_cell_fluxes(1), o.get_cell_fluxes()
    for name in fc:
        scale = max(np.abs(fc[name]).max(), 1e-30)
        assert np.abs(fg[name] - fc[name]).max() / scale < 1e-12, name
        assert np.array_equal(fg[name] == 0, fc[name] == 0), name
    assert np.abs(fc["AdvFlux" + direction[0]]).max() > 0
    ts.close()


@pytest.mark.parametrize("use", ["free_convection", "small_depths", "offsets", "all"])
def test_caller_side_pre_steps(oracle_lib, use):
    """FreeConvection, SmallDepthsMixing_Processes and AddOffSet of WP:14716-14759 / 14833-14858, done by the library
    on the device-resident fields, against the oracle's restatement of the caller loop."""
    I, J, K, N = 41, 33, 8, 3
    case = make_case(I, J, K, nprop=N, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    rng = np.random.default_rng(3)
    shape3 = s["OpenPoints3D"].shape
    density = np.ascontiguousarray(1025.0 + rng.standard_normal(shape3) * 0.05 - 0.02 * np.arange(shape3[0])[:, None, None])
    wcol = np.ascontiguousarray(5.0 + 40.0 * rng.random(shape3[1:]))
    limit = 12.0
    offs = [0.0, 273.15, -3.0]
    fc = use in ("free_convection", "all")
    sd = use in ("small_depths", "all")
    of = use in ("offsets", "all")
    prm = [default_params(4, 4, 4, 4, bc=4, theta_difv=0.5) for _ in range(N)]
    ts = gpu_for(case, g, s)
    ts.set_premix(density if fc else None, wcol if sd else None, limit)
    if of:
        ts.set_offsets(offs)
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    water = np.zeros(shape3, bool)
    water[1:K + 1, 1:J + 1, 1:I + 1] = s["WaterPoints3D"][1:K + 1, 1:J + 1, 1:I + 1] == 1
    for _ in range(2):
        ts.advect_batch(gpu, prm, refs)
        shifted_refs, on = [], None
        for n in range(N):
            off = offs[n] if of else 0.0
            on = o.caller_premix(cpu[n], density if fc else None, wcol if sd else None, limit, off)
            r = refs[n].copy()
            r[water] += off
            shifted_refs.append(r)
        o.set_step(s, small_depths=on)
        o.advect_batch(cpu, prm, shifted_refs)
        for n in range(N):
            if of and offs[n] != 0.0:
                o.caller_premix(cpu[n], offset=-offs[n])
    if sd:
        assert np.array_equal(ts.get_small_depths(), on) and on.sum() > 0
    compare(gpu, cpu, s, 1e-11)
    ts.close()


@pytest.mark.parametrize("method", [1, 4])
def test_orlanski_boundary(oracle_lib, method):
    """BoundaryCondition 6 (AD:5504-5570, OrlanskiCelerity2D MF:4129-4490): boundary rows and the exterior (halo) cells
    the reference writes into the property.  Halo values go through Q_b DT/V, so they agree to round-off, not bitwise."""
    case = make_case(44, 37, 7, nprop=2, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    prm = [default_params(method, 4, method, 4, bc=6) for _ in range(2)]
    ts = gpu_for(case, g, s)
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    for _ in range(3):
        ts.advect_batch(gpu, prm, refs)
        o.advect_batch(cpu, prm, refs)
    w = water_mask(s)
    for a, b, p0 in zip(gpu, cpu, props):
        assert rel_err(a, b, w) <= 3 * TOL_STEP
        changed = (b != p0) & ~w                                       # exterior cells written by the radiation routine
        assert changed.sum() > 0
        assert np.array_equal(a[~w & ~changed], b[~w & ~changed])
        assert np.allclose(a[changed], b[changed], rtol=1e-13, atol=0)
    ts.close()


@pytest.mark.parametrize("docycle", [1, 2])
def test_set_limits_after_the_step(oracle_lib, docycle):
    """SetLimitsProperty (WP:20594-20720) applied by the library after every step, with Mass_created / Mass_Destroid."""
    case = make_case(40, 31, 7, nprop=2, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case, docycle_method=docycle)
    ts = gpu_for(case, g, s, docycle_method=docycle)
    w = water_mask(s)
    lo, hi = np.percentile(props[0][w], [20, 80])
    ts.set_limits([float(lo), None], [float(hi), None])
    prm = [default_params(4, 4, 4, 4, bc=4) for _ in range(2)]
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    mc, md = np.zeros_like(props[0]), np.zeros_like(props[0])
    for _ in range(2):
        ts.advect_batch(gpu, prm, refs)
        o.advect_batch(cpu, prm, refs)
        o.set_limits(cpu[0], float(lo), float(hi), mc, md)
    compare(gpu, cpu, s, 2 * TOL_STEP)
    gmc, gmd = ts.get_limit_mass(0)
    assert mc.max() > 0 and md.min() < 0
    scale = np.abs(mc).max() + np.abs(md).max()
    assert np.abs(gmc - mc).max() <= 1e-11 * scale and np.abs(gmd - md).max() <= 1e-11 * scale
    assert gpu[0][w].min() >= lo and gpu[0][w].max() <= hi
    ts.close()


def test_set_step_keeps_arrays_passed_as_none(oracle_lib):
    """A NULL 3-D array in a later mohid_adt_set_step call keeps the device copy of the previous step."""
    case = make_case(30, 26, 6, nprop=1)
    o, g, s, props, refs = oracle_for(case)
    ts = gpu_for(case, g, s)
    s2 = dict(s)
    s2["Wflux_X"] = np.ascontiguousarray(s["Wflux_X"] * 0.5)
    partial = {k: (v if k == "Wflux_X" else None) for k, v in s2.items()}
    ts.set_step(partial)
    o.set_step(s2)
    prm = [default_params(4, 4, 4, 4)]
    a, b = [props[0].copy()], [props[0].copy()]
    ts.advect_batch(a, prm)
    o.advect_batch(b, prm)
    compare(a, b, s2, TOL_STEP)
    ts.close()
