"""Seeded random configurations of the whole option space of the step -- grid shape and padding, bottom, open / closed
basin, Vertical1D / XZFlow, Docycle method, advection methods and limiters, explicit / implicit vertical advection,
theta of the vertical diffusion, boundary condition, decay time, NullDif, horizontally implicit directions, batch size --
three steps each, CUDA path against the oracle with the bar of tests/test_gpu_parity.py.  The point is the combinations
nobody thought of writing a test for."""
import os

import numpy as np
import pytest

from mohid_b200.synthetic import make_case, default_params
from helpers import oracle_for, water_mask
from test_gpu_parity import gpu_for, compare, TOL_STEP

pytestmark = pytest.mark.gpu


# MOHID_ADT_FUZZ_OFFSET=n: another set of draws for the same test ids (exploration runs; the committed suite uses 0)
OFFSET = int(os.environ.get("MOHID_ADT_FUZZ_OFFSET", "0"))


def draw(seed):
    r = np.random.default_rng(1000 + seed + 100000 * OFFSET)
    I, J, K = int(r.integers(18, 75)), int(r.integers(18, 64)), int(r.choice([1, 2, 3, 5, 8, 12]))
    nprop = int(r.choice([1, 2, 3, 4, 6]))
    case_kw = dict(stepped_bottom=bool(r.integers(2)) and K > 1, closed=bool(r.integers(4) == 0), islands=bool(r.integers(2)),
                   ld=(None if r.integers(2) else I + 2 + int(r.integers(1, 9))))
    opt = {}
    mode = int(r.integers(5))
    if mode == 0 and K > 1:
        opt["vertical1d"] = True
    elif mode == 1:
        opt["xzflow"] = True
    if r.integers(3) == 0:
        opt["docycle_method"] = 2
    mh = int(r.choice([1, 2, 3, 4, 5, 6]))
    lim = int(r.integers(1, 6))
    advv = float(r.choice([0.0, 1.0]))
    mv = mh if not (advv == 1.0 and mh in (2, 3)) else 1
    prm = []
    for n in range(nprop):
        bc = int(r.choice([0, 1, 2, 4, 5, 7, 8, 6])) if not case_kw["closed"] else 0
        p = default_params(mh, lim, mv, lim, bc=bc, impexp_advv=advv, theta_difv=float(r.choice([0.0, 0.35, 1.0])),
                           decay_time=float(r.choice([0.0, 450.0])), schmidt_h=float(r.choice([1.0, 0.7])))
        p["NullDif"] = int(r.integers(5) == 0)
        if mh not in (2, 3) and not opt.get("vertical1d") and r.integers(4) == 0:
            p["ImpExp_Adv" + ("XX" if r.integers(2) and not opt.get("xzflow") else "YY")] = 1.0
        prm.append(p)
    return (I, J, K, nprop), case_kw, opt, prm


@pytest.mark.parametrize("seed", range(64))
def test_random_configuration_matches_the_oracle(oracle_lib, seed):
    (I, J, K, nprop), case_kw, opt, prm = draw(seed)
    case = make_case(I, J, K, nprop=nprop, **case_kw)
    o, g, s, props, refs = oracle_for(case, **opt)
    ts = gpu_for(case, g, s, **opt)
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    for _ in range(3):
        ts.advect_batch(gpu, prm, refs)
        o.advect_batch(cpu, prm, refs)
    # the Orlanski routine writes exterior (halo) cells through Q DT/V: round-off, not bitwise (test_orlanski_boundary)
    w = water_mask(s)
    for n, p in enumerate(prm):
        if p["BoundaryCondition"] == 6:
            ext = ~w & (cpu[n] != props[n])
            assert np.allclose(gpu[n][ext], cpu[n][ext], rtol=1e-12, atol=0)
            gpu[n][ext] = cpu[n][ext]
    compare(gpu, cpu, s, 3 * TOL_STEP)
    assert ts.counters()["zero_pivots"] == 0
    ts.close()


@pytest.mark.parametrize("seed", range(64, 192))
def test_random_configuration_resident_chunked_with_fluxes(oracle_lib, monkeypatch, seed):
    """The same draw, plus: the column walk forced to narrow chunks, the properties resident on the device between the
    steps, and CellFluxes on the last property (compared face by face after the last step)."""
    (I, J, K, nprop), case_kw, opt, prm = draw(seed)
    r = np.random.default_rng(5000 + seed)
    chunk = int(r.choice([0, 0, 4, 9]))
    resident = bool(r.integers(2))
    fluxes = bool(r.integers(2))
    if chunk:
        monkeypatch.setenv("MOHID_ADT_CHUNK_COLS", str(chunk))
    else:
        monkeypatch.delenv("MOHID_ADT_CHUNK_COLS", raising=False)
    if fluxes:
        prm[-1]["CellFluxes"] = 1
    case = make_case(I, J, K, nprop=nprop, **case_kw)
    o, g, s, props, refs = oracle_for(case, **opt)
    ts = gpu_for(case, g, s, **opt)
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    if resident:
        ts.upload(gpu, refs)
        ts.advect_device(prm, 3)
        ts.download(gpu)
    for _ in range(3):
        if not resident:
            ts.advect_batch(gpu, prm, refs)
        o.advect_batch(cpu, prm, refs)
    w = water_mask(s)
    for n, p in enumerate(prm):
        if p["BoundaryCondition"] == 6:
            ext = ~w & (cpu[n] != props[n])
            assert np.allclose(gpu[n][ext], cpu[n][ext], rtol=1e-12, atol=0)
            gpu[n][ext] = cpu[n][ext]
    compare(gpu, cpu, s, 3 * TOL_STEP)
    if fluxes:
        fg, fc = ts.get_cell_fluxes(nprop - 1), o.get_cell_fluxes()
        for name in fc:
            scale = max(np.abs(fc[name]).max(), 1e-30)
            assert np.abs(fg[name] - fc[name]).max() / scale < 1e-11, name
    assert ts.counters()["zero_pivots"] == 0
    ts.close()


@pytest.mark.parametrize("seed", range(192, 256))
def test_random_configuration_with_noflux_cell_lists(oracle_lib, monkeypatch, seed):
    """The same draw with random NoFluxU / V / W cell lists and NoAdvFlux / NoDifFlux flags on some properties (their
    zeroing carries over to later properties of the step, AD:5768-5785), chunked or not."""
    (I, J, K, nprop), case_kw, opt, prm = draw(seed)
    r = np.random.default_rng(9000 + seed)
    chunk = int(r.choice([0, 6]))
    if chunk:
        monkeypatch.setenv("MOHID_ADT_CHUNK_COLS", str(chunk))
    else:
        monkeypatch.delenv("MOHID_ADT_CHUNK_COLS", raising=False)
    for p in prm:
        p["NoAdvFlux"], p["NoDifFlux"] = int(r.integers(3) == 0), int(r.integers(3) == 0)
    case = make_case(I, J, K, nprop=nprop, **case_kw)
    o, g, s, props, refs = oracle_for(case, **opt)
    nf = [np.ascontiguousarray((r.random(s["OpenPoints3D"].shape) < 0.15).astype(np.int32)) for _ in range(3)]
    ts = gpu_for(case, g, s, **opt)
    ts.set_noflux(*nf)
    o.set_noflux(*nf)
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    for _ in range(3):
        ts.advect_batch(gpu, prm, refs)
        o.advect_batch(cpu, prm, refs)
    w = water_mask(s)
    for n, p in enumerate(prm):
        if p["BoundaryCondition"] == 6:
            ext = ~w & (cpu[n] != props[n])
            assert np.allclose(gpu[n][ext], cpu[n][ext], rtol=1e-12, atol=0)
            gpu[n][ext] = cpu[n][ext]
    compare(gpu, cpu, s, 3 * TOL_STEP)
    assert ts.counters()["zero_pivots"] == 0
    ts.close()


@pytest.mark.parametrize("seed", range(256, 304))
def test_random_configuration_with_discharges(oracle_lib, monkeypatch, seed):
    """One property with point discharges (bottom source, uniform over the column, withdrawal; AD:4025-4128) under the
    drawn options -- the discharge terms enter the explicit rows, the line solve of an implicit direction and the column
    solve alike."""
    from test_gpu_parity import _discharge_set
    (I, J, K, nprop), case_kw, opt, prm = draw(seed)
    K = max(K, 2)
    case_kw = dict(case_kw, stepped_bottom=False)
    prm = prm[:1]
    r = np.random.default_rng(11000 + seed)
    chunk = int(r.choice([0, 7]))
    if chunk:
        monkeypatch.setenv("MOHID_ADT_CHUNK_COLS", str(chunk))
    else:
        monkeypatch.delenv("MOHID_ADT_CHUNK_COLS", raising=False)
    case = make_case(I, J, K, nprop=1, **case_kw)
    o, g, s, props, refs = oracle_for(case, **opt)
    dset = _discharge_set(case, s, [float(r.uniform(0, 40)), float(r.uniform(0, 10)), 0.0], [1.0, float(r.integers(2)), 0.5])
    ts = gpu_for(case, g, s, **opt)
    ts.set_discharges(0, dset)
    o.set_discharges(dset)
    gpu, cpu = [props[0].copy()], [props[0].copy()]
    for _ in range(3):
        ts.advect_batch(gpu, prm, refs)
        o.advect_batch(cpu, prm, refs)
    w = water_mask(s)
    if prm[0]["BoundaryCondition"] == 6:
        ext = ~w & (cpu[0] != props[0])
        assert np.allclose(gpu[0][ext], cpu[0][ext], rtol=1e-12, atol=0)
        gpu[0][ext] = cpu[0][ext]
    compare(gpu, cpu, s, 3 * TOL_STEP)
    ts.close()
