"""The in-place shifted step (one buffer per property) walks the columns in chunks; the chunk width
is chosen from the free device memory, so small cases run as one chunk.  Here the width is forced down to a few columns:
the result must not depend on it, over several steps (the shift direction alternates) and on a column slab."""
import numpy as np
import pytest

from helpers import oracle_for, rel_err, water_mask
from mohid_b200.synthetic import make_case, default_params

pytestmark = pytest.mark.gpu


def _run(case, g, s, props, refs, prm, steps, active=None):
    from mohid_b200.advection_diffusion import TransportStep
    ts = TransportStep(case.I, case.J, case.K, case.ld)
    ts.set_grid2d(**g)
    ts.set_step(s)
    if active:
        ts.set_active_columns(*active)
    out = [p.copy() for p in props]
    ts.upload(out, refs)
    for _ in range(steps):
        ts.advect_device(prm, 1)
    ts.download(out)
    zp = ts.counters()["zero_pivots"]
    ts.close()
    return out, zp


@pytest.mark.parametrize("nprop,method,bc", [(3, 4, 4), (4, 1, 1), (2, 4, 7), (1, 1, 0), (2, 2, 4)])
@pytest.mark.parametrize("chunk", [5, 13])
def test_result_does_not_depend_on_the_chunk_width(oracle_lib, monkeypatch, nprop, method, bc, chunk):
    case = make_case(45, 58, 7, nprop=nprop, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    mv = method if method not in (2, 3) else 1
    prm = [default_params(method, 4, mv, 4, bc=bc, decay_time=600.0) for _ in range(nprop)]
    monkeypatch.delenv("MOHID_ADT_CHUNK_COLS", raising=False)
    whole, _ = _run(case, g, s, props, refs, prm, 3)
    monkeypatch.setenv("MOHID_ADT_CHUNK_COLS", str(chunk))
    parts, zp = _run(case, g, s, props, refs, prm, 3)
    assert zp == 0
    for a, b in zip(whole, parts):
        assert np.array_equal(a, b), "chunked and single-chunk steps differ"
    cpu = [p.copy() for p in props]
    for _ in range(3):
        o.advect_batch(cpu, prm, refs)
    w = water_mask(s)
    for a, b in zip(parts, cpu):
        assert np.array_equal(a[~w], b[~w])
        assert rel_err(a, b, w) <= 5e-12


def test_chunks_on_a_column_slab(monkeypatch):
    """Owned columns 9 .. 40 of 58: the columns outside keep their values whatever the chunk width."""
    case = make_case(45, 58, 7, nprop=3, stepped_bottom=True)
    _, g, s, props, refs = oracle_for(case)
    prm = [default_params(4, 4, 4, 4, bc=4) for _ in range(3)]
    monkeypatch.delenv("MOHID_ADT_CHUNK_COLS", raising=False)
    whole, _ = _run(case, g, s, props, refs, prm, 2, active=(9, 32))
    monkeypatch.setenv("MOHID_ADT_CHUNK_COLS", "6")
    parts, _ = _run(case, g, s, props, refs, prm, 2, active=(9, 32))
    for a, b, p0 in zip(whole, parts, props):
        assert np.array_equal(a, b)
        assert np.array_equal(b[:, :9, :], p0[:, :9, :]) and np.array_equal(b[:, 41:, :], p0[:, 41:, :])


def test_implicit_advection_along_i_on_a_column_slab():
    """ImpExp_AdvYY = 1: the line solve (THOMAS_3D, di = 1) runs along i, inside the slab; the owned columns of a slab equal
    those of the undivided run after one step.  Along j (ImpExp_AdvXX = 1) the lines cross the slabs: that needs the
    communicator."""
    from mohid_b200.advection_diffusion import TransportStep
    from mohid_b200.capi import AdtError
    case = make_case(45, 58, 7, nprop=2, stepped_bottom=True)
    _, g, s, props, refs = oracle_for(case)
    prm = [dict(default_params(1, 4, 1, 4, bc=4), ImpExp_AdvYY=1.0) for _ in range(2)]
    whole, _ = _run(case, g, s, props, refs, prm, 1)
    part, _ = _run(case, g, s, props, refs, prm, 1, active=(9, 32))
    for a, b, p0 in zip(whole, part, props):
        assert np.array_equal(a[:, 9:41, :], b[:, 9:41, :])
        assert np.array_equal(b[:, :9, :], p0[:, :9, :]) and np.array_equal(b[:, 41:, :], p0[:, 41:, :])
        assert not np.array_equal(b[:, 9:41, :], p0[:, 9:41, :])
    prm_x = [dict(default_params(1, 4, 1, 4, bc=4), ImpExp_AdvXX=1.0) for _ in range(2)]
    with pytest.raises(AdtError, match="mohid_adt_comm_init"):      # the recurrence passes from rank to rank (test_gpu_multi.py)
        _run(case, g, s, props, refs, prm_x, 1, active=(9, 32))


@pytest.mark.parametrize("chunk", [0, 5])
def test_2d_domain_line_solve_in_chunks_and_on_a_slab(oracle_lib, monkeypatch, chunk):
    """K = 1, horizontally implicit (AD:1758-1841): the line solve's result is moved into the shifted position chunk by
    chunk; along i it also runs on a column slab."""
    case = make_case(45, 58, 1, nprop=2)
    o, g, s, props, refs = oracle_for(case)
    prm = [dict(default_params(4, 4, 4, 4, bc=4), ImpExp_AdvYY=1.0), dict(default_params(4, 4, 4, 4, bc=1), ImpExp_AdvXX=1.0)]
    if chunk:
        monkeypatch.setenv("MOHID_ADT_CHUNK_COLS", str(chunk))
    else:
        monkeypatch.delenv("MOHID_ADT_CHUNK_COLS", raising=False)
    gpu, zp = _run(case, g, s, props, refs, prm, 3)
    assert zp == 0
    cpu = [p.copy() for p in props]
    for _ in range(3):
        o.advect_batch(cpu, prm, refs)
    w = water_mask(s)
    for a, b in zip(gpu, cpu):
        assert np.array_equal(a[~w], b[~w])
        assert rel_err(a, b, w) <= 5e-12
    prm_y = [prm[0], prm[0]]
    whole, _ = _run(case, g, s, props, refs, prm_y, 1)
    part, _ = _run(case, g, s, props, refs, prm_y, 1, active=(9, 32))
    for a, b, p0 in zip(whole, part, props):
        assert np.array_equal(a[:, 9:41, :], b[:, 9:41, :])
        assert np.array_equal(b[:, :9, :], p0[:, :9, :]) and np.array_equal(b[:, 41:, :], p0[:, 41:, :])
