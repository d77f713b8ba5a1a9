"""Parity at BASELINE.json's full single-GPU size (C3: 2048 x 2048 x 40, 10 properties, P2_TVD + SuperBee)
through size-independent properties, plus a window of the full-size run checked against the oracle.

* a constant tracer stays constant on the continuity-consistent synthetic fluxes (MassConservation boundary
  with ReferenceProp = the constant);
* land cells stay exactly null_real, dry columns / closed cells are untouched, PROP(:,:,KUB+1) = 0;
* in a closed basin the volume-weighted mass of every tracer is conserved;
* a 70 x 60 column window of the full-size step equals the oracle run on the same window (the explicit
  horizontal stencil reaches 2 cells per step, so cells >= 2 away from the window edge see identical inputs).
"""
import numpy as np
import pytest
import torch

from mohid_b200.synthetic import make_case, default_params

pytestmark = pytest.mark.gpu

I, J, K, N = 2048, 2048, 40, 10
NULL_REAL = -9.9e15


def _free_gb():
    free, _ = torch.cuda.mem_get_info()
    return free / 2**30


def test_mass_conservation_closed_basin_large():
    if _free_gb() < 60:
        pytest.skip("needs ~50 GB of device memory")
    from mohid_b200.advection_diffusion import TransportStep
    I2, J2, K2, N2 = 1024, 1024, 40, 4
    case = make_case(I2, J2, K2, nprop=N2, device="cuda", make_refs=False, closed=True, volume_change=0.0)
    torch.cuda.synchronize()              # the library copies on its own stream
    ts = TransportStep(I2, J2, K2)
    ts.set_grid2d(**case.grid2d)
    ts.set_step(case.step)
    ts.upload(case.props)
    w = case.step["OpenPoints3D"] == 1
    V = case.step["VolumeZ"]
    m0 = [float((p[w] * V[w]).sum()) for p in case.props]
    prm = [default_params(4, 4, 4, 4) for _ in range(N2)]
    ts.advect_device(prm, nsteps=10)
    out = [torch.empty_like(p) for p in case.props]
    ts.download(out)
    torch.cuda.synchronize()
    for a, b in zip(m0, out):
        m1 = float((b[w] * V[w]).sum())
        assert abs(m1 - a) / abs(a) < 1e-11
    ts.close()


@pytest.fixture(scope="module")
def big():
    if _free_gb() < 120:
        pytest.skip("needs ~110 GB of device memory")
    from mohid_b200.advection_diffusion import TransportStep
    case = make_case(I, J, K, nprop=N, device="cuda", make_refs=False)
    torch.cuda.synchronize()              # the library copies on its own stream
    ts = TransportStep(I, J, K)
    ts.set_grid2d(**case.grid2d)
    ts.set_step(case.step)
    yield case, ts
    ts.close()


def test_constant_tracers_and_mask_semantics_at_c3(big):
    case, ts = big
    land = case.step["LandPoints3D"] == 1
    water = case.step["WaterPoints3D"] == 1
    vals = [3.0 + n for n in range(N)]
    props, refs = [], []
    for v in vals:
        p = torch.where(land, torch.full_like(case.props[0], NULL_REAL), torch.full_like(case.props[0], v))
        p[0] = 0; p[-1] = 0; p[:, 0] = 0; p[:, -1] = 0; p[:, :, 0] = 0; p[:, :, I + 1:] = 0
        props.append(p.contiguous())
        refs.append(torch.full_like(p, v))
    ts.upload(props, refs)
    del refs
    prm = [default_params(4, 4, 4, 4, bc=1) for _ in range(N)]
    ts.advect_device(prm, nsteps=3)
    out = [torch.empty_like(p) for p in props]
    ts.download(out)
    torch.cuda.synchronize()
    for v, o, p0 in zip(vals, out, props):
        assert float((o[water] - v).abs().max()) < 1e-12 * v * 10
        assert bool((o[land] == NULL_REAL).all())
        assert bool((o[-1] == 0).all())
        assert torch.equal(o[~water & ~land], p0[~water & ~land])       # halos untouched
    assert ts.counters()["zero_pivots"] == 0


def test_window_of_full_size_step_matches_oracle(big, oracle_lib):
    case, ts = big
    prm = [default_params(4, 4, 4, 4) for _ in range(N)]
    ts.upload(case.props)
    ts.advect_device(prm, nsteps=1)
    out = [torch.empty_like(p) for p in case.props]
    ts.download(out)
    torch.cuda.synchronize()
    # window around an island corner: global cells i0..i0+wi-1, j0..j0+wj-1
    wi, wj = 70, 60
    i0, j0 = int(0.20 * I) - 30, int(0.25 * J) - 25
    sl3 = (slice(None), slice(j0 - 1, j0 + wj + 1), slice(i0 - 1, i0 + wi + 1))
    sl2 = (slice(j0 - 1, j0 + wj + 1), slice(i0 - 1, i0 + wi + 1))
    g = {k: np.ascontiguousarray(v[sl2].cpu().numpy()) for k, v in case.grid2d.items()}
    s = {k: np.ascontiguousarray(v[sl3].cpu().numpy()) for k, v in case.step.items()}
    # the window's own array halo is not a compute point (sub-domain convention)
    for name in ("OpenPoints3D",):
        s[name][:, 0, :] = 0; s[name][:, -1, :] = 0; s[name][:, :, 0] = 0; s[name][:, :, -1] = 0
    o = oracle_lib.OracleAdvectionDiffusion(wi, wj, K)
    o.set_grid2d(g)
    o.set_step(s)
    cpu = [np.ascontiguousarray(p[sl3].cpu().numpy()) for p in case.props]
    o.advect_batch(cpu, prm)
    m = 3                                                     # margin: stencil reach 2 (+1 for the masked halo)
    inner = (slice(1, K + 1), slice(1 + m, wj + 1 - m), slice(1 + m, wi + 1 - m))
    wmask = s["WaterPoints3D"][inner] == 1
    worst = 0.0
    for a, b in zip(out, cpu):
        ga = a[sl3].cpu().numpy()[inner]
        cb = b[inner]
        assert np.array_equal(ga == NULL_REAL, cb == NULL_REAL)
        d = np.abs(ga[wmask] - cb[wmask]) / np.maximum(np.abs(cb[wmask]), 1.0)
        worst = max(worst, float(d.max()))
    assert worst < 1e-12, worst
