"""Parity at the sizes BASELINE.json names, over the number of steps its tolerance is stated for (1e-10 relative after
100 steps, masks and indices bit-exact), against the CPU restatement of the reference path (oracle/):

* C2 (`configs[1]`): 512 x 512 x 20 sigma grid, 1 passive tracer, upwind horizontal + implicit vertical, 100 steps;
* C1-like (`configs[0]`, Samples/Coastal3D_Operational dimensions 305 x 232 x 75, temperature + salinity, NullGradient
  open boundary as in the sample's data files), P2_TVD + SuperBee, 100 steps -- once on the kernels a batch of two takes
  by default and once forced onto the fused kernel the 10-property configurations take;
* C3 (`configs[2]`): the full 2048 x 2048 x 40 x 10 run for 20 steps, a 70 x 60 column window of it (margin 3 cells
  per step, the explicit stencil reaches 2) against the oracle run on that window;
* a field built to hit the `|dC| < 1e-16` clamp of the limiter argument (MF:10795-10803, 10811-10819), which the
  division-free SuperBee form of the GPU path does not evaluate: plateaus whose neighbours differ by one ulp next to steps.

The oracle is the checker here, never the thing measured.
"""
import numpy as np
import pytest
import torch

from helpers import oracle_for, rel_err, water_mask
from mohid_b200.synthetic import make_case, default_params

pytestmark = pytest.mark.gpu

NULL_REAL = -9.9e15
TOL_100 = 1e-10            # BASELINE.json north_star: "within 1e-10 relative (fp64) after 100 steps"


def _gpu(case, g, s):
    from mohid_b200.advection_diffusion import TransportStep
    ts = TransportStep(case.I, case.J, case.K, case.ld)
    ts.set_grid2d(**g)
    ts.set_step(s)
    return ts


def _check(gpu, cpu, s, tol):
    w = water_mask(s)
    worst = 0.0
    for a, b in zip(gpu, cpu):
        assert np.array_equal(a == NULL_REAL, b == NULL_REAL), "land cells (null_real) differ"
        assert np.array_equal(a[~w], b[~w]), "cells that are not water points must be bit-identical"
        worst = max(worst, rel_err(a, b, w))
    assert worst <= tol, worst
    return worst


def test_c2_full_size_100_steps(oracle_lib):
    case = make_case(512, 512, 20, nprop=1)
    o, g, s, props, refs = oracle_for(case)
    prm = [default_params(1, 4, 1, 4)]
    ts = _gpu(case, g, s)
    gpu = [p.copy() for p in props]
    ts.upload(gpu)
    ts.advect_device(prm, nsteps=100)
    ts.download(gpu)
    cpu = [p.copy() for p in props]
    for _ in range(100):
        o.advect_batch(cpu, prm)
    worst = _check(gpu, cpu, s, TOL_100)
    assert ts.counters()["zero_pivots"] == 0
    print(f"C2 512x512x20, 100 steps: max relative difference {worst:.3e}")
    ts.close()


@pytest.mark.parametrize("fused", [False, True])
def test_c1_like_100_steps_null_gradient(oracle_lib, monkeypatch, fused):
    if fused:
        monkeypatch.setenv("MOHID_ADT_LEAN_ALWAYS", "1")      # two properties take the round-1 kernels by default
    case = make_case(305, 232, 75, nprop=2, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    prm = [default_params(4, 4, 4, 4, bc=4, decay_time=900.0) for _ in range(2)]
    ts = _gpu(case, g, s)
    gpu = [p.copy() for p in props]
    ts.upload(gpu, refs)
    ts.advect_device(prm, nsteps=100)
    ts.download(gpu)
    cpu = [p.copy() for p in props]
    for _ in range(100):
        o.advect_batch(cpu, prm, refs)
    worst = _check(gpu, cpu, s, TOL_100)
    assert ts.counters()["zero_pivots"] == 0
    print(f"C1-like 305x232x75 x 2, BC 4, 100 steps (fused={fused}): max relative difference {worst:.3e}")
    ts.close()


def test_c3_window_20_steps(oracle_lib):
    free, _ = torch.cuda.mem_get_info()
    if free / 2**30 < 120:
        pytest.skip("needs ~110 GB of device memory")
    from mohid_b200.advection_diffusion import TransportStep
    I, J, K, N, steps = 2048, 2048, 40, 10, 20
    case = make_case(I, J, K, nprop=N, device="cuda", make_refs=False)
    torch.cuda.synchronize()
    ts = TransportStep(I, J, K)
    ts.set_grid2d(**case.grid2d)
    ts.set_step(case.step)
    ts.upload(case.props)
    prm = [default_params(4, 4, 4, 4) for _ in range(N)]
    ts.advect_device(prm, nsteps=steps)
    m = 3 * steps                                              # cells per side the window's own edge can have reached
    wi, wj = 70 + 2 * m, 60 + 2 * m
    i0, j0 = int(0.20 * I) - 30 - m, int(0.25 * J) - 25 - m      # around an island corner
    sl3 = (slice(None), slice(j0 - 1, j0 + wj + 1), slice(i0 - 1, i0 + wi + 1))
    sl2 = (slice(j0 - 1, j0 + wj + 1), slice(i0 - 1, i0 + wi + 1))
    full = [torch.empty_like(p) for p in case.props]
    ts.download(full)
    torch.cuda.synchronize()
    out = [np.ascontiguousarray(p[sl3].cpu().numpy()) for p in full]
    del full
    g = {k: np.ascontiguousarray(v[sl2].cpu().numpy()) for k, v in case.grid2d.items()}
    s = {k: np.ascontiguousarray(v[sl3].cpu().numpy()) for k, v in case.step.items()}
    s["OpenPoints3D"][:, 0, :] = 0; s["OpenPoints3D"][:, -1, :] = 0        # the window's own halo is not a compute point
    s["OpenPoints3D"][:, :, 0] = 0; s["OpenPoints3D"][:, :, -1] = 0
    cpu = [np.ascontiguousarray(p[sl3].cpu().numpy()) for p in case.props]
    ts.close()
    del case
    torch.cuda.empty_cache()
    o = oracle_lib.OracleAdvectionDiffusion(wi, wj, K)
    o.set_grid2d(g)
    o.set_step(s)
    for _ in range(steps):
        o.advect_batch(cpu, prm)
    inner = (slice(1, K + 1), slice(1 + m, wj + 1 - m), slice(1 + m, wi + 1 - m))
    wmask = s["WaterPoints3D"][inner] == 1
    assert wmask.sum() > 1000 and (~wmask).sum() > 1000          # the window straddles a coast
    worst = 0.0
    for a, b in zip(out, cpu):
        ga, cb = a[inner], b[inner]
        assert np.array_equal(ga == NULL_REAL, cb == NULL_REAL)
        assert np.array_equal(ga[~wmask], cb[~wmask])
        d = np.abs(ga[wmask] - cb[wmask]) / np.maximum(np.abs(cb[wmask]), 1.0)
        worst = max(worst, float(d.max()))
    assert worst < 2e-11, worst
    print(f"C3 window after {steps} steps: max relative difference {worst:.3e}")


@pytest.mark.parametrize("nprop", [2, 4])
def test_limiter_argument_clamp_on_plateaus(oracle_lib, nprop):
    """`ComputeAdvectionFace` keeps the denominator of r away from zero: |dC| < 1e-16 -> sign(dC) 1e-16.  The explicit
    faces of the GPU path use psi(r) |dP| = max(0, min(|dP|, 2a), min(a, 2|dP|)), which has no denominator; the two differ
    by at most psi (0.5 (1 - Cr)) |dP| <= 1e-16 per face where the clamp acts.  Plateaus with one-ulp ripples (differences
    ~1e-17, dC ~ 1e-20) beside unit steps exercise exactly that, horizontally and vertically."""
    case = make_case(64, 48, 10, nprop=nprop, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    w = water_mask(s)
    rng = np.random.default_rng(7)
    K, nj, ld = props[0].shape
    ii, jj, kk = np.meshgrid(np.arange(ld), np.arange(nj), np.arange(K), indexing="ij")
    ii, jj, kk = ii.transpose(2, 1, 0), jj.transpose(2, 1, 0), kk.transpose(2, 1, 0)
    for n, p in enumerate(props):
        base = 0.1 + 0.1 * (((ii // 9) + (jj // 7) + (kk // 4) + n) % 2)          # plateaus 0.1 / 0.2 with sharp steps
        ulps = rng.integers(-1, 2, size=p.shape)                                    # -1, 0, +1 ulp ripples
        field = base + ulps * np.spacing(base)
        p[w] = field[w]
    prm = [default_params(4, 4, 4, 4, bc=0) for _ in range(nprop)]
    # how often the reference's clamp acts on the first step: differences below 1e-16 * (DX sum) between neighbours
    d = np.abs(np.diff(props[0], axis=1))[:, :, 1:-1]
    assert (d[(d > 0)] < 1e-16).sum() > 1000
    ts = _gpu(case, g, s)
    gpu, cpu = [p.copy() for p in props], [p.copy() for p in props]
    for _ in range(3):
        ts.advect_batch(gpu, prm)
        o.advect_batch(cpu, prm)
    worst = _check(gpu, cpu, s, 1e-13)
    print(f"clamp plateaus, {nprop} properties: max relative difference {worst:.3e}")
    ts.close()


def test_c4_window_across_a_chunk_boundary(oracle_lib):
    """C4 (4096 x 4096 x 40, 10 properties: bench.py's default workload) on ONE GPU, generated in column pieces straight into
    the device mirrors and advanced with the in-place step in 17 chunks of 256 columns; a window that straddles the chunk
    boundary at column 1024 (and an island corner) is read back with the column-window download and compared with the
    oracle run on that window."""
    free, _ = torch.cuda.mem_get_info()
    if free / 2**30 < 165:
        pytest.skip("needs ~150 GB of device memory")
    from mohid_b200.advection_diffusion import TransportStep
    from mohid_b200.synthetic import case_pieces
    I, J, K, N, steps = 4096, 4096, 40, 10, 5
    ts = TransportStep(I, J, K, max_properties=N)
    g2, dt = {}, None
    for j0, pc in case_pieces(I, J, K, N, piece=64, device="cuda"):
        dt = pc.dt
        torch.cuda.synchronize()          # the library copies on its own stream: the generator's kernels must have finished
        ts.set_step_columns(j0, pc.step)
        ts.upload_columns(j0, pc.props)
        for k, v in pc.grid2d.items():
            g2.setdefault(k, []).append(v)
        del pc
    ts.set_grid2d(**{k: torch.cat(v, 0).contiguous() for k, v in g2.items()})
    ts.mark_step_resident()
    del g2
    prm = [default_params(4, 4, 4, 4, dt=dt) for _ in range(N)]
    ts.advect_device(prm, nsteps=steps)
    m = 3 * steps
    wi, wj = 70 + 2 * m, 60 + 2 * m
    i0, j0 = int(0.20 * I) - 30 - m, 1024 - 30 - m                   # window columns j0 .. j0+wj-1 around column 1024
    # the window's own inputs: the same generator, restricted to its columns (identical to the undivided case)
    pcs = list(case_pieces(I, J, K, N, j_lo=j0, j_hi=j0 + wj - 1, piece=wj + 2, device="cpu"))
    assert len(pcs) == 1
    wc = pcs[0][1]
    si = slice(i0 - 1, i0 + wi + 1)
    g = {k: np.ascontiguousarray(v[:, si].numpy()) for k, v in wc.grid2d.items()}
    s = {k: np.ascontiguousarray(v[:, :, si].numpy()) for k, v in wc.step.items()}
    s["OpenPoints3D"][:, 0, :] = 0; s["OpenPoints3D"][:, -1, :] = 0
    s["OpenPoints3D"][:, :, 0] = 0; s["OpenPoints3D"][:, :, -1] = 0
    cpu = [np.ascontiguousarray(p[:, :, si].numpy()) for p in wc.props]
    out_full = [np.zeros((K + 2, wj + 2, I + 2)) for _ in range(N)]
    ts.download_columns(j0 - 1, out_full)                             # local column index = global column (single GPU)
    out = [np.ascontiguousarray(p[:, :, si]) for p in out_full]
    ts.close()
    o = oracle_lib.OracleAdvectionDiffusion(wi, wj, K)
    o.set_grid2d(g)
    o.set_step(s)
    for _ in range(steps):
        o.advect_batch(cpu, prm)
    inner = (slice(1, K + 1), slice(1 + m, wj + 1 - m), slice(1 + m, wi + 1 - m))
    wmask = s["WaterPoints3D"][inner] == 1
    assert wmask.sum() > 1000 and (~wmask).sum() > 1000
    worst = 0.0
    for a, b in zip(out, cpu):
        ga, cb = a[inner], b[inner]
        assert np.array_equal(ga == NULL_REAL, cb == NULL_REAL)
        assert np.array_equal(ga[~wmask], cb[~wmask])
        d = np.abs(ga[wmask] - cb[wmask]) / np.maximum(np.abs(cb[wmask]), 1.0)
        worst = max(worst, float(d.max()))
    assert worst < 1e-11, worst
    print(f"C4 window across the chunk boundary after {steps} steps: max relative difference {worst:.3e}")
