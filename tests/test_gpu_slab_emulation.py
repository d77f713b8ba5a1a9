"""The j-slab decomposition with every rank's handle on ONE GPU: the ranks step in turn and their ghost columns are filled
through mohid_adt_pack_columns / mohid_adt_unpack_columns (what mohid_adt_exchange_halos does around ncclSend/ncclRecv).
Owned columns -- halo rows and, on the edge ranks, the outer halo columns included -- must equal the undivided run bit for
bit, for every open-boundary condition that is served on a slab.  (The NCCL path itself: tests/test_gpu_multi.py.)"""
import numpy as np
import pytest
import torch

from mohid_b200.synthetic import make_case, default_params

pytestmark = pytest.mark.gpu

I, J, K, N, STEPS, G = 53, 66, 6, 3, 4, 2


def _handle(case):
    from mohid_b200.advection_diffusion import TransportStep
    ts = TransportStep(case.I, case.J, case.K)
    torch.cuda.synchronize()            # the generator's kernels run on torch's stream, the library copies on its own
    ts.set_grid2d(**case.grid2d)
    ts.set_step(case.step)
    ts.upload(case.props, case.refs)
    return ts


def _download(ts, case):
    out = [torch.empty_like(p) for p in case.props]
    ts.download(out)
    torch.cuda.synchronize()
    return np.stack([o.cpu().numpy() for o in out])


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("bc,method,extra", [(4, 4, {}), (6, 4, {}), (6, 1, {}), (1, 4, {}), (2, 1, {}), (7, 4, {}),
                                              (4, 4, {"ImpExp_AdvYY": 1.0})])
def test_slab_ranks_on_one_gpu_equal_the_undivided_run(world, bc, method, extra):
    from mohid_b200.partition import SlabDecomposition
    prm = [dict(default_params(method, 4, method, 4, bc=bc, decay_time=700.0), **extra) for _ in range(N)]
    whole_case = make_case(I, J, K, nprop=N, device="cuda", stepped_bottom=True)
    whole = _handle(whole_case)
    whole.advect_device(prm, STEPS)
    glob = _download(whole, whole_case)
    whole.close()

    dec = SlabDecomposition(J, world, ghost=G)
    slabs = [dec.slab(r) for r in range(world)]
    cases = [make_case(I, J, K, nprop=N, device="cuda", stepped_bottom=True, j_range=(sl.j_lo_ext, sl.j_hi_ext)) for sl in slabs]
    ranks = [_handle(c) for c in cases]
    for ts, sl in zip(ranks, slabs):
        ts.set_active_columns(sl.j_begin, sl.n_owned)
    buf = torch.empty(ranks[0].pack_elems(N, G), dtype=torch.float64, device="cuda")      # device rows are padded to 128 B
    for _ in range(STEPS):
        for ts in ranks:
            ts.advect_device(prm, 1)
        torch.cuda.synchronize()
        for r in range(world - 1):
            a, b, sa, sb = ranks[r], ranks[r + 1], slabs[r], slabs[r + 1]
            a.pack_columns(N, sa.j_begin + sa.n_owned - G, G, buf); torch.cuda.synchronize()
            b.unpack_columns(N, sb.j_begin - G, G, buf); torch.cuda.synchronize()
            b.pack_columns(N, sb.j_begin, G, buf); torch.cuda.synchronize()
            a.unpack_columns(N, sa.j_begin + sa.n_owned, G, buf); torch.cuda.synchronize()
    for r, (ts, sl, c) in enumerate(zip(ranks, slabs, cases)):
        part = _download(ts, c)
        jb, n = sl.j_begin, sl.n_owned
        assert np.array_equal(part[:, :, jb:jb + n, :], glob[:, :, sl.j_lo:sl.j_hi + 1, :]), f"rank {r} of {world}"
        if r == 0:
            assert np.array_equal(part[:, :, 0, :], glob[:, :, 0, :])                   # outer halo column (Orlanski writes it)
        if r == world - 1:
            assert np.array_equal(part[:, :, -1, :], glob[:, :, -1, :])
        ts.close()
    assert not np.array_equal(glob, np.stack([p.cpu().numpy() for p in whole_case.props]))


@pytest.mark.parametrize("prm", [dict(default_params(4, 4, 4, 4, bc=8)), dict(default_params(4, 4, 4, 4, bc=4), ImpExp_AdvXX=1.0)])
def test_options_that_couple_the_slabs_need_the_communicator(prm):
    """Cyclic j wrap and lines along j cross the slabs: without mohid_adt_comm_init the step is refused, not silently local."""
    from mohid_b200.capi import AdtError
    from mohid_b200.partition import SlabDecomposition
    sl = SlabDecomposition(J, 2, ghost=G).slab(0)
    case = make_case(I, J, K, nprop=1, device="cuda", j_range=(sl.j_lo_ext, sl.j_hi_ext))
    ts = _handle(case)
    ts.set_active_columns(sl.j_begin, sl.n_owned)
    with pytest.raises(AdtError, match="mohid_adt_comm_init") as e:
        ts.advect_device([prm], 1)
    assert e.value.code == 31           # MOHID_ADT_ERR_STATE
    ts.close()


def _emulated(world, case_kw, opt, prm, dims, steps):
    from mohid_b200.advection_diffusion import TransportStep
    from mohid_b200.partition import SlabDecomposition
    I_, J_, K_, N_ = dims

    def handle(case):
        ts = TransportStep(case.I, case.J, case.K, case.ld, **opt)
        torch.cuda.synchronize()
        ts.set_grid2d(**case.grid2d)
        ts.set_step(case.step)
        ts.upload(case.props, case.refs)
        return ts

    whole_case = make_case(I_, J_, K_, nprop=N_, device="cuda", **case_kw)
    whole = handle(whole_case)
    whole.advect_device(prm, steps)
    glob = _download(whole, whole_case)
    whole.close()
    dec = SlabDecomposition(J_, world, ghost=G)
    slabs = [dec.slab(r) for r in range(world)]
    cases = [make_case(I_, J_, K_, nprop=N_, device="cuda", j_range=(sl.j_lo_ext, sl.j_hi_ext), **case_kw) for sl in slabs]
    ranks = [handle(c) for c in cases]
    for ts, sl in zip(ranks, slabs):
        ts.set_active_columns(sl.j_begin, sl.n_owned)
    buf = torch.empty(ranks[0].pack_elems(N_, G), dtype=torch.float64, device="cuda")
    for _ in range(steps):
        for ts in ranks:
            ts.advect_device(prm, 1)
        torch.cuda.synchronize()
        for r in range(world - 1):
            a, b, sa, sb = ranks[r], ranks[r + 1], slabs[r], slabs[r + 1]
            a.pack_columns(N_, sa.j_begin + sa.n_owned - G, G, buf); torch.cuda.synchronize()
            b.unpack_columns(N_, sb.j_begin - G, G, buf); torch.cuda.synchronize()
            b.pack_columns(N_, sb.j_begin, G, buf); torch.cuda.synchronize()
            a.unpack_columns(N_, sa.j_begin + sa.n_owned, G, buf); torch.cuda.synchronize()
    for r, (ts, sl, c) in enumerate(zip(ranks, slabs, cases)):
        part = _download(ts, c)
        jb, n = sl.j_begin, sl.n_owned
        # (elements beyond I + 1 are row padding: never read, never written, not even initialised by the generator)
        assert np.array_equal(part[:, :, jb:jb + n, :I_ + 2], glob[:, :, sl.j_lo:sl.j_hi + 1, :I_ + 2]), f"rank {r} of {world}"
        ts.close()


@pytest.mark.parametrize("seed", range(40))
def test_random_configuration_on_slabs_equals_the_undivided_run(monkeypatch, seed):
    """The draws of tests/test_gpu_fuzz.py on 2 or 3 slabs (options that need the communicator replaced by their local
    counterparts), chunked or not: owned columns bit-identical to the undivided run."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_fuzz import draw
    dims, case_kw, opt, prm = draw(300 + seed)
    r = np.random.default_rng(7000 + seed)
    for p in prm:
        if p["BoundaryCondition"] == 8:
            p["BoundaryCondition"] = 4
        if p["ImpExp_AdvXX"] == 1.0:
            p["ImpExp_AdvXX"], p["ImpExp_AdvYY"] = 0.0, 1.0
    chunk = int(r.choice([0, 0, 5]))
    if chunk:
        monkeypatch.setenv("MOHID_ADT_CHUNK_COLS", str(chunk))
    else:
        monkeypatch.delenv("MOHID_ADT_CHUNK_COLS", raising=False)
    if opt.get("xzflow"):
        for p in prm:
            p["ImpExp_AdvYY"] = 0.0
    _emulated(int(r.choice([2, 3])), case_kw, opt, prm, dims, 3)
