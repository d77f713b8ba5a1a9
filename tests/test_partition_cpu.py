"""Host-side logic of the j-slab decomposition, on CPU: world_size-2 (and 3) gloo process groups run
the oracle on slabs with the halo exchange after every step and must reproduce the global run
bit for bit on the owned columns."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mohid_b200.partition import SlabDecomposition, exchange_host_arrays
from mohid_b200.synthetic import make_case, default_params

I, J, K, NPROP, STEPS = 30, 46, 6, 2, 4


def test_slab_bounds():
    d = SlabDecomposition(10, 3)
    assert d.bounds == [(1, 4), (5, 7), (8, 10)]
    s0, s1, s2 = d.slab(0), d.slab(1), d.slab(2)
    assert (s0.ghost_left, s0.ghost_right, s0.j_begin, s0.J_local) == (0, 2, 1, 6)
    assert (s1.j_lo_ext, s1.j_hi_ext, s1.j_begin, s1.n_owned, s1.J_local) == (3, 9, 3, 3, 7)
    assert (s2.ghost_left, s2.ghost_right, s2.local_j(8)) == (2, 0, 3)
    assert [d.owner(j) for j in (1, 4, 5, 7, 8, 10)] == [0, 0, 1, 1, 2, 2]
    with pytest.raises(ValueError):
        SlabDecomposition(3, 4)


def test_slab_generation_matches_global_case():
    g = make_case(I, J, K, nprop=NPROP, stepped_bottom=True)
    for lo, hi in [(1, 20), (19, 46), (11, 30)]:
        s = make_case(I, J, K, nprop=NPROP, stepped_bottom=True, j_range=(lo, hi))
        for k, v in s.step.items():
            assert torch.equal(v[:, 1:-1], g.step[k][:, lo:hi + 1, :]), k
        for k, v in s.grid2d.items():
            assert torch.equal(v[1:-1], g.grid2d[k][lo:hi + 1, :]), k
        for a, b in zip(s.props, g.props):
            assert torch.equal(a[:, 1:-1], b[:, lo:hi + 1, :])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import OracleAdvectionDiffusion, case_to_numpy
    dec = SlabDecomposition(J, world, ghost=2)
    sl = dec.slab(rank)
    case = make_case(I, J, K, nprop=NPROP, j_range=(sl.j_lo_ext, sl.j_hi_ext))
    # sub-domain array halos are not compute points (MOHID sub-domains keep the usual 1-cell array halo)
    case.step["OpenPoints3D"][:, 0, :] = 0
    case.step["OpenPoints3D"][:, -1, :] = 0
    g, s, props, refs = case_to_numpy(case)
    o = OracleAdvectionDiffusion(I, case.J, K, nthreads=1)
    o.set_grid2d(g)
    o.set_step(s)
    prm = [default_params(4, 4, 4, 4) for _ in range(NPROP)]
    tprops = [torch.from_numpy(p) for p in props]           # share memory with the numpy arrays
    for _ in range(STEPS):
        o.advect_batch(props, prm, force_optimize=1)
        exchange_host_arrays(tprops, dec, rank)
    jb, n = sl.j_begin, sl.n_owned
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.stack([p[:, jb:jb + n, :] for p in props]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_decomposed_oracle_equals_global_run(oracle_lib, tmp_path, world):
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    from helpers import oracle_for
    case = make_case(I, J, K, nprop=NPROP)
    o, g, s, props, refs = oracle_for(case, nthreads=1)
    prm = [default_params(4, 4, 4, 4) for _ in range(NPROP)]
    for _ in range(STEPS):
        o.advect_batch(props, prm, force_optimize=1)
    glob = np.stack(props)
    dec = SlabDecomposition(J, world)
    for r in range(world):
        lo, hi = dec.bounds[r]
        part = np.load(tmp_path / f"r{r}.npy")
        assert np.array_equal(part, glob[:, :, lo:hi + 1, :]), f"rank {r} differs from the global run"


def test_line_recurrence_passed_from_slab_to_slab_equals_the_whole_line():
    """The scheme of the split line solve (adt_hsolve_kernel with HSolveArgs::split + adt_hsolve_back_kernel): the forward
    recurrence of THOMAS_3D (MF:3790-3801) over each slab's owned cells starting from (W, G) of the neighbour's last cell,
    then the back substitution from the right starting from x of the neighbour's first cell, is the same sequence of
    operations as the undivided recurrence -- bit for bit, whatever the split."""
    rng = np.random.default_rng(11)
    n = 57
    D, F = rng.uniform(-0.4, 0.0, n + 2), rng.uniform(-0.4, 0.0, n + 2)
    E = 1.0 - D - F + rng.uniform(0, 0.2, n + 2)
    TI = rng.uniform(0, 30, n + 2)

    def forward(lo, hi, w, g, W, G):
        for l in range(lo, hi + 1):
            aux = E[l] + D[l] * w
            w, g = -F[l] / aux, (TI[l] - D[l] * g) / aux
            W[l], G[l] = w, g
        return w, g

    def backward(lo, hi, x, X, W, G):
        for l in range(hi, lo - 1, -1):
            x = W[l] * x + G[l]
            X[l] = x
        return x

    W, G, X = np.zeros(n + 2), np.zeros(n + 2), np.zeros(n + 2)
    forward(1, n, 0.0, 0.0, W, G)
    backward(1, n, 0.0, X, W, G)
    for world in (2, 3, 5):
        dec = SlabDecomposition(n, world)
        W2, G2, X2 = np.zeros(n + 2), np.zeros(n + 2), np.zeros(n + 2)
        edge = (0.0, 0.0)
        for r in range(world):                       # left to right: what ncclSend / ncclRecv carry
            edge = forward(*dec.bounds[r], *edge, W2, G2)
        x = 0.0
        for r in reversed(range(world)):             # right to left
            x = backward(*dec.bounds[r], x, X2, W2, G2)
        assert np.array_equal(X, X2) and np.array_equal(W, W2) and np.array_equal(G, G2)
    assert np.allclose(D[1:n + 1] * np.r_[0.0, X[1:n]] + E[1:n + 1] * X[1:n + 1] + F[1:n + 1] * np.r_[X[2:n + 1], 0.0], TI[1:n + 1])
