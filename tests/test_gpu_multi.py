"""Multi-GPU path: j-slab decomposition with NCCL halo exchange must reproduce the single-GPU run
bit for bit on the owned columns (the same kernels run on the same values).  Needs >= 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

from mohid_b200.synthetic import make_case, default_params

pytestmark = pytest.mark.gpu

I, J, K, NPROP, STEPS = 70, 64, 8, 3, 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _params(kind):
    base = default_params(4, 4, 4, 4, bc=4)
    if kind == "implicit":
        # lines along j cross the slabs (the recurrence passes from rank to rank), lines along i lie inside them
        return [dict(base, ImpExp_AdvXX=1.0), dict(base, ImpExp_AdvYY=1.0), dict(base)]
    if kind == "cyclic":
        # Prop_CyclicBoundary joins global column 1 (first rank) and J (last rank)
        return [default_params(4, 4, 4, 4, bc=8) for _ in range(NPROP)]
    return [dict(base) for _ in range(NPROP)]


def _worker(rank, world, port, out_dir, kind):
    import torch.distributed as dist
    from mohid_b200.advection_diffusion import TransportStep
    from mohid_b200.partition import SlabDecomposition, HaloExchanger
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dec = SlabDecomposition(J, world, ghost=2)
    sl = dec.slab(rank)
    case = make_case(I, J, K, nprop=NPROP, device=str(dev), j_range=(sl.j_lo_ext, sl.j_hi_ext))
    ts = TransportStep(I, case.J, K, device=rank)
    ts.set_stream(torch.cuda.current_stream().cuda_stream)
    ts.set_grid2d(**case.grid2d)
    ts.set_step(case.step)
    ts.upload(case.props, case.refs)
    halo = HaloExchanger(ts, dec, rank, NPROP, dev)
    prm = _params(kind)
    for _ in range(STEPS):
        ts.advect_device(prm, 1)
        halo.exchange()
    out = [torch.empty_like(p) for p in case.props]
    ts.download(out)
    torch.cuda.synchronize()
    jb, n = sl.j_begin, sl.n_owned
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.stack([o[:, jb:jb + n, :].cpu().numpy() for o in out]))
    ts.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["explicit", "implicit", "cyclic"])
@pytest.mark.parametrize("world", [2, 4])
def test_slabs_with_nccl_halos_equal_single_gpu(tmp_path, world, kind):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from mohid_b200.advection_diffusion import TransportStep
    from mohid_b200.partition import SlabDecomposition
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path), kind), nprocs=world, join=True, start_method="spawn")
    case = make_case(I, J, K, nprop=NPROP, device="cuda:0")
    ts = TransportStep(I, J, K, device=0)
    ts.set_grid2d(**case.grid2d)
    ts.set_step(case.step)
    ts.upload(case.props, case.refs)
    ts.advect_device(_params(kind), STEPS)
    out = [torch.empty_like(p) for p in case.props]
    ts.download(out)
    torch.cuda.synchronize()
    glob = np.stack([o.cpu().numpy() for o in out])
    dec = SlabDecomposition(J, world)
    for r in range(world):
        lo, hi = dec.bounds[r]
        part = np.load(tmp_path / f"r{r}.npy")
        assert np.array_equal(part, glob[:, :, lo:hi + 1, :]), f"rank {r} differs from the single-GPU run"
    ts.close()
