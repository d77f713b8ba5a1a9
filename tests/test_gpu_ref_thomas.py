"""Row a18 (THOMASZ_NewType2, MF:4026-4123) pinned against REFERENCE code.

The reference's Fortran cannot be built here, but its legacy GPU path holds the same column recurrence in CUDA C++:
``Software/CudaThomas/Thomas.cu`` (``DevThomasIK`` :62-131, reached through ``SolveThomas_C`` :24-52 exactly as
``ModuleCuda.F90:103-111`` binds it).  ``make -C oracle ref`` compiles those sources unmodified into
``oracle/_ref/libcudathomas_ref.so``.  Here the oracle's restatement and the product's column solve are compared with it on
random diagonally dominant systems that include the KUB+1 row, the literal first-row index and non-trivial halos.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def systems(I, J, K, seed):
    rng = np.random.default_rng(seed)
    shape = (K + 2, J + 2, I + 2)
    D = -rng.uniform(0.0, 1.0, shape)
    F = -rng.uniform(0.0, 1.0, shape)
    E = 1.0 + np.abs(D) + np.abs(F) + rng.uniform(0.0, 0.5, shape)       # diagonally dominant, like AD's rows
    TI = rng.uniform(-10.0, 30.0, shape)
    F[K + 1] = rng.uniform(-0.3, 0.3, shape[1:])                          # the reference also eliminates row KUB+1
    res0 = rng.uniform(-1.0, 1.0, shape)                                  # cells outside the solved range are kept
    return D, E, F, TI, res0


@pytest.mark.parametrize("dims", [(37, 21, 9), (70, 45, 40), (16, 16, 1), (130, 7, 75)])
def test_oracle_and_product_column_solve_match_the_reference_cuda_solver(oracle_lib, dims):
    from oracle import ref_thomas
    from oracle.oracle import OracleAdvectionDiffusion
    from mohid_b200.advection_diffusion import TransportStep
    if not ref_thomas.available():
        ref_thomas.build()
    assert ref_thomas.available(), "oracle/_ref/libcudathomas_ref.so is missing: run `make -C oracle ref` where /root/reference exists"
    I, J, K = dims
    D, E, F, TI, res0 = systems(I, J, K, seed=1000 + I)
    water = np.ones(D.shape, np.int32)

    ref = res0.copy()
    rt = ref_thomas.RefThomas(I, J, K)
    rt.solve_z(D, E, F, TI, ref)
    rt.close()

    orc = res0.copy()
    o = OracleAdvectionDiffusion(I, J, K)
    o.thomasz(D, E, F, TI, water, orc)
    assert o.zero_pivots() == 0
    o.close()

    gpu = res0.copy()
    ts = TransportStep(I, J, K)
    ts.solve_thomas_z(D, E, F, TI, gpu, water)
    assert ts.counters()["zero_pivots"] == 0
    ts.close()

    work = np.zeros(D.shape, bool)
    work[1:K + 2, 1:J + 1, 1:I + 1] = True                                # rows 1 .. KUB+1 of the work columns
    # everything outside is untouched by all three
    assert np.array_equal(ref[~work], res0[~work])
    assert np.array_equal(orc[~work], res0[~work])
    assert np.array_equal(gpu[~work], res0[~work])
    scale = np.maximum(np.abs(ref[work]), 1.0)
    e_oracle = float((np.abs(orc[work] - ref[work]) / scale).max())
    e_gpu = float((np.abs(gpu[work] - ref[work]) / scale).max())
    print(f"{dims}: oracle vs reference CUDA solver {e_oracle:.2e}, product vs reference {e_gpu:.2e}")
    # the reference kernel is compiled with FMA contraction, the oracle without, the product uses a reciprocal pivot:
    # agreement to rounding, K+1 dependent rows deep
    assert e_oracle <= 1e-13, e_oracle
    assert e_gpu <= 1e-13, e_gpu


def test_dry_columns_are_skipped_like_the_fortran_solver(oracle_lib):
    """MF:4086: only columns with WaterPoints3D(i,j,KUB) == 1 are solved (the CUDA reference solves every column; the
    product follows the Fortran, as the oracle does)."""
    from oracle.oracle import OracleAdvectionDiffusion
    from mohid_b200.advection_diffusion import TransportStep
    I, J, K = 33, 18, 12
    D, E, F, TI, res0 = systems(I, J, K, seed=7)
    rng = np.random.default_rng(3)
    water = np.ones(D.shape, np.int32)
    dry = rng.uniform(size=D.shape[1:]) < 0.3
    water[:, dry] = 0
    orc, gpu = res0.copy(), res0.copy()
    o = OracleAdvectionDiffusion(I, J, K)
    o.thomasz(D, E, F, TI, water, orc)
    o.close()
    ts = TransportStep(I, J, K)
    ts.solve_thomas_z(D, E, F, TI, gpu, water)
    ts.close()
    assert np.array_equal(gpu[:, dry], res0[:, dry])
    assert np.array_equal(orc[:, dry], res0[:, dry])
    assert float(np.abs(gpu - orc).max()) <= 1e-12
