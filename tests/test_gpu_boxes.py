"""Device-side box budgets (mohid_adt_set_boxes / mohid_adt_box_fluxes) against the restatement of BoxDifFluxes3D, and
the box balance they exist for: the mass a box gains over a step equals what entered through its boundary faces."""
import numpy as np
import pytest

from helpers import oracle_for, water_mask
from mohid_b200.synthetic import make_case, default_params
from oracle.box_dif import box_dif_fluxes_3d

pytestmark = pytest.mark.gpu


def _boxes(case, nbx=3, nby=2):
    K, nj, ld = case.K + 2, case.J + 2, case.ld
    b = np.full((K, nj, ld), -9999999, np.int32)                       # null_int: no box
    jj, ii = np.meshgrid(np.arange(nj), np.arange(ld), indexing="ij")
    box2d = 1 + (np.clip((jj - 1) * nbx // case.J, 0, nbx - 1) * nby + np.clip((ii - 1) * nby // case.I, 0, nby - 1))
    box2d[(jj < 4) | (ii < 4)] = 0                                     # a strip of "environment" (box 0)
    for k in range(1, case.K + 1):
        b[k, 1:case.J + 1, 1:case.I + 1] = box2d[1:case.J + 1, 1:case.I + 1] + (nbx * nby if k > case.K // 2 else 0)
    return b, 2 * nbx * nby


@pytest.mark.parametrize("method", [1, 4])
def test_box_fluxes_match_the_reference_routine(oracle_lib, method):
    from mohid_b200.advection_diffusion import TransportStep
    case = make_case(40, 36, 8, nprop=2, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    boxes, nb = _boxes(case)
    ts = TransportStep(case.I, case.J, case.K, case.ld)
    ts.set_grid2d(**g)
    ts.set_step(s)
    ts.set_boxes(boxes, nb)
    prm = [dict(default_params(method, 4, method, 4, bc=4), CellFluxes=1) for _ in range(2)]
    gpu = [p.copy() for p in props]
    ts.advect_batch(gpu, prm, refs)
    for n in range(2):
        F = ts.box_fluxes(n).T                                         # [OUT, IN]
        fl = ts.get_cell_fluxes(n)
        want = box_dif_fluxes_3d(boxes, s["WaterPoints3D"], s["OpenPoints3D"], fl["AdvFluxX"] + fl["DifFluxX"],
                                 fl["AdvFluxY"] + fl["DifFluxY"], fl["AdvFluxZ"] + fl["DifFluxZ"], nb, case.I, case.J, case.K)
        scale = np.abs(want).max()
        assert scale > 0 and (np.abs(want) > 0).sum() >= 10
        assert np.array_equal(F == 0, want == 0)                       # the same pairs of boxes exchange
        assert np.abs(F - want).max() <= 1e-12 * scale
        assert np.abs(F + F.T).max() <= 1e-12 * scale                  # antisymmetric
    ts.close()
