"""Settling through the library (mohid_adt_free_vertical_movement) against the restatement of
FreeVerticalMovementIteration (ModuleFreeVerticalMovement.F90:1531-1650): implicit and explicit schemes, closed bottom,
cohesive deposition with a probability field, intertidal-zone masking, the FreeConvFlux output, and the mass balance the
closed-bottom case must keep."""
import numpy as np
import pytest

from helpers import oracle_for, rel_err, water_mask
from mohid_b200.synthetic import make_case
from oracle.free_vertical_movement import free_vertical_movement

pytestmark = pytest.mark.gpu


def _setup(nprop=2):
    case = make_case(38, 30, 9, nprop=nprop, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    K, nj, ld = props[0].shape
    rng = np.random.default_rng(3)
    vel = -2.0e-3 * (0.5 + rng.random(props[0].shape))                 # sinking, m/s
    vel[:, ::7, :] *= -0.3                                             # some columns rise
    area = (g["DUX"] * g["DVY"]).astype(np.float64)
    prob = 0.2 + 0.6 * rng.random(area.shape)
    return case, o, g, s, props, vel, area, prob


@pytest.mark.parametrize("impexp,deposition,non_cohesive,intertidal", [(0.0, False, False, False), (0.0, True, False, False),
                                                                       (0.0, True, True, True), (1.0, False, False, False),
                                                                       (1.0, True, False, True)])
def test_settling_matches_the_reference_routine(oracle_lib, impexp, deposition, non_cohesive, intertidal):
    from mohid_b200.advection_diffusion import TransportStep
    case, o, g, s, props, vel, area, prob = _setup()
    ts = TransportStep(case.I, case.J, case.K, case.ld)
    ts.set_grid2d(**g)
    ts.set_step(s)
    gpu = [p.copy() for p in props]
    ts.upload(gpu)
    dt = 20.0 if impexp == 1.0 else 120.0                              # explicit: Courant number below one
    mask = s["WaterPoints3D"] if intertidal else s["OpenPoints3D"]
    w = water_mask(s)
    for n in (1, 0):                                                   # any property of the batch, in any order
        fl = ts.free_vertical_movement(n, vel, area, DepositionProbability=prob, Deposition=deposition,
                                       NonCohesive=non_cohesive, DepositionIntertidalZones=intertidal,
                                       ImpExp_AdvV=impexp, DTProp=dt, want_flux=True)
        want, want_fl = free_vertical_movement(o, props[n], vel, area, s["VolumeZ"], mask, s["LandPoints3D"],
                                               s["WaterPoints3D"], g["KFloorZ"], case.I, case.J, case.K, dep_prob=prob,
                                               deposition=deposition, non_cohesive=non_cohesive, impexp=impexp, dt=dt)
        ts.download(gpu)
        # columns the solver skips (MF:4086) keep every bit; inside solved columns the cells below the floor go through the
        # same arithmetic as water cells in the reference too (VerticalFreeConvection loops k = KLB..KUB of every column
        # whose surface cell is open), so they agree to rounding like the water cells
        solved = np.broadcast_to(s["WaterPoints3D"][case.K] == 1, w.shape)
        assert np.array_equal(gpu[n][~solved], want[~solved])
        assert rel_err(gpu[n], want, solved) <= 1e-12
        assert np.abs(want - props[n])[w].max() > 1e-3               # the step did something
        scale = np.abs(want_fl).max()
        assert scale > 0 and np.abs(fl - want_fl).max() <= 1e-12 * scale
        if not deposition and impexp == 0.0:                           # closed bottom, nothing leaves through the surface
            V = s["VolumeZ"]
            m0, m1 = (props[n][w] * V[w]).sum(), (gpu[n][w] * V[w]).sum()
            cols_open = mask[case.K] == 1
            if cols_open[w.any(axis=0)].all():
                assert abs(m1 - m0) <= 1e-12 * abs(m0)
    ts.close()


def test_settling_argument_checks():
    from mohid_b200.advection_diffusion import TransportStep
    from mohid_b200.capi import AdtError
    case, o, g, s, props, vel, area, prob = _setup(1)
    ts = TransportStep(case.I, case.J, case.K, case.ld)
    ts.set_grid2d(**g)
    ts.set_step(s)
    ts.upload([props[0].copy()])
    with pytest.raises(AdtError, match="ERR04"):                       # VerticalFreeConvection - ... - ERR04
        ts.free_vertical_movement(0, vel, area, ImpExp_AdvV=0.5)
    with pytest.raises(AdtError, match="DepositionProbability"):
        ts.free_vertical_movement(0, vel, area, Deposition=True)
    with pytest.raises(AdtError, match="never uploaded"):
        ts.free_vertical_movement(3, vel, area)
    ts.close()
