"""ModuleHydroIntegration on the device mirrors (mohid_adt_hydro_integration_*) against the restatement of
ReInitalizeIntegration / OneIntegrationStep / EndIntegrationStep, and a transport step that runs on the integrated inputs."""
import numpy as np
import pytest

from helpers import oracle_for, rel_err, water_mask
from mohid_b200.synthetic import make_case, default_params
from oracle.hydro_integration import HydroIntegration

pytestmark = pytest.mark.gpu


def test_integrated_fluxes_and_mapping(oracle_lib):
    from mohid_b200.advection_diffusion import TransportStep
    case = make_case(34, 29, 8, nprop=3, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    rng = np.random.default_rng(11)
    nsub, dt_hydro = 4, 7.5
    ts = TransportStep(case.I, case.J, case.K, case.ld)
    ts.set_grid2d(**g)
    ts.set_step(s)                                           # a first complete call; the integration then rewrites 9 of the mirrors
    ref = HydroIntegration(props[0].shape, case.I, case.J, case.K, g["BoundaryPoints2D"])
    ts.hydro_integration_reinit(s["VolumeZOld"])
    ref.reinit(s["VolumeZOld"])
    disch = np.zeros_like(s["Wflux_X"]); disch[3, 10, 12] = 4.0
    for m in range(nsub):
        fx = s["Wflux_X"] * (0.7 + 0.6 * rng.random(s["Wflux_X"].shape))
        fy = s["Wflux_Y"] * (0.7 + 0.6 * rng.random(s["Wflux_Y"].shape))
        cfu, cfv = s["ComputeFacesU3D"].copy(), s["ComputeFacesV3D"].copy()
        if m == 1:                                           # faces that are compute faces in one sub-step only
            cfu[:, 5:9, :] = 0
        ts.hydro_integration_step(fx, fy, cfu, cfv, disch)
        ref.step(fx, fy, cfu, cfv, disch)
    ts.hydro_integration_end(s["VolumeZ"], s["WaterPoints3D"], nsub * dt_hydro)
    ref.end(s["VolumeZ"], s["WaterPoints3D"], nsub * dt_hydro)
    work = (slice(1, case.K + 2), slice(1, case.J + 2), slice(1, case.I + 2))
    for which, want in ((0, ref.wx), (1, ref.wy), (2, ref.wz)):
        got = ts.step_input(which)
        scale = np.abs(want).max()
        assert scale > 0 and np.abs(got[work] - want[work]).max() <= 1e-13 * scale, which
    assert np.array_equal(ts.step_input(3), ref.v0)
    for which, want in ((14, ref.cfu), (15, ref.cfv), (16, ref.cfw), (11, ref.open)):
        assert np.array_equal(ts.step_input(which)[work], want[work]), which
    assert ref.open.sum() > 1000 and ref.cfw.sum() > 1000
    # a transport step on the integrated inputs == the oracle's step on the restatement's arrays
    s2 = dict(s, Wflux_X=ref.wx, Wflux_Y=ref.wy, Wflux_Z=ref.wz, VolumeZOld=ref.v0, ComputeFacesU3D=ref.cfu,
              ComputeFacesV3D=ref.cfv, ComputeFacesW3D=ref.cfw, OpenPoints3D=ref.open)
    o.set_step(s2)
    prm = [default_params(4, 4, 4, 4, bc=4, dt=nsub * dt_hydro) for _ in range(3)]
    cpu = [p.copy() for p in props]
    o.advect_batch(cpu, prm, refs)
    gpu = [p.copy() for p in props]
    ts.upload(gpu, refs)
    ts.advect_device(prm, 1)
    ts.download(gpu)
    w = water_mask(s)
    for a, b in zip(gpu, cpu):
        assert np.array_equal(a[~w], b[~w])
        assert rel_err(a, b, w) <= 5e-12
    ts.close()
