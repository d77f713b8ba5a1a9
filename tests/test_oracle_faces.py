"""Face-weight function of the oracle against the closed forms of SURVEY.md A.5
(reference: MOHIDBase1/ModuleFunctions.F90:10702-10894, 11045-11141)."""
import numpy as np
import pytest

P4 = np.array([1.0, 2.5, 3.0, 2.0])
DU = np.array([400.0, 500.0, 550.0, 480.0])
V4 = np.array([3.0e5, 3.2e5, 2.9e5, 3.1e5])
DT = 30.0


def face(oracle_lib, q, method, lim=4, near=False, up2=True, vrel=1.5, p4=P4):
    return oracle_lib.advection_face(p4, V4, DU, DT, q, vrel, method, lim, near, up2)


def test_upwind1(oracle_lib):
    assert np.array_equal(face(oracle_lib, +5.0, 1), [0, 1, 0, 0])
    assert np.array_equal(face(oracle_lib, -5.0, 1), [0, 0, 1, 0])
    assert np.array_equal(face(oracle_lib, 0.0, 1), [0, 0, 1, 0])      # Q == 0 -> downstream slot 3 (MF:11127-11141)


def test_quick_and_volume_rel(oracle_lib):
    assert np.allclose(face(oracle_lib, +5.0, 2), [-1 / 8, 6 / 8, 3 / 8, 0], rtol=0, atol=0)
    assert np.allclose(face(oracle_lib, -5.0, 2), [0, 3 / 8, 6 / 8, -1 / 8], rtol=0, atol=0)
    # VolumeRel = max/min of the 3 upwind-side volumes > VolumeRelMax -> first order (MF:10751-10760)
    assert np.array_equal(face(oracle_lib, +5.0, 2, vrel=1.05), [0, 1, 0, 0])
    # near boundary + Upwind2 -> first order; without Upwind2 the reference stops
    assert np.array_equal(face(oracle_lib, +5.0, 2, near=True), [0, 1, 0, 0])
    with pytest.raises(ValueError):
        face(oracle_lib, +5.0, 2, near=True, up2=False)


def test_quickest(oracle_lib):
    q = 2000.0
    cr = q * DT / V4[1]
    c = (1 - 2 * abs(cr)) / 6.0
    a, b, d = 0.5 + c, 0.5 - c, (1 - abs(cr)) / 2.0
    assert np.allclose(face(oracle_lib, q, 3), [-d * b, 1 + d * (b - a), d * a, 0], rtol=1e-15)
    cr = -q * DT / V4[2]
    c = (1 - 2 * abs(cr)) / 6.0
    a, b, d = 0.5 + c, 0.5 - c, (1 - abs(cr)) / 2.0
    assert np.allclose(face(oracle_lib, -q, 3), [0, d * a, 1 + d * (b - a), -d * b], rtol=1e-15)


def test_central(oracle_lib):
    w = face(oracle_lib, 3.0, 5)
    assert np.allclose(w, [0, DU[2] / (DU[1] + DU[2]), DU[1] / (DU[1] + DU[2]), 0], rtol=1e-15)
    assert np.array_equal(face(oracle_lib, 3.0, 6), w)
    # central has no near-boundary test of its own, but near-boundary + Upwind2 wins (MF:10739)
    assert np.array_equal(face(oracle_lib, 3.0, 5, near=True, up2=True), [0, 1, 0, 0])
    assert np.array_equal(face(oracle_lib, 3.0, 5, near=True, up2=False), w)


def psi(lim, r, cr):
    if lim == 1:
        return max(0.0, min(1.0, r))
    if lim == 2:
        return 0.0 if r < 0 else 2 * r / (1 + r)
    if lim == 3:
        return max(0.0, min(2.0, 2 * r, (1 + r) / 2))
    if lim == 4:
        return max(0.0, min(1.0, 2 * r), min(r, 2.0))
    a = 0.5 + (1 - 2 * abs(cr)) / 6
    b = 0.5 - (1 - 2 * abs(cr)) / 6
    aux = a + b * r
    if abs(cr) < 1e-16:
        cr = 1e-16
    return max(0.0, min(aux, 2 / (1 - cr), 2 * r / cr))


@pytest.mark.parametrize("lim", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("q", [1500.0, -1500.0])
def test_tvd_limiters(oracle_lib, lim, q):
    rng = np.random.default_rng(lim)
    for _ in range(50):
        p4 = rng.uniform(0, 10, 4)
        if q > 0:
            cr = q * DT / V4[1]
            dC = (p4[2] - p4[1]) / (DU[2] + DU[1])
            if abs(dC) < 1e-16:
                dC = 1e-16 if dC >= 0 else -1e-16
            r = (p4[1] - p4[0]) / (DU[1] + DU[0]) / dC
        else:
            cr = q * DT / V4[2]                      # signed Courant (quirk A.4-1)
            dC = (p4[1] - p4[2]) / (DU[2] + DU[1])
            if abs(dC) < 1e-16:
                dC = 1e-16 if dC >= 0 else -1e-16
            r = (p4[2] - p4[3]) / (DU[2] + DU[3]) / dC
        ps = psi(lim, r, cr)
        if lim == 5 and abs(cr) < 1e-16:
            cr = 1e-16
        th = 0.5 * ps * (1 - cr)
        want = [0, 1 - th, th, 0] if q > 0 else [0, th, 1 - th, 0]
        got = face(oracle_lib, q, 4, lim, p4=p4)
        assert np.allclose(got, want, rtol=1e-14, atol=1e-15)


def test_tvd_dc_clamp(oracle_lib):
    p4 = np.array([1.0, 2.0, 2.0, 5.0])               # dC == 0 -> +1e-16, r huge -> superbee psi = 2
    q = 1000.0
    cr = q * DT / V4[1]
    th = 0.5 * 2.0 * (1 - cr)
    assert np.allclose(face(oracle_lib, q, 4, 4, p4=p4), [0, 1 - th, th, 0], rtol=1e-15)
