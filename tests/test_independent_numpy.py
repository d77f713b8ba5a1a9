"""An independent, equation-level restatement of the transport step in plain numpy (all six advection methods, P2_TVD with
the five limiters, explicit horizontal terms, theta-weighted vertical diffusion, implicit or explicit vertical advection,
the NullGradient, MassConservation and ImposedValue open boundaries, dense column solve with numpy.linalg.solve)
checked against the C++ oracle.

It shares no code and no structure with oracle/adv_diff_oracle.cpp: it is written from the discrete equations of
SURVEY.md A.3 / A.5 (cell-wise assembly of one matrix per column instead of the reference's pass-by-pass scatter into
D/E/F/TI arrays), so an error in the oracle's bookkeeping (pass order, index shifts, sign conventions) shows up here.
"""
import numpy as np
import pytest

from mohid_b200.synthetic import make_case, default_params
from helpers import oracle_for, water_mask, NULL_REAL


def superbee_theta(q, Puu, Pu, Pd, du_uu, du_u, du_d, dt_over_vu, second_upwind_open, limiter=4, method=4):
    """Weight of the downwind value in the face value, theta = psi(r) (1 - Cr) / 2 (MF:10785-10858): r compares the
    upwind gradient with the face gradient (distance-weighted), Cr = Q DT / V_upwind keeps the sign of Q (quirk A.4),
    faces whose second upwind cell is not an open point fall back to first order (Upwind2)."""
    if not second_upwind_open:
        return 0.0
    if method in (5, 6):                               # central differences / leap-frog: linear interpolation to the face (MF:10773-10783)
        return du_u / (du_u + du_d)
    dc = (Pd - Pu) / (du_u + du_d)
    if abs(dc) < 1e-16:
        dc = 1e-16 if dc >= 0 else -1e-16
    r = ((Pu - Puu) / (du_u + du_uu)) / dc
    cr = q * dt_over_vu
    if limiter == 1:                                   # MinMod (Roe 1986)
        psi = max(0.0, min(1.0, r))
    elif limiter == 2:                                 # van Leer (1974)
        psi = 0.0 if r < 0 else 2.0 * r / (1.0 + r)
    elif limiter == 3:                                 # MUSCL / monotonized central (van Leer 1977)
        psi = max(0.0, min(2.0, 2.0 * r, 0.5 * (1.0 + r)))
    elif limiter == 4:                                 # SuperBee (Roe 1986)
        psi = max(0.0, min(2.0 * r, 1.0), min(r, 2.0))
    else:                                              # PDM: third-order flux bounded by the universal limiter (MF:10842-10853)
        a, b = 0.5 + (1.0 - 2.0 * abs(cr)) / 6.0, 0.5 - (1.0 - 2.0 * abs(cr)) / 6.0
        if abs(cr) < 1e-16:
            cr = 1e-16
        psi = max(0.0, min(a + b * r, 2.0 / (1.0 - cr), 2.0 * r / cr))
    return 0.5 * psi * (1.0 - cr)


def three_point_value(method, q, Puu, Pu, Pd, Vuu, Vu, Vd, dt, second_upwind_open, vrelmax=1.5):
    """Explicit face value of the upwind-biased three-point schemes: method 2 the QUICK weights -1/8, 6/8, 3/8
    (MF:11060-11083), method 3 QUICKEST with the Courant number of the upwind cell (MF:11085-11122).  First-order upwind next
    to a closed cell (Upwind2) and where the three volumes differ by more than VolumeRelMax (MF:10746-10768)."""
    if not second_upwind_open or max(Vuu, Vu, Vd) / min(Vuu, Vu, Vd) > vrelmax:
        return Pu
    if method == 2:
        return -0.125 * Puu + 0.75 * Pu + 0.375 * Pd
    cr = q * dt / Vu
    c = (1.0 - 2.0 * abs(cr)) / 6.0
    a, b, d = 0.5 + c, 0.5 - c, (1.0 - abs(cr)) / 2.0
    return -d * b * Puu + (1.0 + d * (b - a)) * Pu + d * a * Pd


def numpy_step(g, s, P, dt, theta, schmidt_h=1.0, coef_v=1.0, bg_v=1.0e-8, tvd=False, limiter=4, advv_implicit=True,
               null_gradient=False, bc=0, ref=None, decay_time=0.0, vertical_only=False, method=4, method_v=None,
               cyclic=False):
    """One step of one property; arrays are (K+2, J+2, ld) / (J+2, ld), index order [k, j, i]."""
    K, J, I = P.shape[0] - 2, P.shape[1] - 2, g["_I"]         # the i extent may be padded: the work size comes along
    method_v = method if method_v is None else method_v
    Open, Water, Land = s["OpenPoints3D"], s["WaterPoints3D"], s["LandPoints3D"]
    CFU, CFV, CFW = s["ComputeFacesU3D"], s["ComputeFacesV3D"], s["ComputeFacesW3D"]
    V, Vold = s["VolumeZ"], s["VolumeZOld"]
    Qx, Qy, Qz = s["Wflux_X"], s["Wflux_Y"], s["Wflux_Z"]
    DUX, DVY, DZX, DZY = g["DUX"], g["DVY"], g["DZX"], g["DZY"]
    out = P.copy()

    def hflux(k, j, i, dj, di):
        """Flux of property through the low face of cell (k,j,i) in direction (dj,di), positive towards the cell,
        split as (advective, diffusive)."""
        jm, im = j - dj, i - di
        if dj:
            if CFU[k, j, i] != 1:
                return 0.0, 0.0
            q, area, dz = Qx[k, j, i], s["AreaU"][k, j, i], DZX[jm, im]
            dif = schmidt_h * (s["Visc_H"][k, j, i] * DUX[jm, im] + s["Visc_H"][k, jm, im] * DUX[j, i]) / (DUX[j, i] + DUX[jm, im])
        else:
            if CFV[k, j, i] != 1:
                return 0.0, 0.0
            q, area, dz = Qy[k, j, i], s["AreaV"][k, j, i], DZY[jm, im]
            dif = schmidt_h * (s["Visc_H"][k, j, i] * DVY[jm, im] + s["Visc_H"][k, jm, im] * DVY[j, i]) / (DVY[j, i] + DVY[jm, im])
        adv = 0.0
        if Open[k, jm, im] == 1 and Open[k, j, i] == 1:
            adv = q * (P[k, jm, im] if q > 0 else P[k, j, i])
            if tvd:
                du = DUX if dj else DVY
                # cells along the direction: a-2, a-1 | a, a+1 around the face
                c = [(j - 2 * dj, i - 2 * di), (jm, im), (j, i), (j + dj, i + di)]
                c[0] = (max(c[0][0], 0), max(c[0][1], 0))
                if q > 0:
                    uu, u, d = c[0], c[1], c[2]
                else:
                    uu, u, d = c[3], c[2], c[1]
                if method in (2, 3):
                    adv = q * three_point_value(method, q, P[k][uu], P[k][u], P[k][d], V[k][uu], V[k][u], V[k][d], dt, Open[k][uu] == 1)
                else:
                    th = superbee_theta(q, P[k][uu], P[k][u], P[k][d], du[uu], du[u], du[d], dt / V[k][u], Open[k][uu] == 1, limiter, method)
                    adv = q * ((1.0 - th) * P[k][u] + th * P[k][d])
        difflux = -dif * area / dz * (P[k, j, i] - P[k, jm, im])
        return adv, difflux

    for j in range(1, J + 1):
        for i in range(1, I + 1):
            if Water[K, j, i] != 1:
                continue
            n = K + 1                                   # unknowns k = 1 .. K+1 (the last one is the identity halo row)
            A = np.zeros((n, n))
            b = np.zeros(n)
            for k in range(1, K + 1):
                r = k - 1
                dtv = dt / V[k, j, i]
                is_open = Open[k, j, i] == 1
                b[r] = P[k, j, i] * (Vold[k, j, i] / V[k, j, i]) if (is_open and not vertical_only) else P[k, j, i]
                A[r, r] = 1.0
                if is_open and k == K and not vertical_only:
                    A[r, r] += dtv * Qz[K + 1, j, i]
                # horizontal: inflow through the low faces, outflow through the high faces of the cell
                for dj, di in (() if vertical_only else ((1, 0), (0, 1))):
                    a_lo, d_lo = hflux(k, j, i, dj, di)
                    a_hi, d_hi = hflux(k, j + dj, i + di, dj, di) if (j + dj <= J + 1 and i + di <= I + 1) else (0.0, 0.0)
                    if dj and j + 1 > J:
                        a_hi, d_hi = 0.0, 0.0            # face loops of the reference end at JUB / IUB
                    if di and i + 1 > I:
                        a_hi, d_hi = 0.0, 0.0
                    b[r] += (a_lo - a_hi) * dtv + (d_lo - d_hi) * dtv
                # vertical faces: bottom (k) and top (k+1)
                for kf, sign in ((k, +1.0), (k + 1, -1.0)):
                    if kf < 2 or kf > K or CFW[kf, j, i] != 1:
                        continue
                    lo, hi = kf - 1, kf                    # cells below / above the face
                    if True:
                        difz = coef_v * s["Diff_V"][kf, j, i] + bg_v
                        auxk = difz * DUX[j, i] * DVY[j, i] / s["DZZ"][lo, j, i]
                        # flux upwards through the face = -auxk (P_hi - P_lo); implicit share theta
                        other = lo if k == hi else hi
                        A[r, r] += theta * auxk * dtv
                        A[r, other - 1] -= theta * auxk * dtv
                        b[r] += (1.0 - theta) * auxk * dtv * (P[other, j, i] - P[k, j, i])
                    if Open[lo, j, i] == 1 and Open[hi, j, i] == 1 and Open[K, j, i] == 1:
                        q = Qz[kf, j, i]
                        up, dn = (lo, hi) if q > 0 else (hi, lo)   # implicit: flux = q ((1-th) P_up + th P_dn)^{n+1}
                        th = 0.0
                        uu = min(max(up + (up - dn), 0), K + 1)
                        if tvd and method_v not in (1, 2, 3):
                            dwz = s["DWZ"]
                            th = superbee_theta(q, P[uu, j, i], P[up, j, i], P[dn, j, i], dwz[uu, j, i], dwz[up, j, i],
                                                dwz[dn, j, i], dt / V[up, j, i], Open[uu, j, i] == 1, limiter, method_v)
                        if tvd and method_v in (2, 3):             # explicit only (AD:1229-1237 stops the implicit use)
                            assert not advv_implicit
                            b[r] += sign * q * dtv * three_point_value(method_v, q, P[uu, j, i], P[up, j, i], P[dn, j, i], V[uu, j, i],
                                                                       V[up, j, i], V[dn, j, i], dt, Open[uu, j, i] == 1)
                        elif advv_implicit:
                            A[r, up - 1] -= sign * q * dtv * (1.0 - th)
                            A[r, dn - 1] -= sign * q * dtv * th
                        else:                                      # the same face value from the field at time n
                            b[r] += sign * q * dtv * ((1.0 - th) * P[up, j, i] + th * P[dn, j, i])
                if (null_gradient or cyclic) and g["BoundaryPoints2D"][j, i] == 1 and is_open:
                    A[r, :] = 0.0; A[r, r] = 1.0; b[r] = P[k, j, i]        # kept through the solve, replaced below
                if bc in (1, 2, 7) and g["BoundaryPoints2D"][j, i] == 1 and is_open:
                    Bnd = g["BoundaryPoints2D"]
                    tdec = 1.0 / (1.0 + decay_time / dt)                   # relaxation towards the reference field (AD:5418-5419)
                    if bc == 2:
                        # ImposedValue (AD:5440-5500): the cell takes the mean of its interior neighbours at time n, relaxed
                        nb = [(j, i + 1), (j, i - 1), (j + 1, i), (j - 1, i)]
                        inner = [P[k, jj, ii] for jj, ii in nb if Open[k, jj, ii] == 1 and Bnd[jj, ii] != 1]
                        ext = ref[k, j, i] if not inner else (sum(inner) / len(inner)) * (1.0 - tdec) + ref[k, j, i] * tdec
                        A[r, :] = 0.0; A[r, r] = 1.0; b[r] = ext
                    else:
                        # MassConservation (AD:5572-5672): the water the compute faces and the volume change do not account
                        # for crosses the open boundary: leaving, it takes the cell's (new) value along; entering, it
                        # brings the exterior value
                        qb = (Qx[k, j, i] * (CFU[k, j, i] == 1) - Qx[k, j + 1, i] * (CFU[k, j + 1, i] == 1)
                              + Qy[k, j, i] * (CFV[k, j, i] == 1) - Qy[k, j, i + 1] * (CFV[k, j, i + 1] == 1)
                              + Qz[k, j, i] * (CFW[k, j, i] == 1) - Qz[k + 1, j, i] * (CFW[k + 1, j, i] == 1)
                              - (V[k, j, i] - Vold[k, j, i]) / dt)
                        if qb < 0 and bc == 1:
                            b[r] -= qb * dtv * (P[k, j, i] * (1.0 - tdec) + ref[k, j, i] * tdec)
                        elif qb < 0:
                            # MassConservNullGrad (AD:5610-5640): an inflow cell takes the mean of the old field across its
                            # compute faces instead
                            nb = [(CFV[k, j, i + 1], P[k, j, i + 1]), (CFV[k, j, i], P[k, j, i - 1]),
                                  (CFU[k, j + 1, i], P[k, j + 1, i]), (CFU[k, j, i], P[k, j - 1, i])]
                            vals = [v for cf, v in nb if cf == 1]
                            A[r, :] = 0.0; A[r, r] = 1.0; b[r] = sum(vals) / len(vals) if vals else P[k, j, i]
                        else:
                            A[r, r] += qb * dtv
                if Land[k, j, i] == 1:
                    A[r, :] = 0.0; A[r, r] = 1.0; b[r] = NULL_REAL
            A[n - 1, n - 1] = 1.0
            x = np.linalg.solve(A, b)
            out[1:K + 2, j, i] = x
    if null_gradient:
        # boundary cells take the mean of the new values across their compute faces (AD:1926-1987); a compute face never
        # joins two boundary points, so the order does not matter
        new = out.copy()
        for j in range(1, J + 1):
            for i in range(1, I + 1):
                if g["BoundaryPoints2D"][j, i] != 1:
                    continue
                for k in range(abs(int(g["KFloorZ"][j, i])), K + 1):
                    nb = [(CFU[k, j, i], out[k, j - 1, i]), (CFU[k, j + 1, i], out[k, j + 1, i]),
                          (CFV[k, j, i], out[k, j, i - 1]), (CFV[k, j, i + 1], out[k, j, i + 1])]
                    wsum = sum(1 for c, _ in nb if c == 1)
                    if wsum > 0:
                        new[k, j, i] = sum(v for c, v in nb if c == 1) / wsum
        out = new
    if cyclic:
        # Prop_CyclicBoundary (AD:2121-2224): boundary cells <- reference field, then the opposite edges are joined: each
        # boundary column / row takes the values next to the opposite one, from that column's bottom layer up
        Bnd, kf = g["BoundaryPoints2D"], g["KFloorZ"]
        for j in range(1, J + 1):
            for i in range(1, I + 1):
                if Bnd[j, i] == 1:
                    out[1:K + 1, j, i] = ref[1:K + 1, j, i]
        for i in range(2, I):
            if Bnd[1, i] == 1 and Bnd[J, i] == 1:
                out[kf[J - 1, i]:K + 1, 1, i] = out[kf[J - 1, i]:K + 1, J - 1, i]
                out[kf[2, i]:K + 1, J, i] = out[kf[2, i]:K + 1, 2, i]
        for j in range(2, J):
            if Bnd[j, 1] == 1 and Bnd[j, I] == 1:
                out[kf[j, I - 1]:K + 1, j, 1] = out[kf[j, I - 1]:K + 1, j, I - 1]
                out[kf[j, 2]:K + 1, j, I] = out[kf[j, 2]:K + 1, j, 2]
    return out


@pytest.mark.parametrize("limiter", [1, 2, 3, 5])
@pytest.mark.parametrize("advv", [1.0, 0.0])
def test_oracle_matches_equation_level_numpy_limiters_and_explicit_vertical_advection(oracle_lib, limiter, advv):
    """The other four limiters (MF:10823-10853) and the explicit form of the vertical advection (AD:3148-3207)."""
    case = make_case(13, 11, 5, nprop=1, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    g = dict(g); g["_I"] = case.I
    prm = [default_params(4, limiter, 4, limiter, theta_difv=0.6, impexp_advv=advv)]
    a = [props[0].copy()]
    o.advect_batch(a, prm)
    want = numpy_step(g, s, props[0], case.dt, 0.6, tvd=True, limiter=limiter, advv_implicit=advv == 1.0)
    w = water_mask(s)
    scale = np.abs(props[0][w]).max()
    assert np.abs(a[0] - want)[w].max() <= 1e-11 * scale
    assert np.array_equal(a[0][~w], want[~w])


@pytest.mark.parametrize("tvd", [False, True])
def test_oracle_matches_equation_level_numpy_null_gradient_boundary(oracle_lib, tvd):
    """BoundaryCondition = NullGradient (AD:5369-5400 rows, AD:1926-1987 post pass)."""
    case = make_case(13, 11, 5, nprop=1, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    g = dict(g); g["_I"] = case.I
    m = 4 if tvd else 1
    prm = [default_params(m, 4, m, 4, bc=4)]
    a = [props[0].copy()]
    o.advect_batch(a, prm, refs)
    want = numpy_step(g, s, props[0], case.dt, 1.0, tvd=tvd, null_gradient=True)
    w = water_mask(s)
    scale = np.abs(props[0][w]).max()
    assert (g["BoundaryPoints2D"] == 1).sum() > 0 and not np.array_equal(want, numpy_step(g, s, props[0], case.dt, 1.0, tvd=tvd))
    assert np.abs(a[0] - want)[w].max() <= 1e-11 * scale
    assert np.array_equal(a[0][~w], want[~w])


@pytest.mark.parametrize("decay", [0.0, 900.0])
@pytest.mark.parametrize("bc", [1, 2, 7])
def test_oracle_matches_equation_level_numpy_flux_and_value_boundaries(oracle_lib, bc, decay):
    """BoundaryCondition = MassConservation (1), ImposedValue (2) and MassConservNullGrad (7), with and without relaxation
    (DecayTime)."""
    case = make_case(13, 11, 5, nprop=1, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    g = dict(g); g["_I"] = case.I
    prm = [default_params(4, 4, 4, 4, bc=bc, decay_time=decay)]
    a = [props[0].copy()]
    o.advect_batch(a, prm, refs)
    want = numpy_step(g, s, props[0], case.dt, 1.0, tvd=True, bc=bc, ref=refs[0], decay_time=decay)
    w = water_mask(s)
    scale = np.abs(props[0][w]).max()
    assert not np.array_equal(want, numpy_step(g, s, props[0], case.dt, 1.0, tvd=True))
    assert np.abs(a[0] - want)[w].max() <= 1e-11 * scale
    assert np.array_equal(a[0][~w], want[~w])


@pytest.mark.parametrize("theta", [1.0, 0.4])
@pytest.mark.parametrize("tvd", [False, True])
def test_oracle_matches_equation_level_numpy(oracle_lib, theta, tvd):
    case = make_case(13, 11, 5, nprop=1, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    g = dict(g); g["_I"] = case.I
    m = 4 if tvd else 1
    prm = [default_params(m, 4, m, 4, theta_difv=theta)]
    a = [props[0].copy()]
    o.advect_batch(a, prm)
    want = numpy_step(g, s, props[0], case.dt, theta, tvd=tvd)
    w = water_mask(s)
    assert np.array_equal(a[0] == NULL_REAL, want == NULL_REAL)
    scale = np.abs(props[0][w]).max()
    assert np.abs(a[0] - want)[w].max() <= 1e-11 * scale
    assert np.array_equal(a[0][~w], want[~w])


def numpy_lines_implicit(g, s, P, dt, tvd, direction="xx"):
    """ImpExp_AdvXX = 1 or ImpExp_AdvYY = 1: one dense system per line of the implicit direction and level.  Advection along
    the line is implicit (face value from the new field, weights from the old one), the other horizontal terms and the
    volume change explicit; the surface layer's row carries the water flux through its top face.  On a 2-D domain
    (K = 1, AD:1758-1841) this is the whole step, in 3-D (AD:4132-4265) the first half: numpy_step(vertical_only=True)
    continues from its result."""
    out = P.copy()
    K = P.shape[0] - 2
    for k in range(1, K + 1):
        _lines_of_level(g, s, P, dt, tvd, direction, k, K, out)
    return out


def _lines_of_level(g, s, P, dt, tvd, direction, k, K, out):
    J, I = P.shape[1] - 2, g["_I"]
    Open, Land = s["OpenPoints3D"], s["LandPoints3D"]
    V, Vold, Qz = s["VolumeZ"], s["VolumeZOld"], s["Wflux_Z"]
    xx = direction == "xx"
    NL, NC = (J, I) if xx else (I, J)                           # cells along / across the lines
    at = (lambda a, c: (a, c)) if xx else (lambda a, c: (c, a))  # (along, across) -> (j, i)
    # along the line / across it: compute faces, face flows, metrics, face areas
    CFL, QL, DUL, DZL, AL = ((s["ComputeFacesU3D"], s["Wflux_X"], g["DUX"], g["DZX"], s["AreaU"]) if xx else
                             (s["ComputeFacesV3D"], s["Wflux_Y"], g["DVY"], g["DZY"], s["AreaV"]))
    CFC, QC, DUC, DZC, AC = ((s["ComputeFacesV3D"], s["Wflux_Y"], g["DVY"], g["DZY"], s["AreaV"]) if xx else
                             (s["ComputeFacesU3D"], s["Wflux_X"], g["DUX"], g["DZX"], s["AreaU"]))

    def face_weights(q, cells, du):
        """(cell, weight) pairs of the face value of a face with flow q; cells = a-2, a-1 | a, a+1 around the face."""
        uu, u, d = (cells[0], cells[1], cells[2]) if q > 0 else (cells[3], cells[2], cells[1])
        th = superbee_theta(q, P[k][uu], P[k][u], P[k][d], du[uu], du[u], du[d], dt / V[k][u], Open[k][uu] == 1) if tvd else 0.0
        return ((u, 1.0 - th), (d, th))

    def dif_coef(hi, lo, du, dz, area):
        nu = (s["Visc_H"][k][hi] * du[lo] + s["Visc_H"][k][lo] * du[hi]) / (du[hi] + du[lo])
        return nu * area[k][hi] / dz[lo]

    for c in range(1, NC + 1):
        A = np.eye(NL + 2)
        b = np.array([P[k][at(a, c)] for a in range(NL + 2)])
        b[NL + 1] = 0.0                                         # the halo cell behind the line: identity row, 0 (MF:3803)
        for a in range(1, NL + 1):
            me = at(a, c)
            if Land[k][me] == 1 and K == 1:                        # a 3-D step fills the land cells in its vertical half
                b[a] = NULL_REAL
                continue
            if Open[k][me] != 1:
                continue
            dtv = dt / V[k][me]
            b[a] = P[k][me] * Vold[k][me] / V[k][me]
            if k == K:
                A[a, a] += dtv * Qz[(k + 1,) + me]
            for af, sign in ((a, +1.0), (a + 1, -1.0)):          # faces along the line: low (inflow positive) and high
                hi, lo = at(af, c), at(af - 1, c)
                if af > NL or CFL[k][hi] != 1:
                    continue
                b[a] += sign * dtv * (-dif_coef(hi, lo, DUL, DZL, AL) * (P[k][hi] - P[k][lo]))
                if Open[k][lo] == 1 and Open[k][hi] == 1:
                    cells = [at(max(af - 2, 0), c), lo, hi, at(af + 1, c)]
                    pos = {cells[1]: af - 1, cells[2]: af}
                    for cell, wgt in face_weights(QL[k][hi], cells, DUL):
                        A[a, pos[cell]] -= sign * dtv * QL[k][hi] * wgt
            for cf, sign in ((c, +1.0), (c + 1, -1.0)):          # faces across the line: explicit
                hi, lo = at(a, cf), at(a, cf - 1)
                if cf > NC or CFC[k][hi] != 1:
                    continue
                b[a] += sign * dtv * (-dif_coef(hi, lo, DUC, DZC, AC) * (P[k][hi] - P[k][lo]))
                if Open[k][lo] == 1 and Open[k][hi] == 1:
                    cells = [at(a, max(cf - 2, 0)), lo, hi, at(a, cf + 1)]
                    q = QC[k][hi]
                    b[a] += sign * dtv * q * sum(wgt * P[k][cell] for cell, wgt in face_weights(q, cells, DUC))
        x = np.linalg.solve(A[1:, 1:], b[1:])
        for a in range(1, NL + 2):
            out[(k,) + at(a, c)] = x[a - 1]


@pytest.mark.parametrize("direction", ["xx", "yy"])
@pytest.mark.parametrize("tvd", [False, True])
def test_oracle_2d_implicit_line_solve_matches_equation_level_numpy(oracle_lib, tvd, direction):
    case = make_case(12, 14, 1, nprop=1)
    o, g, s, props, refs = oracle_for(case)
    g = dict(g); g["_I"] = case.I
    m = 4 if tvd else 1
    prm = [dict(default_params(m, 4, m, 4), **{"ImpExp_Adv" + direction.upper(): 1.0})]
    a = [props[0].copy()]
    o.advect_batch(a, prm)
    want = numpy_lines_implicit(g, s, props[0], case.dt, tvd, direction)
    w = water_mask(s)
    scale = np.abs(props[0][w]).max()
    assert np.abs(a[0] - want)[w].max() <= 1e-11 * scale
    assert np.array_equal((a[0] == NULL_REAL)[1], (want == NULL_REAL)[1])


@pytest.mark.parametrize("direction", ["xx", "yy"])
@pytest.mark.parametrize("tvd", [False, True])
def test_oracle_split_implicit_step_matches_equation_level_numpy(oracle_lib, tvd, direction):
    """3-D, one horizontal direction implicit (AD:4132-4265): line systems per level, then the vertical half restarted from
    their result (D = 0, E = 1, F = 0, TI = PROP; vertical weights from the intermediate field)."""
    case = make_case(12, 13, 4, nprop=1, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    g = dict(g); g["_I"] = case.I
    m = 4 if tvd else 1
    prm = [dict(default_params(m, 4, m, 4, theta_difv=0.7), **{"ImpExp_Adv" + direction.upper(): 1.0})]
    a = [props[0].copy()]
    o.advect_batch(a, prm)
    mid = numpy_lines_implicit(g, s, props[0], case.dt, tvd, direction)
    want = numpy_step(g, s, mid, case.dt, 0.7, tvd=tvd, vertical_only=True)
    w = water_mask(s)
    scale = np.abs(props[0][w]).max()
    assert np.abs(a[0] - want)[w].max() <= 1e-11 * scale
    assert np.array_equal(a[0] == NULL_REAL, want == NULL_REAL)


@pytest.mark.parametrize("method", [5, 6])
@pytest.mark.parametrize("advv", [1.0, 0.0])
def test_oracle_matches_equation_level_numpy_central_differences(oracle_lib, advv, method):
    """CentralDif and LeapFrog (the same weights, MF:10773-10783): the face value is the distance-weighted mean of the two cells, first-order upwind
    next to a closed cell (Upwind2); horizontally explicit, vertically implicit or explicit."""
    case = make_case(13, 11, 5, nprop=1, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    g = dict(g); g["_I"] = case.I
    prm = [default_params(method, 4, method, 4, theta_difv=0.5, impexp_advv=advv)]
    a = [props[0].copy()]
    o.advect_batch(a, prm)
    want = numpy_step(g, s, props[0], case.dt, 0.5, tvd=True, method=method, advv_implicit=advv == 1.0)
    w = water_mask(s)
    scale = np.abs(props[0][w]).max()
    assert np.abs(a[0] - want)[w].max() <= 1e-11 * scale
    assert np.array_equal(a[0][~w], want[~w])


@pytest.mark.parametrize("method,method_v,advv", [(2, 2, 0.0), (3, 3, 0.0), (2, 1, 1.0), (3, 1, 1.0)])
def test_oracle_matches_equation_level_numpy_three_point_upwind(oracle_lib, method, method_v, advv):
    """UpwindOrder2 (QUICK weights) and UpwindOrder3 (QUICKEST): explicit three-point face values with the VolumeRelMax
    and near-boundary fall-backs to first order; vertically the same explicitly, or first-order upwind implicitly."""
    case = make_case(13, 11, 5, nprop=1, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    g = dict(g); g["_I"] = case.I
    prm = [default_params(method, 4, method_v, 4, theta_difv=0.5, impexp_advv=advv)]
    a = [props[0].copy()]
    o.advect_batch(a, prm)
    want = numpy_step(g, s, props[0], case.dt, 0.5, tvd=True, method=method, method_v=method_v, advv_implicit=advv == 1.0)
    first = numpy_step(g, s, props[0], case.dt, 0.5, tvd=False, advv_implicit=advv == 1.0)
    w = water_mask(s)
    scale = np.abs(props[0][w]).max()
    assert np.abs(want - first)[w].max() > 1e-6 * scale             # the higher order really acted
    assert np.abs(a[0] - want)[w].max() <= 1e-11 * scale
    assert np.array_equal(a[0][~w], want[~w])


def test_oracle_matches_equation_level_numpy_cyclic_boundary(oracle_lib):
    case = make_case(13, 11, 5, nprop=1, stepped_bottom=True)
    o, g, s, props, refs = oracle_for(case)
    g = dict(g); g["_I"] = case.I
    a = [props[0].copy()]
    o.advect_batch(a, [default_params(4, 4, 4, 4, bc=8)], refs)
    want = numpy_step(g, s, props[0], case.dt, 1.0, tvd=True, cyclic=True, ref=refs[0])
    w = water_mask(s)
    scale = np.abs(props[0][w]).max()
    assert not np.array_equal(want, numpy_step(g, s, props[0], case.dt, 1.0, tvd=True, null_gradient=True))
    assert np.abs(a[0] - want)[w].max() <= 1e-11 * scale
    assert np.array_equal(a[0][~w], want[~w])
