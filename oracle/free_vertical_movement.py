"""TEST INFRASTRUCTURE (oracle): numpy restatement of FreeVerticalMovementIteration,
MOHIDWater/ModuleFreeVerticalMovement.F90:1531-1650 (SetMatrixValue fills :1555-1560, BottomBoundary :2141-2227,
VerticalFreeConvection :1651-1775, CalcVerticalFreeConvFlux :1779-1805, land fill :1583-1592, THOMASZ_NewType2 through the
C++ oracle's solver).  Arrays are (K+2, J+2, ld) with element (i, j, k) at [k, j, i].  Only tests import this."""
import numpy as np

FILL = -9.9e15


def free_vertical_movement(o, conc, velocity, area, volume, mask, land, water, kfloor, I, J, K, *, dep_prob=None,
                           deposition=False, non_cohesive=False, impexp=0.0, dt=30.0):
    """`o`: an OracleAdvectionDiffusion (its thomasz is THOMASZ_NewType2).  Returns (new concentration, FreeConvFlux)."""
    vel = velocity.copy()
    for j in range(1, J + 1):                                    # BottomBoundary
        for i in range(1, I + 1):
            if mask[K, j, i] == 1:
                kb = kfloor[j, i]
                if deposition:
                    if not non_cohesive:
                        vel[kb, j, i] = vel[kb, j, i] * dep_prob[j, i]
                else:
                    vel[kb, j, i] = 0.0
    D = np.zeros_like(conc); E = np.ones_like(conc); F = np.zeros_like(conc); TI = conc.copy()
    dfl = np.zeros_like(conc); efl = np.zeros_like(conc)
    flux = np.zeros_like(conc)
    for k in range(1, K + 1):                                    # VerticalFreeConvection
        for j in range(1, J + 1):
            for i in range(1, I + 1):
                if mask[K, j, i] != 1:
                    continue
                dtv = dt / volume[k, j, i]
                w1 = vel[k, j, i] * area[j, i]
                w2 = vel[k + 1, j, i] * area[j, i] if k < K else 0.0
                aw1, aw2 = abs(w1), abs(w2)
                d_flux = -(w1 + aw1) / 2.0
                coef_d = d_flux * dtv
                e_flux = -(w1 - aw1) / 2.0
                coef_e = ((w2 + aw2) / 2.0 + e_flux) * dtv
                coef_f = ((w2 - aw2) / 2.0) * dtv
                dfl[k, j, i], efl[k, j, i] = d_flux, e_flux
                if impexp == 0.0:
                    D[k, j, i] += coef_d; E[k, j, i] += coef_e; F[k, j, i] += coef_f
                if impexp == 1.0:
                    TI[k, j, i] -= coef_d * conc[k - 1, j, i] + coef_e * conc[k, j, i] + coef_f * conc[k + 1, j, i]

    def add_flux(weight, c):                                     # CalcVerticalFreeConvFlux
        for k in range(1, K + 1):
            for j in range(1, J + 1):
                for i in range(1, I + 1):
                    if mask[k, j, i] == 1:
                        flux[k, j, i] -= weight * (dfl[k, j, i] * c[k - 1, j, i] + efl[k, j, i] * c[k, j, i])
    add_flux(impexp, conc)
    inw = np.zeros(conc.shape, bool); inw[1:K + 1, 1:J + 1, 1:I + 1] = True
    TI = np.where(inw, TI * (1.0 - land) + land * FILL, TI)
    if impexp != 1.0:
        res = conc.copy()
        o.thomasz(D, E, F, TI, water, res)
    else:
        res = TI.copy()
    add_flux(1.0 - impexp, res)
    return res, flux
