"""TEST INFRASTRUCTURE (oracle): numpy restatement of BoxDifFluxes3D, MOHIDBase2/ModuleBoxDif.F90:2659-2776, with the
boundary faces of FindAdjacentBoxesBoundaries3D (:1697-1735) and the flux sums ModuleWaterProperties feeds it
(WP:14967-14990: MassFluxes = AdvFlux + DifFlux; mask passed at WP:15001 = OpenPoints3D).  Arrays are (K+2, J+2, ld)
with element (i, j, k) at [k, j, i].  Only tests import this."""
import numpy as np


def box_dif_fluxes_3d(boxes, water, mask, flux_x, flux_y, flux_z, nboxes, I, J, K):
    """Returns F[(OUT, IN)] as a (nboxes+1, nboxes+1) array indexed [OUT, IN]."""
    F = np.zeros((nboxes + 1, nboxes + 1))
    for k in range(1, K + 1):
        for j in range(1, J + 1):
            for i in range(1, I + 1):
                b = boxes[k, j, i]
                if not (b > -55 and water[k, j, i] == 1):            # BoundaryFace3D* (BoxDif:1700-1701)
                    continue
                if mask[k, j, i] != 1:                               # BoxDif:2710, 2730, 2749
                    continue
                for (dk, dj, di, fl) in ((0, 1, 0, flux_x), (0, 0, 1, flux_y), (1, 0, 0, flux_z)):
                    if fl is None:
                        continue
                    nb = boxes[k + dk, j + dj, i + di]
                    if nb != b and nb > -55:
                        F[b, nb] = F[b, nb] + fl[k + dk, j + dj, i + di]     # BoxDif:2718, 2737, 2756
                        F[nb, b] = -F[b, nb]
    return F
