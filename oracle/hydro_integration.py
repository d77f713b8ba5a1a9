"""TEST INFRASTRUCTURE (oracle): numpy restatement of MOHIDBase1/ModuleHydroIntegration.F90 -- ReInitalizeIntegration
(:767-796), OneIntegrationStep (:843-906), EndIntegrationStep (:910-994).  Arrays are (K+2, J+2, ld), element (i, j, k) at
[k, j, i].  Only tests import this."""
import numpy as np


class HydroIntegration:
    def __init__(self, shape, I, J, K, bnd2d):
        self.I, self.J, self.K, self.bnd = I, J, K, bnd2d
        self.shape = shape

    def reinit(self, volume_old):
        z = lambda dt: np.zeros(self.shape, dt)
        self.n = 0
        self.v0 = volume_old.copy()
        self.wx, self.wy, self.wz, self.d = z(np.float64), z(np.float64), z(np.float64), z(np.float64)
        self.cfu, self.cfv, self.cfw, self.open = z(np.int32), z(np.int32), z(np.int32), z(np.int32)

    def step(self, fx, fy, cfu, cfv, disch=None):
        I, J, K = self.I, self.J, self.K
        self.n += 1
        n = float(self.n)
        f = (slice(1, K + 1), slice(1, J + 2), slice(1, I + 2))           # faces ILB..IUB+1, JLB..JUB+1
        self.wx[f] = (self.wx[f] * (n - 1.0) + fx[f]) / n
        self.wy[f] = (self.wy[f] * (n - 1.0) + fy[f]) / n
        c = (slice(1, K + 1), slice(1, J + 1), slice(1, I + 1))
        dd = disch[c] if disch is not None else 0.0
        self.d[c] = (self.d[c] * (n - 1.0) + dd) / n
        self.cfu[f] = np.where(cfu[f] > 0, 1, self.cfu[f])
        self.cfv[f] = np.where(cfv[f] > 0, 1, self.cfv[f])

    def end(self, volume, water, dt):
        I, J, K = self.I, self.J, self.K
        for k in range(1, K + 1):
            for j in range(1, J + 1):
                for i in range(1, I + 1):
                    dvdt = (volume[k, j, i] - self.v0[k, j, i]) / dt
                    self.wz[k + 1, j, i] = (self.wz[k, j, i] + self.wx[k, j, i] - self.wx[k, j + 1, i] + self.wy[k, j, i]
                                            - self.wy[k, j, i + 1] - dvdt + self.d[k, j, i]) * (1.0 - self.bnd[j, i])
        for j in range(1, J + 1):
            for i in range(1, I + 1):
                if self.cfu[K, j, i] + self.cfu[K, j + 1, i] + self.cfv[K, j, i] + self.cfv[K, j, i + 1] > 0:
                    for k in range(2, K + 1):
                        if water[k - 1, j, i] == 1:
                            self.cfw[k, j, i] = 1
        for k in range(1, K + 1):
            for j in range(1, J + 1):
                for i in range(1, I + 1):
                    if (self.cfu[k, j, i] + self.cfu[k, j + 1, i] + self.cfv[k, j, i] + self.cfv[k, j, i + 1]
                            + self.cfw[k, j, i] + self.cfw[k + 1, j, i]) > 0:
                        self.open[k, j, i] = 1
