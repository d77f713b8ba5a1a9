"""ctypes front-end of the CPU oracle (oracle/adv_diff_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py -- never by the product package ``mohid_b200``.
PARITY UNPINNED (see the header of adv_diff_oracle.cpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmohid_oracle.so")
LIB_FAST = os.path.join(HERE, "libmohid_oracle_fast.so")      # -O3 -mavx2 build used for CPU timing


class Size3D(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("ILB", "IUB", "JLB", "JUB", "KLB", "KUB")]


class Params(C.Structure):
    _fields_ = [("Schmidt_H", C.c_double), ("SchmidtCoef_V", C.c_double), ("SchmidtBackground_V", C.c_double),
                ("AdvMethodH", C.c_int), ("TVDLimitationH", C.c_int), ("AdvMethodV", C.c_int),
                ("TVDLimitationV", C.c_int), ("Upwind2H", C.c_int), ("Upwind2V", C.c_int),
                ("VolumeRelMax", C.c_double), ("DTProp", C.c_double), ("ImpExp_AdvV", C.c_double),
                ("ImpExp_DifV", C.c_double), ("ImpExp_AdvXX", C.c_double), ("ImpExp_AdvYY", C.c_double),
                ("ImpExp_DifH", C.c_double), ("NullDif", C.c_int), ("BoundaryCondition", C.c_int),
                ("DecayTime", C.c_double), ("NoAdvFlux", C.c_int), ("NoDifFlux", C.c_int),
                ("CellFluxes", C.c_int), ("Optimize", C.c_int)]


class Options(C.Structure):
    _fields_ = [("Vertical1D", C.c_int), ("XZFlow", C.c_int), ("Docycle_method", C.c_int), ("device", C.c_int),
                ("max_properties", C.c_int), ("reserved", C.c_int * 3)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ -O2 -ffp-contract=off -fopenmp)."""
    src = os.path.join(HERE, "adv_diff_oracle.cpp")
    stale = lambda f: not os.path.exists(f) or os.path.getmtime(f) < os.path.getmtime(src)
    if force or stale(LIB) or stale(LIB_FAST):
        subprocess.run(["make", "-C", HERE, "-B", "all"], check=True, capture_output=True)
    return LIB


_lib = None


def use_fast_build(on: bool = True):
    """Switch to the -O3 -mavx2 build (timing only; must be called before the first oracle call)."""
    global _lib
    build()
    _lib = C.CDLL(LIB_FAST if on else LIB)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
    return _lib


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int))


def make_params(d: dict) -> Params:
    p = Params()
    for k, v in d.items():
        setattr(p, k, v)
    return p


class OracleAdvectionDiffusion:
    """Mirror of the ModuleAdvectionDiffusion public interface on top of the oracle library."""

    def __init__(self, I: int, J: int, K: int, ld: Optional[int] = None, *, vertical1d=False, xzflow=False,
                 docycle_method=1, nthreads: int = 0):
        self.I, self.J, self.K = I, J, K
        self.ld = ld or (I + 2)
        self._keep = {}
        size = Size3D(0, I + 1, 0, J + 1, 0, K + 1)
        work = Size3D(1, I, 1, J, 1, K)
        opt = Options(int(vertical1d), int(xzflow), docycle_method, -1, 0)
        self.h = C.c_int(0)
        rc = lib().mohid_oracle_create(C.byref(self.h), C.byref(size), C.byref(work), C.byref(C.c_int(self.ld)),
                                       C.byref(opt), C.byref(C.c_int(nthreads)))
        if rc:
            raise RuntimeError(f"mohid_oracle_create failed: {rc}")
        self.now = 0.0

    @property
    def nthreads(self) -> int:
        return lib().mohid_oracle_num_threads(C.byref(self.h))

    def _check(self, rc):
        if rc:
            buf = C.create_string_buffer(512)
            lib().mohid_oracle_last_error(C.byref(self.h), buf, C.byref(C.c_int(512)))
            raise RuntimeError(f"oracle error {rc}: {buf.value.decode()}")

    def set_grid2d(self, g: dict):
        self._keep["grid"] = g
        self._check(lib().mohid_oracle_set_grid2d(C.byref(self.h), _dp(g["DUX"]), _dp(g["DVY"]), _dp(g["DZX"]),
                                                  _dp(g["DZY"]), _ip(g["KFloorZ"]), _ip(g["BoundaryPoints2D"])))

    def set_step(self, s: dict, small_depths: Optional[np.ndarray] = None):
        self._keep["step"] = (s, small_depths)
        f = lib().mohid_oracle_set_step
        self._check(f(C.byref(self.h), _dp(s["Wflux_X"]), _dp(s["Wflux_Y"]), _dp(s["Wflux_Z"]),
                      _dp(s["VolumeZOld"]), _dp(s["VolumeZ"]), _dp(s["Visc_H"]), _dp(s["Diff_V"]), _dp(s["DWZ"]),
                      _dp(s["DZZ"]), _dp(s["AreaU"]), _dp(s["AreaV"]), _ip(s["OpenPoints3D"]),
                      _ip(s["LandPoints3D"]), _ip(s["WaterPoints3D"]), _ip(s["ComputeFacesU3D"]),
                      _ip(s["ComputeFacesV3D"]), _ip(s["ComputeFacesW3D"]), _ip(small_depths)))

    def set_noflux(self, u, v, w):
        self._keep["noflux"] = (u, v, w)
        self._check(lib().mohid_oracle_set_noflux(C.byref(self.h), _ip(u), _ip(v), _ip(w)))

    def caller_premix(self, prop, density=None, water_column=None, limit=0.0, offset=0.0):
        """FreeConvection + SmallDepthsMixing_Processes + AddOffSet of WP:14716-14759 on one property (in place);
        returns Me%SmallDepths%ON as an int32 2-D array (None when water_column is None)."""
        on = np.zeros((self.J + 2, self.ld), np.int32) if water_column is not None else None
        self._check(lib().mohid_oracle_caller_premix(C.byref(self.h), _dp(prop), _dp(density), _dp(water_column),
                                                     C.byref(C.c_double(limit)), _ip(on), C.byref(C.c_double(offset))))
        return on

    def set_limits(self, prop, min_value=None, max_value=None, mass_created=None, mass_destroyed=None):
        """SetLimitsProperty (WP:20594-20720) on one property, in place; the mass arrays accumulate."""
        shape = prop.shape
        mc = mass_created if mass_created is not None else np.zeros(shape)
        md = mass_destroyed if mass_destroyed is not None else np.zeros(shape)
        self._check(lib().mohid_oracle_set_limits(
            C.byref(self.h), _dp(prop), C.byref(C.c_int(min_value is not None)), C.byref(C.c_double(min_value or 0.0)),
            C.byref(C.c_int(max_value is not None)), C.byref(C.c_double(max_value or 0.0)), _dp(mc), _dp(md)))
        return mc, md

    def set_discharges(self, d: dict):
        nd, nc = len(d["DischnCells"]), len(d["DischFlow"])
        a = {k: np.ascontiguousarray(v, dtype=(np.float64 if k in ("DischFlow", "DischConc", "DischConcMF") else np.int32))
             for k, v in d.items()}
        self._keep["disch"] = a
        self._check(lib().mohid_oracle_set_discharges(
            C.byref(self.h), C.byref(C.c_int(nd)), C.byref(C.c_int(nc)), _dp(a["DischFlow"]), _dp(a["DischConc"]),
            _ip(a["DischI"]), _ip(a["DischJ"]), _ip(a["DischK"]), _ip(a["DischKmin"]), _ip(a["DischKmax"]),
            _ip(a["DischVert"]), _ip(a["IgnoreDisch"]), _ip(a["DischnCells"]), _ip(a["ByPass"]),
            _dp(a["DischConcMF"])))

    def unset_discharges(self):
        self._check(lib().mohid_oracle_unset_discharges(C.byref(self.h)))

    def advection_diffusion(self, prop: np.ndarray, params: dict, ref: Optional[np.ndarray] = None, *,
                            optimize=False, first_property=True):
        """One AdvectionDiffusion call (AD:1108); ``prop`` is updated in place."""
        p = make_params(params)
        self._check(lib().mohid_oracle_advection_diffusion(
            C.byref(self.h), _dp(prop), _dp(ref), C.byref(p), C.byref(C.c_int(int(optimize))),
            C.byref(C.c_int(int(first_property))), C.byref(C.c_double(self.now))))

    def advect_batch(self, props: Sequence[np.ndarray], params: Sequence[dict],
                     refs: Optional[Sequence[Optional[np.ndarray]]] = None, *, force_optimize: int = -1):
        """The WP:14580-14822 caller loop over the properties of one time step."""
        n = len(props)
        pa = (Params * n)(*[make_params(d) for d in params])
        pp = (C.POINTER(C.c_double) * n)(*[_dp(a) for a in props])
        if refs is None:
            rp = None
        else:
            rp = (C.POINTER(C.c_double) * n)(*[_dp(a) for a in refs])
        self._check(lib().mohid_oracle_advect_batch(C.byref(self.h), C.byref(C.c_int(n)), pp, rp, pa,
                                                    C.byref(C.c_double(self.now)), C.byref(C.c_int(force_optimize))))
        self.now += float(params[0]["DTProp"])

    def get_cell_fluxes(self):
        """GetAdvFlux + GetDifFlux (AD:697-851) of the last call made with CellFluxes = 1."""
        out = {}
        n3 = (self.K + 2) * (self.J + 2) * self.ld
        for w, name in enumerate(("AdvFluxX", "AdvFluxY", "AdvFluxZ", "DifFluxX", "DifFluxY", "DifFluxZ")):
            a = np.zeros((self.K + 2, self.J + 2, self.ld))
            self._check(lib().mohid_oracle_get_cell_flux(C.byref(self.h), C.byref(C.c_int(w)), _dp(a)))
            out[name] = a
        return out

    def thomasz(self, D, E, F, TI, water, res):
        """THOMASZ_NewType2 (MF:4026-4123) on caller-supplied coefficient fields; ``res`` is updated in place."""
        self._check(lib().mohid_oracle_thomasz(C.byref(self.h), _dp(D), _dp(E), _dp(F), _dp(TI), _ip(water), _dp(res)))

    def zero_pivots(self) -> int:
        n = C.c_longlong(0)
        lib().mohid_oracle_zero_pivots(C.byref(self.h), C.byref(n))
        return n.value

    def close(self):
        if self.h.value:
            lib().mohid_oracle_destroy(C.byref(self.h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def advection_face(prop4, v4, du4, dt, q, vrelmax, method, limiter, near_boundary, upwind2) -> np.ndarray:
    out = np.zeros(4)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (prop4, v4, du4)]
    rc = lib().mohid_oracle_advection_face(_dp(a[0]), _dp(a[1]), _dp(a[2]), C.byref(C.c_double(dt)),
                                           C.byref(C.c_double(q)), C.byref(C.c_double(vrelmax)),
                                           C.byref(C.c_int(method)), C.byref(C.c_int(limiter)),
                                           C.byref(C.c_int(int(near_boundary))), C.byref(C.c_int(int(upwind2))),
                                           _dp(out))
    if rc:
        raise ValueError("invalid method/limiter combination (reference would stop)")
    return out


def case_to_numpy(case):
    """Convert a mohid_b200.synthetic.Case (torch, any device) into dicts of numpy arrays."""
    g = {k: v.cpu().numpy() for k, v in case.grid2d.items()}
    s = {k: v.cpu().numpy() for k, v in case.step.items()}
    props = [p.cpu().numpy().copy() for p in case.props]
    refs = [r.cpu().numpy() for r in case.refs]
    return g, s, props, refs
