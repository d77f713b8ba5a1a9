"""ctypes binding of the C-ABI library (include/mohid_adt.h).

This is exactly the binding a reference-side host would write (Fortran: ISO_C_BINDING, see
INTEGRATION.md).  The library is the only compute path: if it is missing, or no CUDA device is
usable, calls raise -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import _build


class Size3D(C.Structure):
    """T_Size3D (ModuleGlobalData.F90:2041-2052)."""
    _fields_ = [(n, C.c_int) for n in ("ILB", "IUB", "JLB", "JUB", "KLB", "KUB")]


class Params(C.Structure):
    """mohid_adt_params: scalar dummies of AdvectionDiffusion (ModuleAdvectionDiffusion.F90:1108-1147)."""
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


class AdtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mohid_adt error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib: Optional[C.CDLL] = None


def load(build_if_missing: bool = False) -> C.CDLL:
    """Load libmohid_adt.so (in-tree).  Fails loudly when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        _build.build()
    if not os.path.exists(_build.LIB):
        raise ImportError(f"{_build.LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA extension is the only compute path; there is no fallback)")
    # MOHID_ADT_LIB: load another build of the same library (A/B experiments with compile-time switches)
    _lib = C.CDLL(os.environ.get("MOHID_ADT_LIB") or _build.LIB)
    return _lib


def last_error(handle: Optional[C.c_int] = None) -> str:
    buf = C.create_string_buffer(1024)
    load().mohid_adt_last_error(C.byref(handle) if handle is not None else None, buf, C.byref(C.c_int(1024)))
    return buf.value.decode(errors="replace")


def check(rc: int, handle: Optional[C.c_int] = None):
    if rc != 0:
        raise AdtError(rc, last_error(handle))


def make_params(d: dict) -> Params:
    p = Params()
    for k, v in d.items():
        setattr(p, k, v)
    return p
