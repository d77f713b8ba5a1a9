"""ctypes front-end of oracle/_ref/libcudathomas_ref.so -- the REFERENCE's own CUDA column solver
(/root/reference/Software/CudaThomas/Thomas.cu + CudaWrapper/*, compiled unmodified by ``make -C oracle ref``).

TEST INFRASTRUCTURE ONLY (needs a GPU: the reference solver is CUDA code).  It pins row a18 of SURVEY.md section 8
(THOMASZ_NewType2, MF:4026-4123) against reference code: tests/test_gpu_ref_thomas.py compares the oracle's
restatement and the product's column solve with it.  The binding follows ModuleCuda.F90:49-121:
ConstructCudaBinding_C, InitializeThomas_C(ObjCudaID, Size), SolveThomas_C(ObjCudaID, ILB, IUB, JLB, JUB, KLB, KUB,
D, E, F, TI, Res, dimension), KillThomas_C.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libcudathomas_ref.so")
REF_SRC = "/root/reference/Software/CudaThomas/Thomas.cu"


class Size3D(C.Structure):          # T_Size3D of CudaWrapper/CuWrapperBinding.h:14-22
    _fields_ = [(n, C.c_int) for n in ("iLowerBound", "iUpperBound", "jLowerBound", "jUpperBound", "kLowerBound",
                                       "kUpperBound")]


def build(force: bool = False) -> str:
    """Compile the reference solver where the reference tree is present; elsewhere the prebuilt file is used."""
    if os.path.exists(REF_SRC) and (force or not os.path.exists(LIB)):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)
    return LIB


def available() -> bool:
    return os.path.exists(LIB)


_lib = None
_next_id = [1]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        _lib.ConstructCudaBinding_C()          # CuInit(0, ignorePitch = true): rows are not padded
    return _lib


class RefThomas:
    """One Thomas instance of the reference (Thomas::CreateInstance) for arrays (0:I+1, 0:J+1, 0:K+1)."""

    def __init__(self, I: int, J: int, K: int):
        self.I, self.J, self.K = I, J, K
        self.id = C.c_int(_next_id[0])
        _next_id[0] += 1
        size = Size3D(0, I + 1, 0, J + 1, 0, K + 1)
        lib().InitializeThomas_C(C.byref(self.id), C.byref(size))

    def solve_z(self, D, E, F, TI, res):
        """SolveThomas_C(..., dimension = Z) (Thomas.cu:24-52 -> SolveThomasZ :445-474 -> DevThomasIK :62-131); ``res``
        in place."""
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        for a in (D, E, F, TI, res):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.shape == (self.K + 2, self.J + 2, self.I + 2)
        one = C.c_int(1)
        lib().SolveThomas_C(C.byref(self.id), C.byref(one), C.byref(C.c_int(self.I)), C.byref(one),
                            C.byref(C.c_int(self.J)), C.byref(one), C.byref(C.c_int(self.K)), dp(D), dp(E), dp(F),
                            dp(TI), dp(res), C.byref(C.c_int(2)))

    def close(self):
        if self.id.value:
            lib().KillThomas_C(C.byref(self.id))
            self.id = C.c_int(0)
