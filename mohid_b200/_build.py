"""In-tree build of the CUDA extension (libmohid_adt.so) for sm_100a with nvcc.

The shared library is a plain C-ABI library (include/mohid_adt.h); it does not link torch.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmohid_adt.so")
SOURCES = [os.path.join(CSRC, "adt_api.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "adt_fused_kernel.cuh"), os.path.join(CSRC, "adt_lean_kernel.cuh"), os.path.join(CSRC, "adt_kernels.cuh"), os.path.join(CSRC, "adt_hsolve_kernel.cuh"), os.path.join(ROOT, "include", "mohid_adt.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--compiler-options", "-fPIC", "-shared"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    if os.environ.get("MOHID_ADT_NO_REBUILD"):      # measure the shipped binary, whatever the source timestamps say
        return False
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source of the package into mohid_b200/libmohid_adt.so."""
    if not force and not needs_build():
        return LIB
    # one builder at a time: the ranks of a torchrun launch all come through here
    import fcntl
    lock = open(LIB + ".lock", "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and not needs_build():            # another rank built it while this one waited
            return LIB
        return _compile(verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _compile(verbose: bool) -> str:
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [nvcc_path(), *NVCC_FLAGS, "-ccbin", ccbin, "-o", tmp, *SOURCES]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, LIB)                                # readers never see a half-written library
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
