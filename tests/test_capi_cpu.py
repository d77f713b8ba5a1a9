"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/mohid_adt.h declares, and fails loudly (no CPU fallback) without a CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "mohid_adt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(mohid_adt_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from mohid_b200 import _build, capi
    _build.build()
    return capi.load()


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("mohid_adt_create", "mohid_adt_destroy", "mohid_adt_set_grid2d", "mohid_adt_set_step",
                 "mohid_adt_advect_batch", "mohid_adt_set_discharges", "mohid_adt_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts_match_header():
    from mohid_b200 import capi
    # mohid_adt_params: 3 doubles, 6 ints, 7 doubles, 2 ints, 1 double, 4 ints
    assert C.sizeof(capi.Params) == 3 * 8 + 6 * 4 + 7 * 8 + 2 * 4 + 8 + 4 * 4
    assert C.sizeof(capi.Size3D) == 24
    assert C.sizeof(capi.Options) == 32


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from mohid_b200 import capi
    from mohid_b200.advection_diffusion import TransportStep
    with pytest.raises(capi.AdtError, match="no CUDA device|no CPU fallback|CUDA"):
        TransportStep(8, 8, 4)


def test_version_string(lib):
    buf = C.create_string_buffer(128)
    assert lib.mohid_adt_version(buf, C.byref(C.c_int(128))) == 0
    assert b"sm_100a" in buf.value


def test_product_does_not_import_the_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "mohid_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "synthetic.py" and "oracle" not in txt, (dirpath, f)


def _build_c_driver(tmp_path, name="c_driver"):
    import subprocess
    exe = str(tmp_path / name)
    libdir = os.path.join(ROOT, "mohid_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pthread", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", name + ".c"), "-L", libdir, "-lmohid_adt", f"-Wl,-rpath,{libdir}", "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_a_c_host_links(lib, tmp_path):
    """include/mohid_adt.h compiles as C99 and a C program links against the library (the Fortran shim's position);
    without a CUDA device the program stops on the library's error, not on a fallback."""
    import subprocess
    import torch
    exe = _build_c_driver(tmp_path)
    exe2 = _build_c_driver(tmp_path, "c_driver_2rank")            # two ranks, NCCL halo exchange through the C-ABI
    if not torch.cuda.is_available():
        for e in (exe, exe2):
            r = subprocess.run([e], capture_output=True, text=True)
            assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_host_runs_a_step(lib, tmp_path):
    import subprocess
    r = subprocess.run([_build_c_driver(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_c_host_two_ranks_equal_one(lib, tmp_path):
    """examples/c_driver_2rank.c: two slabs with the library's NCCL exchange == the undivided run, from plain C
    (prints `skipped` on a one-GPU box)."""
    import subprocess
    r = subprocess.run([_build_c_driver(tmp_path, "c_driver_2rank")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    print(r.stdout.strip())
