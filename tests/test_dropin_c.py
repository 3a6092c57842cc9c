"""The drop-in boundary from the reference's side: a plain C program written against the
reference's API (own CUDA context, pitched matrix - cf. reference test/lapack/cudpotrf.c) compiles
and links against libcholesky_b200.so alone (CPU check), and runs correctly on a B200 (GPU check).
The headers must also parse as C99 and as C++."""
import os
import subprocess
import tempfile

import pytest

import harness as H
from cholesky_b200_loader import load

chol = load()
ROOT = H.ROOT
INC = ["-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include"]
SRC = os.path.join(ROOT, "tests", "c", "dropin_dpotrf.c")
SRC_MGPU = os.path.join(ROOT, "tests", "c", "cumultigpu_unit.c")


def _build(exe, src=SRC):
    libdir = os.path.dirname(chol.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", *INC, src, "-o", exe, "-L" + libdir, "-lcholesky_b200",
           "-L/usr/local/cuda/lib64/stubs", "-lcuda", "-lm", "-Wl,-rpath," + libdir]
    subprocess.check_call(cmd)


def test_headers_parse_as_c_and_cpp(tmp_path):
    src = tmp_path / "h.c"
    src.write_text('#include "lapack.h"\n#include "blas.h"\n#include "cumultigpu.h"\n#include "error.h"\n'
                   '#include "cholesky_b200.h"\nint main(void) { return 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", *INC, str(src)])
    cpp = tmp_path / "h.cpp"
    cpp.write_text(src.read_text())
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", *INC, str(cpp)])


def test_c_caller_compiles_and_links(tmp_path):
    exe = str(tmp_path / "dropin")
    _build(exe)
    out = subprocess.check_output(["nm", "-u", exe]).decode()
    for sym in ("cuDpotrf", "cuDpotri", "cuDlogdet", "cuLAPACKCreate", "cuLAPACKDestroy", "cuCholGetErrorString"):
        assert sym in out  # resolved by our library at run time, nothing else provides them


@pytest.mark.gpu
def test_c_caller_runs_on_gpu(tmp_path):
    exe = str(tmp_path / "dropin")
    _build(exe)
    for n in ("300", "1111"):
        p = subprocess.run([exe, n], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=120)
        out = p.stdout.decode()
        assert p.returncode == 0 and "PASSED" in out, out


def test_multigpu_unit_program_compiles_and_links(tmp_path):
    exe = str(tmp_path / "mgpu_unit")
    _build(exe, SRC_MGPU)
    out = subprocess.check_output(["nm", "-u", exe]).decode()
    for sym in ("cuMultiGPUCreate", "cuMultiGPURunTask", "cuMultiGPUSynchronize", "cuTaskCreate", "cuTaskDestroy",
                "cuMultiGPULAPACKCreate", "cuMultiGPUDpotrf"):  # (errorHandler is data: a copy relocation, not in nm -u)
        assert sym in out


@pytest.mark.gpu
def test_multigpu_unit_program_runs_on_gpu(tmp_path):
    """reference test/multigpu/cumultigpu.c + cutask.c semantics from C, on every visible device"""
    exe = str(tmp_path / "mgpu_unit")
    _build(exe, SRC_MGPU)
    p = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    out = p.stdout.decode()
    assert p.returncode == 0 and "PASSED" in out, out
