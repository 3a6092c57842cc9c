"""The C host drivers (cuda-cholesky_b200/host/*.c) run end to end WITHOUT a GPU on a mock back
end: tests/mock/mock_cuda.c (the ~30 CUDA driver entry points the host code uses; "device" memory is
host memory; streams are asynchronous queues on their own threads with CUDA's event semantics, optional
random stalls and slow uploads), tests/mock/mock_kernels.cpp (the launcher ABI of
csrc/chol_kernels.h as plain loops written from the launchers' documented semantics) and
tests/mock/mock_nccl.c (libnccl.so.2 as a rendezvous of the worker threads).  The product sources are
compiled unchanged; only what they link against differs.  tests/mock/run_host_driver_checks.py then
compares every driver with the CPU oracle, on 1 and on 3 mock devices:

  blocked POTRF with look-ahead and the diagonal-block recursion, TRTRI, LAUUM, POTRI, POTRS, logdet,
  TRSM / TRMM in all side / uplo / trans / diag variants, `info` for non-PD and singular inputs,
  argument errors, context push/pop balance, allocation balance, the CUmultiGPU runtime, the
  distributed POTRF (ownership, broadcast order, block-cyclic update, page-locked overlapped
  download and - CHOL_MGPU_H2D_OVERLAP - upload paths), the slab-distributed inverse
  (CHOL_MGPU_INVERSE=slab), the host-pointer BLAS, a device allocation failure that must fail
  the call on every rank instead of leaving some of them in a broadcast, and the one-rank-per-GPU
  cholDist* form with three ranks (threads standing in for torchrun's processes) on matrices packed
  with bench.py's own layout helper.

Plain-loop kernels say nothing about the CUDA kernels: that is the job of
tests/test_kernel_emulation.py and of the GPU parity tests."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import pytest

import harness as H

ROOT = H.ROOT
MOCK = os.path.join(ROOT, "tests", "mock")
INC = ["-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include"]


@pytest.fixture(scope="module")
def mock_runs(tmp_path_factory):
    if not os.path.exists("/usr/local/cuda/include/cuda.h"):
        pytest.skip("CUDA headers not installed")
    d = str(tmp_path_factory.mktemp("mock"))
    objs = []
    for src in sorted(glob.glob(os.path.join(ROOT, "cuda-cholesky_b200", "host", "*.c"))) + [os.path.join(MOCK, "mock_cuda.c")]:
        o = os.path.join(d, os.path.basename(src)[:-2] + ".o")
        subprocess.check_call(["gcc", "-std=c99", "-O1", "-fPIC", "-D_POSIX_C_SOURCE=200809L", *INC, "-c", src, "-o", o])
        objs.append(o)
    ko = os.path.join(d, "mock_kernels.o")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-c", os.path.join(MOCK, "mock_kernels.cpp"), "-o", ko])
    lib = os.path.join(d, "libcholesky_b200_mock.so")
    subprocess.check_call(["g++", "-shared", "-o", lib, *objs, ko, "-lpthread", "-ldl", "-lm"])
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-fPIC", "-shared", os.path.join(MOCK, "mock_nccl.c"), "-o",
                           os.path.join(d, "libnccl.so.2"), "-lpthread"])
    env = dict(os.environ, LD_LIBRARY_PATH=d + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""),
               OMP_NUM_THREADS="2")  # four groups run side by side: keep the oracle's OpenMP teams small

    def run(group):
        # "<group>+<variant>+...": nopipe = panel solve on the panel stream instead of pipelined on the second one, nosplit = POTRF
        # look-ahead in one piece and no TRTRI look-ahead (the library's A/B switches, read once per process; the defaults
        # are pipelined / split), jitter = every mock stream stalls at random before its operations (ordering races)
        name, *variants = group.split("+")
        e = dict(env)
        if "nopipe" in variants:
            e["CHOL_PANEL_PIPELINE"] = "0"
        if "nosplit" in variants:
            e["CHOL_LOOKAHEAD_SPLIT"] = "0"
            e["CHOL_TRTRI_LOOKAHEAD"] = "0"
        if "jitter" in variants:
            e["MOCK_CUDA_JITTER"] = "1"
        p = subprocess.run([sys.executable, os.path.join(MOCK, "run_host_driver_checks.py"), lib, name], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, env=e, timeout=1500)
        return p.returncode, p.stdout.decode()

    def run_c(src, devices, args=()):
        """the plain-C callers of tests/c, linked against the mock library instead of libcholesky_b200 + libcuda"""
        exe = os.path.join(d, os.path.basename(src)[:-2] + "_%d" % devices)
        subprocess.check_call(["gcc", "-std=c99", "-O1", *INC, src, "-o", exe, "-L" + d, "-lcholesky_b200_mock", "-lm", "-Wl,-rpath," + d])
        p = subprocess.run([exe, *args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=dict(env, MOCK_CUDA_DEVICES=str(devices)),
                           timeout=600)
        return p.returncode, p.stdout.decode()

    groups = ("lapack", "blas", "multi1", "multi3", "lapack+nopipe", "multi3+nopipe", "lapack+jitter", "multi3+jitter",
              "lapack+nosplit+jitter", "multi3+nosplit+nopipe+jitter", "dist3+jitter", "rng+jitter")
    with ThreadPoolExecutor(max_workers=len(groups)) as ex:
        res = dict(zip(groups, ex.map(run, groups)))
    cdir = os.path.join(ROOT, "tests", "c")
    res["c:cumultigpu_unit:1"] = run_c(os.path.join(cdir, "cumultigpu_unit.c"), 1)
    res["c:cumultigpu_unit:3"] = run_c(os.path.join(cdir, "cumultigpu_unit.c"), 3)
    res["c:dropin_dpotrf"] = run_c(os.path.join(cdir, "dropin_dpotrf.c"), 1, ("300",))
    res["c:mgpu_potri_chain:1"] = run_c(os.path.join(cdir, "mgpu_potri_chain.c"), 1, ("500",))
    res["c:mgpu_potri_chain:3"] = run_c(os.path.join(cdir, "mgpu_potri_chain.c"), 3, ("700",))
    return res


@pytest.mark.parametrize("group,min_checks", [("lapack", 700), ("blas", 90), ("multi1", 800), ("multi3", 800),
                                              ("lapack+nopipe", 700), ("multi3+nopipe", 800),
                                              ("lapack+jitter", 700), ("multi3+jitter", 800),
                                              ("lapack+nosplit+jitter", 700), ("multi3+nosplit+nopipe+jitter", 800),
                                              ("dist3+jitter", 20), ("rng+jitter", 340)])
def test_host_drivers_on_mock_backend(mock_runs, group, min_checks):
    rc, out = mock_runs[group]
    assert rc == 0 and "MOCK PASSED" in out, out[-3000:]
    n = int([l for l in out.splitlines() if l.endswith(" checks")][-1].split()[0])
    assert n >= min_checks


@pytest.mark.parametrize("prog", ["c:cumultigpu_unit:1", "c:cumultigpu_unit:3", "c:dropin_dpotrf", "c:mgpu_potri_chain:1",
                                  "c:mgpu_potri_chain:3"])
def test_c_callers_on_mock_backend(mock_runs, prog):
    """tests/c/cumultigpu_unit.c (the reference's cumultigpu.c / cutask.c unit tests plus error latching and
    submission order, then cuMultiGPUDpotrf on every device), tests/c/dropin_dpotrf.c and tests/c/mgpu_potri_chain.c
    (cuMultiGPUDpotrf -> cuMultiGPUDpotri -> cuMultiGPUDgemm, both triangles, A inv(A) against I), without a GPU"""
    rc, out = mock_runs[prog]
    assert rc == 0 and "PASSED" in out, out[-3000:]
