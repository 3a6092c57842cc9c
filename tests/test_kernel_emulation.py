"""Kernel LOGIC on the CPU: the CUDA sources of the diagonal-block kernels (csrc/panel_kernels.cu:
potf2 with its in-kernel look-ahead, the substitution kernel trsm_rows, trti2, lauu2), of the small-tile FP64 GEMM
(csrc/gemm_small.cu), of the FP32 GEMM engine (csrc/gemm_ffma.cu) and of the tcgen05 3xTF32 engine (csrc/gemm_tf32x3.cu, on
tests/emu/tcgen05_emu.h) are compiled as host C++ behind tests/emu/cuda_emu.h - one OS thread per CUDA
thread, real barriers, warp shuffles through a per-warp exchange, cp.async groups that complete as
late as wait_group allows - and run against plain loops.  The FP64 tensor-core engine (csrc/gemm_dmma.cu)
runs the same way on top of tests/emu/ptx_emu.h: mbarriers with phase / arrival / transaction-byte
accounting and blocking waits, TMA bulk rows, cp.async with barrier arrival, the m8n8k4 MMA as a warp
exchange - so the producer / consumer protocol of all three schedules is exercised, and a protocol
deadlock blocks here (watchdog report: who waits on which barrier) instead of on a GPU.

This is how variants written without GPU access get checked before they ever reach hardware; it
also found a latent race in the measured potf2 kernels (a fast warp could raise the not-positive-
definite flag for column c+1 while a slow warp was still testing it for column c, which would have
left the block one barrier short).  It does not replace the GPU parity tests: only races that
follow from the program's barrier structure show up here, and nothing about speed."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import pytest

import harness as H

ROOT = H.ROOT
EMU = os.path.join(ROOT, "tests", "emu")
INC = ["-I" + EMU, "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "cuda-cholesky_b200", "csrc"),
       "-I" + os.path.join(ROOT, "include")]


HARNESSES = {"emu_panel.cpp": 900, "emu_small_gemm.cpp": 600, "emu_ffma.cpp": 1200, "emu_dmma.cpp": 1500, "emu_rng.cpp": 600,
             "emu_tf32x3.cpp": 600}
# a second run of a harness under an adversarial schedule: its output is appended behind a marker line
SECOND_RUN = {"emu_tf32x3.cpp": {"EMU_TC_SLOW_LD_MS": "25"}}  # epilogue a whole tile behind the MMA issuer
# plain-C sources a harness links (compiled as C): the product's polynomial code and the oracle's sequential generator
EXTRA_C = {"emu_rng.cpp": [os.path.join(ROOT, "cuda-cholesky_b200", "host", "rng_poly.c"), os.path.join(ROOT, "oracle", "dsfmt_oracle.c")]}


def _build_and_run(src, tmpdir, timeout):
    exe = os.path.join(tmpdir, os.path.splitext(src)[0])
    try:
        objs = []
        for c in EXTRA_C.get(src, ()):
            objs.append(os.path.join(tmpdir, os.path.basename(c)[:-2] + ".o"))
            subprocess.check_call(["gcc", "-std=c99", "-O2", "-D_POSIX_C_SOURCE=200809L", "-c", c, "-o", objs[-1]])
        subprocess.check_call(["g++", "-std=c++17", "-O1", *INC, os.path.join(EMU, src), *objs, "-o", exe, "-lpthread"])
        p = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout,
                           env=dict(os.environ, EMU_WATCHDOG=str(timeout - 60)))
        rc, out = p.returncode, p.stdout.decode()
        if rc == 0 and src in SECOND_RUN:
            p = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout,
                               env=dict(os.environ, EMU_WATCHDOG=str(timeout - 60), **SECOND_RUN[src]))
            rc, out = p.returncode, out + "\n=== second run ===\n" + p.stdout.decode()
        return rc, out
    except (subprocess.CalledProcessError, subprocess.TimeoutExpired) as e:
        return -1, "%r" % (e,)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    """all harnesses are built and run side by side (they mostly sleep in barriers)"""
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not installed")
    tmpdir = str(tmp_path_factory.mktemp("emu"))
    with ThreadPoolExecutor(max_workers=len(HARNESSES)) as ex:
        futs = {src: ex.submit(_build_and_run, src, tmpdir, t) for src, t in HARNESSES.items()}
        return {src: f.result() for src, f in futs.items()}


def _passed(emu, src):
    rc, out = emu[src]
    assert rc == 0 and "EMU PASSED" in out and "MISMATCH" not in out, out[-4000:]
    return out


def test_panel_kernels_emulated(emu):
    out = _passed(emu, "emu_panel.cpp")
    assert out.count("info, factor vs loops, other triangle untouched") >= 20  # potf2: d, z, s, c x lower/upper x full and ragged blocks
    assert out.count("not PD") >= 12
    assert out.count("trsm_rows<") >= 100  # both shapes x d, z, s, c x storage x transposed operands x in / out of place
    assert out.count("scatter_rows to block-cyclic rows") == 3


def test_small_gemm_kernel_emulated(emu):
    assert "147 cases" in _passed(emu, "emu_small_gemm.cpp")


def test_fp32_gemm_engine_emulated(emu):
    """the FP32 engine (k-tile 16, hoisted loader): ops, conjugation, triangular masks, unit diagonals, k-range
    skip, triangle-only / Hermitian output, ragged edges, aligned and unaligned operands (s and c)"""
    out = _passed(emu, "emu_ffma.cpp")
    assert "deep-k kernels: 86 cases so far, 0 mismatches" in out


def test_device_dsfmt_kernels_emulated(emu):
    """csrc/rng_dsfmt.cu: the warp-parallel refill (64 steps per round through one warp scan), windowed conversion and store in
    all five output forms, Krylov snapshots, and the GF(2) jump kernel with the product's own polynomial rows - bit for bit
    against the oracle's sequential generator"""
    out = _passed(emu, "emu_rng.cpp")
    assert "32 cases, 0 mismatches" in out
    assert out.count("identical to the oracle") >= 7


def test_tcgen05_tf32x3_engine_emulated(emu):
    """csrc/gemm_tf32x3.cu: the warp-specialised persistent tcgen05 kernel (TMA producer / MMA issuer / TMEM allocation / eight
    epilogue warps, two smem stages, two TMEM accumulators) and the operand-split kernels on tests/emu/tcgen05_emu.h - tensor-map
    boxes with the 128-byte swizzle, MMAs that read shared memory through the kernel's own matrix / instruction descriptors (decoded
    from CUTLASS's documented bit fields) and execute as late as the next commit, TMEM that starts as NaN - against double-precision
    loops: s and the complex embedding, triangles, ragged m / n / k, several tiles per CTA.  Second run: slow tcgen05.ld, the
    schedule under which a missing accumulator hand-back wait shows (checked by mutation, as were eight other protocol /
    encoding mutations: profiles/r02_cpu_verification.md)"""
    out = _passed(emu, "emu_tf32x3.cpp")
    assert out.count("16 cases") == 2 and out.count("EMU PASSED") == 2 and "=== second run ===" in out


def test_fp64_tensor_core_engine_emulated(emu):
    """classic (TMA-staged C epilogue), persistent, staggered and staggered-with-ring-staged-C schedules; TMA and generic loaders; d and z;
    triangular masks, triangle-only / Hermitian output, block-cyclic column map, K = 0"""
    out = _passed(emu, "emu_dmma.cpp")
    assert "54 cases" in out
    for sched in ("classic", "persistent", "staggered", "stag+ringC"):
        assert out.count(sched) >= 4
