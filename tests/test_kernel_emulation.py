"""Kernel LOGIC on the CPU: the CUDA sources of the diagonal-block kernels (csrc/panel_kernels.cu:
potf2_inv in its measured and V2 variants, potf2, trti2, lauu2), of the small-tile FP64 GEMM
(csrc/gemm_small.cu) and of the FP32 GEMM engine (csrc/gemm_ffma.cu and its deeper-k variant
csrc/gemm_ffma_deep.cu) are compiled as host C++ behind tests/emu/cuda_emu.h - one OS thread per CUDA
thread, real barriers, warp shuffles through a per-warp exchange, cp.async groups that complete as
late as wait_group allows - and run against plain loops.

This is how variants written without GPU access get checked before they ever reach hardware; it
also found a latent race in the measured potf2 kernels (a fast warp could raise the not-positive-
definite flag for column c+1 while a slow warp was still testing it for column c, which would have
left the block one barrier short).  It does not replace the GPU parity tests: only races that
follow from the program's barrier structure show up here, and nothing about speed."""
import os
import subprocess

import pytest

import harness as H

ROOT = H.ROOT
EMU = os.path.join(ROOT, "tests", "emu")
INC = ["-I" + EMU, "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "cuda-cholesky_b200", "csrc"),
       "-I" + os.path.join(ROOT, "include")]


def _build_and_run(src, tmp_path, timeout):
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not installed")
    exe = str(tmp_path / os.path.splitext(src)[0])
    subprocess.check_call(["g++", "-std=c++17", "-O1", *INC, os.path.join(EMU, src), "-o", exe, "-lpthread"])
    p = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
    out = p.stdout.decode()
    assert p.returncode == 0 and "EMU PASSED" in out, out[-4000:]
    return out


def test_panel_kernels_emulated(tmp_path):
    out = _build_and_run("emu_panel.cpp", tmp_path, 900)
    assert "MISMATCH" not in out
    assert out.count("V2 == measured variant") >= 16  # d, z, s, c x lower/upper x full and ragged blocks
    assert out.count("not PD") >= 6


def test_small_gemm_kernel_emulated(tmp_path):
    out = _build_and_run("emu_small_gemm.cpp", tmp_path, 600)
    assert "147 cases" in out


def test_fp32_gemm_engine_emulated(tmp_path):
    """measured kernel and deep-k variant: ops, conjugation, triangular masks, unit diagonals, k-range
    skip, triangle-only / Hermitian output, ragged edges, aligned and unaligned operands (s and c)"""
    out = _build_and_run("emu_ffma.cpp", tmp_path, 1200)
    assert "measured kernels: 86 cases so far, 0 mismatches" in out
    assert "deep-k kernels: 172 cases so far, 0 mismatches" in out
