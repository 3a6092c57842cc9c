"""Bring-up check of the tcgen05 3xTF32 engine (csrc/gemm_tf32x3.cu) through the launcher ABI:
correctness against an FP64 product on ragged / triangular shapes, then timing of the spotrf-shaped
trailing update.   python scripts/tf32x3_check.py [--time]
"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cholesky_b200_loader  # noqa: E402

cb = cholesky_b200_loader.load()
L = cb.lib()
L.chol_tf32_kpitch.restype = C.c_size_t
L.chol_tf32_kpitch.argtypes = [C.c_size_t]


def split(X, trans=0):
    """X: torch (rows x k) logical operand stored column-major (as .T of a (k, ld) tensor) when trans=0."""
    rows, k = X.shape
    kp = L.chol_tf32_kpitch(k)
    hi = torch.empty((rows, kp), dtype=torch.float32, device="cuda")
    lo = torch.empty_like(hi)
    if trans == 0:
        S = X.t().contiguous()  # (k, rows) row-major == rows x k column-major, lda = rows
        lda = rows
    else:
        S = X.contiguous()      # (rows, k) row-major: element (r, c) at r*lda + c
        lda = k
    rc = L.chol_tf32_split_launch(C.c_int(trans), C.c_size_t(rows), C.c_size_t(k), C.c_void_p(S.data_ptr()), C.c_size_t(lda),
                                  C.c_void_p(hi.data_ptr()), C.c_void_p(lo.data_ptr()), C.c_size_t(kp), C.c_void_p(0))
    assert rc == 0, rc
    return hi, lo, kp


def gemm(m, n, k, alpha, Ah, Al, Bh, Bl, kp, beta, Cm, ldc, uplo):
    rc = L.chol_tf32x3_gemm_launch(C.c_size_t(m), C.c_size_t(n), C.c_size_t(k), C.c_double(alpha), C.c_void_p(Ah.data_ptr()),
                                   C.c_void_p(Al.data_ptr()), C.c_void_p(Bh.data_ptr()), C.c_void_p(Bl.data_ptr()), C.c_size_t(kp),
                                   C.c_double(beta), C.c_void_p(Cm.data_ptr()), C.c_size_t(ldc), C.c_void_p(Cm.data_ptr()),
                                   C.c_size_t(ldc), C.c_int(ord(uplo) if uplo else 0), C.c_void_p(0))
    assert rc == 0, rc


def check(m, n, k, uplo, alpha=-1.0, beta=1.0, trans=0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.rand((m, k), device="cuda", generator=g) - 0.5
    B = torch.rand((n, k), device="cuda", generator=g) - 0.5
    ldc = m + 3
    Ct = torch.rand((n, ldc), device="cuda", generator=g)  # column-major m x n with ld = ldc
    C0 = Ct.clone()
    Ah, Al, kp = split(A, trans)
    Bh, Bl, _ = split(B, trans)
    # split self-check: hi + lo reproduces x to 2^-22 relative
    rec = (Ah[:, :k].double() + Al[:, :k].double() - A.double()).abs().max().item()
    gemm(m, n, k, alpha, Ah, Al, Bh, Bl, kp, beta, Ct, ldc, uplo)
    torch.cuda.synchronize()
    ref = alpha * (A.double() @ B.double().t()) + beta * C0[:, :m].t().double()
    got = Ct[:, :m].t().double()
    ii = torch.arange(m, device="cuda")[:, None]
    jj = torch.arange(n, device="cuda")[None, :]
    keep = torch.ones((m, n), dtype=torch.bool, device="cuda") if not uplo else ((ii >= jj) if uplo == "L" else (ii <= jj))
    err = ((got - ref).abs() * keep).max().item()
    untouched = ((got - C0[:, :m].t().double()).abs() * (~keep)).max().item() if uplo else 0.0
    pad_ok = torch.equal(Ct[:, m:], C0[:, m:])
    bound = (2 * k + 8) * 2.0 ** -21 * 0.25  # |a|,|b| <= 0.5
    ok = err <= bound and untouched == 0.0 and pad_ok
    print("m=%d n=%d k=%d uplo=%s trans=%d: err %.3e (bound %.3e) split-rec %.2e untouched %.1e pad %s  %s"
          % (m, n, k, uplo, trans, err, bound, rec, untouched, pad_ok, "OK" if ok else "FAIL"), flush=True)
    return ok


def timing():
    out = {}
    for (m, k) in ((32768 - 1024, 512), (16384, 512), (32768 - 512, 256), (8192, 512)):
        A = torch.rand((m, k), device="cuda") - 0.5
        Ct = torch.rand((m, m), device="cuda")
        Ah, Al, kp = split(A)
        for _ in range(2):
            gemm(m, m, k, -1.0, Ah, Al, Ah, Al, kp, 1.0, Ct, m, "L")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 5
        for _ in range(reps):
            gemm(m, m, k, -1.0, Ah, Al, Ah, Al, kp, 1.0, Ct, m, "L")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tf = m * (m + 1.0) * k / ms / 1e9
        e0.record()
        for _ in range(reps):
            split(A)
        e1.record()
        torch.cuda.synchronize()
        ms_split = e0.elapsed_time(e1) / reps
        out["syrk_L_m%d_k%d" % (m, k)] = {"ms": ms, "tflops_fp32_equiv": tf, "split_ms": ms_split}
        print("syrk L m=%d k=%d: %.3f ms  %.1f TFLOP/s (FP32-equivalent), split %.3f ms" % (m, k, ms, tf, ms_split), flush=True)
        del A, Ct, Ah, Al
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/tf32x3_timing.json", "w"), indent=1)


if __name__ == "__main__":
    ok = True
    for (m, n, k, u) in ((128, 256, 32, ""), (128, 256, 64, ""), (256, 256, 128, "L"), (300, 300, 96, "L"), (300, 300, 100, "U"),
                         (1000, 520, 200, ""), (1111, 1111, 512, "L"), (777, 777, 40, "U"), (129, 5, 8, "")):
        ok &= check(m, n, k, u)
    ok &= check(515, 300, 64, "", alpha=0.5, beta=0.0)
    ok &= check(400, 400, 64, "L", trans=1)
    print("ALL OK" if ok else "FAILED", flush=True)
    if "--time" in sys.argv and ok:
        timing()
    sys.exit(0 if ok else 1)
