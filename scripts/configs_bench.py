"""BASELINE.json configs[1..3] on one B200: {d}potrf U/L N=16384, {s,c,z}potrf N=32768, dpotri + dlogdet N=32768.
Timings are wall clock around the blocking C-ABI call (it returns with the device idle)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholesky_b200_loader import load  # noqa: E402

chol = load()
DT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}
EPS = {"s": 1.2e-7, "d": 2.2e-16, "c": 1.2e-7, "z": 2.2e-16}
out = {}
lap = chol.LapackHandle()
g = torch.Generator(device="cuda").manual_seed(20121018)


def spd(p, n, k=1024):
    rt = torch.float64 if p in "dz" else torch.float32
    if p in "cz":
        B = torch.complex(torch.rand(n, k, dtype=rt, device="cuda", generator=g) - 0.5,
                          torch.rand(n, k, dtype=rt, device="cuda", generator=g) - 0.5)
    else:
        B = torch.rand(n, k, dtype=rt, device="cuda", generator=g) - 0.5
    A = B @ B.conj().T
    A.diagonal().add_(float(n))
    return A  # Hermitian: row-major storage == column-major conj; use its transpose-consistent form below


def flops_potrf(p, n):
    return (4.0 if p in "cz" else 1.0) * n ** 3 / 3.0


which = sys.argv[1:] or ["cfg1", "cfg2", "cfg3"]
for cfg, n in (("cfg0", 8192), ("cfg1", 16384)):
    if cfg not in which:
        continue
    A0 = spd("d", n)
    for uplo in "LU":
        best = 1e9
        for _ in range(3):
            A = A0.clone()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            info = lap.potrf("d", uplo, n, A.data_ptr(), n)
            best = min(best, time.perf_counter() - t0)
        Lt = torch.triu(A) if uplo == "L" else torch.tril(A)
        R = Lt.T @ Lt if uplo == "L" else Lt @ Lt.T
        resid = (torch.linalg.norm(R - A0) / torch.linalg.norm(A0)).item()
        out["dpotrf_%s_%d" % (uplo, n)] = {"s": best, "tflops": flops_potrf("d", n) / best / 1e12, "resid": resid, "info": info}
        print("dpotrf %s N=%d: %.4f s %.2f TFLOP/s resid %.2e info %d" % (uplo, n, best, flops_potrf("d", n) / best / 1e12, resid, info), flush=True)
    del A0, A, Lt, R
if "cfg2" in which:
    n = 32768
    for p in os.environ.get("CFG2_PRECS", "scz"):
        A0 = spd(p, n)  # Hermitian; a row-major Hermitian matrix read column-major is its conjugate = still HPD
        best = 1e9
        for _ in range(2):
            A = A0.clone()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            info = lap.potrf(p, "L", n, A.data_ptr(), n)
            best = min(best, time.perf_counter() - t0)
        # column-major lower factor L of conj(A0) <=> torch upper triangle of A holds L^T: A0^T = conj(A0) = L L^H
        Lt = torch.triu(A)          # = L^T (plain transpose)
        m = 4096                    # residual on the leading block (factor of a leading block = leading block of the factor)
        Lb = Lt[:m, :m].T           # L (m x m)
        R = Lb @ Lb.conj().T
        ref = A0[:m, :m].conj() if p in "cz" else A0[:m, :m]
        resid = (torch.linalg.norm(R - ref) / torch.linalg.norm(ref)).item()
        tf = flops_potrf(p, n) / best / 1e12
        out["%spotrf_L_%d" % (p, n)] = {"s": best, "real_tflops": tf, "resid_lead4096": resid, "info": info}
        print("%spotrf L N=%d: %.4f s %.2f real-TFLOP/s resid(lead 4096) %.2e info %d" % (p, n, best, tf, resid, info), flush=True)
        del A0, A, Lt, R
if "cfg3" in which:
    n = 32768
    A0 = spd("d", n)
    A = A0.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = lap.potrf("d", "L", n, A.data_ptr(), n)
    t_potrf = time.perf_counter() - t0
    ld = lap.logdet("d", A.data_ptr(), n + 1, n)
    want = 2 * torch.log(torch.diagonal(A)).sum().item()
    F = A.clone()  # the factor
    t_potri = 1e9
    for _ in range(3):  # the first call also grows the handle's workspace; best of 3
        A.copy_(F)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        info2 = lap.potri("d", "L", n, A.data_ptr(), n)
        t_potri = min(t_potri, time.perf_counter() - t0)
    del F
    # column-major lower triangle of A^-1 sits in torch's upper triangle
    U = torch.triu(A)
    Ainv = U + U.T - torch.diag(torch.diagonal(U))
    m = 8192
    chk = (A0[:m, :] @ Ainv[:, :m] - torch.eye(m, dtype=torch.float64, device="cuda")).abs().max().item()
    out["dpotri_L_%d" % n] = {"potrf_s": t_potrf, "potri_s": t_potri, "potri_tflops": 2 * n ** 3 / 3 / t_potri / 1e12,
                              "potrf_tflops": n ** 3 / 3 / t_potrf / 1e12, "max_abs_AAinv_minus_I_lead8192": chk,
                              "logdet": ld, "logdet_relerr": abs(ld - want) / abs(want), "info": [info, info2]}
    print("dpotrf+dpotri L N=%d: potrf %.3f s (%.2f TF/s), potri %.3f s (%.2f TF/s), |A*Ainv-I|max(lead 8192 rows) %.2e, logdet relerr %.1e"
          % (n, t_potrf, n ** 3 / 3 / t_potrf / 1e12, t_potri, 2 * n ** 3 / 3 / t_potri / 1e12, chk, abs(ld - want) / abs(want)), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("CFG_OUT", "configs_bench.json")), "w"), indent=1)
