"""First-light measurements on a B200: FP64 peak probes, the DMMA GEMM/SYRK tile engine, DPOTRF."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholesky_b200_loader import load  # noqa: E402

chol = load()
out = {}
torch.cuda.init()
torch.zeros(1, device="cuda")
for kind, name in ((0, "dmma"), (1, "dfma"), (2, "dmma+dfma"), (3, "dmma_same_operands"), (10, "dmma_1cta_per_sm(1 warp/SMSP)")):
    tf, ms = chol.probe_fp64(kind, 10000)
    out["probe_" + name] = {"tflops": tf, "ms": ms}
    print("probe", name, "%.2f TFLOP/s (%.2f ms)" % (tf, ms), flush=True)


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = torch.Generator(device="cuda").manual_seed(1)
blas = chol.BlasHandle()
for (m, n, k) in ((8192, 8192, 256), (16384, 16384, 256), (16384, 16384, 512), (32768, 32768, 256), (8192, 8192, 8192)):
    A = torch.rand(k, m, dtype=torch.float64, device="cuda", generator=g) - 0.5  # col-major m x k
    B = torch.rand(k, n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    Cm = torch.zeros(n, m, dtype=torch.float64, device="cuda")
    for force in (0, 1):
        def run():
            chol.gemm_desc_launch(prec="d", opA="N", opB="T", m=m, n=n, k=k, alpha=-1.0, beta=1.0, A=A.data_ptr(), lda=m,
                                  B=B.data_ptr(), ldb=n, C=Cm.data_ptr(), ldc=m, D=Cm.data_ptr(), ldd=m, force_generic=force)
        ms = timeit(run)
        tf = 2.0 * m * n * k / (ms * 1e-3) / 1e12
        out["gemm_nt_%d_%d_%d_%s" % (m, n, k, "generic" if force else "tma")] = {"ms": ms, "tflops": tf}
        print("gemm NT %dx%dx%d %s: %.3f ms  %.2f TFLOP/s" % (m, n, k, "generic" if force else "tma", ms, tf), flush=True)
    if m == n:
        def run2():
            blas.rankk("d", "L", "N", n, k, -1.0, A.data_ptr(), m, 1.0, Cm.data_ptr(), m)
        ms = timeit(run2)
        tf = 1.0 * n * n * k / (ms * 1e-3) / 1e12
        out["syrk_%d_%d" % (n, k)] = {"ms": ms, "tflops": tf}
        print("syrk L %dx%d: %.3f ms  %.2f TFLOP/s" % (n, k, ms, tf), flush=True)
    del A, B, Cm

lap = chol.LapackHandle()
for n in (4096, 8192, 16384, 32768):
    kk = 512
    Bm = torch.rand(n, kk, dtype=torch.float64, device="cuda", generator=g) - 0.5
    A0 = Bm @ Bm.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    del Bm
    for nb in (128, 256, 512):
        lap.set_block_size(nb)
        for uplo in "LU":
            best = 1e30
            for rep in range(2):
                A = A0.clone()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                info = lap.potrf("d", uplo, n, A.data_ptr(), n)
                dt = time.perf_counter() - t0
                best = min(best, dt)
            Lt = torch.triu(A) if uplo == "L" else torch.tril(A)
            R = Lt.T @ Lt if uplo == "L" else Lt @ Lt.T
            resid = (torch.linalg.norm(R - A0) / torch.linalg.norm(A0)).item()
            tf = n ** 3 / 3.0 / best / 1e12
            out["dpotrf_%s_%d_nb%d" % (uplo, n, nb)] = {"s": best, "tflops": tf, "resid": resid, "info": info}
            print("dpotrf %s n=%d nb=%d: %.4f s  %.2f TFLOP/s  resid %.2e info %d" % (uplo, n, nb, best, tf, resid, info), flush=True)
            del A, Lt, R
    del A0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "first_light.json"), "w"), indent=1)
