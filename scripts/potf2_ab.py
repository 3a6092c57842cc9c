"""Time the diagonal-block kernels alone: potf2 (one 128 x 128 block, 64 for z) and trsm_rows (X = B L^-H, m rows),
back-to-back launches on one stream between two CUDA events, results checked against torch."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholesky_b200_loader import load  # noqa: E402

chol = load()
lib = chol.lib()
DT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}
reps = int(os.environ.get("AB_REPS", "200"))
for kind, name in ((0, "DFMA"), (1, "DMUL"), (2, "rsqrt(double)"), (3, "shfl 64-bit"), (4, "LDS chase")):
    cyc = C.c_double(0)
    lib.chol_probe_latency(C.c_int(kind), C.byref(cyc))
    print("latency %-14s %.1f cycles" % (name, cyc.value), flush=True)
def potf2_phases(p="d", n=128, uplo="L"):
    """clock64 phase timestamps of ONE potf2 launch (test hook chol_potf2_set_debug)"""
    rt = torch.float64 if p in "dz" else torch.float32
    g = torch.Generator(device="cuda").manual_seed(5)
    B = torch.rand(n, n, dtype=rt, device="cuda", generator=g) - 0.5
    A = B @ B.T + n * torch.eye(n, dtype=rt, device="cuda")
    dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
    dinfo = torch.zeros(1, dtype=torch.int64, device="cuda")
    for rep in range(2):
        A2 = A.clone()
        dbg.zero_()
        lib.chol_potf2_set_debug(C.c_void_p(dbg.data_ptr()))
        lib.chol_potf2_launch(C.c_int(ord(p)), C.c_int(ord(uplo)), C.c_size_t(n), C.c_void_p(A2.data_ptr()), C.c_size_t(n),
                              C.c_void_p(dinfo.data_ptr()), C.c_long(0), C.c_void_p(0))
        torch.cuda.synchronize()
        lib.chol_potf2_set_debug(C.c_void_p(0))
    t = dbg.cpu().tolist()
    t0 = t[0]
    print("potf2 %s n=%d phases (cycles since kernel start): staged+regs %d, panel0 done %d, main loop done %d, stored %d"
          % (p, n, t[1] - t0, t[2] - t0, t[60] - t0, t[61] - t0), flush=True)
    rounds = []
    for q in range(15):
        a, b, c_ = t[4 + 3 * q], t[5 + 3 * q], t[6 + 3 * q]
        if a:
            rounds.append("r%d: start %d update %d panel %d" % (q, a - t0, b - a, c_ - b))
    print("  owner of panel p+1 per round: " + "; ".join(rounds), flush=True)


if os.environ.get("AB_PHASES", "1") != "0":
    potf2_phases("d")
    potf2_phases("s")
for p in os.environ.get("AB_PRECS", "dzsc"):
    n = 64 if p == "z" else 128
    rt = torch.float64 if p in "dz" else torch.float32
    for uplo in "LU":
        g = torch.Generator(device="cuda").manual_seed(5)
        B = torch.rand(n, n, dtype=rt, device="cuda", generator=g) - 0.5
        if p in "cz":
            B = torch.complex(B, torch.rand(n, n, dtype=rt, device="cuda", generator=g) - 0.5)
        A0 = B @ B.conj().T + n * torch.eye(n, dtype=DT[p], device="cuda")  # Hermitian: row-major == conj column-major, still HPD
        A = A0.clone()
        dinfo = torch.zeros(1, dtype=torch.int64, device="cuda")

        def launch():
            return lib.chol_potf2_launch(C.c_int(ord(p)), C.c_int(ord(uplo)), C.c_size_t(n), C.c_void_p(A.data_ptr()), C.c_size_t(n),
                                         C.c_void_p(dinfo.data_ptr()), C.c_long(0), C.c_void_p(0))
        assert launch() == 0
        torch.cuda.synchronize()
        Acm, A0cm = A.T, A0.T  # column-major view of the row-major buffers
        F = torch.tril(Acm) if uplo == "L" else torch.triu(Acm)
        R = F @ F.conj().T if uplo == "L" else F.conj().T @ F
        e_f = (torch.linalg.norm(R - A0cm) / torch.linalg.norm(A0cm)).item()
        A.copy_(A0)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):  # every launch factors a pristine block
            launch()
            A.copy_(A0)
        e1.record()
        for _ in range(reps):
            A.copy_(A0)
        e2.record()
        torch.cuda.synchronize()
        us = (e0.elapsed_time(e1) - e1.elapsed_time(e2)) * 1e3 / reps
        print("%spotf2 %s n=%d: %.1f us per launch (restore copy subtracted), |LL^H-A|/|A| %.1e, info %d"
              % (p, uplo, n, us, e_f, int(dinfo.item())), flush=True)
    # trsm_rows: X = B * L^-H, B m x n column-major (torch: (n, m) row-major)
    g = torch.Generator(device="cuda").manual_seed(6)
    Lr = torch.rand(n, n, dtype=rt, device="cuda", generator=g) - 0.5
    if p in "cz":
        Lr = torch.complex(Lr, torch.rand(n, n, dtype=rt, device="cuda", generator=g) - 0.5)
    Lcm = torch.tril(Lr, -1) + torch.diag(torch.full((n,), 4.0, dtype=rt, device="cuda")).to(DT[p])  # column-major lower L
    Lbuf = Lcm.T.contiguous()  # row-major buffer whose column-major reading is Lcm
    for m in (384, 4096, 16384, 65536):
        Bt = torch.rand(n, m, dtype=rt, device="cuda", generator=g) - 0.5  # buffer (n, m) row-major == column-major m x n
        if p in "cz":
            Bt = torch.complex(Bt, torch.rand(n, m, dtype=rt, device="cuda", generator=g) - 0.5)
        Xt = torch.empty_like(Bt)

        def tl():
            return lib.chol_trsm_rows_launch(C.c_int(ord(p)), C.c_int(ord("L")), C.c_size_t(n), C.c_void_p(Lbuf.data_ptr()), C.c_size_t(n),
                                             C.c_size_t(m), C.c_int(0), C.c_void_p(Bt.data_ptr()), C.c_size_t(m), C.c_int(0),
                                             C.c_void_p(Xt.data_ptr()), C.c_size_t(m), C.c_void_p(0), C.c_void_p(0))
        assert tl() == 0
        torch.cuda.synchronize()
        Bm, Xm = Bt.T, Xt.T  # m x n
        err = (torch.linalg.norm(Xm @ Lcm.conj().T - Bm) / torch.linalg.norm(Bm)).item()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            tl()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        fl = (8.0 if p in "cz" else 2.0) * m * n * n / 2
        print("%strsm_rows n=%d m=%d: %.1f us per launch, %.2f real TFLOP/s, |X L^H - B|/|B| %.1e" % (p, n, m, us, fl / us / 1e6, err), flush=True)
