"""Time the fused diagonal-block kernel (potf2 + inverse of one 128 x 128 block, 64 for z) alone:
N back-to-back launches on one stream between two CUDA events, and check the result against
numpy.  Run once per setting of CHOL_POTF2_V2 (the switch is read once per process):
    for v in 0 1; do CHOL_POTF2_V2=$v python scripts/potf2_ab.py; done"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholesky_b200_loader import load  # noqa: E402

chol = load()
lib = chol.lib()
DT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}
reps = int(os.environ.get("AB_REPS", "200"))
for p in "dzsc":
    n = 64 if p == "z" else 128
    for uplo in "LU":
        g = torch.Generator(device="cuda").manual_seed(5)
        rt = torch.float64 if p in "dz" else torch.float32
        B = torch.rand(n, n, dtype=rt, device="cuda", generator=g) - 0.5
        if p in "cz":
            B = torch.complex(B, torch.rand(n, n, dtype=rt, device="cuda", generator=g) - 0.5)
        A0 = B @ B.conj().T + n * torch.eye(n, dtype=DT[p], device="cuda")  # Hermitian: row-major == conj column-major, still HPD
        A = A0.clone()
        X = torch.zeros_like(A0)
        dinfo = torch.zeros(1, dtype=torch.int64, device="cuda")

        def launch():
            return lib.chol_potf2_inv_launch(C.c_int(ord(p)), C.c_int(ord(uplo)), C.c_size_t(n), C.c_void_p(A.data_ptr()), C.c_size_t(n),
                                             C.c_void_p(X.data_ptr()), C.c_size_t(n), C.c_void_p(dinfo.data_ptr()), C.c_long(0),
                                             C.c_void_p(0))
        assert launch() == 0
        torch.cuda.synchronize()
        # column-major view of the row-major buffers: M_cm = buffer^T
        Acm, Xcm, A0cm = A.T, X.T, A0.T
        F = torch.tril(Acm) if uplo == "L" else torch.triu(Acm)
        R = F @ F.conj().T if uplo == "L" else F.conj().T @ F
        e_f = (torch.linalg.norm(R - A0cm) / torch.linalg.norm(A0cm)).item()
        e_x = torch.linalg.norm(Xcm @ F - torch.eye(n, dtype=DT[p], device="cuda")).item()
        A.copy_(A0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):  # refactoring a factor is not SPD-safe: each launch gets the same pristine block only the first time;
            launch()           # dinfo != 0 would turn later launches into no-ops, so check it stayed 0 below
            A.copy_(A0)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print("%spotf2_inv %s n=%d V2=%s: %.1f us per launch (incl. a 128x128 restore copy), |LL^H-A|/|A| %.1e, |X L - I| %.1e, info %d"
              % (p, uplo, n, os.environ.get("CHOL_POTF2_V2", "0"), us, e_f, e_x, int(dinfo.item())), flush=True)
