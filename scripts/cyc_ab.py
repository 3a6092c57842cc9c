"""Micro-benchmark of one rank's block-cyclic trailing update (P ranks emulated: local N x N/P panel)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholesky_b200_loader import load  # noqa: E402

chol = load()
torch.zeros(1, device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)
tag = "pp=%s tps=%s" % (os.environ.get("CHOL_GEMM_PP", "1"), os.environ.get("CHOL_GEMM_TPS", "8"))
P, nb = 8, int(os.environ.get("NB", "256"))
for m in (16384, 32768, 49152):
    nloc = m // P
    W = torch.rand(nb, m, dtype=torch.float64, device="cuda", generator=g) - 0.5     # column-major m x nb
    D = torch.zeros(nloc, m, dtype=torch.float64, device="cuda")                      # column-major m x nloc

    def run():
        chol.gemm_desc_launch(prec="d", opA="N", opB="T", cuplo="L", m=m, n=nloc, k=nb, alpha=-1.0, beta=1.0, A=W.data_ptr(), lda=m,
                              B=W.data_ptr(), ldb=m, C=D.data_ptr(), ldc=m, D=D.data_ptr(), ldd=m, cyc_nb=nb, cyc_p=P, cyc_off=0)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    flops = 0.0
    for lc in range(nloc // nb):
        gcol = lc * P * nb
        flops += ((m - gcol) * nb - 0.5 * nb * (nb - 1)) * 2.0 * nb
    print("%s cyc P=%d nb=%d m=%d nloc=%d: %.3f ms %.2f TFLOP/s" % (tag, P, nb, m, nloc, ms, flops / ms / 1e9), flush=True)
