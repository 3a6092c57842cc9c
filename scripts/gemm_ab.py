"""A/B timing of the DMMA engine under the CHOL_GEMM_* environment switches (run one config per process)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholesky_b200_loader import load  # noqa: E402

chol = load()
torch.zeros(1, device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)
blas = chol.BlasHandle()
tag = "pp=%s stagec=%s" % (os.environ.get("CHOL_GEMM_PP", "auto"), os.environ.get("CHOL_GEMM_STAGEC", "1"))
for (n, k) in ((16384, 256), (32768, 256), (49152, 256), (32768, 512)):
    A = torch.rand(k, n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    Cm = torch.zeros(n, n, dtype=torch.float64, device="cuda")

    def run():
        blas.rankk("d", "L", "N", n, k, -1.0, A.data_ptr(), n, 1.0, Cm.data_ptr(), n)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%s syrk L %dx%d: %.3f ms %.2f TFLOP/s" % (tag, n, k, ms, 1.0 * n * n * k / ms / 1e9), flush=True)
    del A, Cm
