"""One trailing-update launch of the d engine alone (lower SYRK  C -= W W^T,  n = $SYRK_N, k = $SYRK_K), the program that
`ncu --set full -k regex:gemm_dmma` captures for profiles/r02_ncu_dmma_update.json (scripts/ncu_summary.py)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholesky_b200_loader import load  # noqa: E402

chol = load()
n, k = int(os.environ.get("SYRK_N", "32768")), int(os.environ.get("SYRK_K", "1024"))
g = torch.Generator(device="cuda").manual_seed(1)
W = torch.rand(k, n, dtype=torch.float64, device="cuda", generator=g) - 0.5  # column-major n x k
Cm = torch.zeros(n, n, dtype=torch.float64, device="cuda")
for _ in range(int(os.environ.get("SYRK_REPS", "2"))):
    chol.gemm_desc_launch(prec="d", opA="N", opB="T", cuplo="L", m=n, n=n, k=k, alpha=-1.0, beta=1.0, A=W.data_ptr(), lda=n,
                          B=W.data_ptr(), ldb=n, C=Cm.data_ptr(), ldc=n, D=Cm.data_ptr(), ldd=n)
torch.cuda.synchronize()
print("done")
