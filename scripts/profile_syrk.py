"""The POTRF trailing update in isolation for ncu: C(lower) -= W*W^T, n x n x k (default 32768 x 512)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholesky_b200_loader import load
chol = load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
k = int(sys.argv[2]) if len(sys.argv) > 2 else 512
g = torch.Generator(device="cuda").manual_seed(1)
W = torch.rand(k, n, dtype=torch.float64, device="cuda", generator=g) - 0.5
Cm = torch.zeros(n, n, dtype=torch.float64, device="cuda")
blas = chol.BlasHandle()
for _ in range(3):
    blas.rankk("d", "L", "N", n, k, -1.0, W.data_ptr(), n, 1.0, Cm.data_ptr(), n)
torch.cuda.synchronize()
print("ok")
