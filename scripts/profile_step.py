"""One DPOTRF on cuda:0 for ncu (launch list / full capture of the DMMA trailing-update kernel).
usage: python scripts/profile_step.py [N] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholesky_b200_loader import load  # noqa: E402

chol = load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator(device="cuda").manual_seed(1)
B = torch.rand(512, n, dtype=torch.float64, device="cuda", generator=g) - 0.5
A0 = B.T @ B + n * torch.eye(n, dtype=torch.float64, device="cuda")
h = chol.LapackHandle()
for _ in range(reps):
    A = A0.clone()
    torch.cuda.synchronize()
    info = h.potrf("d", "L", n, A.data_ptr(), n)
    assert info == 0
print("ok", n, reps)
