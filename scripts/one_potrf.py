"""One warm-up + one cuDpotrf (lower, N = $ONE_N, precision $ONE_PREC) - the program profiled for launch lists."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholesky_b200_loader import load  # noqa: E402

chol = load()
n = int(os.environ.get("ONE_N", "16384"))
p = os.environ.get("ONE_PREC", "d")
uplo = os.environ.get("ONE_UPLO", "L")
dt = {"s": torch.float32, "d": torch.float64}[p]
g = torch.Generator(device="cuda").manual_seed(3)
B = torch.rand(n, 1024, dtype=dt, device="cuda", generator=g) - 0.5
A0 = B @ B.T
A0.diagonal().add_(float(n))
lap = chol.LapackHandle()
for _ in range(int(os.environ.get("ONE_REPS", "2"))):
    A = A0.clone()
    torch.cuda.synchronize()
    info = lap.potrf(p, uplo, n, A.data_ptr(), n)
    print("info", info)
