"""Time TRTRI and LAUUM separately (d, lower/upper) at N and nb from argv."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholesky_b200_loader import load
chol = load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
g = torch.Generator(device="cuda").manual_seed(3)
B = torch.rand(1024, n, dtype=torch.float64, device="cuda", generator=g) - 0.5
A0 = B.T @ B
A0.diagonal().add_(float(n))
del B
lap = chol.LapackHandle()
assert lap.potrf("d", "L", n, A0.data_ptr(), n) == 0   # A0 now holds the factor in its (col-major) lower triangle
for nb in [int(x) for x in (sys.argv[2:] or ["256", "512"])]:
    lap.set_block_size(nb)
    for uplo in "L":
        A = A0.clone()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        info = lap.trtri("d", uplo, "N", n, A.data_ptr(), n)
        t1 = time.perf_counter() - t0
        torch.cuda.synchronize(); t0 = time.perf_counter()
        info2 = lap.lauum("d", uplo, n, A.data_ptr(), n)
        t2 = time.perf_counter() - t0
        print("n=%d nb=%d %s: trtri %.3f s (%.2f TF/s)  lauum %.3f s (%.2f TF/s)  info %d %d" % (
            n, nb, uplo, t1, n ** 3 / 3 / t1 / 1e12, t2, n ** 3 / 3 / t2 / 1e12, info, info2), flush=True)
        del A
