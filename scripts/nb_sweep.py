"""Block-size sweep: python scripts/nb_sweep.py <prec> <N> nb1 nb2 ...   (potrf lower, best of 2)"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholesky_b200_loader import load
chol = load()
p, n = sys.argv[1], int(sys.argv[2])
nbs = [int(x) for x in sys.argv[3:]]
g = torch.Generator(device="cuda").manual_seed(5)
rt = torch.float64 if p in "dz" else torch.float32
if p in "cz":
    B = torch.complex(torch.rand(n, 512, dtype=rt, device="cuda", generator=g) - 0.5, torch.rand(n, 512, dtype=rt, device="cuda", generator=g) - 0.5)
else:
    B = torch.rand(n, 512, dtype=rt, device="cuda", generator=g) - 0.5
A0 = B @ B.conj().T
A0.diagonal().add_(float(n))
del B
lap = chol.LapackHandle()
for nb in nbs:
    lap.set_block_size(nb)
    best = 1e9
    for _ in range(2):
        A = A0.clone()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        info = lap.potrf(p, "L", n, A.data_ptr(), n)
        best = min(best, time.perf_counter() - t0)
        del A
    f = (4.0 if p in "cz" else 1.0) * n ** 3 / 3
    print("%spotrf N=%d nb=%d pp=%s: %.4f s %.2f real-TF/s info %d" % (p, n, nb, os.environ.get("CHOL_GEMM_PP", "0"), best, f / best / 1e12, info), flush=True)
