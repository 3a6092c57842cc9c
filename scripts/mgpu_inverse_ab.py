"""A/B of the host-pointer cuMultiGPUDpotri on all visible GPUs: whole matrix staged on the first GPU, the distributed
block-cyclic sweeps (default with more than one GPU), and the exchange-free slab split (CHOL_MGPU_INVERSE=staged /
dist / slab).  Page-locked host matrix.  Prints seconds and max |A A^-1 - I|."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cholesky_b200_loader import load  # noqa: E402

chol = load()
N = int(os.environ.get("AB_N", "16384"))
ngpu = torch.cuda.device_count()


def spd(n):
    g = torch.Generator(device="cuda").manual_seed(20121018)
    B = torch.rand(n, 256, dtype=torch.float64, device="cuda", generator=g) - 0.5
    A = B @ B.T
    A += n * torch.eye(n, dtype=torch.float64, device="cuda")
    return A


A0 = spd(N)
host0 = A0.cpu().numpy()  # symmetric: row-major == column-major
lap = chol.LapackHandle()
F = A0.clone()
assert lap.potrf("d", "L", N, F.data_ptr(), N) == 0  # column-major lower factor == row-major upper triangle
factor = F.cpu().numpy().T.copy(order="F")  # column-major lower factor as a numpy F-array
del F
torch.cuda.synchronize()

MODES = (("staged on GPU 0", {"CHOL_MGPU_INVERSE": "staged"}, list(range(ngpu))),
         ("distributed, all GPUs", {"CHOL_MGPU_INVERSE": "dist"}, list(range(ngpu))),
         ("slab, all GPUs", {"CHOL_MGPU_INVERSE": "slab"}, list(range(ngpu))))
for label, env, devs in MODES:
    if os.environ.get("AB_MODES") and label.split(",")[0].split()[0] not in os.environ["AB_MODES"].split(","):
        continue
    for k in ("CHOL_MGPU_INVERSE", "CHOL_MGPU_INVERSE_SLABS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    mg = chol.MultiGPU(devs)
    best = None
    Xp = torch.empty(N, N, dtype=torch.float64, pin_memory=True)  # page-locked, like the POTRF e2e runs
    X = Xp.numpy().T  # column-major view of the pinned buffer
    for rep in range(2):
        X[:, :] = factor
        t0 = time.perf_counter()
        info = mg.potri("d", "L", N, Xp.data_ptr(), N)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    mg.close()
    Xl = np.tril(X)
    Xs = Xl + np.tril(X, -1).T
    m = min(N, 2048)  # check a leading slice of A * A^-1 on the host
    err = np.abs(host0[:m, :] @ Xs[:, :m] - np.eye(m)).max()
    print("potri N=%d %-22s %d GPU(s): %.3f s  (%.1f TFLOP/s incl. PCIe)  info=%d  max|A A^-1 - I|=%.2e"
          % (N, label, len(devs), best, 2.0 * N ** 3 / 3.0 / best / 1e12, info, err))
