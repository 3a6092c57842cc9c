"""GPU-side helpers for the parity tests: device matrices are torch tensors used purely as HBM
allocations (column-major m x n with leading dimension ld == a row-major (n, ld) tensor)."""
import numpy as np
import torch

import harness as H
from cholesky_b200_loader import load

chol = load()

TORCH_DT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}


def to_dev(A):
    """numpy column-major view (from harness.fmat) -> device tensor (n, ld); returns (tensor, ld)."""
    ld = A.strides[1] // A.itemsize
    n = A.shape[1]
    base = np.lib.stride_tricks.as_strided(A, shape=(n, ld), strides=(A.strides[1], A.itemsize))
    return torch.from_numpy(np.ascontiguousarray(base)).cuda(), ld


def to_host(T, m):
    """device tensor (n, ld) -> numpy column-major m x n view"""
    a = T.cpu().numpy()
    return a.T[:m, :]


def sync():
    torch.cuda.synchronize()
