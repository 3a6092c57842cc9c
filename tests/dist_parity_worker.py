"""Worker of tests/test_gpu_multigpu.py::test_cholDist_torchrun_parity_all_gpus (launched under torch.distributed.run, one
rank per GPU): factor an N x N SPD matrix in the 1-D block-cyclic layout with cholDistDpotrf (NCCL panel broadcast),
gather the factor on rank 0 and compare it elementwise with the CPU reference.  argv: N, output json."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
from bench import exchange_unique_id, owned_columns, combine_info  # noqa: E402
from cholesky_b200_loader import load  # noqa: E402

n, out = int(sys.argv[1]), sys.argv[2]
rank, local_rank, P = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
dist.init_process_group("nccl", device_id=dev)
chol = load()
d = chol.Dist(rank, P, exchange_unique_id(chol.dist_unique_id, rank), nb=256)  # several block columns per rank at this n
nb = d.block_size("d", n)
cols = owned_columns(n, nb, rank, P)
A0 = H.spd_matrix("d", n, k=64)  # same seed on every rank: the same matrix
loc = torch.from_numpy(np.ascontiguousarray(A0[:, cols].T)).to(dev)  # (nloc, n) row-major == column-major n x nloc
info = combine_info(d.potrf("d", n, loc.data_ptr(), n))
parts = [None] * P
dist.all_gather_object(parts, (cols, loc.cpu().numpy()))
d.close()
if rank == 0:
    L = np.zeros((n, n), order="F")
    for gc, blk in parts:
        L[:, gc] = blk.T
    want = A0.copy(order="F")
    lib = H.CpuLib("ref") if H.ref() is not None else H.CpuLib("oracle")
    assert lib.potrf("d", "L", want) == 0
    maxrel = float(np.abs(np.tril(L) - np.tril(want)).max() / np.abs(want).max())
    json.dump({"n": n, "ranks": P, "nb": nb, "info": info, "maxrel": maxrel}, open(out, "w"))
dist.barrier()
dist.destroy_process_group()
