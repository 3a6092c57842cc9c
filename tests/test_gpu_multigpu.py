"""GPU tests of the multi-GPU surface: the host-pointer cuMultiGPU* API on however many GPUs are
visible (1 is enough), the cholDist* SPMD core, and the block-cyclic trailing-update kernel
(emulating every rank of P = 2, 4 on ONE GPU, as separate launches on that rank's packed columns)."""
import os

import numpy as np
import pytest

import harness as H
from harness import PREC

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no GPU", allow_module_level=True)

import gpu_util as G  # noqa: E402

ORA = H.CpuLib("oracle")
NGPU = torch.cuda.device_count()


@pytest.fixture(scope="module")
def mgpu():
    m = G.chol.MultiGPU(list(range(NGPU)))
    yield m
    m.close()


@pytest.mark.parametrize("p", "dzsc")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("n", [5, 256, 700, 1300])
def test_multigpu_potrf_host_pointers(mgpu, p, uplo, n):
    A = H.spd_matrix(p, n, ld=n + 3, k=min(n, 64))
    A0 = A.copy()
    want = A.copy(order="F")
    assert ORA.potrf(p, uplo, want) == 0
    assert mgpu.potrf(p, uplo, n, A.ctypes.data, A.strides[1] // A.itemsize) == 0
    assert np.abs(H.tri(A, uplo) - H.tri(want, uplo)).max() <= 64 * n * H.eps(p) * np.abs(want).max()
    other = np.triu(A, 1) if uplo == "L" else np.tril(A, -1)
    other0 = np.triu(A0, 1) if uplo == "L" else np.tril(A0, -1)
    assert np.array_equal(other, other0)


@pytest.mark.parametrize("uplo", "LU")
def test_multigpu_potrf_pinned_host_overlapped_gather(mgpu, uplo):
    """Page-locked host matrices take the path that streams finished block columns back while the
    factorisation is still running (lower); the result must be the same."""
    n = 1500
    A = H.spd_matrix("d", n, k=64)
    want = A.copy(order="F")
    assert ORA.potrf("d", uplo, want) == 0
    hp = torch.empty(n, n, dtype=torch.float64, pin_memory=True)  # (n, lda) row-major == column-major A
    hp.copy_(torch.from_numpy(np.ascontiguousarray(A.T)))
    assert mgpu.potrf("d", uplo, n, hp.data_ptr(), n) == 0
    got = hp.numpy().T
    assert np.abs(H.tri(got, uplo) - H.tri(want, uplo)).max() <= 64 * n * H.eps("d") * np.abs(want).max()
    other = np.triu(got, 1) if uplo == "L" else np.tril(got, -1)
    other0 = np.triu(A, 1) if uplo == "L" else np.tril(A, -1)
    assert np.array_equal(other, other0)




@pytest.mark.parametrize("gbps", ["50", "0.0001"])
@pytest.mark.parametrize("n", [700, 1500, 3000])
def test_multigpu_potrf_upload_overlap(mgpu, n, gbps, monkeypatch):
    """CHOL_MGPU_H2D_OVERLAP=1: the factorisation starts while block columns are still being
    uploaded; a tiny assumed PCIe rate forces EVERY step through the group-by-group update."""
    monkeypatch.setenv("CHOL_MGPU_H2D_OVERLAP", "1")
    monkeypatch.setenv("CHOL_H2D_GBPS", gbps)
    A = H.spd_matrix("d", n, k=64)
    want = A.copy(order="F")
    assert ORA.potrf("d", "L", want) == 0
    hp = torch.empty(n, n, dtype=torch.float64, pin_memory=True)
    hp.copy_(torch.from_numpy(np.ascontiguousarray(A.T)))
    assert mgpu.potrf("d", "L", n, hp.data_ptr(), n) == 0
    got = hp.numpy().T
    assert np.abs(np.tril(got) - np.tril(want)).max() <= 64 * n * H.eps("d") * np.abs(want).max()
    assert np.array_equal(np.triu(got, 1), np.triu(A, 1))


@pytest.mark.parametrize("p", "dzs")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("n,slabs", [(400, 1), (400, 3), (1500, 4)])
def test_multigpu_inverse_slabs(mgpu, p, uplo, n, slabs, monkeypatch):
    """CHOL_MGPU_INVERSE=slab: TRTRI / LAUUM / POTRI split into independent slabs (several per GPU
    here, so one GPU exercises the slab algebra) against the oracle, other triangle untouched."""
    monkeypatch.setenv("CHOL_MGPU_INVERSE", "slab")
    monkeypatch.setenv("CHOL_MGPU_INVERSE_SLABS", str(slabs))
    A = H.spd_matrix(p, n, k=64)
    ORA.potrf(p, uplo, A)
    cases = [("trtri", "N"), ("trtri", "U"), ("lauum", None), ("potri", None)]
    for name, diag in cases:
        src = A / (np.abs(np.diag(A)).max() * 4) if diag == "U" else A  # keeps the unit-diagonal inverse well conditioned
        want = src.copy(order="F")
        got = src.copy(order="F")
        if name == "trtri":
            ORA.trtri(p, uplo, diag, want)
            assert mgpu.trtri(p, uplo, diag, n, got.ctypes.data, n) == 0
            if diag == "U":
                assert np.array_equal(np.diag(got), np.diag(src))
        else:
            getattr(ORA, name)(p, uplo, want)
            assert getattr(mgpu, name)(p, uplo, n, got.ctypes.data, n) == 0
        assert np.abs(H.tri(got, uplo) - H.tri(want, uplo)).max() <= 64 * n * H.eps(p) * np.abs(H.tri(want, uplo)).max(), (name, diag)
        other = np.triu(got, 1) if uplo == "L" else np.tril(got, -1)
        other0 = np.triu(src, 1) if uplo == "L" else np.tril(src, -1)
        assert np.array_equal(other, other0), (name, diag)


@pytest.mark.parametrize("uplo", "LU")
def test_multigpu_inverse_slabs_zero_pivot(mgpu, uplo, monkeypatch):
    monkeypatch.setenv("CHOL_MGPU_INVERSE", "slab")
    monkeypatch.setenv("CHOL_MGPU_INVERSE_SLABS", "3")
    n = 600
    A = H.spd_matrix("d", n, k=64)
    ORA.potrf("d", uplo, A)
    A[100, 100] = 0.0
    A[450, 450] = 0.0
    want = A.copy(order="F")
    winfo = ORA.trtri("d", uplo, "N", want)
    got = A.copy(order="F")
    assert mgpu.trtri("d", uplo, "N", n, got.ctypes.data, n) == winfo
    assert winfo == (451 if uplo == "L" else 101)


def test_multigpu_potrf_info_and_args(mgpu):
    G.chol.set_xerbla_silent()
    n = 900
    for uplo in "LU":
        A = H.spd_matrix("d", n, k=32)
        A[600, 600] = -2.0
        assert mgpu.potrf("d", uplo, n, A.ctypes.data, n) == 601
    A = H.fmat(4, 4, np.float64)
    assert mgpu.potrf("d", "L", 8, A.ctypes.data, 4, allow_invalid=True) == (-4, 1)
    assert mgpu.potrf("d", "L", 0, A.ctypes.data, 4) == 0


@pytest.mark.parametrize("p", "dz")
@pytest.mark.parametrize("uplo", "LU")
def test_multigpu_trtri_lauum_potri_host_pointers(mgpu, p, uplo):
    n = 400
    A = H.spd_matrix(p, n, k=64)
    ORA.potrf(p, uplo, A)
    for name, ref in (("trtri", lambda X: ORA.trtri(p, uplo, "N", X)), ("lauum", lambda X: ORA.lauum(p, uplo, X)),
                      ("potri", lambda X: ORA.potri(p, uplo, X))):
        want = A.copy(order="F")
        ref(want)
        got = A.copy(order="F")
        if name == "trtri":
            assert mgpu.trtri(p, uplo, "N", n, got.ctypes.data, n) == 0
        else:
            assert getattr(mgpu, name)(p, uplo, n, got.ctypes.data, n) == 0
        assert np.abs(H.tri(got, uplo) - H.tri(want, uplo)).max() <= 64 * n * H.eps(p) * np.abs(H.tri(want, uplo)).max(), name
        other = np.triu(got, 1) if uplo == "L" else np.tril(got, -1)
        other0 = np.triu(A, 1) if uplo == "L" else np.tril(A, -1)
        assert np.array_equal(other, other0)


@pytest.mark.parametrize("p", "dc")
def test_multigpu_blas_host_pointers(mgpu, p):
    cplx = PREC[p]["cplx"]
    m, n, k = 300, 200, 150
    alpha, beta = (0.7 - 0.2j, -0.3 + 0.5j) if cplx else (0.7, -0.3)
    hT = "C" if cplx else "T"
    for tA, tB in (("N", "N"), ("N", hT), (hT, "N")):
        A = H.rand_matrix(p, *((m, k) if tA == "N" else (k, m)), seed=1)
        B = H.rand_matrix(p, *((k, n) if tB == "N" else (n, k)), seed=2)
        Cm = H.rand_matrix(p, m, n, seed=3, ld=m + 2)
        want = Cm.copy(order="F")
        ORA.gemm(p, tA, tB, m, n, k, alpha, A, B, beta, want)
        mgpu.gemm(p, tA, tB, m, n, k, alpha, A.ctypes.data, A.strides[1] // A.itemsize, B.ctypes.data,
                  B.strides[1] // B.itemsize, beta, Cm.ctypes.data, Cm.strides[1] // Cm.itemsize)
        assert np.abs(Cm - want).max() <= (8 * k + 8) * 2 * H.eps(p), (tA, tB)
    for uplo in "LU":
        for trans in ("N", hT):
            A = H.rand_matrix(p, *((n, k) if trans == "N" else (k, n)), seed=4)
            C0 = H.spd_matrix(p, n, seed=5, k=8)
            want = C0.copy(order="F")
            ORA.rankk(p, uplo, trans, n, k, -0.5, A, 1.25, want)
            got = C0.copy(order="F")
            mgpu.rankk(p, uplo, trans, n, k, -0.5, A.ctypes.data, A.strides[1] // A.itemsize, 1.25, got.ctypes.data, n)
            assert np.abs(got - want).max() <= (8 * k + 8) * 2 * H.eps(p) * np.abs(want).max(), (uplo, trans)
    a1 = (0.9 + 0.3j) if cplx else 0.9
    for side in "LR":
        na = m if side == "L" else n
        T0 = H.spd_matrix(p, na, seed=6, k=16)
        ORA.potrf(p, "L", T0)
        T0 = (np.tril(T0) + np.tril(T0, -1).T).copy(order="F")
        for uplo, trans in (("L", "N"), ("U", hT)):
            B0 = H.rand_matrix(p, m, n, seed=7)
            want = B0.copy(order="F")
            ORA.trsm(p, side, uplo, trans, "N", m, n, a1, T0, want)
            got = B0.copy(order="F")
            mgpu.trsm(p, side, uplo, trans, "N", m, n, a1, T0.ctypes.data, na, got.ctypes.data, m)
            assert np.abs(got - want).max() <= 64 * na * H.eps(p) * np.abs(want).max(), ("trsm", side, uplo, trans)
            want = B0.copy(order="F")
            ORA.trmm(p, side, uplo, trans, "N", m, n, a1, T0, want)
            got = B0.copy(order="F")
            mgpu.trmm(p, side, uplo, trans, "N", m, n, a1, T0.ctypes.data, na, got.ctypes.data, m)
            assert np.abs(got - want).max() <= (8 * na + 8) * 2 * H.eps(p) * max(1.0, np.abs(want).max()), ("trmm", side, uplo, trans)


@pytest.mark.parametrize("p", "dz")
@pytest.mark.parametrize("P", [2, 4])
def test_block_cyclic_update_kernel_all_ranks_on_one_gpu(p, P):
    """One step of the distributed trailing update, every rank's launch run back to back on one
    GPU: together they must equal the global lower-triangular SYRK/HERK."""
    nb, n, c = (128, 1100, 1) if p == "z" else (256, 2100, 1)
    j, jb = c * nb, nb
    rest = n - j - jb
    Wm = H.rand_matrix(p, rest, jb, seed=21)
    A = H.spd_matrix(p, n, seed=22, k=8)
    want = A.copy(order="F")
    r0 = (c + 2) * nb  # global row/col where the REST update starts (column block c+1 is look-ahead)
    sub = want[r0:, r0:]
    W2 = np.ascontiguousarray(Wm[r0 - (j + jb):, :])
    upd = sub - W2 @ W2.conj().T
    want[r0:, r0:] = np.tril(upd) + np.triu(sub, 1)
    got = A.copy(order="F")
    dW, ldw = G.to_dev(Wm)
    nblk = (n + nb - 1) // nb
    es = np.dtype(PREC[p]["np"]).itemsize
    for rank in range(P):
        owned = [cb for cb in range(rank, nblk, P)]
        cols = np.concatenate([np.arange(cb * nb, min(n, cb * nb + nb)) for cb in owned])
        loc = H.fmat(n, len(cols), PREC[p]["np"])
        loc[:, :] = A[:, cols]
        dL, ldl = G.to_dev(loc)
        lc2 = max(0, -(-(c + 2 - rank) // P))
        col2 = lc2 * nb
        if col2 >= len(cols):
            continue
        g_first = (lc2 * P + rank) * nb
        Wp = dW.data_ptr() + (g_first - (j + jb)) * es
        Dp = dL.data_ptr() + (col2 * ldl + g_first) * es
        G.chol.gemm_desc_launch(prec=p, opA="N", opB="C", cuplo="L", herm=1 if p == "z" else 0, m=n - g_first, n=len(cols) - col2,
                                k=jb, alpha=-1.0, beta=1.0, A=Wp, lda=ldw, B=Wp, ldb=ldw, C=Dp, ldc=ldl, D=Dp, ldd=ldl,
                                cyc_nb=nb, cyc_p=P, cyc_off=0)
        G.sync()
        got[:, cols] = G.to_host(dL, n)
    assert np.abs(got - want).max() <= (8 * jb + 8) * 2 * H.eps(p) * np.abs(want).max()


@pytest.mark.parametrize("p", "dz")
def test_dist_single_rank_core(p):
    """cholDist* with one rank is the packed layout == the plain column-major matrix."""
    n = 1000
    d = G.chol.Dist(0, 1)
    A = H.spd_matrix(p, n, k=64)
    want = A.copy(order="F")
    ORA.potrf(p, "L", want)
    dA, lda = G.to_dev(A)
    assert d.potrf(p, n, dA.data_ptr(), lda) == 0
    got = G.to_host(dA, n)
    assert np.abs(np.tril(got) - np.tril(want)).max() <= 64 * n * H.eps(p) * np.abs(want).max()
    d.close()


def _cpu_potrf_fast(p, uplo, A):
    """the compiled reference (OpenMP) when it is there, else the oracle port"""
    lib = H.CpuLib("ref") if H.ref() is not None else ORA
    return lib.potrf(p, uplo, A)


@pytest.mark.parametrize("uplo", "LU")
def test_multigpu_potrf_large_all_gpus(mgpu, uplo):
    """In-process cuMultiGPUDpotrf over ALL visible devices (one worker thread per GPU, ncclCommInitAll inside the
    library; the reference's model, lapack/dpotrf.c:298-366, test/lapack/cumultigpudpotrf.c:47-60) at a size with
    several block columns per GPU, elementwise against the reference's CPU dpotrf."""
    n = 6400
    A = H.spd_matrix("d", n, k=64)
    A0 = A.copy()
    want = A.copy(order="F")
    assert _cpu_potrf_fast("d", uplo, want) == 0
    assert mgpu.potrf("d", uplo, n, A.ctypes.data, n) == 0
    assert np.abs(H.tri(A, uplo) - H.tri(want, uplo)).max() <= 64 * n * H.eps("d") * np.abs(want).max()
    other = np.triu(A, 1) if uplo == "L" else np.tril(A, -1)
    other0 = np.triu(A0, 1) if uplo == "L" else np.tril(A0, -1)
    assert np.array_equal(other, other0)
    F = H.tri(A, uplo)
    R = F @ F.T if uplo == "L" else F.T @ F
    assert np.linalg.norm(R - A0) / np.linalg.norm(A0) <= 1e-13


@pytest.mark.skipif(NGPU < 2, reason="needs >= 2 GPUs")
def test_cholDist_torchrun_parity_all_gpus(tmp_path):
    """One process per GPU under torch.distributed.run (the form bench.py --gpus N times): every rank factors its
    block-cyclic columns with cholDistDpotrf over NCCL; rank 0 gathers the factor and compares it elementwise with the
    reference's CPU dpotrf."""
    import subprocess
    import sys
    out = tmp_path / "dist_parity.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(NGPU), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(H.ROOT, "tests", "dist_parity_worker.py"), "5000", str(out)]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert p.returncode == 0, p.stdout.decode()[-4000:]
    import json
    r = json.load(open(out))
    assert r["info"] == 0 and r["ranks"] == NGPU
    assert r["maxrel"] <= 64 * r["n"] * H.eps("d"), r
