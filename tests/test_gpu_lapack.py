"""GPU parity tests of cu{S,D,C,Z}{potrf,trtri,trti22,lauum,potri,logdet} through the C-ABI against
the CPU oracle (which is pinned to the reference, tests/test_oracle_pins.py).

Tolerances: elementwise max|X - X_ref| <= 64 n eps max|X_ref| and residual
||A - L L^H||_F / ||A||_F <= 8 n eps; `info` must be EQUAL to the oracle's."""
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


@pytest.fixture(scope="module")
def lapack():
    h = G.chol.LapackHandle()
    yield h
    h.close()


def elem_tol(p, n):
    return 64 * n * H.eps(p)


SIZES = [1, 2, 31, 64, 100, 128, 129, 256, 257, 300, 513, 777]


@pytest.mark.parametrize("p", "dzsc")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("n", SIZES)
def test_potrf_vs_oracle(lapack, p, uplo, n):
    A = H.spd_matrix(p, n, ld=n + (n % 3), k=min(n, 64))
    A0 = A.copy()
    want = A.copy(order="F")
    assert ORA.potrf(p, uplo, want) == 0
    dA, lda = G.to_dev(A)
    assert lapack.potrf(p, uplo, n, dA.data_ptr(), lda) == 0
    got = G.to_host(dA, n)
    err = np.abs(H.tri(got, uplo) - H.tri(want, uplo)).max() / np.abs(want).max()
    assert err <= elem_tol(p, n), err
    # the opposite (strict) triangle is untouched, bit for bit
    other = np.triu(got, 1) if uplo == "L" else np.tril(got, -1)
    other0 = np.triu(A0, 1) if uplo == "L" else np.tril(A0, -1)
    assert np.array_equal(other, other0)
    L = H.tri(got, uplo).astype(np.complex128 if PREC[p]["cplx"] else np.float64)
    R = L @ L.conj().T if uplo == "L" else L.conj().T @ L
    assert np.linalg.norm(R - A0) / np.linalg.norm(A0) <= 8 * n * H.eps(p)


@pytest.mark.parametrize("p", "dz")
@pytest.mark.parametrize("uplo", "LU")
def test_potrf_golden_fixture(lapack, p, uplo):
    """Directly against factors produced by the reference's own CPU code (tests/golden)."""
    g = np.load(os.path.join(H.GOLDEN, "potrf_%s.npz" % p))
    for n in (33, 100, 257):
        A = H.spd_matrix(p, n, ld=n + 3)
        dA, lda = G.to_dev(A)
        assert lapack.potrf(p, uplo, n, dA.data_ptr(), lda) == 0
        got, want = G.to_host(dA, n), g["%s_%d" % (uplo, n)]
        assert np.abs(H.tri(got, uplo) - H.tri(want, uplo)).max() <= elem_tol(p, n) * np.abs(want).max()


@pytest.mark.parametrize("p", "dzsc")
@pytest.mark.parametrize("uplo", "LU")
def test_potrf_not_positive_definite_info(lapack, p, uplo):
    n = 700
    for bad in (0, 5, 255, 256, 300, 699):
        A = H.spd_matrix(p, n, k=32)
        A[bad, bad] = -1.0
        want = A.copy(order="F")
        i_ref = ORA.potrf(p, uplo, want)
        dA, lda = G.to_dev(A)
        assert lapack.potrf(p, uplo, n, dA.data_ptr(), lda) == i_ref == bad + 1


def test_lapack_argument_errors_and_empty(lapack):
    G.chol.set_xerbla_silent()
    d = torch.zeros(64, dtype=torch.float64, device="cuda").data_ptr()
    assert lapack.potrf("d", "L", 8, d, 4, allow_invalid=True) == (-4, G.chol.CUDA_ERROR_INVALID_VALUE)
    assert lapack.trtri("d", "L", "N", 8, d, 4, allow_invalid=True) == (-5, G.chol.CUDA_ERROR_INVALID_VALUE)
    assert lapack.lauum("d", "U", 8, d, 4, allow_invalid=True) == (-4, G.chol.CUDA_ERROR_INVALID_VALUE)
    assert lapack.potri("z", "U", 8, d, 4, allow_invalid=True) == (-4, G.chol.CUDA_ERROR_INVALID_VALUE)
    assert lapack.potrf("d", "L", 0, d, 0) == 0
    assert lapack.logdet("d", d, 1, 0) == 0.0


@pytest.mark.parametrize("p", "dzsc")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("diag", "NU")
@pytest.mark.parametrize("n", [3, 64, 129, 256, 300, 600])
def test_trtri_vs_oracle(lapack, p, uplo, diag, n):
    A = H.spd_matrix(p, n, k=min(n, 64))
    ORA.potrf(p, uplo, A)
    if diag == "U":
        A /= np.abs(np.diag(A)).max() * 4
    A0 = A.copy()
    want = A.copy(order="F")
    assert ORA.trtri(p, uplo, diag, want) == 0
    dA, lda = G.to_dev(A)
    assert lapack.trtri(p, uplo, diag, n, dA.data_ptr(), lda) == 0
    got = G.to_host(dA, n)
    assert np.abs(H.tri(got, uplo) - H.tri(want, uplo)).max() <= elem_tol(p, n) * np.abs(H.tri(want, uplo)).max()
    other = np.triu(got, 1) if uplo == "L" else np.tril(got, -1)
    other0 = np.triu(A0, 1) if uplo == "L" else np.tril(A0, -1)
    assert np.array_equal(other, other0)
    if diag == "U":
        assert np.array_equal(np.diag(got), np.diag(A0))  # unit diagonal is never referenced or written


@pytest.mark.parametrize("p", "dz")
@pytest.mark.parametrize("uplo", "LU")
def test_trtri_singular_info_and_trti22(lapack, p, uplo):
    n = 600
    A = H.spd_matrix(p, n, k=32)
    ORA.potrf(p, uplo, A)
    A[77, 77] = 0
    A[400, 400] = 0
    want = A.copy(order="F")
    i_ref = ORA.trtri(p, uplo, "N", want)
    dA, lda = G.to_dev(A)
    assert lapack.trtri(p, uplo, "N", n, dA.data_ptr(), lda) == i_ref
    # unblocked out-of-place block inverse
    nb = 64
    Bm = H.spd_matrix(p, nb, seed=3)
    ORA.potrf(p, uplo, Bm)
    wantb = Bm.copy(order="F")
    ORA.trtri(p, uplo, "N", wantb)
    dB, ldb = G.to_dev(Bm)
    dX = torch.zeros_like(dB)
    # the reference's contract (lapack/dtrtri.cu:23-29): info is a device *int* the kernel zeroes itself -
    # neither 8 bytes wide nor pre-zeroed; the guard word behind it must survive
    dinfo = torch.full((2,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
    lapack.trti22(p, uplo, "N", nb, dB.data_ptr(), ldb, dX.data_ptr(), ldb, dinfo.data_ptr())
    G.sync()
    assert dinfo.tolist() == [0, 0x5A5A5A5A]
    assert np.abs(H.tri(G.to_host(dX, nb), uplo) - H.tri(wantb, uplo)).max() <= elem_tol(p, nb) * np.abs(wantb).max()
    # zero pivots: the first one the reference's sweep meets (bottom-up for lower, top-down for upper)
    Bm[10, 10] = 0
    Bm[40, 40] = 0
    dB, ldb = G.to_dev(Bm)
    lapack.trti22(p, uplo, "N", nb, dB.data_ptr(), ldb, dX.data_ptr(), ldb, dinfo.data_ptr())
    G.sync()
    assert dinfo.tolist() == [41 if uplo == "L" else 11, 0x5A5A5A5A]


@pytest.mark.parametrize("p", "dzsc")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("n", [2, 100, 256, 257, 640])
def test_lauum_potri_logdet_vs_oracle(lapack, p, uplo, n):
    A = H.spd_matrix(p, n, k=min(n, 64))
    A0 = A.copy()
    ORA.potrf(p, uplo, A)
    F = A.copy(order="F")
    # lauum
    want = F.copy(order="F")
    ORA.lauum(p, uplo, want)
    dA, lda = G.to_dev(F)
    assert lapack.lauum(p, uplo, n, dA.data_ptr(), lda) == 0
    got = G.to_host(dA, n)
    assert np.abs(H.tri(got, uplo) - H.tri(want, uplo)).max() <= elem_tol(p, n) * np.abs(want).max()
    # logdet of the factor's diagonal
    dF, ldf = G.to_dev(F)
    ld = lapack.logdet(p, dF.data_ptr(), ldf + 1, n)
    ld_ref = ORA.logdet(p, F, F.strides[1] // F.itemsize + 1, n)
    e = H.eps("s") if p != "d" else H.eps("d")  # c/z oracle uses logf (lapack/zlogdet.c:18)
    assert abs(ld - ld_ref) <= 4 * n * e * abs(ld_ref)
    # potri
    want = F.copy(order="F")
    assert ORA.potri(p, uplo, want) == 0
    assert lapack.potri(p, uplo, n, dF.data_ptr(), ldf) == 0
    got = G.to_host(dF, n)
    assert np.abs(H.tri(got, uplo) - H.tri(want, uplo)).max() <= elem_tol(p, n) * np.abs(H.tri(want, uplo)).max()
    T = H.tri(got, uplo)
    Ainv = T + T.conj().T - np.diag(np.diag(T))
    assert np.abs(A0.astype(np.complex128) @ Ainv - np.eye(n)).max() <= 32 * n * H.eps(p)
    other = np.triu(got, 1) if uplo == "L" else np.tril(got, -1)
    other0 = np.triu(F, 1) if uplo == "L" else np.tril(F, -1)
    assert np.array_equal(other, other0)


@pytest.mark.parametrize("uplo", "LU")
def test_dpotrf_large_properties(lapack, uplo):
    """Full-size properties on the GPU (no oracle at this size): residual, leading-block parity
    (the factor of a leading block is the leading block of the factor), logdet consistency."""
    n, k = 8192, 512
    g = torch.Generator(device="cuda").manual_seed(H.SEED)
    B = torch.rand(n, k, dtype=torch.float64, device="cuda", generator=g) - 0.5
    A = B @ B.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    F = A.clone()
    assert lapack.potrf("d", uplo, n, F.data_ptr(), n) == 0
    # torch row-major (n,n) tensor == column-major transpose: lower(col-major) is torch's upper
    Lt = torch.triu(F) if uplo == "L" else torch.tril(F)  # = L^T (row-major view) / U^T
    R = Lt.T @ Lt if uplo == "L" else Lt @ Lt.T
    resid = (torch.linalg.norm(R - A) / torch.linalg.norm(A)).item()
    assert resid <= 1e-13, resid
    m = 1024
    lead = np.asfortranarray(A[:m, :m].cpu().numpy().T)
    assert ORA.potrf("d", uplo, lead) == 0
    got = F[:m, :m].cpu().numpy().T
    assert np.abs(H.tri(got, uplo) - H.tri(lead, uplo)).max() <= elem_tol("d", m) * np.abs(lead).max()
    ld = lapack.logdet("d", F.data_ptr(), n + 1, n)
    want = 2 * torch.log(torch.diagonal(F)).sum().item()
    assert abs(ld - want) <= 1e-12 * abs(want)


@pytest.mark.parametrize("p,n", [("d", 4096), ("z", 2048), ("s", 4096), ("c", 2048)])
@pytest.mark.parametrize("uplo", "LU")
def test_potrf_full_elementwise_vs_oracle_mid_size(lapack, p, uplo, n):
    """SURVEY.md section 8d, config 2 style: the WHOLE factor against the CPU oracle (which is pinned to the
    reference's dpotrf), elementwise max|L - L_ref| <= 64 n eps max|L_ref|, at a size where several
    outer blocks, the look-ahead and the triangular-grid update kernel are all in play."""
    A = H.spd_matrix(p, n, k=256)
    want = A.copy(order="F")
    assert ORA.potrf(p, uplo, want) == 0
    dA, lda = G.to_dev(A)
    assert lapack.potrf(p, uplo, n, dA.data_ptr(), lda) == 0
    got = G.to_host(dA, n)
    err = np.abs(H.tri(got, uplo) - H.tri(want, uplo)).max() / np.abs(want).max()
    assert err <= elem_tol(p, n), err


@pytest.mark.slow
@pytest.mark.parametrize("uplo", "LU")
def test_potrf_full_elementwise_n16384_vs_reference(lapack, uplo):
    """BASELINE.json configs[1] at its stated size: dpotrf N = 16384, the WHOLE factor elementwise against the
    reference's own CPU dpotrf (oracle/_ref, OpenMP on the box's host cores: ~1 min per triangle), tolerance
    64 n eps max|L_ref| (SURVEY.md section 8d cfg 2), plus one non-PD case with equal info."""
    if H.ref() is None:
        pytest.skip("compiled reference (oracle/_ref) not available")
    import ctypes
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(ctypes.c_int(len(os.sched_getaffinity(0))))
    except OSError:
        pass
    REF = H.CpuLib("ref")
    n = 16384
    g = torch.Generator(device="cuda").manual_seed(H.SEED)
    B = torch.rand(n, 1024, dtype=torch.float64, device="cuda", generator=g) - 0.5
    A = B @ B.T
    A.diagonal().add_(float(n))
    del B
    want = np.asfortranarray(A.cpu().numpy())  # symmetric: row-major == column-major
    assert REF.potrf("d", uplo, want) == 0
    F = A.clone()
    assert lapack.potrf("d", uplo, n, F.data_ptr(), n) == 0
    got = F.cpu().numpy().T  # column-major view
    err = np.abs(H.tri(got, uplo) - H.tri(want, uplo)).max() / np.abs(H.tri(want, uplo)).max()
    assert err <= elem_tol("d", n), err
    # not positive definite at a column deep inside: same info as the reference (leading block only, to keep it short)
    m = 3000
    Ab = np.asfortranarray(A[:m, :m].cpu().numpy())
    Ab[2500, 2500] = -1.0
    wantb = Ab.copy(order="F")
    winfo = REF.potrf("d", uplo, wantb)
    dA, lda = G.to_dev(Ab)
    assert lapack.potrf("d", uplo, m, dA.data_ptr(), lda) == winfo == 2501


@pytest.mark.parametrize("p", "dzsc")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("n,nrhs", [(1, 1), (100, 7), (300, 64), (700, 130)])
def test_potrs_solves_the_system(lapack, p, uplo, n, nrhs):
    """cu*potrs (extension): X from the GPU factor against the oracle's two triangular solves on the
    oracle's factor, and the residual ||A X - B|| / (||A|| ||X||)."""
    A0 = H.spd_matrix(p, n, k=min(n, 64))
    B0 = H.rand_matrix(p, n, nrhs, seed=11, ld=n + 2)
    F = A0.copy(order="F")
    assert ORA.potrf(p, uplo, F) == 0
    want = B0.copy(order="F")
    hT = "C" if PREC[p]["cplx"] else "T"
    for trans in (("N", hT) if uplo == "L" else (hT, "N")):
        ORA.trsm(p, "L", uplo, trans, "N", n, nrhs, 1.0, F, want)
    dA, lda = G.to_dev(A0)
    dB, ldb = G.to_dev(B0)
    assert lapack.potrf(p, uplo, n, dA.data_ptr(), lda) == 0
    assert lapack.potrs(p, uplo, n, nrhs, dA.data_ptr(), lda, dB.data_ptr(), ldb) == 0
    got = G.to_host(dB, n)[:, :nrhs]
    assert np.abs(got - want[:, :nrhs]).max() <= 64 * n * H.eps(p) * np.abs(want).max()
    wide = np.complex128 if PREC[p]["cplx"] else np.float64
    R = A0.astype(wide) @ got.astype(wide) - B0[:, :nrhs].astype(wide)
    assert np.linalg.norm(R) <= 16 * n * H.eps(p) * np.linalg.norm(A0) * np.linalg.norm(got)


@pytest.mark.parametrize("p", "dzsc")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("n", [1, 100, 129, 777, 2000])
def test_potrf_fused_logdet(lapack, p, uplo, n):
    """cuLAPACKPotrfLogdet (extension): the log-determinant the diagonal-block kernels accumulate
    while they factor equals the reference-style logdet of the factor's diagonal; a failed factorisation has none."""
    A = H.spd_matrix(p, n, k=min(n, 64))
    F = A.copy(order="F")
    assert ORA.potrf(p, uplo, F) == 0
    want = ORA.logdet(p, F, F.strides[1] // F.itemsize + 1, n)
    dA, lda = G.to_dev(A)
    assert lapack.potrf(p, uplo, n, dA.data_ptr(), lda) == 0
    got = lapack.potrf_logdet()
    tol = 1e-12 if p == "d" else 2e-6  # the reference's c AND z logdet use logf (lapack/zlogdet.c:18), s / c are FP32 anyway
    assert abs(got - want) <= tol * max(1.0, abs(want)), (got, want)
    sep = lapack.logdet(p, dA.data_ptr(), lda + 1, n)
    assert abs(got - sep) <= (1e-12 if p in "dz" else 1e-5) * max(1.0, abs(sep))
    if n >= 100:
        A[n // 2, n // 2] = -1.0
        dA, lda = G.to_dev(A)
        assert lapack.potrf(p, uplo, n, dA.data_ptr(), lda) == n // 2 + 1
        with pytest.raises(G.chol.CholeskyError):
            lapack.potrf_logdet()


def test_potrs_argument_errors(lapack):
    G.chol.set_xerbla_silent()
    A = H.spd_matrix("d", 8)
    dA, lda = G.to_dev(A)
    assert lapack.potrs("d", "L", 16, 1, dA.data_ptr(), 8, dA.data_ptr(), 16, allow_invalid=True) == (-5, 1)
    assert lapack.potrs("d", "L", 8, 1, dA.data_ptr(), lda, dA.data_ptr(), 4, allow_invalid=True) == (-7, 1)
    assert lapack.potrs("d", "L", 0, 3, dA.data_ptr(), lda, dA.data_ptr(), lda) == 0


@pytest.mark.parametrize("p", "dzsc")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("n", [2, 64, 300, 1000])
def test_reference_test_inputs_latmc(lapack, p, uplo, n):
    """The reference's LAPACK tests factor latmc(n, c = 2) after srand(0) and check `info` only
    (test/lapack/dpotrf.c:56-67): same inputs here, info AND values, for the whole chain
    potrf -> trtri / lauum / potri -> logdet."""
    A0 = H.latmc(p, n, ld=n + 2)
    want = A0.copy(order="F")
    assert ORA.potrf(p, uplo, want) == 0
    dA, lda = G.to_dev(A0)
    assert lapack.potrf(p, uplo, n, dA.data_ptr(), lda) == 0
    got = G.to_host(dA, n)
    assert np.abs(H.tri(got, uplo) - H.tri(want, uplo)).max() <= elem_tol(p, n) * np.abs(want).max()
    if n == 64 and uplo in "LU":  # frozen copy of the reference's own factor of this input
        g = np.load(os.path.join(H.GOLDEN, "latmc.npz"))
        ref = g["%s_%d_%s" % (p, n, uplo)]
        assert np.abs(H.tri(got, uplo) - ref).max() <= elem_tol(p, n) * np.abs(ref).max()
    inv = want.copy(order="F")
    assert ORA.potri(p, uplo, inv) == 0
    assert lapack.potri(p, uplo, n, dA.data_ptr(), lda) == 0
    gotinv = G.to_host(dA, n)
    assert np.abs(H.tri(gotinv, uplo) - H.tri(inv, uplo)).max() <= elem_tol(p, n) * np.abs(H.tri(inv, uplo)).max()
