"""GPU parity tests of the BLAS-3 surface (cu{S,D,C,Z}{gemm2,syrk|herk,trmm2,trmm,trsm}) through the
C-ABI, against the CPU oracle, with the reference's own per-element flop-count bounds
(test/blas/dgemm.c:160-168: |err| <= flops_per_element * 2 eps, scaled by the data magnitude)."""
import itertools

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
TRANS = {"s": "NT", "d": "NT", "c": "NTC", "z": "NTC"}


@pytest.fixture(scope="module")
def blas():
    h = G.chol.BlasHandle()
    yield h
    h.close()


def bound(p, k, scale=1.0, extra=8):
    per = (8 * k) if PREC[p]["cplx"] else (2 * k)
    return (per + extra) * 2 * H.eps(p) * scale


@pytest.mark.parametrize("p", "dzsc")
@pytest.mark.parametrize("shape", [(128, 64, 16), (256, 192, 64), (130, 70, 48), (37, 29, 53), (1, 1, 1), (300, 5, 17), (64, 333, 128)])
def test_gemm2_all_ops(blas, p, shape):
    m, n, k = shape
    cplx = PREC[p]["cplx"]
    alpha, beta = (0.7 - 0.2j, -0.3 + 0.5j) if cplx else (0.7, -0.3)
    for tA, tB in itertools.product(TRANS[p], TRANS[p]):
        A = H.rand_matrix(p, *((m, k) if tA == "N" else (k, m)), seed=1, ld=(m if tA == "N" else k) + (m + k) % 2)
        B = H.rand_matrix(p, *((k, n) if tB == "N" else (n, k)), seed=2)
        Cm = H.rand_matrix(p, m, n, seed=3, ld=m + 2)
        want = Cm.copy(order="F")
        ORA.gemm(p, tA, tB, m, n, k, alpha, A, B, beta, want)
        dA, lda = G.to_dev(A)
        dB, ldb = G.to_dev(B)
        dC, ldc = G.to_dev(Cm)
        dD = torch.zeros_like(dC)
        blas.gemm2(p, tA, tB, m, n, k, alpha, dA.data_ptr(), lda, dB.data_ptr(), ldb, beta, dC.data_ptr(), ldc,
                   dD.data_ptr(), ldc)
        G.sync()
        got = G.to_host(dD, m)
        assert np.abs(got - want).max() <= bound(p, k), (tA, tB)
        assert torch.equal(dC.cpu(), G.to_dev(Cm)[0].cpu())  # C untouched when D != C


@pytest.mark.parametrize("p", "dz")
def test_gemm_tma_path_taken_and_matches_generic(p):
    """The hot NT configuration must run on the TMA-fed kernel, and agree bit-for-bit with the
    generic-loader kernel (same tile arithmetic, different operand path)."""
    m, n, k = 384, 320, 256
    A = H.rand_matrix(p, m, k, seed=11)
    B = H.rand_matrix(p, n, k, seed=12)
    Cm = H.rand_matrix(p, m, n, seed=13)
    dA, lda = G.to_dev(A)
    dB, ldb = G.to_dev(B)
    outs = []
    for force in (0, 1):
        dC, ldc = G.to_dev(Cm)
        used = G.chol.gemm_desc_launch(prec=p, opA="N", opB="C", m=m, n=n, k=k, alpha=-1.0, beta=1.0, A=dA.data_ptr(), lda=lda,
                                       B=dB.data_ptr(), ldb=ldb, C=dC.data_ptr(), ldc=ldc, D=dC.data_ptr(), ldd=ldc,
                                       force_generic=force)
        assert used == (1 - force)
        G.sync()
        outs.append(G.to_host(dC, m).copy())
    assert np.array_equal(outs[0], outs[1])
    want = Cm.copy(order="F")
    ORA.gemm(p, "N", "C", m, n, k, -1.0, A, B, 1.0, want)
    assert np.abs(outs[0] - want).max() <= bound(p, k)


@pytest.mark.parametrize("p", "dz")
@pytest.mark.parametrize("ops", [("N", "T"), ("T", "N"), ("N", "N")])
def test_gemm_large_grid_pingpong_kernel(blas, p, ops):
    """Grids with >= 4 tiles per SM run on the ping-pong (two consumer groups, one producer) kernel:
    check it, TMA-fed and generic, against the oracle at a size with ragged edges."""
    tA, tB = ops
    if PREC[p]["cplx"]:
        tA, tB = tA.replace("T", "C"), tB.replace("T", "C")
    m, n, k = (4226, 2114, 80) if p == "d" else (2114, 2114, 40)
    A = H.rand_matrix(p, *((m, k) if tA == "N" else (k, m)), seed=31)
    B = H.rand_matrix(p, *((k, n) if tB == "N" else (n, k)), seed=32)
    Cm = H.rand_matrix(p, m, n, seed=33)
    want = Cm.copy(order="F")
    ORA.gemm(p, tA, tB, m, n, k, -1.0, A, B, 1.0, want)
    dA, lda = G.to_dev(A)
    dB, ldb = G.to_dev(B)
    dC, ldc = G.to_dev(Cm)
    blas.gemm2(p, tA, tB, m, n, k, -1.0, dA.data_ptr(), lda, dB.data_ptr(), ldb, 1.0, dC.data_ptr(), ldc, dC.data_ptr(), ldc)
    G.sync()
    assert np.abs(G.to_host(dC, m) - want).max() <= bound(p, k)


@pytest.mark.parametrize("p", "dzsc")
@pytest.mark.parametrize("n,k", [(200, 64), (129, 33), (64, 256), (513, 48)])
def test_rankk(blas, p, n, k):
    for uplo in "LU":
        for trans in ("N", "C" if PREC[p]["cplx"] else "T"):
            A = H.rand_matrix(p, *((n, k) if trans == "N" else (k, n)), seed=4)
            C0 = H.spd_matrix(p, n, seed=5, k=8)
            want = C0.copy(order="F")
            ORA.rankk(p, uplo, trans, n, k, -0.5, A, 1.25, want)
            dA, lda = G.to_dev(A)
            dC, ldc = G.to_dev(C0)
            blas.rankk(p, uplo, trans, n, k, -0.5, dA.data_ptr(), lda, 1.25, dC.data_ptr(), ldc)
            G.sync()
            got = G.to_host(dC, n)
            assert np.abs(got - want).max() <= bound(p, k, scale=np.abs(want).max()), (uplo, trans)
            # the opposite triangle is never written
            other = np.triu(got, 1) if uplo == "L" else np.tril(got, -1)
            other0 = np.triu(C0, 1) if uplo == "L" else np.tril(C0, -1)
            assert np.array_equal(other, other0)


@pytest.mark.parametrize("p", "dzsc")
@pytest.mark.parametrize("m,n", [(41, 23), (200, 130), (130, 260)])
def test_trmm2_trmm_trsm(blas, p, m, n):
    cplx = PREC[p]["cplx"]
    alpha = (0.9 + 0.3j) if cplx else 0.9
    for side in "LR":
        na = m if side == "L" else n
        T0 = H.spd_matrix(p, na, seed=6, k=16)
        ORA.potrf(p, "L", T0)
        T0 = (np.tril(T0) + np.tril(T0, -1).T).copy(order="F")
        for uplo, trans, diag in itertools.product("LU", TRANS[p], "NU"):
            Tm = T0.copy(order="F")
            if diag == "U":
                Tm /= 8 * np.abs(Tm).max()
            dT, ldt = G.to_dev(Tm)
            B0 = H.rand_matrix(p, m, n, seed=7, ld=m + 1)
            # trmm2
            want = H.fmat(m, n, PREC[p]["np"])
            ORA.trmm2(p, side, uplo, trans, diag, m, n, alpha, Tm, B0, want)
            dB, ldb = G.to_dev(B0)
            dX = torch.zeros_like(dB)
            blas.trmm2(p, side, uplo, trans, diag, m, n, alpha, dT.data_ptr(), ldt, dB.data_ptr(), ldb, dX.data_ptr(), ldb)
            G.sync()
            assert np.abs(G.to_host(dX, m) - want).max() <= bound(p, na, scale=max(1.0, np.abs(want).max())), ("trmm2", side, uplo, trans, diag)
            # trmm in place
            blas.trmm(p, side, uplo, trans, diag, m, n, alpha, dT.data_ptr(), ldt, dB.data_ptr(), ldb)
            G.sync()
            assert np.abs(G.to_host(dB, m) - want).max() <= bound(p, na, scale=max(1.0, np.abs(want).max())), ("trmm", side, uplo, trans, diag)
            # trsm
            want = B0.copy(order="F")
            ORA.trsm(p, side, uplo, trans, diag, m, n, alpha, Tm, want)
            dB, ldb = G.to_dev(B0)
            blas.trsm(p, side, uplo, trans, diag, m, n, alpha, dT.data_ptr(), ldt, dB.data_ptr(), ldb)
            G.sync()
            got = G.to_host(dB, m)
            assert np.abs(got - want).max() <= 64 * na * H.eps(p) * np.abs(want).max(), ("trsm", side, uplo, trans, diag)


@pytest.mark.parametrize("p", "ds")
def test_trsm_singular_operand_gives_non_finite_like_the_blas(blas, p):
    """A zero on the triangular operand's diagonal: the reference's substitution TRSM (blas/dtrsm.c) divides by
    it and returns Inf/NaN.  The inverse-based TRSM here must not return finite garbage instead, and one
    singular 128-block must not disable the blocks after it (the later, unaffected rows stay correct)."""
    m, n = 300, 40
    T0 = H.spd_matrix(p, m, seed=6, k=16)
    ORA.potrf(p, "L", T0)
    T0 = np.tril(T0).copy(order="F")
    Tm = T0.copy(order="F")
    Tm[150, 150] = 0.0  # inside the second 128-block
    B0 = H.rand_matrix(p, m, n, seed=7, ld=m + 1)
    want = B0.copy(order="F")
    ORA.trsm(p, "L", "L", "N", "N", m, n, 1.0, Tm, want)
    dT, ldt = G.to_dev(Tm)
    dB, ldb = G.to_dev(B0)
    blas.trsm(p, "L", "L", "N", "N", m, n, 1.0, dT.data_ptr(), ldt, dB.data_ptr(), ldb)
    G.sync()
    got = G.to_host(dB, m)
    assert not np.isfinite(want[150:]).all()  # the oracle's own result is Inf/NaN from the singular row on
    assert not np.isfinite(got[150:]).all(), "singular TRSM returned finite values"
    # rows before the singular block row depend on nothing singular (forward substitution): still exact
    assert np.abs(got[:128] - want[:128]).max() <= 64 * m * H.eps(p) * np.abs(want[:128]).max()


def test_blas_argument_errors(blas):
    G.chol.set_xerbla_silent()
    d = torch.zeros(64, dtype=torch.float64, device="cuda").data_ptr()
    inv = G.chol.CUDA_ERROR_INVALID_VALUE
    assert blas.gemm2("d", "N", "N", 8, 8, 8, 1.0, d, 4, d, 8, 0.0, d, 8, d, 8, allow_invalid=True) == inv  # lda (8)
    assert blas.gemm2("d", "N", "N", 8, 8, 8, 1.0, d, 8, d, 8, 0.0, d, 8, d, 4, allow_invalid=True) == inv  # ldd (15)
    assert blas.rankk("z", "L", "T", 4, 4, 1.0, d, 4, 0.0, d, 4, allow_invalid=True) == inv  # herk trans=T (2)
    assert blas.trsm("d", "L", "L", "N", "N", 8, 2, 1.0, d, 4, d, 8, allow_invalid=True) == inv  # lda (9)
    assert blas.gemm2("d", "N", "N", 0, 8, 8, 1.0, d, 8, d, 8, 0.0, d, 8, d, 8) == 0  # m == 0: nothing to do
