"""Pin the CPU oracle (oracle/*.c, our restatement) to the reference's own CPU code.

Two anchors (SURVEY.md §8c says the reference stores no golden vectors, so they are produced by
running the reference's sources, compiled unmodified by oracle/Makefile, in the build container):
  * live:   oracle/_ref/libref_oracle.so, when present (container and GPU box - it travels)
  * frozen: tests/golden/*.npz written by tests/golden/make_golden.py from that same library
"""
import json
import os

import numpy as np
import pytest

import harness as H
from harness import CpuLib, PREC

ORA = CpuLib("oracle")
HAVE_REF = H.ref() is not None
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref_oracle.so not built")


def tol(p, n, factor=64):
    return factor * n * H.eps(p)


def test_dsfmt_matches_reference_generator():
    with open(os.path.join(H.GOLDEN, "dsfmt19937_seed20121018.json")) as f:
        g = json.load(f)
    got = H.dsfmt_uniform(g["seed"], len(g["close_open"]))
    assert got.tolist() == g["close_open"]  # bit exact
    odd = H.dsfmt_uniform(g["seed"], 7)
    assert odd.tolist() == g["close_open"][:7]


@pytest.mark.parametrize("kind", list(H.DSFMT_KINDS))
def test_dsfmt_output_kinds_match_reference_generator(kind):
    """get / getCloseOpen / getOpenClose / getOpenOpen / getCloseClose of the reference's rng/dsfmt.c as bit patterns, from the
    start of the stream and after a first call of odd size (which drops the high half of its last word).  The reference's `get`
    (rng/dsfmt.c:127-146) stores whole words while more than two values remain and then writes the low half of one more word to
    res[i] with i counting WORDS: for n = 1001 that lands on value 500 instead of value 1000.  The golden list holds its first
    1000 values, so entry 500 must equal the restatement's value 1000 and everything else must match in place."""
    with open(os.path.join(H.GOLDEN, "dsfmt19937_kinds.json")) as f:
        g = json.load(f)
    want = g["kinds"][kind]
    def same(got, ref):
        got = ["%016x" % v for v in got]
        if kind == "raw":
            assert len(ref) == g["n"] - 1 and ref[g["n"] // 2] == got[g["n"] - 1]
            ref = ref[:g["n"] // 2] + [got[g["n"] // 2]] + ref[g["n"] // 2 + 1:]
        else:
            assert len(ref) == g["n"]
        assert got[:len(ref)] == ref

    gen = H.OracleDsfmt(g["seed"])
    same(gen.fill(g["n"], kind), want["first"])
    gen = H.OracleDsfmt(g["seed"])
    gen.fill(g["skip"], kind)
    same(gen.fill(g["n"], kind), want["after_skip"])
    if kind == "close_open":  # the older single-kind entry point is the same stream
        assert np.array_equal(H.dsfmt_uniform(g["seed"], 100).view(np.uint64), H.OracleDsfmt(g["seed"]).fill(100, kind))


@pytest.mark.parametrize("p", "sdcz")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 100, 257])
def test_potrf_golden(p, uplo, n):
    g = np.load(os.path.join(H.GOLDEN, "potrf_%s.npz" % p))
    key = "%s_%d" % (uplo, n)
    if key not in g:
        pytest.skip("no golden for " + key)
    A = H.spd_matrix(p, n, ld=n + 3)
    assert ORA.potrf(p, uplo, A) == 0
    want = g[key]
    err = np.abs(H.tri(A, uplo) - H.tri(want, uplo)).max() / np.abs(want).max()
    assert err <= tol(p, n), err


@needs_ref
@pytest.mark.parametrize("p", "sdcz")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("n", [1, 17, 64, 200, 513])
def test_potrf_vs_ref(p, uplo, n):
    REF = CpuLib("ref")
    A = H.spd_matrix(p, n, ld=n + 1)
    A0 = A.copy()
    B = A.copy(order="F")
    i1, i2 = ORA.potrf(p, uplo, A), REF.potrf(p, uplo, B)
    assert i1 == i2 == 0
    other = np.triu(A, 1) if uplo == "L" else np.tril(A, -1)
    other0 = np.triu(A0, 1) if uplo == "L" else np.tril(A0, -1)
    assert np.array_equal(other, other0)  # opposite triangle untouched
    err = np.abs(H.tri(A, uplo) - H.tri(B, uplo)).max() / np.abs(B).max()
    assert err <= tol(p, n), err
    L = H.tri(A, uplo).astype(np.complex128 if PREC[p]["cplx"] else np.float64)
    R = L @ L.conj().T if uplo == "L" else L.conj().T @ L
    resid = np.linalg.norm(R - A0) / np.linalg.norm(A0)
    assert resid <= 8 * n * H.eps(p)


@needs_ref
@pytest.mark.parametrize("p", "dz")
@pytest.mark.parametrize("uplo", "LU")
def test_potrf_info_nonpd_vs_ref(p, uplo):
    REF = CpuLib("ref")
    n = 150
    for bad in (0, 40, 149):
        A = H.spd_matrix(p, n)
        A[bad, bad] = -1.0
        B = A.copy(order="F")
        i1, i2 = ORA.potrf(p, uplo, A), REF.potrf(p, uplo, B)
        assert i1 == i2 == bad + 1


def test_potrf_argument_error():
    A = H.fmat(4, 4, np.float64, ld=4)
    import ctypes as C
    info = C.c_long(0)
    H.oracle().oracle_dpotrf(C.c_int(ord("L")), C.c_size_t(8), A.ctypes.data_as(C.c_void_p), C.c_size_t(4), C.byref(info))
    assert info.value == -4  # lapack/dpotrf.c:87-93


@needs_ref
@pytest.mark.parametrize("p", "sdcz")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("diag", "NU")
@pytest.mark.parametrize("n", [3, 63, 64, 65, 300])
def test_trtri_vs_ref(p, uplo, diag, n):
    REF = CpuLib("ref")
    A = H.spd_matrix(p, n)
    assert ORA.potrf(p, uplo, A) == 0
    if diag == "U":
        A /= np.abs(np.diag(A)).max() * 4  # keep the unit-diagonal inverse well conditioned
    B = A.copy(order="F")
    assert ORA.trtri(p, uplo, diag, A) == REF.trtri(p, uplo, diag, B) == 0
    err = np.abs(H.tri(A, uplo) - H.tri(B, uplo)).max() / np.abs(H.tri(B, uplo)).max()
    assert err <= tol(p, n), err


@needs_ref
@pytest.mark.parametrize("p", "dz")
@pytest.mark.parametrize("uplo", "LU")
def test_trtri2_and_singular_vs_ref(p, uplo):
    REF = CpuLib("ref")
    n = 130
    A = H.spd_matrix(p, n)
    ORA.potrf(p, uplo, A)
    B1 = H.fmat(n, n, PREC[p]["np"], ld=n)
    B2 = H.fmat(n, n, PREC[p]["np"], ld=n)
    assert ORA.trtri2(p, uplo, "N", A, B1) == REF.trtri2(p, uplo, "N", A, B2) == 0
    assert np.abs(H.tri(B1, uplo) - H.tri(B2, uplo)).max() <= tol(p, n) * np.abs(B2).max()
    A[77, 77] = 0
    C1, C2 = A.copy(order="F"), A.copy(order="F")
    assert ORA.trtri(p, uplo, "N", C1) == REF.trtri(p, uplo, "N", C2) == 78


@needs_ref
@pytest.mark.parametrize("p", "sdcz")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("n", [2, 31, 32, 33, 250])
def test_lauum_potri_vs_ref(p, uplo, n):
    REF = CpuLib("ref")
    A = H.spd_matrix(p, n)
    A0 = A.copy()
    ORA.potrf(p, uplo, A)
    L1, L2 = A.copy(order="F"), A.copy(order="F")
    assert ORA.lauum(p, uplo, L1) == REF.lauum(p, uplo, L2) == 0
    assert np.abs(H.tri(L1, uplo) - H.tri(L2, uplo)).max() <= tol(p, n) * np.abs(L2).max()
    P1, P2 = A.copy(order="F"), A.copy(order="F")
    assert ORA.potri(p, uplo, P1) == REF.potri(p, uplo, P2) == 0
    assert np.abs(H.tri(P1, uplo) - H.tri(P2, uplo)).max() <= tol(p, n) * np.abs(H.tri(P2, uplo)).max()
    # A * A^-1 = I with the symmetrised inverse
    T = H.tri(P1, uplo)
    Ainv = T + T.conj().T - np.diag(np.diag(T))
    assert np.abs(A0.astype(np.complex128) @ Ainv - np.eye(n)).max() <= 32 * n * H.eps(p)


@needs_ref
@pytest.mark.parametrize("p", "sdcz")
def test_logdet_vs_ref(p):
    REF = CpuLib("ref")
    n = 300
    A = H.spd_matrix(p, n, ld=n + 2)
    ORA.potrf(p, "L", A)
    lda = A.strides[1] // A.itemsize
    a, b = ORA.logdet(p, A, lda + 1, n), REF.logdet(p, A, lda + 1, n)
    # c/z use the single-precision log of the real part (lapack/zlogdet.c:18): float accuracy only
    e = H.eps("s") if PREC[p]["cplx"] else H.eps(p)
    assert abs(a - b) <= 4 * n * e * abs(b)
    want = 2 * np.log(np.real(np.diag(A)).astype(np.float64)).sum()
    assert abs(a - want) <= (2e-4 if p != "d" else 1e-12) * abs(want)
    assert ORA.logdet(p, A, 1, 0) == 0


TRANS = {"s": "NT", "d": "NT", "c": "NTC", "z": "NTC"}


@needs_ref
@pytest.mark.parametrize("p", "sdcz")
def test_gemm_vs_ref(p):
    REF = CpuLib("ref")
    m, n, k = 37, 29, 53
    cplx = PREC[p]["cplx"]
    alpha, beta = (0.7 - 0.2j, -0.3 + 0.5j) if cplx else (0.7, -0.3)
    for tA in TRANS[p]:
        for tB in TRANS[p]:
            A = H.rand_matrix(p, *((m, k) if tA == "N" else (k, m)), seed=1)
            B = H.rand_matrix(p, *((k, n) if tB == "N" else (n, k)), seed=2)
            C1 = H.rand_matrix(p, m, n, seed=3, ld=m + 2)
            C2 = C1.copy(order="F")
            ORA.gemm(p, tA, tB, m, n, k, alpha, A, B, beta, C1)
            REF.gemm(p, tA, tB, m, n, k, alpha, A, B, beta, C2)
            assert np.abs(C1 - C2).max() <= (2 * k + 3) * 4 * H.eps(p), (tA, tB)


@needs_ref
@pytest.mark.parametrize("p", "sdcz")
def test_rankk_vs_ref(p):
    REF = CpuLib("ref")
    n, k = 45, 33
    for uplo in "LU":
        for trans in ("N", "C" if PREC[p]["cplx"] else "T"):
            A = H.rand_matrix(p, *((n, k) if trans == "N" else (k, n)), seed=4)
            C1 = H.spd_matrix(p, n, seed=5)
            C2 = C1.copy(order="F")
            ORA.rankk(p, uplo, trans, n, k, -0.5, A, 1.25, C1)
            REF.rankk(p, uplo, trans, n, k, -0.5, A, 1.25, C2)
            assert np.abs(C1 - C2).max() <= (2 * k + 3) * 4 * H.eps(p) * np.abs(C2).max()


@needs_ref
@pytest.mark.parametrize("p", "sdcz")
def test_trsm_trmm_vs_ref(p):
    REF = CpuLib("ref")
    m, n = 41, 23
    cplx = PREC[p]["cplx"]
    alpha = (0.9 + 0.3j) if cplx else 0.9
    for side in "LR":
        na = m if side == "L" else n
        T0 = H.spd_matrix(p, na, seed=6)
        ORA.potrf(p, "L", T0)
        T0 = (np.tril(T0) + np.tril(T0, -1).T).copy(order="F")  # both triangles well conditioned
        for uplo in "LU":
            for trans in TRANS[p]:
                for diag in "NU":
                    Tm = T0.copy(order="F")
                    if diag == "U":
                        Tm /= 8 * np.abs(Tm).max()
                    B1 = H.rand_matrix(p, m, n, seed=7, ld=m + 1)
                    B2 = B1.copy(order="F")
                    ORA.trsm(p, side, uplo, trans, diag, m, n, alpha, Tm, B1)
                    REF.trsm(p, side, uplo, trans, diag, m, n, alpha, Tm, B2)
                    assert np.abs(B1 - B2).max() <= 64 * na * H.eps(p) * np.abs(B2).max(), (side, uplo, trans, diag)
                    B1 = H.rand_matrix(p, m, n, seed=8)
                    B2 = B1.copy(order="F")
                    X1 = H.fmat(m, n, PREC[p]["np"])
                    X2 = H.fmat(m, n, PREC[p]["np"])
                    ORA.trmm2(p, side, uplo, trans, diag, m, n, alpha, Tm, B1, X1)
                    REF.trmm2(p, side, uplo, trans, diag, m, n, alpha, Tm, B2, X2)
                    assert np.abs(X1 - X2).max() <= 16 * na * H.eps(p) * max(1, np.abs(X2).max()), (side, uplo, trans, diag)
                    ORA.trmm(p, side, uplo, trans, diag, m, n, alpha, Tm, B1)
                    REF.trmm(p, side, uplo, trans, diag, m, n, alpha, Tm, B2)
                    assert np.abs(B1 - B2).max() <= 16 * na * H.eps(p) * max(1, np.abs(B2).max()), (side, uplo, trans, diag)


# ---- the reference's own test inputs: latmc(n, c = 2) after srand(0) (test/lapack/dpotrf.c:56) ----
@pytest.mark.parametrize("p", "sdcz")
@pytest.mark.parametrize("n", [2, 64])
def test_latmc_restatement_matches_reference_generator(p, n):
    """tests/harness.py::latmc against matrices dumped from the reference's {s,d,c,z}latmc.c
    (oracle/pin_latmc.c); a few ulps of slack for the compiler's FMA contraction."""
    g = np.load(os.path.join(H.GOLDEN, "latmc.npz"))
    want = g["%s_%d_A" % (p, n)]
    got = H.latmc(p, n)
    assert got.dtype == want.dtype
    assert np.abs(got - want).max() <= 4 * H.eps(p) * 2.0
    w = np.linalg.eigvalsh(want.astype(np.complex128 if PREC[p]["cplx"] else np.float64))
    assert abs(w.min() - 1.0) <= 1e-5 and abs(w.max() - 2.0) <= 1e-5  # condition number 2, as asked


@pytest.mark.parametrize("p", "sdcz")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("n", [2, 64])
def test_oracle_potrf_on_reference_inputs_golden(p, uplo, n):
    """the oracle's potrf of the reference's latmc input against the reference's CPU potrf of it"""
    g = np.load(os.path.join(H.GOLDEN, "latmc.npz"))
    A = H.fmat(n, n, PREC[p]["np"], n + 2)
    A[:, :] = g["%s_%d_A" % (p, n)]
    assert ORA.potrf(p, uplo, A) == 0
    want = g["%s_%d_%s" % (p, n, uplo)]
    assert np.abs(H.tri(A, uplo) - want).max() <= tol(p, n) * np.abs(want).max()
