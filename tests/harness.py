"""Test harness: ctypes bindings for the CPU oracle (oracle/libchol_oracle.so), the compiled
reference (oracle/_ref/libref_oracle.so, when present) and seeded input builders.

TEST INFRASTRUCTURE ONLY.  Nothing here is imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libchol_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref_oracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")

SEED = 20121018  # fixed harness seed


class cfloat(C.Structure):  # ABI-compatible with C99 `float complex` passed by value
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


class cdouble(C.Structure):  # ABI-compatible with C99 `double complex`
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


PREC = {
    "s": dict(np=np.float32, real=np.float32, scalar=C.c_float, rscalar=C.c_float, cplx=False),
    "d": dict(np=np.float64, real=np.float64, scalar=C.c_double, rscalar=C.c_double, cplx=False),
    "c": dict(np=np.complex64, real=np.float32, scalar=cfloat, rscalar=C.c_float, cplx=True),
    "z": dict(np=np.complex128, real=np.float64, scalar=cdouble, rscalar=C.c_double, cplx=True),
}


def to_scalar(p, v):
    info = PREC[p]
    if info["cplx"]:
        v = complex(v)
        return info["scalar"](v.real, v.imag)
    return info["scalar"](v)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def ensure_oracle_built():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        ensure_oracle_built()
        _oracle = C.CDLL(ORACLE_SO, mode=C.RTLD_LOCAL)
    return _oracle


def ref():
    """The reference's own CPU code (compiled unmodified), or None if it was never built."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            return None
        for cand in ("libcuda.so.1", "/usr/local/cuda/compat/libcuda.so.1"):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                break
            except OSError:
                continue
        _ref = C.CDLL(REF_SO, mode=C.RTLD_LOCAL)
    return _ref


class CpuLib:
    """Uniform numpy-level view of either CPU library.  kind='oracle' -> oracle_<p><name>,
    kind='ref' -> the reference's un-prefixed CPU names (include/lapack.h, include/blas.h)."""

    def __init__(self, kind):
        self.kind = kind
        self.lib = oracle() if kind == "oracle" else ref()
        if self.lib is None:
            raise RuntimeError("reference library not built")

    def _f(self, p, name, restype=None):
        f = getattr(self.lib, ("oracle_" if self.kind == "oracle" else "") + p + name)
        f.restype = restype
        return f

    # LAPACK level -------------------------------------------------------------------------
    def potrf(self, p, uplo, A):
        n, lda = A.shape[1], A.strides[1] // A.itemsize
        info = C.c_long(0)
        self._f(p, "potrf")(C.c_int(ord(uplo)), C.c_size_t(n), _ptr(A), C.c_size_t(lda), C.byref(info))
        return info.value

    def trtri(self, p, uplo, diag, A):
        n, lda = A.shape[1], A.strides[1] // A.itemsize
        info = C.c_long(0)
        self._f(p, "trtri")(C.c_int(ord(uplo)), C.c_int(ord(diag)), C.c_size_t(n), _ptr(A), C.c_size_t(lda), C.byref(info))
        return info.value

    def trtri2(self, p, uplo, diag, A, B):
        n, lda, ldb = A.shape[1], A.strides[1] // A.itemsize, B.strides[1] // B.itemsize
        info = C.c_long(0)
        self._f(p, "trtri2")(C.c_int(ord(uplo)), C.c_int(ord(diag)), C.c_size_t(n), _ptr(A), C.c_size_t(lda),
                             _ptr(B), C.c_size_t(ldb), C.byref(info))
        return info.value

    def lauum(self, p, uplo, A):
        n, lda = A.shape[1], A.strides[1] // A.itemsize
        info = C.c_long(0)
        self._f(p, "lauum")(C.c_int(ord(uplo)), C.c_size_t(n), _ptr(A), C.c_size_t(lda), C.byref(info))
        return info.value

    def potri(self, p, uplo, A):
        n, lda = A.shape[1], A.strides[1] // A.itemsize
        info = C.c_long(0)
        self._f(p, "potri")(C.c_int(ord(uplo)), C.c_size_t(n), _ptr(A), C.c_size_t(lda), C.byref(info))
        return info.value

    def logdet(self, p, x, incx, n):
        rt = C.c_float if PREC[p]["real"] == np.float32 else C.c_double
        return self._f(p, "logdet", rt)(_ptr(x), C.c_size_t(incx), C.c_size_t(n))

    # BLAS level ---------------------------------------------------------------------------
    def gemm(self, p, tA, tB, m, n, k, alpha, A, B, beta, Cm):
        self._f(p, "gemm", C.c_int if self.kind == "oracle" else None)(
            C.c_int(ord(tA)), C.c_int(ord(tB)), C.c_size_t(m), C.c_size_t(n), C.c_size_t(k), to_scalar(p, alpha),
            _ptr(A), C.c_size_t(A.strides[1] // A.itemsize), _ptr(B), C.c_size_t(B.strides[1] // B.itemsize),
            to_scalar(p, beta), _ptr(Cm), C.c_size_t(Cm.strides[1] // Cm.itemsize))

    def rankk(self, p, uplo, trans, n, k, alpha, A, beta, Cm):
        name = "herk" if PREC[p]["cplx"] else "syrk"
        rs = PREC[p]["rscalar"]
        self._f(p, name)(C.c_int(ord(uplo)), C.c_int(ord(trans)), C.c_size_t(n), C.c_size_t(k), rs(alpha), _ptr(A),
                         C.c_size_t(A.strides[1] // A.itemsize), rs(beta), _ptr(Cm),
                         C.c_size_t(Cm.strides[1] // Cm.itemsize))

    def trsm(self, p, side, uplo, trans, diag, m, n, alpha, A, B):
        self._f(p, "trsm")(C.c_int(ord(side)), C.c_int(ord(uplo)), C.c_int(ord(trans)), C.c_int(ord(diag)),
                           C.c_size_t(m), C.c_size_t(n), to_scalar(p, alpha), _ptr(A),
                           C.c_size_t(A.strides[1] // A.itemsize), _ptr(B), C.c_size_t(B.strides[1] // B.itemsize))

    def trmm(self, p, side, uplo, trans, diag, m, n, alpha, A, B):
        self._f(p, "trmm")(C.c_int(ord(side)), C.c_int(ord(uplo)), C.c_int(ord(trans)), C.c_int(ord(diag)),
                           C.c_size_t(m), C.c_size_t(n), to_scalar(p, alpha), _ptr(A),
                           C.c_size_t(A.strides[1] // A.itemsize), _ptr(B), C.c_size_t(B.strides[1] // B.itemsize))

    def trmm2(self, p, side, uplo, trans, diag, m, n, alpha, A, B, Xo):
        self._f(p, "trmm2")(C.c_int(ord(side)), C.c_int(ord(uplo)), C.c_int(ord(trans)), C.c_int(ord(diag)),
                            C.c_size_t(m), C.c_size_t(n), to_scalar(p, alpha), _ptr(A),
                            C.c_size_t(A.strides[1] // A.itemsize), _ptr(B), C.c_size_t(B.strides[1] // B.itemsize),
                            _ptr(Xo), C.c_size_t(Xo.strides[1] // Xo.itemsize))


# ---------------------------------------------------------------------------------------------
# seeded inputs
# ---------------------------------------------------------------------------------------------
def dsfmt_uniform(seed, n):
    """n doubles in [0,1) from the dSFMT-19937 restatement (oracle/dsfmt_oracle.c)."""
    lib = oracle()
    lib.oracle_dsfmt_state_size.restype = C.c_size_t
    st = C.create_string_buffer(lib.oracle_dsfmt_state_size())
    lib.oracle_dsfmt_seed(st, C.c_uint64(seed))
    out = np.empty(n, dtype=np.float64)
    lib.oracle_dsfmt_fill_close_open(st, _ptr(out), C.c_size_t(n))
    return out


DSFMT_KINDS = {"raw": 0, "close_open": 1, "open_close": 2, "open_open": 3, "close_close": 4}


class OracleDsfmt:
    """The oracle's sequential dSFMT-19937 (oracle/dsfmt_oracle.c) as a stateful generator: fill(n, kind) continues the
    stream like successive calls of the reference's rng64Get* do (an odd n drops the high half of its last word)."""

    def __init__(self, seed):
        self.lib = oracle()
        self.lib.oracle_dsfmt_state_size.restype = C.c_size_t
        self.st = C.create_string_buffer(self.lib.oracle_dsfmt_state_size())
        self.lib.oracle_dsfmt_seed(self.st, C.c_uint64(seed))

    def fill(self, n, kind="close_open"):
        """n values as raw 64-bit patterns (np.uint64); .view(np.float64) gives the doubles"""
        out = np.empty(n, dtype=np.uint64)
        self.lib.oracle_dsfmt_fill(self.st, _ptr(out), C.c_size_t(n), C.c_int(DSFMT_KINDS[kind]))
        return out


def fmat(m, n, dtype, ld=None):
    """Column-major m x n array with leading dimension ld (>= m)."""
    ld = ld or max(m, 1)
    buf = np.zeros((n, ld), dtype=dtype)
    return buf.T[:m, :]  # shape (m, n), strides (itemsize, ld*itemsize)


def rand_matrix(p, m, n, seed=SEED, ld=None):
    """Entries u - 0.5 (complex: independent real and imaginary parts), column-major."""
    dt = PREC[p]["np"]
    A = fmat(m, n, dt, ld)
    if PREC[p]["cplx"]:
        u = dsfmt_uniform(seed, 2 * m * n) - 0.5
        A[:, :] = (u[0::2] + 1j * u[1::2]).reshape(n, m).T.astype(dt)
    else:
        A[:, :] = (dsfmt_uniform(seed, m * n) - 0.5).reshape(n, m).T.astype(dt)
    return A


def spd_matrix(p, n, seed=SEED, ld=None, k=None):
    """A = B*B^H + n*I with B (n x k) from the seeded generator (BASELINE.json configs[0]);
    both triangles filled, computed in double precision and rounded to the target type."""
    k = k or n
    hi = "z" if PREC[p]["cplx"] else "d"
    B = np.ascontiguousarray(rand_matrix(hi, n, k, seed))
    Ah = B @ B.conj().T + n * np.eye(n)
    A = fmat(n, n, PREC[p]["np"], ld)
    A[:, :] = Ah.astype(PREC[p]["np"])
    return A


_LIBC = None


def latmc(p, n, c=2.0, ld=None, seed=0):
    """The reference's own test-matrix generator (test/lapack/util/{s,d,c,z}latmc.c:11-70): SPD / HPD
    with condition number c - a diagonal with entries in [1, c] (1 and c forced at (0,0), (1,1)),
    then one similarity  A -= u w^H + w u^H  with a random vector u (w = t v - s u, v = A u,
    t = 2 / u^H u, s = t^2 u^H v / 2).  Random numbers are libc rand() after srand(seed), drawn in
    the reference's order; every reference LAPACK test uses c = 2 and srand(0)."""
    global _LIBC
    if n < 2:
        raise ValueError("latmc needs n >= 2 (as the reference)")
    if _LIBC is None:
        _LIBC = C.CDLL("libc.so.6")
    _LIBC.srand(C.c_uint(seed))
    rmax = 2147483647.0
    real, dt, cplx = PREC[p]["real"], PREC[p]["np"], PREC[p]["cplx"]
    c = real(c)
    diag = np.empty(n, dtype=real)
    diag[0], diag[1] = 1.0, c
    for j in range(2, n):
        diag[j] = real(_LIBC.rand() / rmax) * (c - real(1.0)) + real(1.0)
    u = np.empty(n, dtype=dt)
    for j in range(n):
        # complex: `rand()/RAND_MAX + rand()/RAND_MAX * I` leaves the order of the two calls to the
        # compiler; gcc (the reference's compiler, and oracle/pin_latmc.c here) draws the imaginary part first
        r1 = real(_LIBC.rand() / rmax)
        u[j] = real(_LIBC.rand() / rmax) + 1j * r1 if cplx else r1
    v = (diag * u).astype(dt)
    t, sacc = real(0.0), dt(0.0)
    for j in range(n):  # sequential accumulation, as the reference's loop
        t = real(t + (np.conj(u[j]) * u[j]).real)
        sacc = dt(sacc + np.conj(u[j]) * v[j])
    t = real(2.0) / t
    sacc = dt(t * t * sacc / real(2.0))
    w = (t * v - sacc * u).astype(dt)
    A = fmat(n, n, dt, ld)
    A[:, :] = np.diag(diag).astype(dt)
    A[:, :] -= (np.outer(u, np.conj(w)) + np.outer(w, np.conj(u))).astype(dt)
    return A


def tri(A, uplo):
    return np.tril(A) if uplo == "L" else np.triu(A)


def eps(p):
    return float(np.finfo(PREC[p]["real"]).eps)
