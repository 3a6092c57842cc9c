"""cuda-cholesky_b200 — Python-side binding of libcholesky_b200.so (the C-ABI drop-in boundary).

The product is the shared library built from csrc/*.cu (hand-written sm_100a kernels) and
host/*.c (C drivers) that exports the reference's cu*/cuMultiGPU* API (include/lapack.h,
include/blas.h, include/cumultigpu.h).  This module only binds those C entry points with ctypes so
that tests, bench.py and smoke() can call them with device pointers (e.g. torch tensors'
data_ptr()).  There is NO Python or CPU fallback: if the library is missing, loading fails loudly.

The directory name has a dash (the repo layout asks for `cuda-cholesky_b200/`), so import it with
`cholesky_b200_loader.load()` (repo root) which registers it as module `cuda_cholesky_b200`.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcholesky_b200.so")

CUDA_SUCCESS = 0
CUDA_ERROR_INVALID_VALUE = 1


class cfloat(C.Structure):  # ABI-compatible with C99 float complex by value
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


class cdouble(C.Structure):  # ABI-compatible with C99 double complex by value
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


class GemmDesc(C.Structure):  # csrc/chol_kernels.h chol_gemm_desc
    _fields_ = [("prec", C.c_int), ("opA", C.c_int), ("opB", C.c_int), ("triA", C.c_int), ("triB", C.c_int),
                ("unitA", C.c_int), ("unitB", C.c_int), ("cuplo", C.c_int), ("herm", C.c_int),
                ("m", C.c_size_t), ("n", C.c_size_t), ("k", C.c_size_t), ("alpha", C.c_double * 2),
                ("beta", C.c_double * 2), ("A", C.c_void_p), ("lda", C.c_size_t), ("B", C.c_void_p),
                ("ldb", C.c_size_t), ("C", C.c_void_p), ("ldc", C.c_size_t), ("D", C.c_void_p), ("ldd", C.c_size_t),
                ("stream", C.c_void_p), ("force_generic", C.c_int),
                ("cyc_nb", C.c_int), ("cyc_p", C.c_int), ("cyc_off", C.c_int)]


_SCALAR = {"s": C.c_float, "d": C.c_double, "c": cfloat, "z": cdouble}
_REAL = {"s": C.c_float, "d": C.c_double, "c": C.c_float, "z": C.c_double}
_U = {"s": "S", "d": "D", "c": "C", "z": "Z"}

_lib = None


class CholeskyError(RuntimeError):
    def __init__(self, fn, code):
        super().__init__("%s returned CUresult %d" % (fn, code))
        self.code = code


def lib():
    """The product library.  Raises if it has not been built - there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libcholesky_b200.so is missing (%s): run `python cuda-cholesky_b200/build.py` "
                               "or __graft_entry__.build(); there is no CPU/PyTorch fallback" % LIB_PATH)
        # the driver library: system copy on a GPU box; the toolkit's compat copy lets the library
        # be loaded (symbols checked, no compute) on a build box without a GPU
        for cand in ("libcuda.so.1", "/usr/local/cuda/compat/libcuda.so.1"):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                break
            except OSError:
                continue
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.chol_diag_max.restype = C.c_int
        _lib.chol_gemm_uses_tma.restype = C.c_int
    return _lib


def _scalar(p, v):
    if p in "cz":
        v = complex(v)
        return _SCALAR[p](v.real, v.imag)
    return _SCALAR[p](v)


def _check(name, rc, allow_invalid=False):
    if rc != CUDA_SUCCESS and not (allow_invalid and rc == CUDA_ERROR_INVALID_VALUE):
        raise CholeskyError(name, rc)
    return rc


def set_xerbla_silent():
    """Replace the default stderr xerbla hook with NULL (tests of argument errors)."""
    C.c_void_p.in_dll(lib(), "xerbla").value = None


class BlasHandle:
    """CUBLAShandle (include/blas.h).  All calls are asynchronous on `stream` (an int CUstream)."""

    def __init__(self):
        self.h = C.c_void_p()
        _check("cuBLASCreate", lib().cuBLASCreate(C.byref(self.h)))

    def close(self):
        if self.h:
            lib().cuBLASDestroy(self.h)
            self.h = C.c_void_p()

    def gemm2(self, p, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, Cp, ldc, D, ldd, stream=0, allow_invalid=False):
        f = getattr(lib(), "cu%sgemm2" % _U[p])
        return _check(f.__name__, f(self.h, C.c_int(ord(tA)), C.c_int(ord(tB)), C.c_size_t(m), C.c_size_t(n), C.c_size_t(k),
                                    _scalar(p, alpha), C.c_void_p(A), C.c_size_t(lda), C.c_void_p(B), C.c_size_t(ldb),
                                    _scalar(p, beta), C.c_void_p(Cp), C.c_size_t(ldc), C.c_void_p(D), C.c_size_t(ldd),
                                    C.c_void_p(stream)), allow_invalid)

    def rankk(self, p, uplo, trans, n, k, alpha, A, lda, beta, Cp, ldc, stream=0, allow_invalid=False):
        f = getattr(lib(), "cu%s%s" % (_U[p], "herk" if p in "cz" else "syrk"))
        return _check(f.__name__, f(self.h, C.c_int(ord(uplo)), C.c_int(ord(trans)), C.c_size_t(n), C.c_size_t(k),
                                    _REAL[p](alpha), C.c_void_p(A), C.c_size_t(lda), _REAL[p](beta), C.c_void_p(Cp),
                                    C.c_size_t(ldc), C.c_void_p(stream)), allow_invalid)

    def trmm2(self, p, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, X, ldx, stream=0, allow_invalid=False):
        f = getattr(lib(), "cu%strmm2" % _U[p])
        return _check(f.__name__, f(self.h, C.c_int(ord(side)), C.c_int(ord(uplo)), C.c_int(ord(trans)), C.c_int(ord(diag)),
                                    C.c_size_t(m), C.c_size_t(n), _scalar(p, alpha), C.c_void_p(A), C.c_size_t(lda),
                                    C.c_void_p(B), C.c_size_t(ldb), C.c_void_p(X), C.c_size_t(ldx), C.c_void_p(stream)),
                      allow_invalid)

    def trmm(self, p, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, stream=0, allow_invalid=False):
        f = getattr(lib(), "cu%strmm" % _U[p])
        return _check(f.__name__, f(self.h, C.c_int(ord(side)), C.c_int(ord(uplo)), C.c_int(ord(trans)), C.c_int(ord(diag)),
                                    C.c_size_t(m), C.c_size_t(n), _scalar(p, alpha), C.c_void_p(A), C.c_size_t(lda),
                                    C.c_void_p(B), C.c_size_t(ldb), C.c_void_p(stream)), allow_invalid)

    def trsm(self, p, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, stream=0, allow_invalid=False):
        f = getattr(lib(), "cu%strsm" % _U[p])
        return _check(f.__name__, f(self.h, C.c_int(ord(side)), C.c_int(ord(uplo)), C.c_int(ord(trans)), C.c_int(ord(diag)),
                                    C.c_size_t(m), C.c_size_t(n), _scalar(p, alpha), C.c_void_p(A), C.c_size_t(lda),
                                    C.c_void_p(B), C.c_size_t(ldb), C.c_void_p(stream)), allow_invalid)


class LapackHandle:
    """CULAPACKhandle (include/lapack.h).  potrf/trtri/lauum/potri return `info` and are complete
    (device idle) on return; logdet returns the value after synchronising its stream."""

    def __init__(self, nb=0):
        self.h = C.c_void_p()
        _check("cuLAPACKCreate", lib().cuLAPACKCreate(C.byref(self.h)))
        if nb:
            self.set_block_size(nb)

    def set_block_size(self, nb):
        _check("cuLAPACKSetBlockSize", lib().cuLAPACKSetBlockSize(self.h, C.c_size_t(nb)))

    def profile(self, enable=True):
        """Time every trailing-update launch of subsequent calls with CUDA events (roofline)."""
        _check("cuLAPACKProfile", lib().cuLAPACKProfile(self.h, C.c_int(1 if enable else 0)))

    def profile_read(self):
        ms, fl, n = C.c_double(0), C.c_double(0), C.c_int(0)
        _check("cuLAPACKProfileRead", lib().cuLAPACKProfileRead(self.h, C.byref(ms), C.byref(fl), C.byref(n)))
        return ms.value, fl.value, n.value

    def profile_history(self, cap=4096):
        ms, fl = (C.c_float * cap)(), (C.c_double * cap)()
        n = lib().cuLAPACKProfileHistory(self.h, ms, fl, C.c_int(cap))
        n = min(n, cap)
        return [(ms[i], fl[i]) for i in range(n)]

    def close(self):
        if self.h:
            lib().cuLAPACKDestroy(self.h)
            self.h = C.c_void_p()

    def _simple(self, name, p, uplo, n, A, lda, allow_invalid):
        f = getattr(lib(), "cu%s%s" % (_U[p], name))
        info = C.c_long(0)
        rc = _check(f.__name__, f(self.h, C.c_int(ord(uplo)), C.c_size_t(n), C.c_void_p(A), C.c_size_t(lda), C.byref(info)),
                    allow_invalid)
        return (info.value, rc) if allow_invalid else info.value

    def potrf(self, p, uplo, n, A, lda, allow_invalid=False):
        return self._simple("potrf", p, uplo, n, A, lda, allow_invalid)

    def lauum(self, p, uplo, n, A, lda, allow_invalid=False):
        return self._simple("lauum", p, uplo, n, A, lda, allow_invalid)

    def potri(self, p, uplo, n, A, lda, allow_invalid=False):
        return self._simple("potri", p, uplo, n, A, lda, allow_invalid)

    def trtri(self, p, uplo, diag, n, A, lda, allow_invalid=False):
        f = getattr(lib(), "cu%strtri" % _U[p])
        info = C.c_long(0)
        rc = _check(f.__name__, f(self.h, C.c_int(ord(uplo)), C.c_int(ord(diag)), C.c_size_t(n), C.c_void_p(A),
                                  C.c_size_t(lda), C.byref(info)), allow_invalid)
        return (info.value, rc) if allow_invalid else info.value

    def potrs(self, p, uplo, n, nrhs, A, lda, B, ldb, allow_invalid=False):
        """Solve A X = B in place from the factor left by potrf (extension, include/cholesky_b200.h)."""
        f = getattr(lib(), "cu%spotrs" % _U[p])
        info = C.c_long(0)
        rc = _check(f.__name__, f(self.h, C.c_int(ord(uplo)), C.c_size_t(n), C.c_size_t(nrhs), C.c_void_p(A), C.c_size_t(lda),
                                  C.c_void_p(B), C.c_size_t(ldb), C.byref(info)), allow_invalid)
        return (info.value, rc) if allow_invalid else info.value

    def potrf_logdet(self):
        """log det A accumulated by the diagonal-block kernels during the last potrf on this handle (extension)."""
        v = C.c_double(0)
        _check("cuLAPACKPotrfLogdet", lib().cuLAPACKPotrfLogdet(self.h, C.byref(v)))
        return v.value

    def trti22(self, p, uplo, diag, n, A, lda, B, ldb, dinfo, stream=0):
        f = getattr(lib(), "cu%strti22" % _U[p])
        return _check(f.__name__, f(self.h, C.c_int(ord(uplo)), C.c_int(ord(diag)), C.c_size_t(n), C.c_void_p(A),
                                    C.c_size_t(lda), C.c_void_p(B), C.c_size_t(ldb), C.c_void_p(dinfo), C.c_void_p(stream)))

    def logdet(self, p, x, incx, n, stream=0):
        f = getattr(lib(), "cu%slogdet" % _U[p])
        res = _REAL[p](0)
        _check(f.__name__, f(self.h, C.c_void_p(x), C.c_size_t(incx), C.c_size_t(n), C.byref(res), C.c_void_p(stream)))
        _check("cuStreamSynchronize", _driver().cuStreamSynchronize(C.c_void_p(stream)))
        return res.value


_drv = None


def _driver():
    global _drv
    if _drv is None:
        try:
            _drv = C.CDLL("libcuda.so.1")
        except OSError:
            _drv = C.CDLL("/usr/local/cuda/compat/libcuda.so.1")
    return _drv


def gemm_desc_launch(**kw):
    """Direct access to the tile engine (csrc/chol_kernels.h chol_gemm_launch) for kernel tests."""
    d = GemmDesc()
    for k, v in kw.items():
        if k in ("alpha", "beta"):
            v = complex(v)
            getattr(d, k)[0], getattr(d, k)[1] = v.real, v.imag
        elif k in ("prec", "opA", "opB", "triA", "triB", "cuplo") and isinstance(v, str):
            setattr(d, k, ord(v))
        else:
            setattr(d, k, v)
    uses_tma = lib().chol_gemm_uses_tma(C.byref(d))
    rc = lib().chol_gemm_launch(C.byref(d))
    if rc != 0:
        raise CholeskyError("chol_gemm_launch", rc)
    return uses_tma


def launch_count(reset=False):
    """Kernels launched by the library so far."""
    f = lib().chol_launch_count
    f.restype = C.c_ulonglong
    return int(f(C.c_int(1 if reset else 0)))


def probe_fp64(kind, iters=20000):
    """Measured FP64 peak of this GPU in TFLOP/s: kind 0 = DMMA (tensor core), 1 = DFMA, 2 = both."""
    tf, ms = C.c_double(0), C.c_float(0)
    rc = lib().chol_probe_fp64(C.c_int(kind), C.c_int(iters), C.byref(tf), C.byref(ms))
    if rc != 0:
        raise CholeskyError("chol_probe_fp64", rc)
    return tf.value, ms.value


# ---------------------------------------------------------------------------------------------
# device random numbers: CUrng64 (include/rng.h; reference include/rng.h:326-413)
# ---------------------------------------------------------------------------------------------
RNG_KINDS = {"raw": "cuRng64Get", "close_open": "cuRng64GetCloseOpen", "open_close": "cuRng64GetOpenClose",
             "open_open": "cuRng64GetOpenOpen", "close_close": "cuRng64GetCloseClose"}


class Rng64:
    """CUrng64 of type cu_dsfmt_19937_t: dSFMT-19937 on the device, the same stream as the reference's CPU generator.
    `x` is a device pointer (int) to n 8-byte values; calls are asynchronous on the generator's stream (default 0)."""

    def __init__(self, seed=None):
        self.h = C.c_void_p()
        t = C.c_void_p.in_dll(lib(), "cu_dsfmt_19937_t")
        _check("cuRng64Create", lib().cuRng64Create(C.byref(self.h), t))
        if seed is not None:
            self.set(seed)

    def close(self):
        if self.h:
            _check("cuRng64Destroy", lib().cuRng64Destroy(self.h))
            self.h = C.c_void_p()

    def set(self, seed):
        _check("cuRng64Set", lib().cuRng64Set(self.h, C.c_uint64(seed)))

    def set_stream(self, stream):
        _check("cuRng64SetStream", lib().cuRng64SetStream(self.h, C.c_void_p(stream)))

    def get(self, x, n, kind="close_open", allow_invalid=False):
        f = getattr(lib(), RNG_KINDS[kind])
        return _check(f.__name__, f(self.h, C.c_void_p(x), C.c_size_t(n)), allow_invalid)

    def discard(self, n):
        _check("cuRng64Discard", lib().cuRng64Discard(self.h, C.c_size_t(n)))

    @property
    def position(self):
        f = lib().cuRng64Position
        f.restype = C.c_ulonglong
        return int(f(self.h))

    def set_segment_bias(self, bias):
        _check("cuRng64SetSegmentBias", lib().cuRng64SetSegmentBias(self.h, C.c_int(bias)))

    def profile(self, enable=True):
        _check("cuRng64Profile", lib().cuRng64Profile(self.h, C.c_int(1 if enable else 0)))

    def profile_read(self):
        """(jump ms, generate ms, segments, refills per segment) of the last get since profile(True)"""
        j, g, s, l = C.c_float(0), C.c_float(0), C.c_int(0), C.c_int(0)
        _check("cuRng64ProfileRead", lib().cuRng64ProfileRead(self.h, C.byref(j), C.byref(g), C.byref(s), C.byref(l)))
        return j.value, g.value, s.value, l.value


# ---------------------------------------------------------------------------------------------
# multi-GPU: the reference's host-pointer API (one process, one worker thread per GPU) ...
# ---------------------------------------------------------------------------------------------
class MultiGPU:
    """CUmultiGPU + CUmultiGPULAPACKhandle + CUmultiGPUBLAShandle (include/cumultigpu.h, lapack.h, blas.h).
    Matrices are HOST arrays (numpy, column-major); pass their ctypes address."""

    def __init__(self, devices):
        self.m = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        _check("cuMultiGPUCreate", lib().cuMultiGPUCreate(C.byref(self.m), arr, C.c_int(len(devices))))
        self.lap = C.c_void_p()
        _check("cuMultiGPULAPACKCreate", lib().cuMultiGPULAPACKCreate(C.byref(self.lap), self.m))
        self.blas = C.c_void_p()
        _check("cuMultiGPUBLASCreate", lib().cuMultiGPUBLASCreate(C.byref(self.blas), self.m))

    def count(self):
        return lib().cuMultiGPUGetContextCount(self.m)

    def close(self):
        if self.m:
            lib().cuMultiGPUBLASDestroy(self.blas)
            lib().cuMultiGPULAPACKDestroy(self.lap)
            lib().cuMultiGPUDestroy(self.m)
            self.m = C.c_void_p()

    def set_block_size(self, nb):
        lib().chol_mgpu_rank_lapack.restype = C.c_void_p
        for i in range(self.count()):
            h = lib().chol_mgpu_rank_lapack(self.m, C.c_int(i))
            _check("cuLAPACKSetBlockSize", lib().cuLAPACKSetBlockSize(C.c_void_p(h), C.c_size_t(nb)))

    def _lapack(self, name, p, uplo, n, A, lda, diag=None, allow_invalid=False):
        f = getattr(lib(), "cuMultiGPU%s%s" % (_U[p], name))
        info = C.c_long(0)
        if diag is None:
            rc = f(self.lap, C.c_int(ord(uplo)), C.c_size_t(n), C.c_void_p(A), C.c_size_t(lda), C.byref(info))
        else:
            rc = f(self.lap, C.c_int(ord(uplo)), C.c_int(ord(diag)), C.c_size_t(n), C.c_void_p(A), C.c_size_t(lda), C.byref(info))
        _check(f.__name__, rc, allow_invalid)
        return (info.value, rc) if allow_invalid else info.value

    def potrf(self, p, uplo, n, A, lda, **kw):
        return self._lapack("potrf", p, uplo, n, A, lda, **kw)

    def lauum(self, p, uplo, n, A, lda, **kw):
        return self._lapack("lauum", p, uplo, n, A, lda, **kw)

    def potri(self, p, uplo, n, A, lda, **kw):
        return self._lapack("potri", p, uplo, n, A, lda, **kw)

    def trtri(self, p, uplo, diag, n, A, lda, **kw):
        return self._lapack("trtri", p, uplo, n, A, lda, diag=diag, **kw)

    def gemm(self, p, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, Cp, ldc):
        f = getattr(lib(), "cuMultiGPU%sgemm" % _U[p])
        return _check(f.__name__, f(self.blas, C.c_int(ord(tA)), C.c_int(ord(tB)), C.c_size_t(m), C.c_size_t(n), C.c_size_t(k),
                                    _scalar(p, alpha), C.c_void_p(A), C.c_size_t(lda), C.c_void_p(B), C.c_size_t(ldb),
                                    _scalar(p, beta), C.c_void_p(Cp), C.c_size_t(ldc)))

    def rankk(self, p, uplo, trans, n, k, alpha, A, lda, beta, Cp, ldc):
        f = getattr(lib(), "cuMultiGPU%s%s" % (_U[p], "herk" if p in "cz" else "syrk"))
        return _check(f.__name__, f(self.blas, C.c_int(ord(uplo)), C.c_int(ord(trans)), C.c_size_t(n), C.c_size_t(k),
                                    _REAL[p](alpha), C.c_void_p(A), C.c_size_t(lda), _REAL[p](beta), C.c_void_p(Cp), C.c_size_t(ldc)))

    def _tr(self, name, p, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb):
        f = getattr(lib(), "cuMultiGPU%s%s" % (_U[p], name))
        return _check(f.__name__, f(self.blas, C.c_int(ord(side)), C.c_int(ord(uplo)), C.c_int(ord(trans)), C.c_int(ord(diag)),
                                    C.c_size_t(m), C.c_size_t(n), _scalar(p, alpha), C.c_void_p(A), C.c_size_t(lda),
                                    C.c_void_p(B), C.c_size_t(ldb)))

    def trmm(self, *a):
        return self._tr("trmm", *a)

    def trsm(self, *a):
        return self._tr("trsm", *a)


# ---------------------------------------------------------------------------------------------
# ... and one process per GPU (torchrun): the SPMD core behind cuMultiGPU*potrf, C-ABI cholDist*
# ---------------------------------------------------------------------------------------------
def dist_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, the launcher ships it to the other ranks)."""
    buf = C.create_string_buffer(128)
    _check("cholDistGetUniqueId", lib().cholDistGetUniqueId(buf))
    return buf.raw


def dist_local_cols(n, nb, rank, nranks):
    f = lib().cholDistLocalCols
    f.restype = C.c_size_t
    return int(f(C.c_size_t(n), C.c_size_t(nb), C.c_int(rank), C.c_int(nranks)))


class Dist:
    """One rank of the distributed factorisation.  The rank's block columns (global block
    c = lc*nranks + rank) are packed side by side in device memory, n rows, leading dimension ldl."""

    def __init__(self, rank, nranks, unique_id=None, nb=0):
        self.h = C.c_void_p()
        idbuf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        _check("cholDistCreate", lib().cholDistCreate(C.byref(self.h), C.c_int(rank), C.c_int(nranks), idbuf))
        self.rank, self.nranks = rank, nranks
        if nb:
            _check("cholDistSetBlockSize", lib().cholDistSetBlockSize(self.h, C.c_size_t(nb)))
        lib().cholDistLapackHandle.restype = C.c_void_p
        self.lapack = LapackHandle.__new__(LapackHandle)
        self.lapack.h = C.c_void_p(lib().cholDistLapackHandle(self.h))

    def block_size(self, p, n):
        f = lib().cholDistBlockSize
        f.restype = C.c_size_t
        return int(f(self.h, C.c_int(ord(p)), C.c_size_t(n)))

    def potrf(self, p, n, Aloc, ldl):
        f = getattr(lib(), "cholDist%spotrf" % _U[p])
        info = C.c_long(0)
        _check(f.__name__, f(self.h, C.c_size_t(n), C.c_void_p(Aloc), C.c_size_t(ldl), C.byref(info)))
        return info.value

    def close(self):
        if self.h:
            lib().cholDistDestroy(self.h)
            self.h = C.c_void_p()


def tile_schedule(cuplo, m, n, cyc_nb=0, cyc_p=0, cyc_off=0):
    """Host-side decode of the d-engine's tile schedule (test hook).  Returns (grid, [(bi, bj), ...])."""
    cap = ((m + 127) // 128 + 8) * ((n + 63) // 64 + 8)
    out = (C.c_int * (2 * cap))()
    nv = C.c_longlong(0)
    f = lib().chol_test_tile_schedule
    f.restype = C.c_longlong
    grid = f(C.c_int(ord(cuplo) if cuplo else 0), C.c_int(m), C.c_int(n), C.c_int(cyc_nb), C.c_int(cyc_p), C.c_int(cyc_off),
             out, C.c_longlong(cap), C.byref(nv))
    return int(grid), [(out[2 * i], out[2 * i + 1]) for i in range(nv.value)]
