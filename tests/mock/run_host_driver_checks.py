"""Run the C host drivers of cuda-cholesky_b200/host on the MOCK back end (tests/mock: a synchronous
CPU stand-in for the CUDA driver, plain-loop implementations of the kernel launcher ABI and a
thread-rendezvous libnccl) and compare with the CPU oracle.  TEST INFRASTRUCTURE ONLY - executed in
its own process by tests/test_host_drivers_mock.py with
    argv[1] = the mock library,  LD_LIBRARY_PATH = the directory holding the mock libnccl.so.2.
The mock streams are real asynchronous queues (one thread each) with CUDA's event semantics, uploads can
be made slow and every stream can be made to stall at random, so a missing dependency between the
panel / update / upload / download streams shows up as a wrong result (checked by mutation: dropping the
update's wait for the panel, or a group update's wait for its upload, fails these checks).
What this covers: blocking, look-ahead bookkeeping, workspace carving, recursion of the diagonal
block, TRSM/TRMM/TRTRI/LAUUM/POTRI/POTRS drivers, argument errors, context push/pop balance,
allocation balance, the CUmultiGPU runtime with several (mock) devices, the distributed POTRF incl.
the page-locked overlapped paths, the slab-distributed inverse, the host-pointer BLAS.  What it
cannot cover: the kernels themselves (tests/test_kernel_emulation.py, and the GPU parity tests) and
timing."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
from harness import PREC  # noqa: E402
from cholesky_b200_loader import load  # noqa: E402

chol = load()
mock = C.CDLL(sys.argv[1], mode=C.RTLD_GLOBAL)
mock.chol_diag_max.restype = C.c_int
mock.chol_gemm_uses_tma.restype = C.c_int
mock.mock_cuda_live_allocs.restype = C.c_long
chol._lib = mock  # the test owns this substitution; the product wrapper knows nothing about a mock
chol._drv = mock  # ... and the wrapper's own driver calls (cuStreamSynchronize) go to the stand-in too
ORA = H.CpuLib("oracle")
chol.set_xerbla_silent()
checks = 0


def ok(cond, what):
    global checks
    checks += 1
    if not cond:
        print("FAILED:", what)
        sys.exit(1)


def tol(p, n):
    return 64 * n * H.eps(p)


def close(got, want, p, n, what):
    scale = max(1.0, float(np.abs(want).max()))
    err = float(np.abs(got - want).max())
    ok(err <= tol(p, n) * scale, "%s: err %.3e > %.3e" % (what, err, tol(p, n) * scale))


def other_untouched(got, orig, uplo, what):
    a = np.triu(got, 1) if uplo == "L" else np.tril(got, -1)
    b = np.triu(orig, 1) if uplo == "L" else np.tril(orig, -1)
    ok(np.array_equal(a, b), what + ": other triangle changed")


def ptr(A):
    return A.ctypes.data


def ld(A):
    return A.strides[1] // A.itemsize


# ------------------------------------------------------------------------------------------------
def single_device_lapack():
    allocs0 = mock.mock_cuda_live_allocs()
    lap = chol.LapackHandle()  # no current context: the handle makes device 0's primary context current, as documented
    depth0 = mock.mock_cuda_ctx_depth()
    for p in "dzsc":
        for uplo in "LU":
            for n, nb in ((1, 0), (5, 0), (130, 0), (300, 0), (700, 0), (700, 128), (900, 512)):
                lap.set_block_size(nb) if hasattr(lap, "set_block_size") else mock.cuLAPACKSetBlockSize(lap.h, C.c_size_t(nb))
                A0 = H.spd_matrix(p, n, ld=n + 3, k=min(n, 48))
                want = A0.copy(order="F")
                ok(ORA.potrf(p, uplo, want) == 0, "oracle potrf")
                A = A0.copy(order="F")
                ok(lap.potrf(p, uplo, n, ptr(A), ld(A)) == 0, "potrf info")
                tag = "%spotrf %s n=%d nb=%d" % (p, uplo, n, nb)
                close(H.tri(A, uplo), H.tri(want, uplo), p, n, tag)
                other_untouched(A, A0, uplo, tag)
                if n in (5, 300, 700) and nb in (0, 128):
                    # logdet of the factor, and the one the diagonal-block launches accumulated while factoring
                    got_ld = lap.logdet(p, ptr(A), ld(A) + 1, n)
                    want_ld = ORA.logdet(p, want, ld(want) + 1, n)
                    ok(abs(got_ld - want_ld) <= 1e-5 * abs(want_ld) + 1e-5, tag + " logdet")
                    fused = lap.potrf_logdet()
                    ok(abs(fused - want_ld) <= 1e-5 * abs(want_ld) + 1e-5, tag + " fused logdet %r vs %r" % (fused, want_ld))
                    # potrs: solve A X = B from the factor
                    nrhs = 7
                    B0 = H.rand_matrix(p, n, nrhs, seed=5, ld=n + 1)
                    Bw = B0.copy(order="F")
                    hT = "C" if PREC[p]["cplx"] else "T"
                    for trans in (("N", hT) if uplo == "L" else (hT, "N")):
                        ORA.trsm(p, "L", uplo, trans, "N", n, nrhs, 1.0, want, Bw)
                    Bg = B0.copy(order="F")
                    ok(lap.potrs(p, uplo, n, nrhs, ptr(A), ld(A), ptr(Bg), ld(Bg)) == 0, tag + " potrs info")
                    close(Bg, Bw, p, n, tag + " potrs")
                    # trtri (both diag), lauum, potri on the factor
                    for diag in "NU":
                        src = A / (np.abs(np.diag(A)).max() * 4) if diag == "U" else A
                        w2 = src.copy(order="F")
                        ok(ORA.trtri(p, uplo, diag, w2) == 0, "oracle trtri")
                        g2 = src.copy(order="F")
                        ok(lap.trtri(p, uplo, diag, n, ptr(g2), ld(g2)) == 0, tag + " trtri info")
                        close(H.tri(g2, uplo), H.tri(w2, uplo), p, n, tag + " trtri " + diag)
                        other_untouched(g2, src, uplo, tag + " trtri " + diag)
                    for name in ("lauum", "potri"):
                        w2 = A.copy(order="F")
                        getattr(ORA, name)(p, uplo, w2)
                        g2 = A.copy(order="F")
                        ok(getattr(lap, name)(p, uplo, n, ptr(g2), ld(g2)) == 0, tag + " " + name + " info")
                        close(H.tri(g2, uplo), H.tri(w2, uplo), p, n, tag + " " + name)
                        other_untouched(g2, A, uplo, tag + " " + name)
    mock.cuLAPACKSetBlockSize(lap.h, C.c_size_t(0))
    # not positive definite: info equals the oracle's, for a failure in the first, a middle and the last block
    for uplo in "LU":
        for bad in (0, 300, 699):
            A0 = H.spd_matrix("d", 700, k=32)
            A0[bad, bad] = -2.0
            want = A0.copy(order="F")
            winfo = ORA.potrf("d", uplo, want)
            A = A0.copy(order="F")
            ok(lap.potrf("d", uplo, 700, ptr(A), 700) == winfo == bad + 1, "not-PD info at %d" % bad)
    # argument errors: info, CUDA_ERROR_INVALID_VALUE, and nothing left on the context stack
    A = H.fmat(4, 4, np.float64)
    ok(lap.potrf("d", "L", 8, ptr(A), 4, allow_invalid=True) == (-4, 1), "potrf lda error")
    ok(lap.trtri("d", "L", "N", 8, ptr(A), 4, allow_invalid=True) == (-5, 1), "trtri lda error")
    ok(lap.potrs("d", "L", 8, 1, ptr(A), 4, ptr(A), 8, allow_invalid=True) == (-5, 1), "potrs lda error")
    ok(lap.potrs("d", "L", 4, 1, ptr(A), 4, ptr(A), 2, allow_invalid=True) == (-7, 1), "potrs ldb error")
    ok(lap.potrf("d", "L", 0, ptr(A), 4) == 0, "n = 0")
    ok(mock.mock_cuda_ctx_depth() == depth0, "context stack depth changed: %d -> %d" % (depth0, mock.mock_cuda_ctx_depth()))
    lap.close()
    ok(mock.mock_cuda_live_allocs() == allocs0, "device allocations leaked: %d" % (mock.mock_cuda_live_allocs() - allocs0))


def single_device_blas():
    blas = chol.BlasHandle()
    depth0 = mock.mock_cuda_ctx_depth()
    for p in "dc":
        cplx = PREC[p]["cplx"]
        hT = "C" if cplx else "T"
        alpha = (0.9 + 0.3j) if cplx else 0.9
        m, n = 300, 170
        for side in "LR":
            na = m if side == "L" else n
            T0 = H.spd_matrix(p, na, seed=6, k=16)
            ORA.potrf(p, "L", T0)
            T0 = (np.tril(T0) + np.tril(T0, -1).conj().T).copy(order="F")  # both triangles hold a usable factor
            for uplo in "LU":
                for trans in ("N", "T", hT) if cplx else ("N", "T"):
                    for diag in "NU":
                        Tm = T0 / (np.abs(np.diag(T0)).max() * 4) if diag == "U" else T0
                        Tm = np.asfortranarray(Tm)
                        B0 = H.rand_matrix(p, m, n, seed=7, ld=m + 1)
                        for name in ("trsm", "trmm"):
                            want = B0.copy(order="F")
                            getattr(ORA, name)(p, side, uplo, trans, diag, m, n, alpha, Tm, want)
                            got = B0.copy(order="F")
                            getattr(blas, name)(p, side, uplo, trans, diag, m, n, alpha, ptr(Tm), ld(Tm), ptr(got), ld(got))
                            close(got, want, p, na, "%s%s %s%s%s%s" % (p, name, side, uplo, trans, diag))
    ok(mock.mock_cuda_ctx_depth() == depth0, "BLAS left contexts on the stack")
    # two TRSMs / TRMMs on DIFFERENT streams of one handle share the handle's scratch: the library orders them
    # (an event), so both must still be right when the streams run side by side
    s1, s2 = C.c_void_p(), C.c_void_p()
    mock.cuStreamCreate(C.byref(s1), C.c_uint(1))
    mock.cuStreamCreate(C.byref(s2), C.c_uint(1))
    m, n = 400, 90
    T0 = H.spd_matrix("d", m, seed=8, k=16)
    ORA.potrf("d", "L", T0)
    for rep in range(6):
        Bs = [H.rand_matrix("d", m, n, seed=20 + 2 * rep + i) for i in range(2)]
        wants = [b.copy(order="F") for b in Bs]
        names = ("trsm", "trmm") if rep % 2 else ("trsm", "trsm")
        for w, nm in zip(wants, names):
            getattr(ORA, nm)("d", "L", "L", "N", "N", m, n, 1.0, T0, w)
        gots = [b.copy(order="F") for b in Bs]
        for g, st, nm in zip(gots, (s1, s2), names):
            getattr(blas, nm)("d", "L", "L", "N", "N", m, n, 1.0, ptr(T0), ld(T0), ptr(g), ld(g), stream=st.value)
        mock.cuStreamSynchronize(s1)
        mock.cuStreamSynchronize(s2)
        for g, w, nm in zip(gots, wants, names):
            close(g, w, "d", m, "two streams, one handle: %s" % nm)
    mock.cuStreamDestroy_v2(s1)
    mock.cuStreamDestroy_v2(s2)
    blas.close()


def multi_device(P):
    os.environ["MOCK_CUDA_DEVICES"] = str(P)
    for k in ("CHOL_MGPU_H2D_OVERLAP", "CHOL_H2D_GBPS", "CHOL_MGPU_INVERSE", "CHOL_MGPU_INVERSE_SLABS", "MOCK_CUDA_MEM_LIMIT"):
        os.environ.pop(k, None)
    allocs0 = mock.mock_cuda_live_allocs()
    mg = chol.MultiGPU(list(range(P)))
    ok(mg.count() == P, "context count")

    def potrf_cases(label, pinned):
        for p in "dz":
            for uplo in "LU":
                for n, nb in ((5, 0), (300, 0), (700, 128), (1000, 0), (650, 128)):
                    mg.set_block_size(nb)
                    A0 = H.spd_matrix(p, n, ld=n + 2, k=min(n, 48))
                    want = A0.copy(order="F")
                    ok(ORA.potrf(p, uplo, want) == 0, "oracle")
                    A = A0.copy(order="F")
                    if pinned:
                        mock.mock_cuda_register_pinned(C.c_void_p(A.ctypes.data - 0), C.c_size_t(A.strides[1] * A.shape[1]))
                    info = mg.potrf(p, uplo, n, ptr(A), ld(A))
                    if pinned:
                        mock.mock_cuda_unregister_pinned(C.c_void_p(A.ctypes.data))
                    tag = "%s cuMultiGPU%spotrf %s n=%d nb=%d P=%d" % (label, p, uplo, n, nb, P)
                    ok(info == 0, tag + " info %d" % info)
                    close(H.tri(A, uplo), H.tri(want, uplo), p, n, tag)
                    other_untouched(A, A0, uplo, tag)
        mg.set_block_size(0)

    potrf_cases("pageable", False)
    potrf_cases("page-locked (download overlapped)", True)
    os.environ["CHOL_MGPU_H2D_OVERLAP"] = "1"
    os.environ["MOCK_CUDA_SLOW_H2D_US"] = "30000"  # uploads lag behind the factorisation: every missing wait shows
    for gbps in ("50", "0.000001"):  # the tiny rate sends EVERY step through the group-by-group update
        os.environ["CHOL_H2D_GBPS"] = gbps
        potrf_cases("page-locked, upload overlapped (assumed %s GB/s)" % gbps, True)
    os.environ.pop("CHOL_MGPU_H2D_OVERLAP")
    os.environ.pop("CHOL_H2D_GBPS")
    os.environ.pop("MOCK_CUDA_SLOW_H2D_US")

    # not positive definite: the first failing column over all ranks
    for uplo in "LU":
        A = H.spd_matrix("d", 900, k=32)
        A[600, 600] = -2.0
        A[100, 100] = -1.0
        ok(mg.potrf("d", uplo, 900, ptr(A), 900) == 101, "multi-GPU not-PD info")

    # host-pointer TRTRI / LAUUM / POTRI: staged on one device, the distributed block-cyclic sweeps (the default with
    # more than one device), then the slab split
    for mode in ("staged", "dist", "slab", "slab x2"):
        os.environ.pop("CHOL_MGPU_INVERSE", None)
        os.environ.pop("CHOL_MGPU_INVERSE_SLABS", None)
        if mode in ("staged", "dist"):
            os.environ["CHOL_MGPU_INVERSE"] = mode
        else:
            os.environ["CHOL_MGPU_INVERSE"] = "slab"
        if mode == "slab x2":
            os.environ["CHOL_MGPU_INVERSE_SLABS"] = "2"
        for p in "dz":
            for uplo in "LU":
                for n in (70, 400, 900) + ((1300,) if mode == "dist" and p == "d" else ()):
                    F = H.spd_matrix(p, n, k=48)
                    ORA.potrf(p, uplo, F)
                    for name, diag in (("trtri", "N"), ("trtri", "U"), ("lauum", None), ("potri", None)):
                        src = F / (np.abs(np.diag(F)).max() * 4) if diag == "U" else F
                        want = src.copy(order="F")
                        got = src.copy(order="F")
                        if name == "trtri":
                            ORA.trtri(p, uplo, diag, want)
                            info = mg.trtri(p, uplo, diag, n, ptr(got), n)
                        else:
                            getattr(ORA, name)(p, uplo, want)
                            info = getattr(mg, name)(p, uplo, n, ptr(got), n)
                        tag = "%s cuMultiGPU%s%s %s %s n=%d P=%d" % (mode, p, name, uplo, diag or "", n, P)
                        ok(info == 0, tag + " info")
                        close(H.tri(got, uplo), H.tri(want, uplo), p, n, tag)
                        other_untouched(got, src, uplo, tag)
        # zero pivots: the reference's sweep direction decides which one is reported
        for uplo in "LU":
            F = H.spd_matrix("d", 600, k=48)
            ORA.potrf("d", uplo, F)
            F[100, 100] = 0.0
            F[450, 450] = 0.0
            want = F.copy(order="F")
            winfo = ORA.trtri("d", uplo, "N", want)
            got = F.copy(order="F")
            ok(mg.trtri("d", uplo, "N", 600, ptr(got), 600) == winfo, mode + " zero-pivot info")
    os.environ.pop("CHOL_MGPU_INVERSE", None)
    os.environ.pop("CHOL_MGPU_INVERSE_SLABS", None)

    # host-pointer BLAS-3
    for p in "dc":
        cplx = PREC[p]["cplx"]
        hT = "C" if cplx else "T"
        m, n, k = 300, 200, 150
        alpha, beta = ((0.7 - 0.2j, -0.3 + 0.5j) if cplx else (0.7, -0.3))
        for tA, tB in (("N", "N"), ("N", hT), (hT, "N")):
            A = H.rand_matrix(p, *((m, k) if tA == "N" else (k, m)), seed=1)
            B = H.rand_matrix(p, *((k, n) if tB == "N" else (n, k)), seed=2)
            Cm = H.rand_matrix(p, m, n, seed=3, ld=m + 2)
            want = Cm.copy(order="F")
            ORA.gemm(p, tA, tB, m, n, k, alpha, A, B, beta, want)
            mg.gemm(p, tA, tB, m, n, k, alpha, ptr(A), ld(A), ptr(B), ld(B), beta, ptr(Cm), ld(Cm))
            close(Cm, want, p, k, "cuMultiGPU gemm %s%s" % (tA, tB))
        for uplo in "LU":
            for trans in ("N", hT):
                A = H.rand_matrix(p, *((n, k) if trans == "N" else (k, n)), seed=4)
                C0 = H.spd_matrix(p, n, seed=5, k=8)
                want = C0.copy(order="F")
                ORA.rankk(p, uplo, trans, n, k, -0.5, A, 1.25, want)
                got = C0.copy(order="F")
                mg.rankk(p, uplo, trans, n, k, -0.5, ptr(A), ld(A), 1.25, ptr(got), n)
                close(got, want, p, k, "cuMultiGPU rank-k %s%s" % (uplo, trans))
        a1 = (0.9 + 0.3j) if cplx else 0.9
        for side in "LR":
            na = m if side == "L" else n
            T0 = H.spd_matrix(p, na, seed=6, k=16)
            ORA.potrf(p, "L", T0)
            T0 = (np.tril(T0) + np.tril(T0, -1).conj().T).copy(order="F")
            for uplo, trans in (("L", "N"), ("U", hT), ("L", hT), ("U", "N")):
                B0 = H.rand_matrix(p, m, n, seed=7)
                for name in ("trsm", "trmm"):
                    want = B0.copy(order="F")
                    getattr(ORA, name)(p, side, uplo, trans, "N", m, n, a1, T0, want)
                    got = B0.copy(order="F")
                    getattr(mg, name)(p, side, uplo, trans, "N", m, n, a1, ptr(T0), na, ptr(got), m)
                    close(got, want, p, na, "cuMultiGPU %s %s%s%s" % (name, side, uplo, trans))

    # the one-process-per-GPU entry points with a single rank (what bench.py times at N = 1): the packed
    # layout is then the plain column-major matrix, resident in "device" memory
    if P == 1:
        for p in "dz":
            for n, nb in ((700, 0), (1000, 128)):
                d = chol.Dist(0, 1, nb=nb)
                A0 = H.spd_matrix(p, n, k=48)
                want = A0.copy(order="F")
                ORA.potrf(p, "L", want)
                A = A0.copy(order="F")
                ok(d.potrf(p, n, ptr(A), ld(A)) == 0, "cholDist info")
                close(np.tril(A), np.tril(want), p, n, "cholDist%spotrf n=%d nb=%d" % (p, n, nb))
                other_untouched(A, A0, "L", "cholDist potrf")
                ok(d.block_size(p, n) == (nb or (128 if p == "z" else 256)), "cholDist block size")
                d.close()

    # an allocation failure on the devices fails the call cleanly (no rank is left in a broadcast)
    os.environ["MOCK_CUDA_MEM_LIMIT"] = str(1 << 20)
    A = H.spd_matrix("d", 2000, k=16)  # larger than anything factored so far: the resident block columns must grow
    try:
        mg.potrf("d", "L", 2000, ptr(A), 2000)
        ok(False, "expected CUDA_ERROR_OUT_OF_MEMORY")
    except chol.CholeskyError as e:
        ok(e.code == 2, "out-of-memory code %d" % e.code)
    os.environ.pop("MOCK_CUDA_MEM_LIMIT")
    A = H.spd_matrix("d", 500, k=16)
    ok(mg.potrf("d", "L", 500, ptr(A), 500) == 0, "usable again after the failed call")
    mg.close()
    ok(mock.mock_cuda_live_allocs() == allocs0, "multi-GPU device allocations leaked: %d" % (mock.mock_cuda_live_allocs() - allocs0))


def dist_ranks(P):
    """cholDist* with P ranks (threads of this process standing in for torchrun's processes): every rank packs its
    block-cyclic share the way bench.py does, all factor together, the pieces are put back and compared."""
    import threading
    sys.path.insert(0, ROOT)
    from bench import owned_columns  # the launcher-side layout helper that bench.py itself uses
    for p in "dz":
        for n, nb in ((900, 128), (1300, 0), (700, 256)):
            uid = chol.dist_unique_id()
            A0 = H.spd_matrix(p, n, k=48)
            want = A0.copy(order="F")
            ORA.potrf(p, "L", want)
            got = A0.copy(order="F")
            infos, errs = [None] * P, []

            def rank_main(r):
                try:
                    d = chol.Dist(r, P, uid, nb=nb)
                    bs = d.block_size(p, n)
                    cols = owned_columns(n, bs, r, P)
                    loc = np.asfortranarray(A0[:, cols]) if len(cols) else H.fmat(n, 1, PREC[p]["np"])
                    infos[r] = d.potrf(p, n, ptr(loc), ld(loc))
                    if len(cols):
                        got[:, cols] = loc
                    d.close()
                except Exception as e:  # noqa: BLE001 - reported below
                    errs.append(repr(e))

            ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(P)]
            for t in ts:
                t.start()
            for t in ts:
                t.join(600)
            ok(not errs and all(not t.is_alive() for t in ts), "cholDist ranks: %s" % errs)
            ok(all(i == 0 for i in infos), "cholDist infos %s" % infos)
            close(np.tril(got), np.tril(want), p, n, "cholDist%spotrf P=%d n=%d nb=%d" % (p, P, n, nb))
            other_untouched(got, A0, "L", "cholDist P=%d" % P)


def device_rng():
    """cuRng64* (host/rng.c + host/rng_poly.c) against the oracle's sequential dSFMT-19937: segmenting on the absolute grid,
    windows that start / end inside a refill, odd counts (half a word dropped), every output form, a stream that continues
    over many calls, re-seeding, Discard, two generators with different seeds (the second one's state forces the process-wide
    polynomial to be extended, after which the first must rebuild its jump table), a user stream, argument errors."""
    allocs0 = mock.mock_cuda_live_allocs()
    mock.mock_cuda_ctx_depth.restype = C.c_int
    stream = C.c_void_p()
    ok(mock.cuInit(0) == 0, "cuInit")
    g = chol.Rng64()
    depth0 = mock.mock_cuda_ctx_depth()
    buf = np.zeros(1, dtype=np.uint64)
    try:
        g.get(ptr(buf), 1)
        ok(False, "Get before Set must fail")
    except chol.CholeskyError as e:
        ok(e.code == 3, "Get before Set: CUDA_ERROR_NOT_INITIALIZED")
    ok(mock.cuStreamCreate(C.byref(stream), 0) == 0, "stream")
    seed = H.SEED
    for use_stream in (False, True):
        g.set_stream(stream.value if use_stream else 0)
        for kind in H.DSFMT_KINDS:
            g.set(seed)
            ora = H.OracleDsfmt(seed)
            # sizes chosen to cross refill (382 values) and segment (16 refills = 6112 values) boundaries, odd ones included
            for n in (1, 2, 3, 381, 382, 383, 1000, 6111, 6113, 20001, 7, 150000, 2, 400001, 5):
                out = np.full(n + 2, 0x5555555555555555, dtype=np.uint64)
                g.get(ptr(out[1:]), n, kind)
                mock.cuStreamSynchronize(stream if use_stream else None)
                want = ora.fill(n, kind)
                ok(np.array_equal(out[1:n + 1], want), "rng %s n=%d stream=%s: stream differs from the oracle" % (kind, n, use_stream))
                ok(out[0] == 0x5555555555555555 and out[n + 1] == 0x5555555555555555, "rng %s n=%d: wrote outside the request" % (kind, n))
            if kind == "close_open":
                d = out[1:n + 1].view(np.float64)
                ok(bool((d >= 0).all() and (d < 1).all()), "close_open range")
    # a burst of requests on the user stream with no synchronisation in between: each one needs other polynomial rows, staged
    # through the same page-locked buffer (the library must wait for the previous upload before it refills the buffer)
    g.set_stream(stream.value)
    g.set(seed)
    ora = H.OracleDsfmt(seed)
    sizes = (50000, 3, 250001, 40000, 777777, 12)
    outs = [np.zeros(n, dtype=np.uint64) for n in sizes]
    for o_, n in zip(outs, sizes):
        g.get(ptr(o_), n, "open_close")
    mock.cuStreamSynchronize(stream)
    for o_, n in zip(outs, sizes):
        ok(np.array_equal(o_, ora.fill(n, "open_close")), "burst of requests: n=%d" % n)
    # Discard = a Get without the values; Position counts 128-bit words
    g.set_stream(0)
    g.set(seed)
    ora = H.OracleDsfmt(seed)
    g.discard(12345)
    ora.fill(12345)
    ok(g.position == 6173, "position after discarding 12345 values")
    out = np.zeros(777, dtype=np.uint64)
    g.get(ptr(out), 777)
    ok(np.array_equal(out, ora.fill(777)), "rng after Discard")
    # the segment length is a tuning parameter only: every bias yields the same values; the profiling hook reports the split
    g.set_stream(0)
    for bias in (-3, -1, 1, 2, 4):
        g.set_segment_bias(bias)
        g.set(seed)
        ora = H.OracleDsfmt(seed)
        g.profile(True)
        for n in (100001, 4, 654321):
            out = np.zeros(n, dtype=np.uint64)
            g.get(ptr(out), n, "open_open")
            ok(np.array_equal(out, ora.fill(n, "open_open")), "segment bias %d, n=%d" % (bias, n))
        jm, gm, nseg, seglen = g.profile_read()
        ok(nseg >= 1 and seglen >= 16 and (seglen & (seglen - 1)) == 0, "profile_read: %d segments of %d refills" % (nseg, seglen))
        g.profile(False)
    g.set_segment_bias(0)
    # three generators fill the block columns r, r + 3, ... of one column-major matrix from ONE global stream
    m_, n_, nb_, P_ = 640, 1100, 128, 3
    want = H.OracleDsfmt(seed).fill(m_ * n_).reshape(n_, m_)
    got = np.zeros((n_, m_), dtype=np.uint64)
    for r in range(P_):
        gr = chol.Rng64(seed)
        pos = 0
        for c0 in range(r * nb_, n_, P_ * nb_):
            cols = min(nb_, n_ - c0)
            gr.discard(c0 * m_ - pos)
            gr.get(ptr(got[c0:]), cols * m_)
            pos = (c0 + cols) * m_
        ok(gr.position == pos // 2, "position of rank %d" % r)
        gr.close()
    ok(np.array_equal(got, want), "block-cyclic windows of one global stream")
    # far ahead: a window 3e8 values into the stream (checkpointed polynomial walk), then back to the start with a new seed
    g.set(seed)
    far = 300_000_001
    g.discard(far)
    out = np.zeros(5000, dtype=np.uint64)
    g.get(ptr(out), 5000)
    ora = H.OracleDsfmt(seed)
    left = far
    while left > 0:  # the oracle has to walk there (an even chunk size keeps the word alignment; the odd total drops half a word)
        step = min(left, 1 << 22)
        ora.fill(step)
        left -= step
    ok(np.array_equal(out, ora.fill(5000)), "rng window 3e8 values into the stream")
    # a second generator whose seed needs a larger polynomial than the first one's (seed 20121018: degree 19991, seed 12345:
    # 19993): the first generator, in the middle of its stream, must notice, rebuild its jump table and carry on
    g.set(seed)
    ora = H.OracleDsfmt(seed)
    out = np.zeros(10001, dtype=np.uint64)
    g.get(ptr(out), 10001)
    ok(np.array_equal(out, ora.fill(10001)), "first generator before the polynomial changes")
    g2 = chol.Rng64(12345)
    o2 = H.OracleDsfmt(12345)
    out = np.zeros(30001, dtype=np.uint64)
    g2.get(ptr(out), 30001)
    ok(np.array_equal(out, o2.fill(30001)), "second generator (seed 12345)")
    for n in (30001, 999):
        out = np.zeros(n, dtype=np.uint64)
        g.get(ptr(out), n)
        ok(np.array_equal(out, ora.fill(n)), "first generator after the polynomial was extended, n=%d" % n)
    g3 = chol.Rng64(seed)  # same seed, fresh object, interleaved with g
    o3 = H.OracleDsfmt(seed)
    out = np.zeros(4097, dtype=np.uint64)
    g3.get(ptr(out), 4097, "open_open")
    ok(np.array_equal(out, o3.fill(4097, "open_open")), "third generator")
    # argument errors
    ok(g.get(0, 4, allow_invalid=True) == 1, "NULL pointer")
    ok(g.get(ptr(out) + 4, 4, allow_invalid=True) == 1, "misaligned pointer")
    ok(g.get(ptr(out), 0) == 0, "n = 0")
    ok(mock.mock_cuda_ctx_depth() == depth0, "context stack balanced")
    for x in (g, g2, g3):
        x.close()
    mock.cuStreamDestroy_v2(stream)
    ok(mock.mock_cuda_live_allocs() == allocs0, "device allocations balanced: %d vs %d" % (mock.mock_cuda_live_allocs(), allocs0))


if __name__ == "__main__":
    which = sys.argv[2] if len(sys.argv) > 2 else "all"
    if which in ("all", "rng"):
        device_rng()
    if which in ("all", "lapack"):
        single_device_lapack()
    if which in ("all", "blas"):
        single_device_blas()
    if which in ("all", "multi1"):
        multi_device(1)
    if which in ("all", "multi3"):
        multi_device(3)
    if which in ("all", "dist3"):
        dist_ranks(3)
    print("%d checks" % checks)
    print("MOCK PASSED")
