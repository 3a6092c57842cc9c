"""mock_fuzz.py — randomised sweep of the multi-GPU HOST drivers on the mock back end (no GPU).

    python scripts/mock_fuzz.py [--seconds 600] [--seed 1] [--devices 2,4,5,8] [--jitter]

TEST INFRASTRUCTURE ONLY.  Builds the product's host/*.c against tests/mock (the same recipe as
tests/test_host_drivers_mock.py), then draws random (routine, precision, uplo, n, nb, device count, page-locked or
not, leading dimension) cases for cuMultiGPU{potrf,trtri,lauum,potri} and the cholDist ranks form, and compares each
with the CPU oracle under the tolerances of the committed tests (64 n eps elementwise, the other triangle bit-untouched,
info equal).  The committed suite pins 1 and 3 devices at a handful of sizes; this sweep is for the shapes in between:
fewer blocks than devices, n = k nb +- 1, nb larger than n, odd device counts.  Each device count runs in its own
child process (the mock reads MOCK_CUDA_DEVICES once).  Prints one line per failure and a JSON summary.
"""
import argparse
import ctypes as C
import glob
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "mock")
INC = ["-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include"]


def build(d):
    objs = []
    for src in sorted(glob.glob(os.path.join(ROOT, "cuda-cholesky_b200", "host", "*.c"))) + [os.path.join(MOCK, "mock_cuda.c")]:
        o = os.path.join(d, os.path.basename(src)[:-2] + ".o")
        subprocess.check_call(["gcc", "-std=c99", "-O1", "-fPIC", "-D_POSIX_C_SOURCE=200809L", *INC, "-c", src, "-o", o])
        objs.append(o)
    ko = os.path.join(d, "mock_kernels.o")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-c", os.path.join(MOCK, "mock_kernels.cpp"), "-o", ko])
    lib = os.path.join(d, "libcholesky_b200_mock.so")
    subprocess.check_call(["g++", "-shared", "-o", lib, *objs, ko, "-lpthread", "-ldl", "-lm"])
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-fPIC", "-shared", os.path.join(MOCK, "mock_nccl.c"), "-o",
                           os.path.join(d, "libnccl.so.2"), "-lpthread"])
    return lib


def child_single(lib, seconds, seed):
    """one device, the cu* entry points: potrf (+ not-PD), potrs, trtri, lauum, potri, trsm, trmm at random n / nb / ld"""
    import numpy as np
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness as H
    from cholesky_b200_loader import load
    from harness import PREC
    chol = load()
    mock = C.CDLL(lib, mode=C.RTLD_GLOBAL)
    mock.mock_cuda_live_allocs.restype = C.c_long
    chol._lib = mock
    chol._drv = mock
    chol.set_xerbla_silent()
    ORA = H.CpuLib("oracle")
    rng = np.random.default_rng(seed * 1000)
    allocs0 = mock.mock_cuda_live_allocs()
    lap, blas = chol.LapackHandle(), chol.BlasHandle()
    fails, cases, by = [], 0, {}

    def dup(M):
        out = H.fmat(M.shape[0], M.shape[1], M.dtype, M.strides[1] // M.itemsize)
        out[:, :] = M
        return out

    def ldof(M):
        return M.strides[1] // M.itemsize

    def close(tag, got, want, p, n, cond=1.0):
        nonlocal cases
        cases += 1
        by[tag.split()[0]] = by.get(tag.split()[0], 0) + 1
        scale = max(1.0, float(np.abs(want).max()))
        err = float(np.abs(got - want).max())
        if not err <= 64 * max(n, 1) * H.eps(p) * scale * cond:
            fails.append("%s: err %.3e" % (tag, err))

    t_end = time.time() + seconds
    while time.time() < t_end and not fails:
        p = "dzsc"[int(rng.choice([0, 0, 0, 1, 1, 2, 3]))]
        cplx = PREC[p]["cplx"]
        hT = "C" if cplx else "T"
        uplo = "LU"[int(rng.integers(0, 2))]
        nb = int(rng.choice([0, 0, 0, 128, 256, 384, 512, 1024]))
        nbe = nb or 128
        kind = int(rng.integers(0, 4))
        n = int(rng.integers(1, 40)) if kind == 0 else (max(1, int(rng.integers(1, 9)) * nbe + int(rng.integers(-2, 3))) if kind == 1
                                                        else int(rng.integers(40, 1400)))
        n = min(n, 1500 if p in "ds" else 1000)
        ldx = n + int(rng.choice([0, 0, 1, 4]))
        routine = str(rng.choice(["potrf", "potrf_bad", "potrs", "trtri", "lauum", "potri", "trsm", "trmm"]))
        tag = "%s %s %s n=%d nb=%d ld=%d" % (routine, p, uplo, n, nb, ldx)
        if os.environ.get("MOCK_FUZZ_VERBOSE"):
            sys.stderr.write(tag + "\n")
        try:
            lap.set_block_size(nb)
            A0 = H.spd_matrix(p, n, ld=ldx, k=min(n, 32), seed=int(rng.integers(1, 1 << 30)))
            if routine in ("trsm", "trmm"):
                side = "LR"[int(rng.integers(0, 2))]
                trans = ("N", "T", hT)[int(rng.integers(0, 3))]
                diag = "NU"[int(rng.integers(0, 2))]
                other = int(rng.integers(1, 300))
                m_, n_ = (n, other) if side == "L" else (other, n)
                T0 = A0
                ORA.potrf(p, "L", T0)
                Tm = dup(T0)
                Tm[:, :] = np.tril(T0) + np.tril(T0, -1).conj().T
                if diag == "U":
                    Tm[:, :] = Tm / (np.abs(np.diag(Tm)).max() * 4)
                alpha = (0.9 + 0.3j) if cplx else 0.9
                B0 = H.rand_matrix(p, m_, n_, seed=int(rng.integers(1, 1 << 30)), ld=m_ + int(rng.choice([0, 1, 3])))
                want, got = dup(B0), dup(B0)
                getattr(ORA, routine)(p, side, uplo, trans, diag, m_, n_, alpha, Tm, want)
                getattr(blas, routine)(p, side, uplo, trans, diag, m_, n_, alpha, Tm.ctypes.data, ldof(Tm), got.ctypes.data, ldof(got))
                mock.cuStreamSynchronize(None)
                close(tag + " %s%s%s m=%d n=%d" % (side, trans, diag, m_, n_), got, want, p, n, cond=4.0 if routine == "trsm" else 1.0)
                continue
            if routine == "potrf_bad":
                j = int(rng.integers(0, n))
                A0[j, j] = -1.0
            want = dup(A0)
            winfo = ORA.potrf(p, uplo, want)
            A = dup(A0)
            info = lap.potrf(p, uplo, n, A.ctypes.data, ldof(A))
            if info != winfo:
                fails.append("%s: info %r oracle %r" % (tag, info, winfo))
                continue
            if winfo != 0:
                cases += 1
                by["potrf_bad"] = by.get("potrf_bad", 0) + 1
                continue
            close(tag, H.tri(A, uplo), H.tri(want, uplo), p, n)
            oth = (np.triu(A, 1), np.triu(A0, 1)) if uplo == "L" else (np.tril(A, -1), np.tril(A0, -1))
            if not np.array_equal(*oth):
                fails.append(tag + ": other triangle changed")
            if routine == "potrs":
                nrhs = int(rng.integers(1, 40))
                B0 = H.rand_matrix(p, n, nrhs, seed=5, ld=n + 1)
                Bw, Bg = dup(B0), dup(B0)
                for trans in (("N", hT) if uplo == "L" else (hT, "N")):
                    ORA.trsm(p, "L", uplo, trans, "N", n, nrhs, 1.0, want, Bw)
                if lap.potrs(p, uplo, n, nrhs, A.ctypes.data, ldof(A), Bg.ctypes.data, ldof(Bg)) != 0:
                    fails.append(tag + ": potrs info")
                close(tag + " nrhs=%d" % nrhs, Bg, Bw, p, n, cond=4.0)
            elif routine in ("trtri", "lauum", "potri"):
                diag = "NU"[int(rng.integers(0, 2))] if routine == "trtri" else "N"
                src = want
                if diag == "U":
                    src = dup(want)
                    src[:, :] = want / (np.abs(np.diag(want)).max() * 4)
                w2, g2 = dup(src), dup(src)
                if routine == "trtri":
                    wi = ORA.trtri(p, uplo, diag, w2)
                    gi = lap.trtri(p, uplo, diag, n, g2.ctypes.data, ldof(g2))
                else:
                    wi = getattr(ORA, routine)(p, uplo, w2) or 0
                    gi = getattr(lap, routine)(p, uplo, n, g2.ctypes.data, ldof(g2))
                if wi != gi:
                    fails.append("%s: info %r oracle %r" % (tag, gi, wi))
                close(tag + " " + diag, H.tri(g2, uplo), H.tri(w2, uplo), p, n)
                oth = (np.triu(g2, 1), np.triu(src, 1)) if uplo == "L" else (np.tril(g2, -1), np.tril(src, -1))
                if not np.array_equal(*oth):
                    fails.append(tag + ": other triangle changed")
        except Exception as e:  # noqa: BLE001
            fails.append("%s: exception %r" % (tag, e))
    lap.close()
    blas.close()
    leaked = mock.mock_cuda_live_allocs() - allocs0
    if leaked:
        fails.append("single device: %d device allocations leaked" % leaked)
    print(json.dumps({"devices": "single (cu*)", "cases": cases, "by_routine": by, "failures": fails}), flush=True)
    return 1 if fails else 0


def child(lib, P, seconds, seed):
    if P == 0:
        return child_single(lib, seconds, seed)
    import threading

    import numpy as np
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness as H
    from bench import owned_columns
    from cholesky_b200_loader import load
    from harness import PREC
    chol = load()
    mock = C.CDLL(lib, mode=C.RTLD_GLOBAL)
    mock.mock_cuda_live_allocs.restype = C.c_long
    chol._lib = mock
    chol._drv = mock
    chol.set_xerbla_silent()
    ORA = H.CpuLib("oracle")
    rng = np.random.default_rng(seed * 1000 + P)
    allocs0 = mock.mock_cuda_live_allocs()
    mg = chol.MultiGPU(list(range(P)))
    fails, cases, by = [], 0, {}

    def check(tag, got, want, orig, uplo, p, n, info, winfo):
        nonlocal cases
        cases += 1
        by[tag.split()[0]] = by.get(tag.split()[0], 0) + 1
        if info != winfo:
            fails.append("%s: info %r, oracle %r" % (tag, info, winfo))
            return
        if winfo != 0:
            return
        scale = max(1.0, float(np.abs(want).max()))
        err = float(np.abs(H.tri(got, uplo) - H.tri(want, uplo)).max())
        if not err <= 64 * n * H.eps(p) * scale:
            fails.append("%s: err %.3e" % (tag, err))
        a = np.triu(got, 1) if uplo == "L" else np.tril(got, -1)
        b = np.triu(orig, 1) if uplo == "L" else np.tril(orig, -1)
        if not np.array_equal(a, b):
            fails.append("%s: other triangle changed" % tag)

    def check2(tag, got, want, p, kdim, slack=1.0):
        nonlocal cases
        cases += 1
        by[tag.split()[0]] = by.get(tag.split()[0], 0) + 1
        scale = max(1.0, float(np.abs(want).max()))
        err = float(np.abs(got - want).max())
        if not err <= 64 * max(kdim, 1) * H.eps(p) * scale * slack:
            fails.append("%s: err %.3e" % (tag, err))

    def dup(M):
        """copy that KEEPS the leading dimension (ndarray.copy() of a padded view is compact)"""
        ldm = M.strides[1] // M.itemsize
        out = H.fmat(M.shape[0], M.shape[1], M.dtype, ldm)
        out[:, :] = M
        return out

    def ldof(M):
        return M.strides[1] // M.itemsize

    def draw_n(nb_eff):
        kind = rng.integers(0, 5)
        if kind == 0:
            return int(rng.integers(1, 40))
        if kind == 1:  # around a multiple of the block size
            return max(1, int(rng.integers(1, 3 * P + 2)) * nb_eff + int(rng.integers(-2, 3)))
        if kind == 2:  # fewer blocks than devices
            return max(1, int(rng.integers(1, max(2, P))) * nb_eff - int(rng.integers(0, nb_eff)))
        return int(rng.integers(40, 1500))

    t_end = time.time() + seconds
    while time.time() < t_end:
        p = "dzsc"[int(rng.choice([0, 0, 0, 1, 1, 2, 3]))]
        uplo = "LU"[int(rng.integers(0, 2))]
        nb = int(rng.choice([0, 0, 128, 128, 256, 384, 512]))
        nb_eff = nb or (128 if p in "zc" else 256)
        n = min(draw_n(nb_eff), 1800 if p in "ds" else 1200)
        ldx = n + int(rng.choice([0, 0, 1, 5]))
        routine = str(rng.choice(["potrf", "potrf", "potrf_bad", "trtri", "lauum", "potri", "dist", "trtri_zero", "gemm", "rankk", "trsm", "trmm"]))
        if routine == "dist" and p in "sc":
            p = "d"
        if routine in ("gemm", "rankk", "trsm", "trmm"):  # the host-pointer BLAS-3 tilers (host/multigpu_blas.c)
            cplx = PREC[p]["cplx"]
            hT = "C" if cplx else "T"
            sd = int(rng.integers(1, 1 << 30))
            m_, n_, k_ = (int(rng.integers(1, 40)) if rng.integers(0, 4) == 0 else int(rng.integers(1, 700)) for _ in range(3))
            alpha, beta = ((0.7 - 0.2j, -0.3 + 0.5j) if cplx else (0.7, -0.3))
            if rng.integers(0, 5) == 0:
                beta = 0.0
            tag = "%s %s m=%d n=%d k=%d P=%d" % (routine, p, m_, n_, k_, P)
            if os.environ.get("MOCK_FUZZ_VERBOSE"):
                sys.stderr.write(tag + "\n")
                sys.stderr.flush()
            try:
                if routine == "gemm":
                    tA, tB = (("N", "T", hT)[int(rng.integers(0, 3))] for _ in range(2))
                    A = H.rand_matrix(p, *((m_, k_) if tA == "N" else (k_, m_)), seed=sd, ld=(m_ if tA == "N" else k_) + int(rng.integers(0, 3)))
                    B = H.rand_matrix(p, *((k_, n_) if tB == "N" else (n_, k_)), seed=sd + 1)
                    Cm = H.rand_matrix(p, m_, n_, seed=sd + 2, ld=m_ + int(rng.integers(0, 4)))
                    want = dup(Cm)
                    ORA.gemm(p, tA, tB, m_, n_, k_, alpha, A, B, beta, want)
                    mg.gemm(p, tA, tB, m_, n_, k_, alpha, A.ctypes.data, ldof(A), B.ctypes.data, ldof(B), beta, Cm.ctypes.data, ldof(Cm))
                    check2(tag + " %s%s" % (tA, tB), Cm, want, p, k_)
                elif routine == "rankk":
                    trans = ("N", hT)[int(rng.integers(0, 2))]
                    A = H.rand_matrix(p, *((n_, k_) if trans == "N" else (k_, n_)), seed=sd)
                    C0 = H.spd_matrix(p, n_, seed=sd + 1, k=min(n_, 8), ld=n_ + int(rng.integers(0, 3)))
                    want, got = dup(C0), dup(C0)
                    a_r, b_r = -0.5, (0.0 if beta == 0.0 else 1.25)  # real scalars (HERK)
                    ORA.rankk(p, uplo, trans, n_, k_, a_r, A, b_r, want)
                    mg.rankk(p, uplo, trans, n_, k_, a_r, A.ctypes.data, ldof(A), b_r, got.ctypes.data, ldof(got))
                    check2(tag + " %s%s" % (uplo, trans), H.tri(got, uplo), H.tri(want, uplo), p, k_)
                    oth = (np.triu(got, 1), np.triu(C0, 1)) if uplo == "L" else (np.tril(got, -1), np.tril(C0, -1))
                    if not np.array_equal(*oth):
                        fails.append(tag + ": other triangle changed")
                else:
                    side = "LR"[int(rng.integers(0, 2))]
                    trans = ("N", "T", hT)[int(rng.integers(0, 3))]
                    diag = "NU"[int(rng.integers(0, 2))]
                    na = m_ if side == "L" else n_
                    T0 = H.spd_matrix(p, na, seed=sd, k=min(na, 16))
                    ORA.potrf(p, "L", T0)
                    Tm = dup(T0)
                    Tm[:, :] = np.tril(T0) + np.tril(T0, -1).conj().T
                    if diag == "U":
                        Tm[:, :] = Tm / (np.abs(np.diag(Tm)).max() * 4)
                    B0 = H.rand_matrix(p, m_, n_, seed=sd + 1, ld=m_ + int(rng.integers(0, 3)))
                    want, got = dup(B0), dup(B0)
                    getattr(ORA, routine)(p, side, uplo, trans, diag, m_, n_, alpha, Tm, want)
                    getattr(mg, routine)(p, side, uplo, trans, diag, m_, n_, alpha, Tm.ctypes.data, ldof(Tm), got.ctypes.data, ldof(got))
                    check2(tag + " %s%s%s%s" % (side, uplo, trans, diag), got, want, p, na, 4.0 if routine == "trsm" else 1.0)
            except Exception as e:  # noqa: BLE001
                fails.append("%s: exception %r" % (tag, e))
                break
            continue
        tag = "%s %s %s n=%d nb=%d ld=%d P=%d" % (routine, p, uplo, n, nb, ldx, P)
        if os.environ.get("MOCK_FUZZ_VERBOSE"):
            sys.stderr.write(tag + "\n")
            sys.stderr.flush()
        try:
            mg.set_block_size(nb)
            A0 = H.spd_matrix(p, n, ld=ldx, k=min(n, 32), seed=int(rng.integers(1, 1 << 30)))
            if routine in ("potrf", "potrf_bad"):
                if routine == "potrf_bad":
                    for _ in range(int(rng.integers(1, 3))):
                        j = int(rng.integers(0, n))
                        A0[j, j] = -1.0
                want = dup(A0)
                winfo = ORA.potrf(p, uplo, want)
                A = dup(A0)
                pinned = bool(rng.integers(0, 2))
                if pinned:
                    mock.mock_cuda_register_pinned(C.c_void_p(A.ctypes.data), C.c_size_t(A.strides[1] * (A.shape[1] - 1) + n * A.itemsize))
                info = mg.potrf(p, uplo, n, A.ctypes.data, ldof(A))
                if pinned:
                    mock.mock_cuda_unregister_pinned(C.c_void_p(A.ctypes.data))
                check(tag + (" pinned" if pinned else ""), A, want, A0, uplo, p, n, info, winfo)
            elif routine == "dist":
                uid = chol.dist_unique_id()
                want = A0.copy(order="F")
                winfo = ORA.potrf(p, "L", want)
                got = A0.copy(order="F")
                infos, errs = [None] * P, []

                def rank_main(r):
                    try:
                        d = chol.Dist(r, P, uid, nb=nb)
                        cols = owned_columns(n, d.block_size(p, n), r, P)
                        loc = np.asfortranarray(A0[:, cols]) if len(cols) else H.fmat(n, 1, PREC[p]["np"])
                        infos[r] = d.potrf(p, n, loc.ctypes.data, loc.strides[1] // loc.itemsize)
                        if len(cols):
                            got[:, cols] = loc
                        d.close()
                    except Exception as e:  # noqa: BLE001
                        errs.append(repr(e))

                ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(P)]
                for t in ts:
                    t.start()
                for t in ts:
                    t.join(300)
                if errs or any(t.is_alive() for t in ts):
                    fails.append("%s: ranks %s alive %s" % (tag, errs, [t.is_alive() for t in ts]))
                    if any(t.is_alive() for t in ts):
                        break
                pos = [i for i in infos if i]
                check(tag, got, want, A0, "L", p, n, min(pos) if pos else 0, winfo)
            else:
                F = A0
                ORA.potrf(p, uplo, F)
                os.environ.pop("CHOL_MGPU_INVERSE", None)
                if routine == "trtri_zero":
                    for _ in range(int(rng.integers(1, 3))):
                        j = int(rng.integers(0, n))
                        F[j, j] = 0.0
                diag = "NU"[int(rng.integers(0, 2))] if routine == "trtri" else "N"
                src = F
                if diag == "U":
                    src = dup(F)
                    src[:, :] = F / (np.abs(np.diag(F)).max() * 4)
                want = dup(src)
                got = dup(src)
                if routine in ("trtri", "trtri_zero"):
                    winfo = ORA.trtri(p, uplo, diag, want)
                    info = mg.trtri(p, uplo, diag, n, got.ctypes.data, got.strides[1] // got.itemsize)
                else:
                    winfo = getattr(ORA, routine)(p, uplo, want) or 0
                    info = getattr(mg, routine)(p, uplo, n, got.ctypes.data, got.strides[1] // got.itemsize)
                check(tag + " " + diag, got, want, src, uplo, p, n, info, winfo)
        except Exception as e:  # noqa: BLE001
            fails.append("%s: exception %r" % (tag, e))
            break
    mg.set_block_size(0)
    mg.close()
    leaked = mock.mock_cuda_live_allocs() - allocs0
    if leaked:
        fails.append("P=%d: %d device allocations leaked" % (P, leaked))
    print(json.dumps({"devices": P, "cases": cases, "by_routine": by, "failures": fails}), flush=True)
    return 1 if fails else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=600.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--devices", default="2,4,5,8", help="mock device counts, one child each; 0 = the single-device cu* entry points")
    ap.add_argument("--jitter", action="store_true", help="random stalls on every mock stream (ordering races)")
    ap.add_argument("--child", nargs=2, metavar=("LIB", "P"))
    args = ap.parse_args()
    if args.child:
        return child(args.child[0], int(args.child[1]), args.seconds, args.seed)
    d = tempfile.mkdtemp(prefix="mockfuzz")
    lib = build(d)
    devs = [int(x) for x in args.devices.split(",")]
    procs = []
    for P in devs:
        env = dict(os.environ, LD_LIBRARY_PATH=d + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""), OMP_NUM_THREADS="1",
                   MOCK_CUDA_DEVICES=str(max(P, 1)))
        if args.jitter:
            env["MOCK_CUDA_JITTER"] = "1"
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--child", lib, str(P), "--seconds",
                                       str(args.seconds), "--seed", str(args.seed)], env=env))
    rc = 0
    for pr in procs:
        try:
            rc |= pr.wait(timeout=args.seconds + 900)
        except subprocess.TimeoutExpired:
            pr.kill()
            print(json.dumps({"failures": ["child hung (killed)"]}))
            rc |= 1
    return rc


if __name__ == "__main__":
    sys.exit(main())
