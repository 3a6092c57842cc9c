"""CPU-only tests (no GPU): the C-ABI library loads and exports every symbol include/*.h declares,
the task objects follow the reference's unit tests (test/multigpu/cutask.c), the tile scheduler
covers exactly the needed tiles, and the multi-process layout logic works at world_size 2 (gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import harness as H
from cholesky_b200_loader import load

chol = load()
ROOT = H.ROOT


def test_library_loads_and_exports_every_declared_symbol():
    lib = chol.lib()
    decl = set()
    for f in ("blas.h", "lapack.h", "cumultigpu.h", "error.h", "cholesky_b200.h", "rng.h"):
        s = open(os.path.join(ROOT, "include", f)).read()
        s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
        s = re.sub(r"#define[^\n]*(\\\n[^\n]*)*", "", s)
        decl |= set(re.findall(r"\b((?:cu[A-Z]|cholDist|chol_)\w+)\s*\(", s))
    decl = {d for d in decl if not d.startswith(("cuMem", "cuCtx", "cuStream", "cuEvent", "cuDevice", "cuInit", "cuGetError"))}
    assert len(decl) >= 90
    for name in sorted(decl):
        assert hasattr(lib, name), name
    assert {"cuRng64Create", "cuRng64Destroy", "cuRng64Set", "cuRng64Get", "cuRng64GetOpenOpen", "cuRng64GetOpenClose",
            "cuRng64GetCloseOpen", "cuRng64GetCloseClose"} <= decl  # reference include/rng.h:346-410
    for var in ("xerbla", "errorHandler", "cu_dsfmt_19937_t"):
        C.c_void_p.in_dll(lib, var)
    # the product library has no CPU path: none of the reference's un-prefixed CPU names
    out = subprocess.check_output(["nm", "-D", "--defined-only", chol.LIB_PATH]).decode()
    syms = {l.split()[-1] for l in out.splitlines() if l.strip()}
    for cpu_name in ("dpotrf", "dgemm", "dtrsm", "dsyrk", "dpotri", "dlogdet", "zpotrf", "spotrf"):
        assert cpu_name not in syms
    # ... and it does not link the oracle
    ldd = subprocess.check_output(["readelf", "-d", chol.LIB_PATH]).decode()
    assert "oracle" not in ldd


def test_dsfmt_jump_polynomial_host_code():
    """host/rng_poly.c without a GPU (the functions are exported by the product library): the seed state equals the oracle's,
    the host refill equals the oracle's stream, the minimal polynomial annihilates the seed state (checked exactly inside
    chol_dsfmt_poly_ensure), and x^r mod p applied to the first D states gives the state after r refills - the identity the
    device jump kernel relies on.  Seed 20121018 has a polynomial of degree 19991; a second seed needs 19993, which must
    extend the process-wide polynomial (generation changes) and still serve the first seed."""
    import numpy as np
    lib = chol.lib()
    lib.chol_dsfmt_poly_reset()
    W = 192
    State = C.c_uint64 * (2 * W)

    def seeded(seed):
        st = State()
        lib.chol_dsfmt_seed_state(C.c_uint64(seed), st)
        return st

    def words(st):
        return np.frombuffer(st, dtype=np.uint64).copy()

    s0 = seeded(H.SEED)
    cur = State.from_buffer_copy(s0)
    ora = H.OracleDsfmt(H.SEED)
    for _ in range(3):  # one refill = 382 raw values = the first 191 words of the new state
        lib.chol_dsfmt_refill_host(cur)
        assert np.array_equal(words(cur)[:382], ora.fill(382, "raw"))
    D, gen = C.c_int(0), C.c_uint(0)
    assert lib.chol_dsfmt_poly_ensure(s0, C.byref(D), C.byref(gen)) == 0
    assert D.value == 19991
    gen0 = gen.value
    # Krylov block and direct states
    nk = D.value
    K = np.empty((nk, 2 * W), dtype=np.uint64)
    cur = State.from_buffer_copy(s0)
    for i in range(nk):
        K[i] = words(cur)
        lib.chol_dsfmt_refill_host(cur)
    pw32 = 2 * ((D.value + 63) // 64)
    rows = (C.c_uint32 * (3 * pw32))()
    level, j0 = 10, 18  # refills 18432, 19456, 20480: below and above D
    assert lib.chol_dsfmt_poly_rows(C.c_uint(gen0), C.c_int(level), C.c_uint64(j0), C.c_size_t(3), rows, C.c_size_t(pw32)) == 0
    r_have, state = nk, words(cur)  # `cur` is T^nk s0
    a = np.frombuffer(rows, dtype=np.uint32).reshape(3, pw32)
    for i in range(3):
        r = (j0 + i) << level
        bits = np.unpackbits(a[i].view(np.uint8), bitorder="little")[:nk].astype(bool)
        combo = np.bitwise_xor.reduce(K[bits], axis=0)
        if r < nk:
            want = K[r]
        else:
            while r_have < r:
                lib.chol_dsfmt_refill_host(cur)
                r_have += 1
            want = words(cur)
        assert np.array_equal(combo, want), r
    # stale generation is refused; a seed that needs a larger polynomial extends it
    assert lib.chol_dsfmt_poly_rows(C.c_uint(gen0 + 1), C.c_int(level), C.c_uint64(0), C.c_size_t(1), rows, C.c_size_t(pw32)) != 0
    s1 = seeded(12345)
    assert lib.chol_dsfmt_poly_ensure(s1, C.byref(D), C.byref(gen)) == 0
    assert D.value == 19993 and gen.value != gen0
    assert lib.chol_dsfmt_poly_ensure(s0, C.byref(D), C.byref(gen)) == 0 and D.value == 19993  # the larger polynomial serves both
    lib.chol_dsfmt_poly_reset()


TASKFN = C.CFUNCTYPE(C.c_int, C.c_void_p)


def test_cutask_semantics():
    """reference test/multigpu/cutask.c: NULL function -> INVALID_VALUE, huge args -> OUT_OF_MEMORY,
    the task's result is plumbed back through cuTaskDestroy."""
    lib = chol.lib()
    task = C.c_void_p()
    seen = []

    @TASKFN
    def fn(args):
        seen.append(C.cast(args, C.POINTER(C.c_int))[0])
        return 7

    val = C.c_int(42)
    assert lib.cuTaskCreate(C.byref(task), None, C.byref(val), C.c_size_t(4)) == 1  # CUDA_ERROR_INVALID_VALUE
    assert lib.cuTaskCreate(C.byref(task), fn, None, C.c_size_t(4)) == 1
    assert lib.cuTaskCreate(C.byref(task), fn, C.byref(val), C.c_size_t(1 << 50)) == 2  # CUDA_ERROR_OUT_OF_MEMORY
    assert lib.cuTaskCreate(C.byref(task), fn, C.byref(val), C.c_size_t(4)) == 0
    val.value = 0  # the task owns a COPY of its arguments (multigpu.c:336-342)
    assert lib.cuTaskExecute(task) == 0
    res = C.c_int(-1)
    assert lib.cuTaskDestroy(task, C.byref(res)) == 0
    assert res.value == 7 and seen == [42]


def test_multigpu_argument_errors_without_gpu():
    lib = chol.lib()
    m = C.c_void_p()
    assert lib.cuMultiGPUCreate(C.byref(m), None, C.c_int(1)) == 1
    dev = (C.c_int * 1)(0)
    assert lib.cuMultiGPUCreate(C.byref(m), dev, C.c_int(0)) == 1
    assert lib.cuMultiGPUGetContextCount(None) == 0


@pytest.mark.parametrize("cuplo", ["", "L", "U"])
@pytest.mark.parametrize("m,n", [(1, 1), (128, 64), (129, 65), (1000, 1000), (3000, 700), (700, 3000), (4096, 4096), (16384, 16384)])
def test_tile_schedule_covers_needed_tiles_once(cuplo, m, n):
    BM, BN = 128, 64
    grid, tiles = chol.tile_schedule(cuplo, m, n)
    assert len(tiles) == len(set(tiles))
    want = set()
    for bi in range((m + BM - 1) // BM):
        for bj in range((n + BN - 1) // BN):
            r0, r1, c0, c1 = bi * BM, min(m, bi * BM + BM) - 1, bj * BN, min(n, bj * BN + BN) - 1
            if cuplo == "L" and bi * BM + BM - 1 < c0:
                continue  # (the kernel tests against the full tile height)
            if cuplo == "U" and r0 > bj * BN + BN - 1:
                continue
            want.add((bi, bj))
    assert set(tiles) == want
    assert grid >= len(tiles)
    if m >= 16384 and cuplo:
        assert grid <= 1.15 * len(tiles)  # the triangular grid wastes little (-> 0 as n grows)


@pytest.mark.parametrize("P", [2, 4, 8])
def test_block_cyclic_schedule_matches_global_triangle(P):
    """A rank's packed local columns with the ColMap must select exactly the tiles whose global
    coordinates touch the lower triangle."""
    nb, n = 256, 4096
    for rank in range(P):
        for c in (0, 3, 9):  # current step: rows start at global block c+1; this rank's first column >= c+2
            lc2 = max(0, -(-(c + 2 - rank) // P))
            ncols = chol.dist_local_cols(n, nb, rank, P) - lc2 * nb
            if ncols <= 0:
                continue
            g_first = (lc2 * P + rank) * nb
            mrows = n - g_first
            grid, tiles = chol.tile_schedule("L", mrows, ncols, nb, P, 0)
            want = set()
            for bi in range((mrows + 127) // 128):
                for bj in range((ncols + 63) // 64):
                    nloc = bj * 64
                    gcol = ((nloc // nb) * P) * nb + nloc % nb  # relative to g_first
                    if bi * 128 + 127 >= gcol:
                        want.add((bi, bj))
            assert set(tiles) == want and len(tiles) == len(set(tiles))


def _wanted_tiles(cuplo, m, n, nb=0, P=0, off=0):
    """tiles of the 128 x 64 grid that touch the (block-cyclic) triangle - the rule the kernel's masks assume"""
    want = set()
    for bj in range((n + 63) // 64):
        nloc = bj * 64
        c0 = (((nloc // nb) * P + off) * nb + nloc % nb) if nb else nloc  # global column of the tile's first column
        for bi in range((m + 127) // 128):
            if cuplo == "L" and bi * 128 + 127 < c0:
                continue
            if cuplo == "U" and bi * 128 > c0 + 63:
                continue
            want.add((bi, bj))
    return want


def test_tile_schedule_random_shapes():
    """property test of the DMMA engine's host-built tile schedule (csrc/gemm_common.cuh: super-rows of 16 row blocks, prefix
    table): for random m, n, triangle and block-cyclic map, the decoded tile list is exactly the set of tiles that touch the
    triangle, each once, and the grid is no smaller than the list"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None, derandomize=True)
    @given(st.sampled_from(["", "L", "U"]), st.integers(1, 9000), st.integers(1, 9000))
    def plain(cuplo, m, n):
        grid, tiles = chol.tile_schedule(cuplo, m, n)
        assert len(tiles) == len(set(tiles)) and set(tiles) == _wanted_tiles(cuplo, m, n) and grid >= len(tiles)

    @settings(max_examples=100, deadline=None, derandomize=True)
    @given(st.integers(1, 9000), st.integers(1, 4000), st.sampled_from([64, 128, 256, 512, 1024]), st.integers(1, 8), st.data())
    def cyclic(m, n, nb, P, data):
        off = data.draw(st.integers(0, P - 1))
        grid, tiles = chol.tile_schedule("L", m, n, nb, P, off)
        assert len(tiles) == len(set(tiles)) and set(tiles) == _wanted_tiles("L", m, n, nb, P, off) and grid >= len(tiles)

    plain()
    cyclic()


def test_local_cols_partition():
    for n, nb in ((1000, 256), (4096, 256), (65536, 256), (777, 128)):
        for P in (1, 2, 3, 8):
            assert sum(chol.dist_local_cols(n, nb, r, P) for r in range(P)) == n


GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import bench
from cholesky_b200_loader import load
chol = load()
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, P = dist.get_rank(), dist.get_world_size()
# 1. the NCCL-id exchange path of bench.py, with a stand-in id
uid = bench.exchange_unique_id(lambda: bytes(range(128)), rank)
assert uid == bytes(range(128))
# 2. block-cyclic ownership: every global column is owned exactly once, in the packed order
n, nb = 2000, 256
cols = bench.owned_columns(n, nb, rank, P)
assert len(cols) == chol.dist_local_cols(n, nb, rank, P)
allc = [None] * P
dist.all_gather_object(allc, cols.tolist())
flat = sorted(c for l in allc for c in l)
assert flat == list(range(n))
# 3. info reduction: smallest positive failure wins
info = bench.combine_info(0 if rank == 0 else 513)
assert info == 513
info = bench.combine_info(700 if rank == 0 else 513)
assert info == 513
assert bench.combine_info(0) == 0
# 4. max-over-ranks timing reduction
t = bench.max_over_ranks(1.0 + rank)
assert t == float(P)
# 5. the leading-block parity check of bench.py: every rank holds its block-cyclic columns of a factor (here numpy's),
#    rank 0 assembles the leading block and compares it with the CPU reference's dpotrf; a corrupted shard must show
n, nb = 600, 128
rng = np.random.default_rng(5)
Bm = rng.random((n, 64)) - 0.5
A0 = Bm @ Bm.T + n * np.eye(n)
L = np.linalg.cholesky(A0)
cols = bench.owned_columns(n, nb, rank, P)
fac = torch.from_numpy(np.ascontiguousarray(L[:, cols].T))     # (nloc, n) row-major == column-major n x nloc
inp = torch.from_numpy(np.ascontiguousarray(A0[:, cols].T))
res = bench.lead_block_parity(fac, inp, cols, n, P, rank)
if rank == 0:
    assert res[0] <= 64 * n * 2.3e-16 and res[0] <= res[1], res
else:
    assert res is None
if rank == 1:
    fac[0, 300] += 1e-6  # one entry of the lower triangle (column 128 of the matrix, row 300)
res = bench.lead_block_parity(fac, inp, cols, n, P, rank)
if rank == 0:
    assert res[0] > res[1], res
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


def test_multiprocess_host_logic_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "OK" in o, o


@pytest.mark.parametrize("profile", [0, 1, 2])
@pytest.mark.parametrize("n,nslabs", [(1, 1), (63, 3), (400, 3), (1500, 4), (32768, 8), (65536, 5)])
def test_inverse_slab_bounds_cover_and_balance(profile, n, nslabs):
    """Slab boundaries of the distributed TRTRI / LAUUM: monotone, multiples of 64 (except n itself),
    cover [0, n) exactly, and equalise the cost profile when the matrix is large enough."""
    lib = chol.lib()
    b = (C.c_size_t * (nslabs + 1))()
    lib.chol_inverse_slab_bounds(C.c_int(profile), C.c_size_t(n), C.c_int(nslabs), b)
    b = list(b)
    assert b[0] == 0 and b[-1] == n
    assert all(b[i] <= b[i + 1] for i in range(nslabs))
    assert all(x % 64 == 0 or x == n for x in b)
    if n >= 32768:
        x = np.arange(n) + 0.5
        dens = [(n - x) ** 2, x ** 2, x * (n - x)][profile]
        work = np.array([dens[b[i]:b[i + 1]].sum() for i in range(nslabs)])
        assert work.max() <= 1.06 * work.mean()


def _slab_bounds(profile, n, nslabs):
    b = (C.c_size_t * (nslabs + 1))()
    chol.lib().chol_inverse_slab_bounds(C.c_int(profile), C.c_size_t(n), C.c_int(nslabs), b)
    return list(b)


def _tri(A, uplo, unit=False):
    T = np.tril(A) if uplo == "L" else np.triu(A)
    return T - np.diag(np.diag(T)) + np.eye(len(T)) if unit else T


@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("n,nslabs", [(200, 1), (200, 3), (700, 4), (130, 8)])
def test_slab_inverse_decomposition_model(uplo, n, nslabs):
    """The algebra host/multigpu_blas.c distributes (CHOL_MGPU_INVERSE=slab), replayed in numpy with
    the library's own slab boundaries: every slab reads only the ORIGINAL matrix, writes only its own
    output slab, and together they give inv(T) / T^H T with the other triangle untouched."""
    rng = np.random.default_rng(3)
    M = rng.standard_normal((n, n))
    F = np.linalg.cholesky(M @ M.T + n * np.eye(n))
    junk = rng.standard_normal((n, n))
    St = np.tril(F) + np.triu(junk, 1) if uplo == "L" else np.triu(F.T) + np.tril(junk, -1)

    def put_triangle(out, lo, X):
        for c in range(len(X)):
            if uplo == "L":
                out[lo + c:lo + len(X), lo + c] = X[c:, c]
            else:
                out[lo:lo + c + 1, lo + c] = X[:c + 1, c]

    # TRTRI: lower columns J need T = L[lo:, lo:]; upper columns J need T = U[:hi, :hi]
    b = _slab_bounds(0 if uplo == "L" else 1, n, nslabs)
    out = St.copy()
    for lo, hi in zip(b[:-1], b[1:]):
        w = hi - lo
        if w == 0:
            continue
        if uplo == "L":
            T = St[lo:, lo:]
            Tww, Trr, Trw = T[:w, :w], T[w:, w:], T[w:, :w]
        else:
            T = St[:hi, :hi]
            Tww, Trr, Trw = T[lo:, lo:], T[:lo, :lo], T[:lo, lo:]
        X = np.linalg.inv(_tri(Tww, uplo))
        if len(Trr):
            Y = -np.linalg.inv(_tri(Trr, uplo)) @ (Trw @ _tri(X, uplo))
            if uplo == "L":
                out[hi:, lo:hi] = Y
            else:
                out[:lo, lo:hi] = Y
        put_triangle(out, lo, X)
    want = np.linalg.inv(_tri(St, uplo))
    assert np.abs(_tri(out, uplo) - _tri(want, uplo)).max() <= 1e-10 * np.abs(want).max()
    strict = (lambda A: np.triu(A, 1)) if uplo == "L" else (lambda A: np.tril(A, -1))
    assert np.array_equal(strict(out), strict(St))

    # LAUUM: lower rows I need R = L[lo:, :hi]; upper columns I need R = U[:hi, lo:]
    b = _slab_bounds(2, n, nslabs)
    out = St.copy()
    for lo, hi in zip(b[:-1], b[1:]):
        if hi == lo:
            continue
        if uplo == "L":
            R = St[lo:, :hi]
            Aii = np.tril(R[:, lo:hi])
            out[lo:hi, :lo] = Aii.T @ R[:, :lo]
            put_triangle(out, lo, Aii.T @ Aii)
        else:
            R = St[:hi, lo:]
            Aii = np.triu(R[lo:hi, :])
            out[:lo, lo:hi] = R[:lo, :] @ Aii.T
            put_triangle(out, lo, Aii @ Aii.T)
    T = _tri(St, uplo)
    want = T.T @ T if uplo == "L" else T @ T.T
    assert np.abs(_tri(out, uplo) - _tri(want, uplo)).max() <= 1e-10 * np.abs(want).max()
    assert np.array_equal(strict(out), strict(St))


def test_default_block_sizes_are_consistent():
    """chol_default_nb (host/lapack.c, through cholDistBlockSize without a handle): every default is a multiple of the
    diagonal-block size of its precision (128; z 64) and of the 64-column tile, does not decrease with n, and keeps
    the pipelined panel solve within its event ring (nb / db <= 16)."""
    lib = chol.lib()
    f = lib.cholDistBlockSize
    f.restype = C.c_size_t
    lib.chol_diag_max.restype = C.c_int
    for p in "sdcz":
        db = lib.chol_diag_max(C.c_int(ord(p)))
        last = 0
        for n in (1, 100, 1000, 4096, 6144, 8192, 12288, 16384, 24576, 32768, 49152, 65536, 131072):
            nb = int(f(None, C.c_int(ord(p)), C.c_size_t(n)))
            assert nb % db == 0 and nb % 64 == 0 and 128 <= nb <= 1024, (p, n, nb)
            assert nb // db <= 16
            assert nb >= last, (p, n, nb, last)
            last = nb


def test_bench_reference_arm_runs_on_cpu(tmp_path):
    """`bench.py --impl reference` (the arm the driver runs next to ours): one JSON line with the contract's keys, the
    reference's CPU dpotrf timed on a bounded sample, e2e repeating the value with zero transfer bytes."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--order", "1024"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, check=True).stdout.decode()
    line = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "GFLOP/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["sample_resid"] < 1e-13 and "sample_N" in line["config"]
    # a rank other than 0 under torchrun prints nothing and exits 0
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=dict(os.environ, RANK="1", WORLD_SIZE="2"), stdout=subprocess.PIPE, timeout=120)
    assert p.returncode == 0 and p.stdout.decode().strip() == ""
