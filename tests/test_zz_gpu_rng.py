"""GPU parity tests of the device generator cuRng64* (include/rng.h) through the C-ABI: the values the
B200 produces must be, bit for bit, the stream of the oracle's sequential dSFMT-19937 (oracle/dsfmt_oracle.c, pinned to
the reference's rng/dsfmt_19937.c in tests/test_oracle_pins.py) for the same seed and the same sequence of calls.
Integer work: every comparison is exact.

(The file name sorts after the other GPU test files on purpose: this is the newest entry point.)"""
import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no GPU", allow_module_level=True)

import gpu_util as G  # noqa: E402

PAT = 0x5555555555555555


@pytest.fixture(scope="module")
def rng():
    g = G.chol.Rng64()
    yield g
    g.close()


def device_get(g, n, kind="close_open", misalign=False):
    """n values from the device generator as np.uint64 bit patterns; checks that nothing outside the request was written"""
    buf = torch.full((n + 4,), PAT, dtype=torch.int64, device="cuda")
    off = 1 if misalign else 2  # torch allocations are 256-byte aligned: +8 bytes = only 8-byte aligned, +16 = 16-byte aligned
    g.get(buf.data_ptr() + 8 * off, n, kind)
    out = buf.cpu().numpy().view(np.uint64)
    assert (out[:off] == PAT).all() and (out[off + n:] == PAT).all(), "wrote outside the request"
    return out[off:off + n]


@pytest.mark.parametrize("kind", list(H.DSFMT_KINDS))
def test_stream_equals_reference_generator(rng, kind):
    """one stream over many calls: sizes cross refill (382 values) and segment boundaries, odd sizes drop half a word exactly
    as rng/dsfmt.c:245-251 does, output pointers with 8- and 16-byte alignment"""
    rng.set(H.SEED)
    ora = H.OracleDsfmt(H.SEED)
    for k, n in enumerate((1, 2, 3, 381, 382, 383, 1000, 6111, 6113, 20001, 7, 150000, 2, 400001, 5, 3000000, 11)):
        got = device_get(rng, n, kind, misalign=bool(k & 1))
        assert np.array_equal(got, ora.fill(n, kind)), (kind, n)
    if kind == "close_open":
        d = got.view(np.float64)
        assert (d >= 0).all() and (d < 1).all()


def test_large_fill_equals_reference_generator(rng):
    """2^27 + 1 doubles (1 GiB) in one call: thousands of segments, every one started by the jump kernel"""
    n = (1 << 27) + 1
    rng.set(H.SEED)
    buf = torch.empty(n, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rng.get(buf.data_ptr(), n)  # first call of this shape: polynomial rows are computed on the host and uploaded
    rng.set(H.SEED)
    torch.cuda.synchronize()
    e0.record()
    rng.get(buf.data_ptr(), n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("device dSFMT: %d doubles in %.3f ms = %.1f GB/s" % (n, ms, 8e-6 * n / ms))
    got = buf.cpu().numpy().view(np.uint64)
    want = H.OracleDsfmt(H.SEED).fill(n)
    assert np.array_equal(got, want)


def test_harness_input_matrix_generated_on_device(rng):
    """the BASELINE input recipe (B = u - 0.5 from the seeded stream, column-major) made on the device equals the matrix the
    CPU harness builds from the oracle's generator, exactly"""
    m, n = 1000, 333
    want = np.ascontiguousarray(H.rand_matrix("d", m, n).T)  # (n, m) row-major == column-major m x n
    rng.set(H.SEED)
    B = torch.empty((n, m), dtype=torch.float64, device="cuda")
    rng.get(B.data_ptr(), m * n)
    B -= 0.5
    assert np.array_equal(B.cpu().numpy(), want)


def test_discard_and_block_cyclic_windows():
    """every rank of a multi-GPU job can produce ITS block columns of one global stream: Discard moves the position without
    producing values.  Three generators fill the block columns c = r, r + 3, ... of a 640 x 1100 column-major matrix; together they
    equal the matrix one generator fills in a single call."""
    m, n, nb, P = 640, 1100, 128, 3
    want = H.OracleDsfmt(H.SEED).fill(m * n).reshape(n, m)
    got = torch.zeros((n, m), dtype=torch.int64, device="cuda")
    gens = [G.chol.Rng64(H.SEED) for _ in range(P)]
    for r, g in enumerate(gens):
        pos = 0
        for c0 in range(r * nb, n, P * nb):
            cols = min(nb, n - c0)
            g.discard(c0 * m - pos)  # m is even: whole words
            g.get(got[c0].data_ptr(), cols * m)
            pos = (c0 + cols) * m
        assert g.position == pos // 2
        g.close()
    assert np.array_equal(got.cpu().numpy().view(np.uint64), want)


def test_far_window_and_second_seed(rng):
    """a window 3e8 values into the stream (the host walks the polynomial over its checkpoints); then a second generator whose
    seed needs a larger minimal polynomial (degree 19993 instead of 19991): the first one, in mid-stream, rebuilds its jump
    table and continues exactly"""
    rng.set(H.SEED)
    ora = H.OracleDsfmt(H.SEED)
    assert np.array_equal(device_get(rng, 10001), ora.fill(10001))
    g2 = G.chol.Rng64(12345)
    assert np.array_equal(device_get(g2, 30001), H.OracleDsfmt(12345).fill(30001))
    for n in (30001, 999):
        assert np.array_equal(device_get(rng, n), ora.fill(n)), n
    g2.close()
    far = 300_000_001
    rng.set(H.SEED)
    rng.discard(far)
    got = device_get(rng, 5000)
    ora = H.OracleDsfmt(H.SEED)
    left = far
    while left > 0:
        step = min(left, 1 << 22)
        ora.fill(step)
        left -= step
    assert np.array_equal(got, ora.fill(5000))


def test_tuning_hooks_do_not_change_the_stream():
    """the segment length (cuRng64SetSegmentBias) is a tuning parameter only; the profiling hook reports the jump / generation split"""
    g = G.chol.Rng64()
    for bias in (-3, -2, -1, 1, 2, 4):
        g.set_segment_bias(bias)
        g.set(H.SEED)
        g.profile(True)
        ora = H.OracleDsfmt(H.SEED)
        for n in (1500001, 7, 4000000, 123457):
            assert np.array_equal(device_get(g, n), ora.fill(n)), (bias, n)
        jump_ms, gen_ms, nseg, seglen = g.profile_read()
        assert jump_ms > 0 and gen_ms >= 0 and nseg >= 1 and seglen >= 16 and seglen & (seglen - 1) == 0
        g.profile(False)
    g.close()


def test_user_stream_burst_and_argument_errors():
    g = G.chol.Rng64()
    with pytest.raises(G.chol.CholeskyError) as e:
        g.get(torch.empty(4, dtype=torch.float64, device="cuda").data_ptr(), 4)
    assert e.value.code == 3  # CUDA_ERROR_NOT_INITIALIZED: no default seed, as on the CPU
    s = torch.cuda.Stream()
    g.set_stream(s.cuda_stream)
    g.set(H.SEED)
    ora = H.OracleDsfmt(H.SEED)
    sizes = (50000, 3, 250001, 40000, 777777, 12)
    bufs = [torch.empty(n, dtype=torch.int64, device="cuda") for n in sizes]
    torch.cuda.synchronize()  # the allocations' fills, if any, are done before the side stream writes
    for b, n in zip(bufs, sizes):
        g.get(b.data_ptr(), n, "open_close")  # no synchronisation in between: the staging buffer is reused
    s.synchronize()
    for b, n in zip(bufs, sizes):
        assert np.array_equal(b.cpu().numpy().view(np.uint64), ora.fill(n, "open_close")), n
    x = bufs[0].data_ptr()
    assert g.get(0, 4, allow_invalid=True) == 1
    assert g.get(x + 4, 4, allow_invalid=True) == 1
    assert g.get(x, 0) == 0
    g.set_stream(0)
    g.close()
