"""One process, one GPU call: the device generator's GPU parity tests (pytest in-process), then its timing over segment-length
bias (RNG_BIASES) x request size (RNG_SIZES) with the jump / generation split from the library's own events (cuRng64Profile).
Every timed configuration is also compared with the oracle on its first 2^20 values.  Prints one JSON line per configuration.

GPU calls 20 - 22 of round 2 ran earlier forms of this script that also switched between kernel variants through a test hook
(generation: read-back stores / register stores / four segments per warp; jump: 16 / 32 / 48 / 64 segments per CTA; pipelined
pieces on a second stream) - commits 612ab67, 6f93a39 and the one before this note; the records are profiles/r02_call2{0,1,2}_summary.txt
(fields "gen", "jump", "chunks").  The variants that lost were removed together with the hook."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
t_start = time.time()
import numpy as np  # noqa: E402
import pytest  # noqa: E402
import torch  # noqa: E402

import harness as H  # noqa: E402
from cholesky_b200_loader import load  # noqa: E402

chol = load()
rc = 0
if not os.environ.get("RNG_FINAL_SKIP_TESTS"):
    rc = int(pytest.main(["-q", "-x", "-p", "no:cacheprovider", os.path.join(ROOT, "tests", "test_zz_gpu_rng.py")]))
    print(json.dumps({"what": "pytest tests/test_zz_gpu_rng.py", "rc": rc, "elapsed_s": time.time() - t_start}), flush=True)

deadline = t_start + float(os.environ.get("RNG_FINAL_BUDGET_S", "40"))
sizes = [int(x) for x in os.environ.get("RNG_SIZES", "%d,%d" % (1 << 30, 1 << 26)).split(",")]
want = H.OracleDsfmt(H.SEED).fill(1 << 20)
g = chol.Rng64()
g.set(H.SEED)
g.profile(True)
buf = torch.empty(max(sizes), dtype=torch.float64, device="cuda")
biases = [int(v) for v in os.environ.get("RNG_BIASES", "0,1,-1").split(",")]
for n in sizes:
    for bias in biases:
        if time.time() > deadline:
            print(json.dumps({"what": "time budget reached", "n": n, "bias": bias}), flush=True)
            continue
        g.set_segment_bias(bias)
        g.set(H.SEED)
        g.get(buf.data_ptr(), n)  # first call of this shape: polynomial rows computed on the host and uploaded
        best = None
        for _ in range(2):
            g.set(H.SEED)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.get(buf.data_ptr(), n)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            jm, gm, nseg, seglen = g.profile_read()
            if best is None or ms < best["ms"]:
                best = {"ms": ms, "jump_ms": jm, "generate_ms": gm, "segments": nseg, "segment_refills": seglen}
        k = min(n, 1 << 20)
        ok = bool(np.array_equal(buf[:k].cpu().numpy().view(np.uint64), want[:k]))
        best.update({"what": "cuRng64GetCloseOpen", "n": n, "bias": bias, "GBps": 8e-6 * n / best["ms"],
                     "frac_of_hbm_peak_6550": 8e-6 * n / best["ms"] / 6550.4, "equal_oracle_first_1M": ok})
        print(json.dumps(best), flush=True)
        if not ok:
            rc = rc or 3
g.close()
sys.exit(rc)
