"""Device dSFMT-19937 (cuRng64GetCloseOpen): achieved HBM write bandwidth at the BASELINE input shapes, with the two kernels
(jump, generate) timed by CUDA events on the stream they run on, after warm-up; each timed fill is larger than L2 or
preceded by one.  Prints one JSON line per shape.  RNG_SIZES = comma-separated counts of doubles (default: 65536*1024 - the
N = 65536 input panel B -, 16384^2, 32768^2)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cholesky_b200_loader import load  # noqa: E402

chol = load()
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:  # noqa: BLE001
    pass
sizes = [int(x) for x in os.environ.get("RNG_SIZES", "%d,%d,%d" % (65536 * 1024, 16384 ** 2, 32768 ** 2)).split(",")]
reps = int(os.environ.get("RNG_REPS", "5"))
g = chol.Rng64()
t = time.time()
g.set(20121018)
torch.cuda.synchronize()
print(json.dumps({"what": "cuRng64Set (first in the process: minimal polynomial + Krylov block)", "ms": 1e3 * (time.time() - t)}))
t = time.time()
g.set(20121018)
torch.cuda.synchronize()
print(json.dumps({"what": "cuRng64Set (again: exact check of the polynomial on the host + Krylov block on the device)", "ms": 1e3 * (time.time() - t)}))
for n in sizes:
    buf = torch.empty(n, dtype=torch.float64, device="cuda")
    t = time.time()
    g.set(20121018)
    g.get(buf.data_ptr(), n)
    torch.cuda.synchronize()
    first_ms = 1e3 * (time.time() - t)
    ms = []
    for _ in range(reps):
        g.set(20121018)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.get(buf.data_ptr(), n)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    med = float(np.median(ms))
    line = {"what": "cuRng64GetCloseOpen", "n": n, "bytes": 8 * n, "ms_median": med, "ms_all": ms, "GBps": 8e-6 * n / med,
            "first_call_wall_ms": first_ms, "launches_per_call": 2}
    hbm = peaks.get("hbm_gbs")
    if hbm:
        line["frac_of_measured_copy_peak"] = line["GBps"] / float(hbm)
    # spot check against the oracle so that a timing is never reported for a wrong stream
    import harness as H  # noqa: E402
    k = min(n, 1 << 20)
    line["first_1M_equal_oracle"] = bool(np.array_equal(buf[:k].cpu().numpy().view(np.uint64), H.OracleDsfmt(20121018).fill(k)))
    print(json.dumps(line))
    del buf
g.close()
