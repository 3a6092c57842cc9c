"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <program>): per-kernel counts and
serialised time of ONE step of bench.py (from the last restore copy of the input to the end), and the dominant kernel's share.
usage: launch_list_summary.py X.csv [marker-substring-of-the-step-boundary-kernel]"""
import collections
import csv
import re
import sys

f = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "direct_copy"  # torch's A.copy_(A0) that opens every step
rows = list(csv.DictReader([l for l in open(f) if not l.startswith("==")]))
names = [r["Kernel Name"] for r in rows]
ours = lambda n: any(k in n for k in ("chol::", "gemm_dmma", "potf2", "trsm_rows", "dgemm_small", "copy_kernel", "fill_kernel", "trti2", "tf32"))
starts = [i for i, n in enumerate(names) if marker in n and i + 1 < len(names) and ours(names[i + 1])]
if not starts:
    starts = [0]
a = starts[-2] + 1 if len(starts) >= 2 else starts[-1] + 1
b = starts[-1] if len(starts) >= 2 else len(rows)
sel = [r for r in rows[a:b] if ours(r["Kernel Name"])]
agg, tot = collections.OrderedDict(), 0.0
for r in sel:
    n = re.sub(r"\(.*", "", r["Kernel Name"])
    t = float(r["Metric Value"].replace(",", "")) / 1e3
    e = agg.setdefault(n, [0, 0.0])
    e[0] += 1
    e[1] += t
    tot += t
print("%s: one step = launches %d..%d, %d library kernels, serialised %.1f ms" % (f, a, b, len(sel), tot / 1e3))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-64s %6d %12.1f us  avg %9.1f  share %.4f" % (n[:64], c, t, t / c, t / tot))
# the trailing updates are the large gemm_dmma launches (one per step of the factorisation): everything above 2 ms here
big = sum(float(r["Metric Value"].replace(",", "")) / 1e3 for r in sel if "gemm_dmma" in r["Kernel Name"] and float(r["Metric Value"].replace(",", "")) > 2e6)
print("  trailing updates (gemm_dmma launches > 2 ms): %.1f ms = %.4f of the serialised step" % (big / 1e3, big / tot))
