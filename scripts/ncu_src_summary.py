"""Summarise the source page of an ncu report (ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > f.csv):
stall-reason totals, samples per barrier-delimited phase, hottest instructions.  usage: ncu_src_summary.py f.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = {s: 0 for s in stalls}
total, data = 0, []
for r in rows[h + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    try:
        n = int(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    total += n
    for s in stalls:
        tot[s] += int(r[ix[s]] or 0)
    data.append((n, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), {s: int(r[ix[s]] or 0) for s in stalls}))
print("total samples", total)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-28s %6d  %.1f%%" % (s, v, 100.0 * v / max(total, 1)))
bars = [i for i, d in enumerate(data) if "BAR.SYNC" in d[1]]
print("instructions", len(data), "barriers at", bars)
prev = 0
for k, b in enumerate(bars + [len(data) - 1]):
    print("  segment %2d (instr %5d..%5d): %6d samples, %7d warp-instr" % (k, prev, b, sum(d[0] for d in data[prev:b + 1]), sum(d[2] for d in data[prev:b + 1])))
    prev = b + 1
print("hottest instructions:")
for n, src, ie, st in sorted(data, key=lambda d: -d[0])[:top]:
    why = ",".join("%s=%d" % (k[6:], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v)
    print("  %6d  exec %6d  %-70s %s" % (n, ie, src[:70], why))
