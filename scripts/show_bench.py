"""Print the key numbers of bench.py JSON lines found in the given files."""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        r = d.get("roofline", {})
        print("%s: gpus %s nb %s value %.0f GF/s ms/step %.1f | update %.2f TF/s/rank frac %.3f share %.3f | e2e %.0f | launches %s" % (
            path, d.get("n_gpus"), d.get("config", {}).get("nb"), d["value"], d["ms_per_step"], r.get("achieved") or 0,
            r.get("frac") or 0, r.get("share_of_step") or 0, d.get("e2e", {}).get("value", 0), d.get("gpu_launches")))
