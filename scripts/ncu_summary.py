"""ncu report -> the JSON bench.py reads for roofline.traffic:  ncu_summary.py X.ncu-rep out.json "launch description" alg_bytes
Takes the LAST captured launch of the report (the warm one)."""
import csv
import json
import subprocess
import sys

rep, out, desc, alg = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, check=True).stdout.decode()
rows = list(csv.reader(raw.splitlines()))
hdr, units, last = rows[0], rows[1], rows[-1]
ix = {h: i for i, h in enumerate(hdr)}


def val(name):
    if name not in ix:
        return None
    v, u = float(last[ix[name]].replace(",", "")), units[ix[name]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return v * scale


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
res = {"kernel": last[ix["Kernel Name"]], "launch": desc, "duration_s": val("gpu__time_duration.sum"),
       "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": (rd or 0) + (wr or 0), "algorithmic_bytes": alg,
       "tensor_pipe_active_pct": val("sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active")
       or val("sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active")
       or val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
       "registers_per_thread": val("launch__registers_per_thread"), "source": rep}
res["tensor_like_metrics"] = {h: last[ix[h]] for h in hdr if "tensor" in h and "pct" in h}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
