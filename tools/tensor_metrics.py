"""Summarise an ncu --csv capture of (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed) per kernel family -> JSON on stdout."""
import csv, json, re, sys
from collections import defaultdict

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
per = defaultdict(dict)   # launch id -> metric -> value
name = {}
for r in rows:
    v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
    u = r.get("Metric Unit", "")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    per[r["ID"]][r["Metric Name"]] = v * scale
    name[r["ID"]] = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("mmdit::", "")
fam = defaultdict(lambda: dict(launches=0, total_ms=0.0, dram_read_GB=0.0, dram_write_GB=0.0, tw=0.0))
for i, m in per.items():
    n = re.sub(r"<.*", "", name[i])
    f = fam[n]
    ms = m.get("gpu__time_duration.sum", 0.0)
    f["launches"] += 1
    f["total_ms"] += ms
    f["dram_read_GB"] += m.get("dram__bytes_read.sum", 0.0) / 1e9
    f["dram_write_GB"] += m.get("dram__bytes_write.sum", 0.0) / 1e9
    f["tw"] += ms * next((v for k, v in m.items() if k.startswith("sm__pipe_tensor")), 0.0)
out = {}
for n, f in sorted(fam.items(), key=lambda kv: -kv[1]["total_ms"]):
    if f["total_ms"] < 0.05:
        continue
    b = (f["dram_read_GB"] + f["dram_write_GB"]) * 1e9
    out[n] = dict(launches=f["launches"], total_ms=round(f["total_ms"], 3), dram_read_GB=round(f["dram_read_GB"], 3),
                  dram_write_GB=round(f["dram_write_GB"], 3), dram_bytes_per_launch=round(b / f["launches"]),
                  dram_GBps=round(b / (f["total_ms"] * 1e-3) / 1e9, 1),
                  tensor_pipe_active_pct_time_weighted=round(f["tw"] / f["total_ms"], 2))
print(json.dumps(out, indent=1))
