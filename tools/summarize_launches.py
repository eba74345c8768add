"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<.*", "", name)
    rows.append((name, ns))
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e6:.3f} ms total (serialised, cold-cache: compare shares)")
print(f"{'kernel':60s} {'count':>6s} {'ms':>9s} {'share':>7s} {'avg us':>9s}")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{n[:60]:60s} {c:6d} {ns / 1e6:9.3f} {100 * ns / tot:6.1f}% {ns / c / 1e3:9.1f}")
