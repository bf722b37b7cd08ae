"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import collections
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"].split("(")[0][:70], v * scale))
rows = rows[skip:]
tot = sum(v for _, v in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in rows:
    agg[k][0] += 1
    agg[k][1] += v
print(f"{len(rows)} launches, {tot/1e3:.3f} ms total")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v/tot*100:6.2f}%  {v/1e3:9.3f} ms  {n:5d} x {v/n:9.1f} us  {k}")
