"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        v *= 1e3
    elif unit in ("ms", "msecond"):
        v *= 1e6
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, v, r.get("Grid Size", ""), r.get("Block Size", "")))
tot = sum(v for _, v, _, _ in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, v, _, _ in rows:
    agg[n][0] += 1
    agg[n][1] += v
print(f"{len(rows)} launches, total {tot/1e6:.3f} ms (serialised, cold-cache: compare shares)")
print(f"{'kernel':60s} {'n':>5s} {'total us':>10s} {'share':>7s} {'avg us':>8s}")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:60]:60s} {c:5d} {v/1e3:10.1f} {100*v/tot:6.1f}% {v/1e3/c:8.2f}")
if len(sys.argv) > 2:
    print("\nTop individual launches:")
    for n, v, g, b in sorted(rows, key=lambda r: -r[1])[: int(sys.argv[2])]:
        print(f"{n[:50]:50s} {v/1e3:9.1f} us grid {g} block {b}")
