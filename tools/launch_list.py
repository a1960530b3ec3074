"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, average duration and share per kernel.
usage: python tools/launch_list.py gpurun_out/launches_bench.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
head = rows[0]
ki, vi, ui = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
acc = defaultdict(lambda: [0, 0.0])
mi = head.index("Metric Name") if "Metric Name" in head else None
for r in rows[1:]:
    if mi is not None and r[mi] != "gpu__time_duration.sum":
        continue  # lists captured with more than one metric
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    acc[r[ki]][0] += 1
    acc[r[ki]][1] += v
tot = sum(v[1] for v in acc.values())
print(f"{'kernel':90s} {'launches':>8s} {'avg_us':>9s} {'share':>7s}")
for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:90]:90s} {n:8d} {t / n:9.2f} {100 * t / tot:6.1f}%")
