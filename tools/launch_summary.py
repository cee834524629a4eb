"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel: share of summed kernel time."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    a = agg[r[ki][:90]]
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} ms of kernel time (cold-cache, serialised)")
print(f"{'share':>7s} {'total us':>11s} {'n':>6s} {'avg us':>9s}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 45]:
    print(f"{v[1] / tot * 100:6.2f}% {v[1]:11.1f} {v[0]:6d} {v[1] / v[0]:9.1f}  {k}")
