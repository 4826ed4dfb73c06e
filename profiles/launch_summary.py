"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, mean time and share per kernel.
usage: python profiles/launch_summary.py launches.csv [max_name_chars]"""
import csv, sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
width = int(sys.argv[2]) if len(sys.argv) > 2 else 150
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    ns = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[ui], 1.0)
    a = agg[r[ki]]
    a[0] += 1
    a[1] += ns
total = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {total / 1e6:.2f} ms of device time")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    own = "*" if "kz_" in name else " "
    print(f"{own}{n:5d} launches  avg {ns / n / 1e3:9.2f} us  share {100 * ns / total:5.1f}%  {name[:width]}")
own = sum(a[1] for k, a in agg.items() if "kz_" in k)
print(f"own kernels (*, kz_*): {100 * own / total:.1f}% of the device time")
