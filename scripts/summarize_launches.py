#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
launch count, total / mean device time and share of the captured region."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"nsecond": 1, "ns": 1, "usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6, "second": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        rows.append((name, ns, r["Grid Size"], r["Block Size"]))
    tot = sum(r[1] for r in rows)
    agg = defaultdict(lambda: [0, 0.0, None, None])
    for name, ns, g, b in rows:
        a = agg[name]
        a[0] += 1; a[1] += ns; a[2] = g; a[3] = b
    print(f"# {path}: {len(rows)} launches, {tot/1e6:.3f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':60s} {'n':>5s} {'total_ms':>10s} {'mean_us':>10s} {'share':>7s}  grid block")
    for name, (n, ns, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} {n:5d} {ns/1e6:10.3f} {ns/n/1e3:10.1f} {100*ns/tot:6.1f}%  {g} {b}")


if __name__ == "__main__":
    main(sys.argv[1])
