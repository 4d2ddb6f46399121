#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict


def main(path, skip=0):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
    rows = rows[skip:]
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns, grid, block in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        agg[short][0] += 1
        agg[short][1] += ns
    total = sum(v[1] for v in agg.values())
    print("launches: %d, total device time %.3f ms (per-launch times under ncu are cold-cache and serialised)" % (len(rows), total / 1e6))
    print("%-70s %8s %12s %8s %10s" % ("kernel", "count", "total_us", "share", "avg_us"))
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %8d %12.1f %7.2f%% %10.2f" % (k[:70], n, ns / 1e3, 100 * ns / total, ns / 1e3 / n))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
