#!/usr/bin/env python
"""Aggregate warp-stall samples per source line from `ncu --page source --csv --print-source cuda,sass` output."""
import csv
import sys
from collections import defaultdict


def main(path, kernel_substr, top=30):
    cur_file, cur_func = None, None
    per_line = defaultdict(lambda: [0, 0, ""])
    total = 0
    hdr = None
    with open(path, newline="") as f:
        for row in csv.reader(f):
            if not row:
                continue
            if row[0] == "File Path":
                cur_file = row[1]; hdr = None; continue
            if row[0] == "Function Name":
                cur_func = row[1]; continue
            if row[0] == "Line No":
                hdr = row; continue
            if hdr is None or kernel_substr not in (cur_func or ""):
                continue
            try:
                line = int(row[0])
            except ValueError:
                continue
            d = dict(zip(hdr, row))
            try:
                samples = int(d.get("# Samples", "0") or 0)
                inst = int(d.get("Instructions Executed", "0") or 0)
            except ValueError:
                continue
            key = (cur_file.split("/")[-1], line)
            per_line[key][0] += samples
            per_line[key][1] += inst
            if not per_line[key][2]:
                per_line[key][2] = row[1].strip()[:110]
            total += samples
    print("kernel ~ %s: %d stall samples" % (kernel_substr, total))
    for (fn, ln), (s, inst, src) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%%  %-16s:%-4d inst=%-10d %s" % (100.0 * s / max(total, 1), fn, ln, inst, src))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
