#!/usr/bin/env python
"""Top SASS instructions by stall samples with the dominant stall reasons (from `ncu --page source --csv --print-source sass`)."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    si = hdr.index("# Samples")
    ii = hdr.index("Instructions Executed")
    data = []
    tot_reason = {hdr[i]: 0 for i in stall_cols}
    for n, r in enumerate(rows[hi + 1:]):
        if len(r) < len(hdr):
            continue
        try:
            s = int(r[si])
        except ValueError:
            continue
        reasons = {}
        for i in stall_cols:
            try:
                v = int(r[i])
            except ValueError:
                v = 0
            if v:
                reasons[hdr[i][6:]] = v
                tot_reason[hdr[i]] += v
        data.append((s, n, r[1].strip(), int(r[ii] or 0), reasons))
    total = sum(d[0] for d in data)
    print("total samples %d; by reason: %s" % (total, ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(total, 1))
                                                                  for k, v in sorted(tot_reason.items(), key=lambda kv: -kv[1]) if v * 100 > total)))
    for s, n, src, inst, reasons in sorted(data, key=lambda d: -d[0])[:top]:
        rs = " ".join("%s:%d" % kv for kv in sorted(reasons.items(), key=lambda kv: -kv[1])[:3])
        print("%5.2f%%  #%-5d inst=%-9d %-58s %s" % (100.0 * s / max(total, 1), n, inst, src[:58], rs))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
