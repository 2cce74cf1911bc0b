#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (count, total, share)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-58s %6s %12s %8s %7s" % ("kernel", "count", "total_us", "avg_us", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-58s %6d %12.1f %8.1f %6.1f%%" % (n[:58], c, t, t / c, 100 * t / tot))
    print("%-58s %6d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))


if __name__ == "__main__":
    main(sys.argv[1])
