#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].split("<")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| {k} | {c} | {t/1e3:.1f} | {t/tot:.1%} | {t/c/1e3:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
