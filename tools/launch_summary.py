#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path, top=16):
    lines = [ln for ln in open(path) if ln.startswith('"')]
    tot = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0][:64]
        v = float(row["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
        t = tot.setdefault(name, [0, 0.0, 0.0])
        t[0] += 1
        t[1] += v
        t[2] = max(t[2], v)
    s = sum(v[1] for v in tot.values())
    print(f"{'kernel':66s} {'n':>5s} {'total ms':>10s} {'avg us':>10s} {'max ms':>9s} share")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k:66s} {v[0]:5d} {v[1] / 1e6:10.3f} {v[1] / v[0] / 1e3:10.2f} {v[2] / 1e6:9.3f} {100 * v[1] / s:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16)
