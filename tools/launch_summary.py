#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  usage: launch_summary.py launches.csv [skip_first_n]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[hdr + 1 + skip:]:
        if len(r) > vi:
            try:
                d[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
    tot = sum(sum(v) for v in d.values())
    print(f"{'kernel':40s} {'n':>6s} {'sum ms':>9s} {'share':>6s} {'median us':>10s} {'min us':>8s} {'max us':>10s}")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        s = sorted(v)
        print(f"{k:40s} {len(v):6d} {sum(v) / 1e6:9.3f} {sum(v) / tot:6.1%} {s[len(v) // 2] / 1e3:10.2f} {s[0] / 1e3:8.2f} {s[-1] / 1e3:10.2f}")
    print(f"total {tot / 1e6:.3f} ms over {sum(len(v) for v in d.values())} launches")


if __name__ == "__main__":
    main()
