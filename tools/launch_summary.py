#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total and mean time, share."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        us = v / 1e3 if r[iu] in ("nsecond", "ns") else v * 1e3 if r[iu] in ("msecond", "ms") else v
        name = re.sub(r"\(.*$", "", r[ik].replace("sb200::", "").replace("<unnamed>::", "").replace("(int)", "").replace("void ", ""))
        agg[name][0] += 1
        agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    print("kernel,launches,total_us,mean_us,share")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k},{n},{t:.1f},{t / n:.1f},{t / tot:.4f}")


if __name__ == "__main__":
    main(sys.argv[1])
