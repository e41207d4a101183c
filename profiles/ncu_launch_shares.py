#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python profiles/ncu_launch_shares.py launches.csv"""
import collections
import csv
import re
import sys


def main():
    rows = []
    with open(sys.argv[1]) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            name = re.sub(r"\(.*$", "", r["Kernel Name"]).strip()
            rows.append((name, float(r["Metric Value"]) / 1e3))
    agg = collections.defaultdict(list)
    for n, us in rows:
        agg[n].append(us)
    tot = sum(us for _, us in rows)
    print("kernel | launches | mean us | share of all listed time")
    for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%s | %d | %.1f | %.1f%%" % (n, len(v), sum(v) / len(v), 100 * sum(v) / tot))


if __name__ == "__main__":
    main()
