#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per kernel name: launches, average duration, share, DRAM MB per launch.  Usage: summarise_launches.py file.csv [skip_first_n]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    lines = [l for l in open(path) if not l.startswith("==")]
    recs = collections.OrderedDict()
    for row in csv.DictReader(lines):
        recs.setdefault(int(row["ID"]), {"name": row["Kernel Name"]})[row["Metric Name"]] = (row["Metric Value"], row["Metric Unit"])
    unit = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    agg = collections.OrderedDict()
    for i, m in recs.items():
        if i < skip:
            continue
        name = re.sub(r"\(.*", "", m["name"]).replace("void ", "").replace("<unnamed>::", "")
        get = lambda k: float(m[k][0].replace(",", "")) * unit[m[k][1]]
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += get("gpu__time_duration.sum"); a[2] += get("dram__bytes_read.sum"); a[3] += get("dram__bytes_write.sum")
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | avg us | share | DRAM read MB/launch | DRAM write MB/launch |")
    print("|---|---:|---:|---:|---:|---:|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.2f | %.1f%% | %.2f | %.2f |" % (name, a[0], a[1] / a[0], 100 * a[1] / total, a[2] / a[0], a[3] / a[0]))
    print("\ntotal %.1f us over %d launches" % (total, sum(a[0] for a in agg.values())))


if __name__ == "__main__":
    main()
