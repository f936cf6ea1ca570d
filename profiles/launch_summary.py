#!/usr/bin/env python
"""ncu launch list (`--metrics gpu__time_duration.sum --csv`) -> per-kernel summary CSV:
kernel, launches, total_ms, avg_us, share_pct.   usage: launch_summary.py launches.csv > summary.csv"""
import collections
import csv
import re
import sys


def main(path):
    tot, cnt = collections.Counter(), collections.Counter()
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    head = rows[0]
    k_name, k_metric, k_unit, k_val = (head.index(x) for x in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    for r in rows[1:]:
        if r[k_metric] != "gpu__time_duration.sum":
            continue
        ns = float(r[k_val].replace(",", "")) * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0}[r[k_unit]]
        name = re.sub(r"\(.*", "", r[k_name])[:60]
        tot[name] += ns
        cnt[name] += 1
    all_ns = sum(tot.values())
    out = csv.writer(sys.stdout, lineterminator="\n")
    out.writerow(["kernel", "launches", "total_ms", "avg_us", "share_pct"])
    for name, ns in tot.most_common():
        out.writerow([name, cnt[name], "%.3f" % (ns / 1e6), "%.1f" % (ns / cnt[name] / 1e3), "%.1f" % (100.0 * ns / all_ns)])


if __name__ == "__main__":
    main(sys.argv[1])
