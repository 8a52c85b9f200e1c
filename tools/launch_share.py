#!/usr/bin/env python
"""Per-kernel share of a run from an ncu launch list: `ncu --metrics gpu__time_duration.sum --clock-control none --csv
--log-file launches.csv <command>; python tools/launch_share.py launches.csv [first_kernel_substring]`.  One row per kernel
name: launches, total / mean duration, share of the summed kernel time.  With a second argument the table starts at the first
launch whose name contains it (skips scene upload / BVH repack kernels).  ncu serialises launches and replays them with cold
caches, so the SHARES are what this table is for, not the absolute times."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, start=None):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    h = rows[0]
    name_i, val_i, unit_i, metric_i = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Metric Name")
    to_us = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    agg, started = OrderedDict(), start is None
    for r in rows[1:]:
        if r[metric_i] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[name_i]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        if not started:
            if start not in name:
                continue
            started = True
        us = float(r[val_i].replace(",", "")) * to_us[r[unit_i]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"{'kernel':60s} {'launches':>8s} {'total us':>10s} {'mean us':>9s} {'share':>7s}")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[-60:]:60s} {n:8d} {us:10.1f} {us / n:9.1f} {100.0 * us / total:6.2f}%")
    print(f"{'total':60s} {sum(a[0] for a in agg.values()):8d} {total:10.1f}")


if __name__ == "__main__":
    main(*sys.argv[1:3])
