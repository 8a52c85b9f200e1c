#!/usr/bin/env python
"""profiles/traffic.json["path"][key] from an ncu launch list of one path-traced batch captured with
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X.csv \
      python tools/prof_path.py --scene c4 --spp 8
usage: ncu_path_traffic.py X.csv key camera_samples [first_kernel_substring]
Sums DRAM bytes (read + write) over every launch from the first k_raygen on; kernel_share = each kernel's share of the summed
kernel time (ncu serialises launches and replays them cold, so the SHARES are what is meaningful, and the DRAM byte counts —
which do not depend on timing — are what bench.py scales to the timed frame)."""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

path, key, camera_samples = sys.argv[1], sys.argv[2], int(sys.argv[3])
start = sys.argv[4] if len(sys.argv) > 4 else "k_raygen"
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
h = rows[0]
idc, name_i, metric_i, val_i, unit_i = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
to_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
to_us = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
launch = OrderedDict()
for r in rows[1:]:
    e = launch.setdefault(r[idc], {"name": re.sub(r"\(.*", "", r[name_i]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", ""),
                                   "bytes": 0.0, "us": 0.0})
    v = float(r[val_i].replace(",", ""))
    if r[metric_i].startswith("dram__bytes"):
        e["bytes"] += v * to_b[r[unit_i]]
    elif r[metric_i] == "gpu__time_duration.sum":
        e["us"] += v * to_us[r[unit_i]]
started, tot_b, tot_us, per = False, 0.0, 0.0, OrderedDict()
for e in launch.values():
    if not started:
        if start not in e["name"]:
            continue
        started = True
    tot_b += e["bytes"]
    tot_us += e["us"]
    p = per.setdefault(e["name"].split("::")[-1], [0.0, 0.0, 0])
    p[0] += e["bytes"]; p[1] += e["us"]; p[2] += 1
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tj_path = os.path.join(root, "profiles", "traffic.json")
tj = json.load(open(tj_path))
tj.setdefault("path", {})[key] = {
    "dram_bytes": tot_b, "camera_samples": camera_samples, "dram_bytes_per_camera_sample": tot_b / camera_samples,
    "source": os.path.basename(path) + " (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one batch)",
    "kernel_share": {k: {"launches": v[2], "time_share": v[1] / tot_us, "dram_bytes": v[0]} for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])}}
json.dump(tj, open(tj_path, "w"), indent=1)
print(json.dumps(tj["path"][key], indent=1)[:1500])
