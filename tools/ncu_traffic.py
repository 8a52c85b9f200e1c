#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of bench.py: DRAM bytes (read + write) per launch of the three
traversal launches of one C3 step, in launch order closest_primary, any_shadow, closest_bounce.
usage: ncu_traffic.py rep label:launch_index ..."""
import csv, io, json, os, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
def col(k):
    i = h.index(k)
    unit = rows[1][i]
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return [float(r[i]) * mul for r in rows[2:]]
rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
names = [r[h.index("Kernel Name")] for r in rows[2:]]
out = {"source": os.path.basename(rep) + " (ncu --set full --clock-control none, one bench.py step)", "launches": {}, "kernels": {}}
for spec in sys.argv[2:]:
    label, idx = spec.split(":")
    i = int(idx)
    out["launches"][label] = rd[i] + wr[i]
    out["kernels"][label] = names[i].split("(")[0]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(out, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
