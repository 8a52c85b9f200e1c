#!/usr/bin/env python
"""Per-SASS-instruction executed counts, lanes and stall samples of the first kernel in an .ncu-rep, grouped into runs
of equal (executions, lanes) = basic-block regions.  usage: ncu_sass.py rep [full]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
full = len(sys.argv) > 2
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# first kernel only
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hdr_i[0]; end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
h = rows[start]
ix = {k: h.index(k) for k in ("Source", "# Samples", "Instructions Executed", "Avg. Threads Executed", "stall_long_sb", "stall_wait")}
body = [r for r in rows[start + 1:end] if len(r) > ix["stall_wait"]]
tot_s = sum(int(r[ix["# Samples"]] or 0) for r in body); tot_i = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
print(f"total samples {tot_s} warp-instr {tot_i}")
blocks = []
for n, r in enumerate(body):
    ex = int(r[ix["Instructions Executed"]] or 0); th = float(r[ix["Avg. Threads Executed"]] or 0); sm = int(r[ix["# Samples"]] or 0)
    lsb = int(r[ix["stall_long_sb"]] or 0)
    if full:
        print(f"{n:4d} {ex/1e6:7.3f}M thr={th:5.1f} samp={100*sm/tot_s:5.2f}% lsb={100*lsb/tot_s:5.2f}% {r[ix['Source']].strip()}")
    key = (round(ex / 1e4), round(th, 1))
    if blocks and blocks[-1][0] == key:
        blocks[-1][1] += 1; blocks[-1][2] += ex; blocks[-1][3] += sm; blocks[-1][4] += lsb
    else:
        blocks.append([key, 1, ex, sm, lsb, n])
if not full:
    for key, cnt, ex, sm, lsb, n in blocks:
        if ex / tot_i > 0.004 or sm / tot_s > 0.004:
            print(f"@{n:4d} {cnt:4d} instr x {key[0]/100:6.2f}M exec, lanes {key[1]:5.1f}: {100*ex/tot_i:5.1f}% of issued, {100*sm/tot_s:5.1f}% of samples (long_sb {100*lsb/tot_s:4.1f}%)")
