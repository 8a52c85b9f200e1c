#!/usr/bin/env python
"""Executed warp instructions and stall samples of one kernel per SOURCE LINE and per enclosing function.
The `ncu --page source --csv` export carries SASS rows without line numbers; this joins them (by instruction offset) with
`nvdisasm -g` of the object the kernel was compiled into (built with -lineinfo).
usage: ncu_lines.py SASS_CSV[.gz] OBJECT KERNEL_SUBSTRING [top_n]"""
import csv, gzip, io, os, re, subprocess, sys, tempfile, collections

csv_path, obj, ksub = sys.argv[1:4]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
op = gzip.open if csv_path.endswith(".gz") else open
rows = list(csv.reader(io.TextIOWrapper(op(csv_path, "rb"))))
h = rows[1]
body = [r for r in rows[2:] if len(r) > 40]
A, S, E, N, T = (h.index(k) for k in ("Address", "Source", "Instructions Executed", "# Samples", "Avg. Threads Executed"))
NOI = h.index("stall_no_inst"); LSB = h.index("stall_long_sb")
base = int(body[0][A], 16)
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
# the kernel's section
line_of = {}
inside = False; cur = None
for l in dis:
    if l.startswith("//--------------------- .text."):
        inside = ksub in l
        continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m: line_of[int(m.group(1), 16)] = cur
assert line_of, "kernel not found in the object"
src_cache = {}
def func_of(fl):
    if fl is None: return "?"
    f, ln = fl
    if f not in src_cache:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), f)
        src_cache[f] = open(p).read().split("\n") if os.path.exists(p) else []
    src = src_cache[f]
    for i in range(min(ln, len(src)) - 1, -1, -1):
        s = src[i]
        if s and not s[0].isspace() and not s.startswith(("//", "#", "}", "template", "{")) and "(" in s:
            m = re.search(r"([A-Za-z_][A-Za-z_0-9<>:]*)\s*\(", s)
            return f"{f}:{m.group(1) if m else s[:30]}"
    return f
by_line = collections.defaultdict(lambda: [0, 0, 0, 0, 0]); by_fn = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
tot_e = tot_s = 0
for r in body:
    off = int(r[A], 16) - base
    fl = line_of.get(off)
    e, s, ni, lsb = int(r[E] or 0), int(r[N] or 0), int(r[NOI] or 0), int(r[LSB] or 0)
    tot_e += e; tot_s += s
    for d, k in ((by_line, fl), (by_fn, func_of(fl))):
        d[k][0] += e; d[k][1] += s; d[k][2] += 1; d[k][3] += ni; d[k][4] += lsb
print(f"{len(body)} SASS instructions, {tot_e/1e6:.1f} M warp instructions executed, {tot_s} stall samples")
print("== by function (static SASS count, share of executed warp instructions, share of stall samples [no_inst, long_sb]) ==")
for k, v in sorted(by_fn.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"{k:55s} sass {v[2]:6d}  exec {100*v[0]/tot_e:5.1f}%  samples {100*v[1]/tot_s:5.1f}% [{100*v[3]/tot_s:4.1f} {100*v[4]/tot_s:4.1f}]")
print("== by line ==")
for k, v in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"{str(k):40s} sass {v[2]:5d}  exec {100*v[0]/tot_e:5.1f}%  samples {100*v[1]/tot_s:5.1f}% [{100*v[3]/tot_s:4.1f} {100*v[4]/tot_s:4.1f}]")
