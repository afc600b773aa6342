#!/usr/bin/env python
"""Stall samples and executed instructions of one captured kernel by CUDA source line: joins the SASS page of an
.ncu-rep with the line table of the object file (nvdisasm -g).
usage: tools/ncu_lines.py report.ncu-rep object.o mangled_kernel_name [units_per_launch]"""
import csv, io, os, re, subprocess, sys, tempfile

rep, obj, kern = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
amap, line, inside = {}, None, False
for l in dis:
    if l.startswith(".text."):
        inside = l.strip().rstrip(":") == ".text." + kern
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
    if m and line:
        amap[int(m.group(1), 16)] = line
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
R = [r for r in rows[2:] if len(r) >= len(hdr)]
agg, tot = {}, 0.0
for k, r in enumerate(R):
    ln = amap.get(16 * k, ("?", 0))
    s, i = float(r[ix["# Samples"]] or 0), float(r[ix["Instructions Executed"]] or 0)
    g = agg.setdefault(ln, [0.0, 0.0])
    g[0] += s
    g[1] += i
    tot += s
srcs = {}
print(f"total stall samples {tot:.0f}")
for ln, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    if ln[0] not in srcs:
        cand = [os.path.join(dp, ln[0]) for dp, _, fs in os.walk(os.path.dirname(os.path.abspath(obj)) + "/..") if ln[0] in fs]
        srcs[ln[0]] = open(cand[0]).read().splitlines() if cand else []
    txt = srcs[ln[0]][ln[1] - 1].strip()[:100] if 0 < ln[1] <= len(srcs[ln[0]]) else ""
    print(f"{ln[0]:16s}:{ln[1]:4d} {100 * s / max(tot, 1):5.1f}%  inst/unit {i / units:8.1f}  {txt}")
