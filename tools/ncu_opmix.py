#!/usr/bin/env python
"""fp32 flop a kernel EXECUTES per input sample, from the opcode mix of an ncu capture (SASS page):
packed FFMA2 = 4 flop per lane, FADD2 / FMUL2 = 2, scalar FFMA = 2, FADD / FMUL = 1; x 32 lanes / samples per launch.
usage: tools/ncu_opmix.py workload report.ncu-rep samples_per_launch [workload report samples ...]   (merges into
profiles/ncu_opmix.json, which bench.py reads for roofline.fp32_frac_executed)"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "profiles", "ncu_opmix.json")
FLOP = {"FFMA2": 4, "FADD2": 2, "FMUL2": 2, "FFMA": 2, "FADD": 1, "FMUL": 1}

out = json.load(open(PATH)) if os.path.exists(PATH) else {}
args = sys.argv[1:]
for k in range(0, len(args), 3):
    name, rep, samples = args[k], args[k + 1], float(args[k + 2])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    mix, total = {}, 0.0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        toks = r[ix["Source"]].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        ie = float(r[ix["Instructions Executed"]] or 0)
        mix[op] = mix.get(op, 0.0) + ie
        total += ie
    flop = sum(FLOP[o] * c for o, c in mix.items() if o in FLOP) * 32.0
    packed = sum(mix.get(o, 0.0) for o in ("FFMA2", "FADD2", "FMUL2"))
    out[name] = {"kernel": kernel, "report": os.path.basename(rep), "samples_per_launch": samples,
                 "flop_per_sample_executed": flop / samples, "warp_instructions_per_sample": total / samples,
                 "packed_fp32_warp_instructions_per_sample": packed / samples,
                 "opcodes": {o: c for o, c in sorted(mix.items(), key=lambda kv: -kv[1])[:12]}}
    print(name, kernel[:60], "flop/sample", round(flop / samples, 2), "warp instr/sample", round(total / samples, 3))
json.dump(out, open(PATH, "w"), indent=1)
