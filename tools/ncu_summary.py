#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch captured with --set full --import-source on) into the
text committed under profiles/: headline metrics, stall breakdown, opcode mix.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout

rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k, r in enumerate(rows[2:]):
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"== launch {k}: {name}")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:70s} {r[i]:>16s} {units[i]}")

rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {s: 0 for s in stalls}
    n_inst = samples = 0
    ops = {}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            ie, sm = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        except ValueError:
            continue
        for s in stalls:
            try:
                tot[s] += int(r[ix[s]])
            except ValueError:
                pass
        n_inst += ie
        samples += sm
        t = r[ix["Source"]].split()
        op = t[0] if t else "?"
        if op.startswith("@") and len(t) > 1:
            op = t[1]
        op = op.split(".")[0]
        o = ops.setdefault(op, [0, 0])
        o[0] += ie
        o[1] += sm
    print(f"\n== source page: {n_inst} warp instructions executed, {samples} stall samples")
    print("-- warp stall reasons (share of samples)")
    for s, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
        print(f"{s:28s} {v:7d} {100 * v / max(samples, 1):5.1f}%")
    print("-- opcode mix (share of executed warp instructions / of stall samples)")
    for op, (ie, sm) in sorted(ops.items(), key=lambda x: -x[1][0])[:24]:
        print(f"{op:10s} {ie:10d} {100 * ie / max(n_inst, 1):5.1f}%   {100 * sm / max(samples, 1):5.1f}%")
