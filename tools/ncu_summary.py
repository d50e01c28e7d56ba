#!/usr/bin/env python
"""Summarise an ncu report of the step kernel: headline metrics, stall mix, opcode mix, hot source lines.
Usage: python tools/ncu_summary.py gpurun_out/<tag>/prof_step.ncu-rep [n_hot_lines]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, r = rows[0], rows[1], rows[2]
m = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]
for k in keys:
    if k in m: print(f"{k:70s} {m[k][0]:>16s} {m[k][1]}")
print("-- stall cycles per issued instruction")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        v = float(m[h][0])
        if v > 0.02: print(f"   {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:22s} {v:6.2f}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
ops, samp = collections.Counter(), collections.Counter(); tot = tots = 0
for row in rows[2:]:
    try: n = int(row[ix["Instructions Executed"]]); s = int(row[ix["# Samples"]])
    except Exception: continue
    t = row[ix["Source"]].split(); op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += n; samp[op] += s; tot += n; tots += s
print(f"-- opcode mix: {tot} warp instructions, {len(rows) - 2} static")
for op, n in ops.most_common(16): print(f"   {op:10s} {100 * n / tot:5.1f}% of instr  {100 * samp[op] / tots:5.1f}% of samples")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
# multiple files: sections start with "File Name"
cur = None; lines = []
for row in csv.reader(io.StringIO(src)):
    if len(row) >= 2 and row[0] == "File Name": cur = row[1].split("/")[-1]; hdr2 = None; continue
    if row and row[0] == "Line No": hdr2 = {h: i for i, h in enumerate(row)}; continue
    if cur and hdr2 and len(row) > 5:
        try: lines.append((int(row[hdr2["# Samples"]]), int(row[hdr2["Instructions Executed"]]), cur, row[hdr2["Line No"]], row[hdr2["Source"]].strip()[:110]))
        except Exception: pass
tots = sum(l[0] for l in lines) or 1
print("-- hottest source lines (samples %, instr)")
for s, n, f, ln, text in sorted(lines, reverse=True)[:nhot]: print(f"   {100 * s / tots:5.1f}% {n:>11d}  {f}:{ln}  {text}")
