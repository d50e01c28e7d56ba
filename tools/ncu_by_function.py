#!/usr/bin/env python
"""Per-device-function breakdown of an ncu report of the step kernel: executed warp instructions, stall samples and
cycles per issued instruction of every out-of-line function inside the kernel (nvdisasm labels matched to the SASS
page of the report by instruction index).
Usage: python tools/ncu_by_function.py <report.ncu-rep> <library.so> [kernel-substring]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else "env_step_kernel_tILb1"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
dis = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# instruction index -> function label, inside the wanted kernel
labels, idx, inside = [], 0, False
for line in dis:
    m = re.match(r"^(\S+):\s*$", line)
    if m:
        name = m.group(1)
        if name.startswith(".text."):
            inside = want in name
            if inside: labels.append((idx, "<kernel body>"))
        elif inside and name.startswith("$") and "$" in name[1:]:
            labels.append((idx, name.split("$")[-1]))
        elif inside and name.startswith("$__internal"):
            labels.append((idx, name.strip("$")))
        elif not name.startswith(".L") and not name.startswith("$") and not name.startswith(".text"):
            if inside and want not in name: inside = False
        continue
    if inside and re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+\S", line): idx += 1
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
k, cur = 0, "<kernel body>"; bounds = labels + [(10**9, None)]; b = 0
for row in rows[2:]:
    try: n = int(row[ix["Instructions Executed"]]); s = int(row[ix["# Samples"]])
    except Exception: continue
    while b + 1 < len(bounds) and k >= bounds[b + 1][0]: b += 1
    cur = bounds[b][1]
    a = agg[cur]; a["instr"] += n; a["samples"] += s; a["static"] += 1
    for c in stall_cols: a[c] += int(row[ix[c]] or 0)
    k += 1
print(f"static instructions in report {k}, in disassembly {idx}")
tot_i = sum(a["instr"] for a in agg.values()); tot_s = sum(a["samples"] for a in agg.values())
print(f"{'function':60s} {'static':>7s} {'instr%':>7s} {'samp%':>7s} {'cyc/instr':>9s}  top stalls")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    if a["samples"] < 0.002 * tot_s: continue
    rel = (a["samples"] / tot_s) / max(a["instr"] / tot_i, 1e-12)
    top = sorted(((a[c], c[6:]) for c in stall_cols), reverse=True)[:4]
    print(f"{name[:60]:60s} {a['static']:7d} {100 * a['instr'] / tot_i:6.1f}% {100 * a['samples'] / tot_s:6.1f}% {rel:9.2f}  " +
          " ".join(f"{t}:{100 * v / max(a['samples'], 1):.0f}%" for v, t in top))
