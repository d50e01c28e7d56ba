#!/usr/bin/env python
"""Per-source-line breakdown of one out-of-line device function inside the step kernel: stall samples and executed
instructions of an ncu report, attributed to the line of the function body each SASS instruction was inlined into
(nvdisasm -gi line tables matched to the SASS page of the report by instruction index).
Usage: python tools/ncu_by_line.py <report.ncu-rep> <library.so> <function-substring> [kernel-substring] [top-N]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, func = sys.argv[1], sys.argv[2], sys.argv[3]
want = sys.argv[4] if len(sys.argv) > 4 else "env_step_kernel_tILb1"
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 60
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
dis = subprocess.run(["nvdisasm", "-c", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
loc_of, idx, inside, cur_fn, group, top = [], 0, False, "<kernel body>", [], None
for line in dis:
    m = re.match(r"^(\S+):\s*$", line)
    if m:
        name = m.group(1)
        if name.startswith(".text."): inside = want in name; cur_fn = "<kernel body>"
        elif inside and name.startswith("$") and "$" in name[1:] and not name.startswith("$__internal"): cur_fn = name.split("$")[-1]
        elif inside and name.startswith("$__internal"): cur_fn = name.strip("$")
        elif not name.startswith(".L") and not name.startswith("$") and not name.startswith(".text"):
            if inside and want not in name: inside = False
        continue
    m = re.match(r'^\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        if m.group(3) is None: top = (os.path.basename(m.group(1)), int(m.group(2)))
        else: top = (os.path.basename(m.group(3)), int(m.group(4)))   # the last "inlined at" of a group is the outermost frame
        continue
    if inside and re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        loc_of.append((cur_fn, top)); idx += 1
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(collections.Counter); k = 0; tot = collections.Counter()
for row in rows[2:]:
    try: n = int(row[ix["Instructions Executed"]]); s = int(row[ix["# Samples"]])
    except Exception: continue
    fn, loc = loc_of[k] if k < len(loc_of) else ("?", None); k += 1
    tot["samples"] += s; tot["instr"] += n
    if func not in fn: continue
    a = agg[loc]; a["instr"] += n; a["samples"] += s; a["static"] += 1
    for c in stall_cols: a[c] += int(row[ix[c]] or 0)
print(f"static instructions in report {k}, in disassembly {idx}")
fs = sum(a["samples"] for a in agg.values()); fi = sum(a["instr"] for a in agg.values())
print(f"function '{func}': {100 * fs / tot['samples']:.1f}% of the kernel's samples, {100 * fi / tot['instr']:.1f}% of its instructions")
src = {}
def text(loc):
    if loc is None: return ""
    f, l = loc
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
print(f"{'line':>22s} {'static':>6s} {'samp%':>6s} {'cyc/i':>6s}  stalls | source")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:topn]:
    top3 = sorted(((a[c], c[6:]) for c in stall_cols), reverse=True)[:3]
    where = f"{loc[0]}:{loc[1]}" if loc else "?"
    print(f"{where:>22s} {a['static']:6d} {100 * a['samples'] / fs:5.1f}% {a['samples'] / max(a['static'], 1) / (fs / max(sum(x['static'] for x in agg.values()), 1)):6.2f}  " +
          " ".join(f"{t}:{100 * v / max(a['samples'], 1):.0f}%" for v, t in top3) + " | " + text(loc))
