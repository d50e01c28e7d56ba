#!/usr/bin/env python
"""Static SASS of one device function of the library: instruction count, opcode mix, and the same for every loop (backward
branch) inside it -- to see before a GPU run whether an inner loop spills or carries converge-and-retry collectives.
Usage: python tools/sass_loops.py <library.so> <kernel-substring> <function-substring>"""
import collections, os, re, subprocess, sys, tempfile
lib, kern, func = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
lines = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(lines) if l.endswith(":") and kern in l and func in l and l.startswith("$"))
end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith(".text") or re.match(r"^\$_Z\S*\$\S+:$", lines[i]) or re.match(r"^\$__internal\S*:$", lines[i]))
fn = lines[start:end]
isins = lambda l: re.match(r"^\s+/\*[0-9a-f]{4,}\*/", l) is not None
op = lambda l: re.sub(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?", "", l).split()[0].split(".")[0]
print(lines[start]); print("static instructions", sum(map(isins, fn)), dict(collections.Counter(op(l) for l in fn if isins(l)).most_common(14)))
labels = {m.group(1): i for i, l in enumerate(fn) for m in [re.match(r"^(\.L_x_\d+):", l)] if m}
for i, l in enumerate(fn):
    m = re.search(r"BRA.*`\((\.L_x_\d+)\)", l)
    if m and m.group(1) in labels and labels[m.group(1)] < i:
        body = [x for x in fn[labels[m.group(1)]:i] if isins(x)]
        if len(body) > 40: print(f"loop of {len(body)} instructions:", dict(collections.Counter(map(op, body)).most_common(12)))
