#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` capture of the step kernel: per-launch DRAM traffic and pipe activity,
stamped with the hash of the device code they were measured on (bench.py reports them only while that hash matches).
Usage: python tools/ncu_traffic.py <report.ncu-rep> <workload> <n_env> <summary file under profiles/>"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
rep, workload, n_env, source = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, r = rows[0], rows[1], rows[2]
m = {h: (float(r[i].replace(",", "")) if r[i].replace(",", "").replace(".", "", 1).replace("e+", "").replace("-", "").isdigit() else r[i], units[i]) for i, h in enumerate(hdr)}
def to_bytes(key):
    v, u = m[key]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d[workload] = {"n_env": n_env, "traffic_bytes": int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")),
               "fp64_pipe_active_pct": round(m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0], 2),
               "issue_active_pct": round(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0], 2),
               "kernel_ms_under_ncu": round(m["gpu__time_duration.sum"][0] * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[m["gpu__time_duration.sum"][1]], 4),
               "kernel_source_sha": bench.kernel_source_sha(), "source": source}
json.dump(d, open(path, "w"), indent=1)
print(json.dumps(d[workload]))
