#!/usr/bin/env python
"""How many projected Gauss-Seidel iterations the reference algorithm needs per solve on a workload, and what a warp of
32/L envs pays when its envs stay in the sweep until the last one has converged (oracle = CPU restatement; diagnostics).
Usage: python tools/pgs_iteration_stats.py [workload] [n_env] [steps]"""
import ctypes, sys
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from jiminy_b200 import scenarios
from oracle import oracle as O
name = sys.argv[1] if len(sys.argv) > 1 else "anymal"
n_env = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
sc = scenarios.make(name, n_env, contact_model="constraint")
orc = O.OracleBatch(sc.robot, sc.options, n_env)
if sc.kp is not None: orc.set_pd_controller(sc.kp, sc.kd)
orc.set_command(sc.target0)
assert not orc.start(sc.q0, sc.v0).any()
L = O.lib()
L.orc_pgs_history_enable.argtypes = [ctypes.c_void_p, ctypes.c_int32]
L.orc_pgs_history.restype = ctypes.c_int64
L.orc_pgs_history.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64]
L.orc_pgs_history_enable(orc._h, 1)
for k in range(steps):
    orc.set_command(sc.sample_targets(k)); orc.step(sc.step_dt)
H = []
for e in range(n_env):
    n = L.orc_pgs_history(orc._h, e, None, 0)
    buf = np.zeros(n, np.int32); L.orc_pgs_history(orc._h, e, buf.ctypes.data, n); H.append(buf)
m = min(len(h) for h in H); H = np.stack([h[:m] for h in H])
print(f"{name}: {m} solves per env over {steps} env-steps; iterations per solve: mean {H.mean():.1f}, median {np.median(H):.0f}, p90 {np.percentile(H, 90):.0f}, max {H.max()}")
for g in (4, 8, 16, 32):
    if n_env % g == 0:
        print(f"  max over groups of {g:2d} envs: mean {H.reshape(n_env // g, g, m).max(axis=1).mean():.1f}")
print("  histogram (iterations: share):", {int(k): round(float(v), 3) for k, v in zip(*np.unique(np.minimum(H, 100) // 10 * 10, return_counts=True)) for v in [v / H.size]})
