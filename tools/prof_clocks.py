#!/usr/bin/env python
"""Cycle accounting of the full body (development build, tools/build_prof.sh): where a warp spends its cycles on the
`constraint` contact model -- sweeps, bound update, solver set-up, PGS sweep, refresh -- and how often its envs were in a
solve together.  Usage: python tools/prof_clocks.py [workload] [n_env] [steps]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jiminy_b200 import scenarios, core
name = sys.argv[1] if len(sys.argv) > 1 else "anymal"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
api = core.Api(C.CDLL(os.path.join(ROOT, "jiminy_b200", "libjiminy_b200_prof.so")))
api.dll.jb_debug_prof.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
sc = scenarios.make(name, n, contact_model="constraint")
eng = core.BatchedEngine(sc.robot, sc.options, n, api_=api)
if sc.kp is not None: eng.set_pd_controller(sc.kp, sc.kd)
eng.set_command(sc.target0); eng.start(sc.q0, sc.v0)
out = (C.c_double * 16)()
for k in range(2):
    eng.set_command(sc.sample_targets(k)); eng.step(sc.step_dt)
api.dll.jb_debug_prof(eng._h, out)          # clear
for k in range(steps):
    eng.set_command(sc.sample_targets(2 + k)); eng.step(sc.step_dt)
api.dll.jb_debug_prof(eng._h, out)
p = np.array(out[:]); warps = p[7] / steps
per = lambda i: p[i] / p[7]                    # per warp and launch
print(f"{name} x {n}, contacts.model = constraint: per warp and env-step ({int(warps)} warps, {steps} steps)")
print(f"  kernel                {per(6) / 1e6:8.2f} M cycles")
print(f"  sweeps (rhs)          {per(0) / 1e6:8.2f} M   bound update {per(1) / 1e6:6.2f} M   votes + solver {per(2) / 1e6:6.2f} M")
print(f"  solver: set-up        {per(3) / 1e6:8.2f} M   sweep loop   {per(4) / 1e6:6.2f} M   refresh        {per(5) / 1e6:6.2f} M")
print(f"  rhs() calls           {per(8):8.1f}     lanes present per call {p[9] / max(p[8], 1):5.1f}")
print(f"  solves: whole warp    {per(10):8.1f}     partial warp {per(11):8.1f}     sweep iterations {per(12):9.1f}  ({per(12) / max(per(10) + per(11), 1):.1f} per solve)")
it = max(per(12), 1)
print(f"  cycles per sweep iteration {per(4) / it:8.0f}   per set-up {per(3) / max(per(10) + per(11), 1):8.0f}   per rhs sweeps {per(0) / max(per(8), 1):8.0f}")
print(f"  inside the sweep, per iteration: normal-force loop {p[13] / max(p[12], 1):7.0f}   friction loop {p[14] / max(p[12], 1):7.0f}   stopping criterion {p[15] / max(p[12], 1):7.0f}")

