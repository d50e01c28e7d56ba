#!/usr/bin/env python
"""Device-vs-oracle deviation of the `constraint` contact path per env-step (development probe; the assertions live in
tests/test_gpu_parity.py).  Usage: python tools/dev/constraint_deviation.py [n_env] [steps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_common as pc
from jiminy_b200 import scenarios
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = scenarios.make("anymal", n, seed=2)
sc.options["contacts"]["model"] = "constraint"
eng, orc = pc.make_pair(sc, None)
for k in range(steps):
    act = sc.sample_targets(k)
    eng.set_command(act); orc.set_command(act)
    eng.step(sc.step_dt); orc.step(sc.step_dt, parallel=True)
    t, q, v, a = eng.get_state(); to, qo, vo, ao = orc.get_state()
    s, so = eng.get_sensors(), orc.get_sensors()
    worst = np.unravel_index(np.abs(v - vo).argmax(), v.shape)
    print(f"step {k}: q {np.abs(q - qo).max():.3e}  v {np.abs(v - vo).max():.3e} (env {worst[0]})  a {np.abs(a - ao).max():.3e}  sensors {np.abs(s - so).max():.3e}  status {eng.get_status().max()}", flush=True)
