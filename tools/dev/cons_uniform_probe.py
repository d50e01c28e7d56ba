"""Development probe: constraint-contact ANYmal step time with identical envs (no PGS trip-count divergence)
versus the regular per-env perturbed batch."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from jiminy_b200 import core, scenarios

for uniform in (True, False):
    sc = scenarios.make("anymal", 4096, contact_model="constraint")
    if uniform:
        sc.q0[:] = sc.q0[0]
    eng = core.BatchedEngine(sc.robot, sc.options, sc.n_env)
    eng.set_pd_controller(sc.kp, sc.kd)
    eng.set_command(sc.target0)
    eng.start(sc.q0, sc.v0)
    ts = []
    for k in range(6):
        act = sc.sample_targets(k)
        if uniform:
            act[:] = act[0]
        eng.set_command(act)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.step(0.04)
        eng.get_status()
        ts.append(time.perf_counter() - t0)
    print("uniform" if uniform else "perturbed", ["%.1f ms" % (1e3 * t) for t in ts], eng.get_state()[1][:2, 2])
