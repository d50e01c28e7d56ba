#!/usr/bin/env python
"""SURVEY.md 8(d) config 3 as written: ANYmal driven by raw torque actions U(-20, 20) Nm per env-step, on the ORACLE
(CPU), for the steppers the survey lists -- RK4 at dtMax in {1e-3, 5e-4, 2.5e-4}, explicit Euler at 1e-4, Dormand-Prince
-- to record which of them survive a 10 s episode (250 env-steps) and which diverge (NaN / too many failed iterations /
joints through their bounds).  The result decides the stepper of the benchmark workload."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jiminy_b200 import scenarios      # noqa: E402
from oracle.oracle import OracleBatch  # noqa: E402


def run(solver, dt_max, n_env=16, n_steps=250, contact="spring_damper", seed=0):
    sc = scenarios.make("anymal", n_env, seed=seed, solver=solver, dt_max=dt_max, contact_model=contact)
    orc = OracleBatch(sc.robot, sc.options, n_env)
    OracleBatch.use_all_cores()
    orc.set_command(np.zeros((n_env, sc.robot.nmotors)))
    assert not orc.start(sc.q0, sc.v0).any()
    rng = np.random.default_rng(seed)
    alive = np.ones(n_env, dtype=bool)
    first_fail = np.full(n_env, -1)
    bounds_hit = np.zeros(n_env, dtype=bool)
    t0 = time.time()
    for k in range(n_steps):
        orc.set_command(rng.uniform(-20.0, 20.0, size=(n_env, sc.robot.nmotors)))
        orc.step(sc.step_dt, parallel=True)
        st = orc.get_status()
        _, q, v, _ = orc.get_state()
        bad = ((st & ~8) != 0) | ~np.isfinite(q).all(axis=1) | (np.abs(v).max(axis=1) > 1e3)
        bounds_hit |= (st & 8) != 0
        newly = alive & bad
        first_fail[newly] = k
        alive &= ~bad
        if not alive.any():
            break
    return {"solver": solver, "dt_max": dt_max, "contact": contact, "n_env": n_env, "survived_10s": int(alive.sum()),
            "first_failure_step_median": float(np.median(first_fail[first_fail >= 0])) if (first_fail >= 0).any() else None,
            "envs_that_reached_a_joint_bound": int(bounds_hit.sum()), "seconds": round(time.time() - t0, 1)}


if __name__ == "__main__":
    out = []
    for contact in ("spring_damper", "constraint"):
        for solver, dt in (("runge_kutta_4", 1e-3), ("runge_kutta_4", 5e-4), ("runge_kutta_4", 2.5e-4), ("euler_explicit", 1e-4),
                           ("runge_kutta_dopri", 1e-3)):
            r = run(solver, dt, contact=contact)
            print(json.dumps(r), flush=True)
            out.append(r)
    with open(os.path.join(ROOT, "profiles", "r02_oracle_stability_sweep.json"), "w") as fh:
        json.dump(out, fh, indent=1)
