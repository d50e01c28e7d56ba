"""Long-horizon device-vs-oracle probe under the warp emulator (not part of the test suite: minutes of CPU).

Drives ANYmal on `constraint` contacts with violent random PD targets so that feet make and break contact at every
env-step, and prints per env-step the deviation from the oracle and whether the enabled constraint sets agree.  What
to expect: identical enabled sets at every step, and a deviation that grows smoothly (the motion is chaotic: about a
factor 1.3 per env-step from 1e-13); a jump of several orders of magnitude within one step would be a discrete
mismatch worth chasing.  Run from the repo root:  python tools/dev/long_horizon_probe.py [n_steps]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emul")):
    sys.path.insert(0, p)
from emul import emul_api          # noqa: E402
from jiminy_b200 import scenarios  # noqa: E402
import parity_common as pc         # noqa: E402


def main(n_steps: int = 50, n_env: int = 8, noise: float = 0.25) -> None:
    sc = scenarios.make("anymal", n_env, seed=11, solver="euler_explicit", dt_max=0.005, contact_model="constraint")
    eng, orc = pc.make_pair(sc, emul_api())
    rng = np.random.default_rng(3)
    for k in range(n_steps):
        act = sc.target0 + rng.uniform(-noise, noise, size=sc.target0.shape)
        eng.set_command(act)
        orc.set_command(act)
        eng.step(sc.step_dt)
        orc.step(sc.step_dt, parallel=True)
        d = np.abs(eng.get_state()[1] - orc.get_state()[1]).max(axis=1)
        c1, c0 = eng.get_constraints(), orc.get_constraints()
        same = np.array_equal(c1[0], c0[0]) and np.array_equal(c1[2], c0[2])
        print(k, "max |dq| per env", np.array2string(d, precision=1), "enabled sets equal:", same, "contacts", c0[2].sum(axis=1), flush=True)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 50)
