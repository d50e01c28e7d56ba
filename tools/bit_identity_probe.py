#!/usr/bin/env python
"""How close to BIT-identical is the device code to the (uncontracted, -ffp-contract=off) oracle when nvcc's
multiply-add contraction is taken out of the picture?  The kernel source compiled for the host WITHOUT contraction
(tests/emul, `-ffp-contract=off`) executes the same IEEE operations in the same order as a `-fmad=false` device build
(apart from libm: sincos / tanh / sqrt), so the comparison runs on the CPU.  For every BASELINE robot: one dynamics
evaluation on random states, and one env-step; reports the share of bit-equal outputs and the distance in ulps.
Usage: python tools/bit_identity_probe.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emul")):
    sys.path.insert(0, p)
from emul import emul_api           # noqa: E402
from jiminy_b200 import scenarios   # noqa: E402
import parity_common as pc          # noqa: E402


def ulps(a, b):
    """|a - b| in units of the spacing of doubles at the largest magnitude of the env's vector (a per-element ulp count is
    meaningless for components that are zero up to rounding)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.spacing(np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-300))
    return np.rint(np.abs(a - b) / scale).astype(np.int64)


def main():
    rows = []
    for fma in (False, True):
        api = emul_api(fma=fma)
        for name in ("double_pendulum", "cartpole", "anymal", "atlas"):
            sc = scenarios.make(name, 8, seed=1)
            eng, orc = pc.make_pair(sc, api)
            eng.set_command(sc.sample_targets(0)); orc.set_command(sc.sample_targets(0))
            eng.step(sc.step_dt); orc.step(sc.step_dt)
            (_, q1, v1, _), (_, q0, v0, _) = eng.get_state(), orc.get_state()
            u_q, u_v = ulps(q1, q0), ulps(v1, v0)
            rng = np.random.default_rng(0)
            q, v = pc.random_states(sc.robot, 8, rng)
            cmd = rng.uniform(-5, 5, size=(8, max(sc.robot.nmotors, 1)))
            a1 = eng.compute_dynamics(q, v, cmd)[0]
            a0 = orc.compute_dynamics(q, v, cmd)[0]
            u_rhs = ulps(a1, a0)
            rows.append((("contracted (GPU-like)" if fma else "uncontracted (-fmad=false)"), name,
                         float((u_rhs == 0).mean()), int(np.median(u_rhs)), int(u_rhs.max()),
                         float((u_q == 0).mean()), int(u_q.max()), float((u_v == 0).mean()), int(np.median(u_v))))
    print(f"{'device rounding':28s} {'robot':16s} {'ddq bit-equal':>13s} {'median ulp':>10s} {'max ulp':>9s} | after one env-step: "
          f"{'q bit-equal':>11s} {'max ulp':>9s} {'v bit-equal':>11s} {'median ulp':>10s}")
    for r in rows:
        print(f"{r[0]:28s} {r[1]:16s} {100 * r[2]:12.1f}% {r[3]:10d} {r[4]:9d} | {'':20s}{100 * r[5]:10.1f}% {r[6]:9d} {100 * r[7]:10.1f}% {r[8]:10d}")


if __name__ == "__main__":
    main()
