"""Shared parity harness: drives a jiminy_b200 BatchedEngine (CUDA library, or -- in the CPU-only
suite -- the same kernel source under the thread emulator of tests/emul) and the oracle through the
same scenario and compares them."""
import numpy as np

from jiminy_b200 import scenarios
from jiminy_b200.core import BatchedEngine
from oracle.oracle import OracleBatch


def random_states(robot, n, rng, base_height=0.45):
    q = np.tile(robot.neutral(), (n, 1))
    for i in range(n):
        for j in range(1, robot.njoints):
            t, iq = int(robot.joint_type[j]), int(robot.idx_q[j])
            if t == 13:
                quat = rng.normal(size=4)
                q[i, iq:iq + 3] = rng.normal(size=3) * 0.1 + [0, 0, base_height]
                q[i, iq + 3:iq + 7] = quat / np.linalg.norm(quat)
            elif t in (5, 6, 7, 8):
                a = rng.uniform(-3, 3)
                q[i, iq], q[i, iq + 1] = np.cos(a), np.sin(a)
            else:
                lo, hi = max(robot.q_lower[iq], -0.5), min(robot.q_upper[iq], 0.5)
                q[i, iq] = rng.uniform(lo, hi)
    v = rng.normal(size=(n, robot.nv)) * 0.5
    return q, v


def make_pair(sc, api=None, device=0):
    eng = BatchedEngine(sc.robot, sc.options, sc.n_env, device=device, api_=api)
    orc = OracleBatch(sc.robot, sc.options, sc.n_env)
    if sc.kp is not None:
        eng.set_pd_controller(sc.kp, sc.kd)
        orc.set_pd_controller(sc.kp, sc.kd)
    eng.set_command(sc.target0)
    orc.set_command(sc.target0)
    eng.start(sc.q0, sc.v0)
    assert not orc.start(sc.q0, sc.v0).any()
    return eng, orc


def compare(eng, orc, tol_state, tol_sens):
    (t1, q1, v1, a1), (t0, q0, v0, a0) = eng.get_state(), orc.get_state()
    np.testing.assert_allclose(t1, t0, rtol=0, atol=1e-15)
    np.testing.assert_allclose(q1, q0, rtol=0, atol=tol_state)
    np.testing.assert_allclose(v1, v0, rtol=0, atol=tol_state * max(1.0, np.abs(v0).max()))
    np.testing.assert_allclose(a1, a0, rtol=0, atol=tol_sens * max(1.0, np.abs(a0).max()))
    s1, s0 = eng.get_sensors(), orc.get_sensors()
    if s0.size:
        np.testing.assert_allclose(s1, s0, rtol=0, atol=tol_sens * max(1.0, np.abs(s0).max()))
    np.testing.assert_array_equal(eng.get_iters()[0], orc.get_iters()[0])
    np.testing.assert_array_equal(eng.get_status(), orc.get_status())
    compare_extra_terms(eng, orc, max(tol_sens, 1e-11))


def run_scenario(name, n_env, n_steps, api=None, tol_state=1e-9, tol_sens=1e-7, **kw):
    sc = scenarios.make(name, n_env, **kw)
    eng, orc = make_pair(sc, api)
    compare(eng, orc, 1e-13, 1e-12)
    for k in range(n_steps):
        act = sc.sample_targets(k)
        eng.set_command(act)
        orc.set_command(act)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
        compare(eng, orc, tol_state, tol_sens)
    return eng, orc, sc


def compare_extra_terms(eng, orc, tol=1e-9):
    """computeExtraTerms outputs: energies, joint spatial accelerations, joint internal wrenches."""
    e1, a1, f1 = eng.get_extra_terms()
    e0, a0, f0 = orc.get_extra_terms()
    np.testing.assert_allclose(e1, e0, rtol=0, atol=tol * max(1.0, np.abs(e0).max()))
    np.testing.assert_allclose(a1, a0, rtol=0, atol=tol * max(1.0, np.abs(a0).max()))
    np.testing.assert_allclose(f1, f0, rtol=0, atol=tol * max(1.0, np.abs(f0).max()))
