"""The reference's analytical checks run on a jiminy_b200 BatchedEngine (device path; in the CPU suite
the same kernel source under the warp emulator).  Mirrors tests/test_oracle_analytic.py, which pins the
oracle with the same physics -- here no oracle is involved at all: CUDA result vs closed form."""
import os

import numpy as np
import scipy.linalg

from jiminy_b200 import model as M
from jiminy_b200.core import BatchedEngine

from conftest import DATA


def _opt(**stepper):
    opt = M.default_engine_options()
    opt["contacts"]["model"] = "spring_damper"
    opt["stepper"].update(stepper)
    return opt


def armature_spring(api=None):
    """test_simple_pendulum.py:100-141: rotor inertia enters the ABA joint-space inertia."""
    r = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    M.attach_motor(r, "PendulumJoint", "PendulumJoint", enableVelocityLimit=False, enableEffortLimit=False,
                   enableArmature=True, armature=0.1)
    opt = _opt(odeSolver="runge_kutta_dopri", tolAbs=1e-8, tolRel=1e-8)
    opt["world"]["gravity"] = [0.0] * 6
    eng = BatchedEngine(r, opt, 3, api_=api)
    eng.set_joint_springs([500.0], [0.0])
    q0 = np.array([[0.1], [0.05], [-0.2]])
    ts, qs, vs, _ = eng.simulate(1.0, q0, np.zeros((3, 1)))
    A = np.array([[0.0, 1.0], [-500.0 / (5.0 + 0.1), 0.0]])
    for e in range(3):
        xa = np.stack([scipy.linalg.expm(A * t) @ np.array([q0[e, 0], 0.0]) for t in ts[:, e]])
        np.testing.assert_allclose(np.c_[qs[:, e], vs[:, e]], xa, rtol=1e-5, atol=1e-7)


def two_masses(api=None, period=1e-3):
    """test_double_spring_mass.py:85-130 (prismatic chain, discrete periods, adaptive DOPRI)."""
    r = M.build_robot_table(os.path.join(DATA, "linear_two_masses.urdf"), False)
    eng = BatchedEngine(r, _opt(odeSolver="runge_kutta_dopri", tolAbs=1e-8, tolRel=1e-8, sensorsUpdatePeriod=period,
                                controllerUpdatePeriod=period), 2, api_=api)
    k, nu, m = np.array([200.0, 20.0]), np.array([0.1, 0.2]), np.array([1.0, 2.5])
    eng.set_joint_springs(k, nu)
    Iq = 1.0 / m[1] + 1.0 / m[0]
    A = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [-k[0] / m[0], k[1] / m[0], -nu[0] / m[0], nu[1] / m[0]],
                  [k[0] / m[0], -k[1] * Iq, nu[0] / m[0], -nu[1] * Iq]])
    x0 = np.array([0.1, -0.1, 0.0, 0.0])
    ts, qs, vs, _ = eng.simulate(1.0, np.tile(x0[:2], (2, 1)), np.tile(x0[2:], (2, 1)))
    idx = np.linspace(0, len(ts) - 1, 25).astype(int)
    xa = np.stack([scipy.linalg.expm(A * t) @ x0 for t in ts[idx, 0]])
    np.testing.assert_allclose(np.c_[qs[idx, 1], vs[idx, 1]], xa, rtol=1e-5, atol=1e-7)


def contact_equilibrium_and_friction(api=None):
    """test_simple_mass.py:113-176 and :248-344: equilibrium depth = weight / k, sensors, friction steady state."""
    r = M.build_robot_table(os.path.join(DATA, "point_mass.urdf"), True)
    r.add_contact_points(["MassBody"])
    M.attach_sensor(r, "ContactSensor", "MassBody", frame_name="MassBody")
    M.attach_sensor(r, "ForceSensor", "F", frame_name="MassBody")
    opt = _opt(odeSolver="runge_kutta_4", dtMax=2e-4, controllerUpdatePeriod=1e-3, sensorsUpdatePeriod=1e-3)
    opt["contacts"].update(stiffness=1e6, damping=2e3, transitionEps=1e-6, friction=2.0, transitionVelocity=5e-2)
    eng = BatchedEngine(r, opt, 2, api_=api)
    q0 = np.tile(r.neutral(), (2, 1))
    q0[:, 2] = [0.0, 1e-3]
    eng.start(q0, np.zeros((2, 6)))
    for _ in range(40):
        eng.step(0.01)
    _, q, v, _ = eng.get_state()
    np.testing.assert_allclose(-q[:, 2], 9.81 / 1e6, atol=1e-7)
    s = eng.get_sensors()
    np.testing.assert_allclose(s[:, 2], 9.81, atol=1e-6)       # force sensor FZ
    np.testing.assert_allclose(s[:, 8], 9.81, atol=1e-6)       # contact sensor FZ
    # friction: horizontal gravity as a constant push -> v = Fx / (mu * weight)
    opt["world"]["gravity"] = [5.0, 0.0, -9.81, 0.0, 0.0, 0.0]
    eng = BatchedEngine(r, opt, 2, api_=api)
    q0[:, 2] = 0.0
    eng.start(q0, np.zeros((2, 6)))
    for _ in range(80):
        eng.step(0.01)
    _, _, v, a = eng.get_state()
    np.testing.assert_allclose(v[:, 0], 5.0 / (2.0 * 9.81), atol=1e-6)


def energy_conservation(api=None):
    """core/unit/engine_sanity_check.cc:47-165 on the device: double pendulum, zero torque."""
    from jiminy_b200 import robots as R
    robot, opt = R.load_robot("double_pendulum")
    opt = R.baseline_options("double_pendulum", opt)
    opt["stepper"].update(odeSolver="runge_kutta_dopri", tolAbs=1e-11, tolRel=1e-11, dtMax=0.02,
                          sensorsUpdatePeriod=1e-3, controllerUpdatePeriod=1e-3)
    eng = BatchedEngine(robot, opt, 2, api_=api)
    eng.start(np.array([[0.0, 0.1], [0.5, -0.3]]), np.zeros((2, 2)))
    e0 = eng.get_extra_terms()[0].sum(axis=1)
    for _ in range(50):
        eng.step(0.02)
    e1 = eng.get_extra_terms()[0].sum(axis=1)
    np.testing.assert_allclose(e1, e0, atol=1e-9)
