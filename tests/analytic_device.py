"""The reference's analytical checks run on a jiminy_b200 BatchedEngine (device path; in the CPU suite
the same kernel source under the warp emulator).  Mirrors tests/test_oracle_analytic.py, which pins the
oracle with the same physics -- here no oracle is involved at all: CUDA result vs closed form."""
import os

import numpy as np
import scipy.integrate
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


def velocity_bounds_criteria(ts, vel, acc, motor, tau=50.0, inertia=5.0, v_max=15.0, slope=0.5, tol=1e-7):
    """The assertions of test_simple_pendulum.py:143-211 (`test_velocity_bounds`): a constant command against
    SimpleMotor's velocity-dependent effort limit (basic_motors.cc:98-128)."""
    assert np.all(np.abs(vel) < v_max)                          # never beyond the limit ...
    assert v_max - abs(vel[-1]) < 1e-6 and abs(acc[-1]) < tol   # ... which is reached, where the acceleration vanishes
    acc_thr = tau / inertia
    start = next(i for i, a in enumerate(acc) if a < acc_thr)
    end = next(i for i, a in enumerate(acc) if a < 0.1)
    rate = np.diff(np.log(acc[start:end] / acc_thr)) / np.diff(ts[start:end])
    assert end - start > 5 and np.all(np.abs(rate - rate.mean()) < 1e-5)      # exponential decay of the acceleration
    v_th_max = max(motor.velocity_limit - slope * motor.effort_limit, 0.0)
    v_th = motor.velocity_limit - (tau / motor.effort_limit) * (motor.velocity_limit - v_th_max)
    assert vel[start - 1] < v_th < vel[start]                   # the taper starts where expected


def velocity_bounds_robot():
    r = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    M.attach_motor(r, "PendulumJoint", "PendulumJoint", enableEffortLimit=True, enableVelocityLimit=True,
                   velocityEffortInvSlope=0.5, velocityLimitFromUrdf=False, velocityLimit=15.0)
    r.q_lower[:], r.q_upper[:] = -1000.0, 1000.0
    opt = _opt(odeSolver="runge_kutta_dopri", tolAbs=1e-12, tolRel=1e-12)
    opt["world"]["gravity"] = [0.0] * 6
    return r, opt


def velocity_bounds(api=None):
    r, opt = velocity_bounds_robot()
    eng = BatchedEngine(r, opt, 2, api_=api)
    eng.set_command(np.array([[50.0], [0.0]]))
    ts, qs, vs, as_ = eng.simulate(4.0, np.zeros((2, 1)), np.zeros((2, 1)))
    velocity_bounds_criteria(ts[:, 0], vs[:, 0, 0], as_[:, 0, 0], r.motors[0])
    assert np.abs(vs[:, 1]).max() == 0.0                        # env 1: no command, nothing moves


def foot_pendulum_robot():
    """The robot of unit_py/test_foot_pendulum.py (hand-written fixture tests/data/foot_pendulum.urdf): contact points
    at the vertices of the foot's collision box, sensors as in its hardware file."""
    r = M.build_robot_table(os.path.join(DATA, "foot_pendulum.urdf"), True)
    box = r.links["Foot"].collision_boxes[0]
    names = []
    for i, xyz in enumerate(np.stack(np.meshgrid(*[0.5 * v * np.array([-1.0, 1.0]) for v in box.size], indexing="ij"), -1).reshape(-1, 3)):
        names.append(f"Foot_CollisionBox_0_{i}")
        r.add_frame(names[-1], "Foot", M.SE3(np.eye(3), xyz))
    r.add_contact_points(names)
    M.attach_motor(r, "PendulumJoint", "PendulumJoint")
    M.attach_sensor(r, "ImuSensor", "Foot", frame_name="Foot")
    M.attach_sensor(r, "ForceSensor", "Foot", frame_name="Foot")
    for i in (0, 2, 4, 6):
        M.attach_sensor(r, "ContactSensor", names[i], frame_name=names[i])
    opt = M.default_engine_options()
    opt["stepper"].update(odeSolver="runge_kutta_4", dtMax=1e-5, sensorsUpdatePeriod=0.0, controllerUpdatePeriod=1e-3)
    opt["contacts"].update(model="constraint", stabilizationFreq=0.0)
    opt["constraints"]["regularization"] = 1e-9
    return r, opt


def foot_pendulum_criteria(engine, r, t_end=1.0, tol=1.0e-5):
    """test_foot_pendulum.py:25-107 (`test_init_and_consistency`, its TOLERANCE = 1e-5): started exactly on its unstable
    equilibrium the pendulum does not move; no discontinuity at initialisation; IMU, force and contact sensors read the
    static values.  `engine`: an oracle or a BatchedEngine with one env."""
    q0 = np.array([[0.0, 0.0, 0.005, 0.0, 0.0, 0.0, 1.0, 0.0]])
    engine.set_command(np.zeros((1, 1)))
    rc = engine.start(q0, np.zeros((1, r.nv)))
    assert rc is None or not np.any(rc)
    a = engine.get_state()[3]
    assert np.all(np.abs(a) < tol)
    s, lay = engine.get_sensors()[0], r.sensor_layout()
    imu, force, contact = (s[lay[k][0]:lay[k][0] + lay[k][1] * lay[k][2]] for k in ("ImuSensor", "ForceSensor", "ContactSensor"))
    mass = float(r.inertia[1:, 0].sum())
    assert np.allclose(imu[:3], 0.0, atol=tol) and np.allclose(imu[3:], [0.0, 0.0, 9.81], atol=tol)
    assert np.allclose(force, [0.0, 0.0, 9.81 * mass, 0.0, 0.0, 0.0], atol=tol)
    c = contact.reshape(3, 4)                                       # field-major: FX, FY, FZ of the four sensors
    for i in range(3):
        assert np.allclose(c[:, i], c[:, i + 1], atol=tol)
    assert abs(c[2].sum() - 9.81 * mass) < 1e-3                     # the four bottom vertices carry the weight
    engine.step(t_end)
    _, q, v, a = engine.get_state()
    assert np.all(np.abs(v) < tol) and np.all(np.abs(a) < tol)


def foot_pendulum(api=None, t_end=1.0):
    r, opt = foot_pendulum_robot()
    foot_pendulum_criteria(BatchedEngine(r, opt, 1, api_=api), r, t_end)


def joint_position_limits_robot():
    r = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    M.attach_motor(r, "PendulumJoint", "PendulumJoint", enableVelocityLimit=False)
    r.q_lower[0], r.q_upper[0] = -0.002, 0.002
    opt = M.default_engine_options()
    opt["stepper"].update(odeSolver="euler_explicit", dtMax=1e-5, tolAbs=1e-9, tolRel=1e-8)
    opt["constraints"]["regularization"] = 0.0
    opt["contacts"]["transitionEps"] = 1e-4
    return r, opt, 0.002, 1e-4


def joint_position_limits_criteria(engine, joint_limit, transition_eps, t_end=0.05, step_dt=1e-5):
    """unit_py/test_dense_pole.py:160-199 (`test_joint_position_limits`): the bound constraint of a joint thrown at its
    limits must be enabled beyond the limit, keep its state inside the transition band, be disabled farther inside --
    and the run must visit all five cases."""
    engine.set_command(np.zeros((1, 1)))
    rc = engine.start(np.array([[0.0]]), np.array([[1.0]]))
    assert rc is None or not np.any(rc)
    branches = set()
    is_enabled = bool(engine.get_constraints()[0][0, 1])
    for _ in range(int(np.round(t_end / step_dt))):
        engine.step(step_dt)
        theta = engine.get_state()[1][0, 0]
        now = bool(engine.get_constraints()[0][0, 1])
        if joint_limit - abs(theta) <= 0.0:
            assert now
            branches.add(0 if is_enabled else 1)
        elif joint_limit - abs(theta) < transition_eps:
            assert now == is_enabled
            branches.add(2 if is_enabled else 3)
        else:
            assert not now
            branches.add(4)
        is_enabled = now
    assert branches == {0, 1, 2, 3, 4}


def joint_position_limits(api=None):
    r, opt, lim, eps = joint_position_limits_robot()
    joint_position_limits_criteria(BatchedEngine(r, opt, 1, api_=api), lim, eps)


def two_masses(api=None, period=1e-3, t_end=1.0):
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
    ts, qs, vs, _ = eng.simulate(t_end, np.tile(x0[:2], (2, 1)), np.tile(x0[2:], (2, 1)))
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


# ---------------------------------------------------------------------------------------------
# test_simple_pendulum.py:540-660: impulse forces are integration breakpoints, impulse-momentum theorem
IMPULSES = [
    dict(t=0.0, dt=2.0e-3, F=[1.0e3, 0.0, 0.0, 0.0, 0.0, 0.0]),
    dict(t=0.1, dt=1.0e-3, F=[0.0, 1.0e3, 0.0, 0.0, 0.0, 0.0]),
    dict(t=0.2, dt=2.0e-5, F=[-1.0e5, 0.0, 0.0, 0.0, 0.0, 0.0]),
    dict(t=0.2, dt=2.0e-4, F=[0.0, 0.0, 1.0e4, 0.0, 0.0, 0.0]),
    dict(t=0.4, dt=1.0e-5, F=[0.0, 0.0, 0.0, 0.0, 2.0e4, 0.0]),
    dict(t=0.4, dt=1.0e-5, F=[1.0e3, 1.0e4, 3.0e4, 0.0, 0.0, 0.0]),
    dict(t=0.6, dt=1.0e-6, F=[0.39e6, 1.72e6, 0.82e6, 0.36e6, -0.61e6, 1.17e6]),
    dict(t=0.8, dt=2.0e-6, F=[0.0, 0.0, 2.0e5, 0.0, 0.0, 0.0]),
]


def pendulum_impulse_reference(ts, m=5.0, l=1.0):
    """Independent solution of the gravity-free pendulum under the piecewise-constant wrenches above:
    ddq = (l F . d(q) + tau_y) / (m l^2), d(q) = (cos q, 0, -sin q), integrated segment by segment between
    consecutive force breakpoints with a tight-tolerance scipy solver."""
    brk = sorted({0.0, float(ts[-1])} | {f["t"] for f in IMPULSES} | {f["t"] + f["dt"] for f in IMPULSES})
    brk = [b for b in brk if b <= ts[-1] + 1e-12]
    out = np.zeros((len(ts), 2))
    x = np.zeros(2)
    for t0, t1 in zip(brk[:-1], brk[1:]):
        tm = 0.5 * (t0 + t1)
        act = [f for f in IMPULSES if f["t"] <= tm < f["t"] + f["dt"]]
        F = np.sum([f["F"] for f in act], axis=0) if act else np.zeros(6)

        def rhs(_t, y):
            d = np.array([np.cos(y[0]), 0.0, -np.sin(y[0])])
            return [y[1], (l * F[:3].dot(d) + F[4]) / (m * l * l)]
        sel = np.where((ts > t0 - 1e-13) & (ts <= t1 + 1e-13))[0]
        sol = scipy.integrate.solve_ivp(rhs, (t0, t1), x, method="DOP853", rtol=1e-13, atol=1e-13,
                                        t_eval=np.clip(ts[sel], t0, t1) if len(sel) else None, dense_output=False)
        if len(sel):
            out[sel] = sol.y.T
        x = scipy.integrate.solve_ivp(rhs, (t0, t1), x, method="DOP853", rtol=1e-13, atol=1e-13).y[:, -1]
    return out


def force_impulse(api=None, t_end=1.0):
    """Gravity-free pendulum under the impulse forces above, continuous and discrete (1 ms) scheduling, plus a
    second env whose impulses are shifted and scaled (per-env schedules)."""
    r = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    for period in (0.0, 1e-3):
        opt = _opt(sensorsUpdatePeriod=period, controllerUpdatePeriod=period)
        opt["world"]["gravity"] = [0.0] * 6
        eng = BatchedEngine(r, opt, 2, api_=api)
        for f in IMPULSES:
            eng.register_impulse_force("PendulumLink", [f["t"], f["t"]], f["dt"], np.array([f["F"], [0.0] * 6]))
        ts, qs, vs, _ = eng.simulate(t_end, np.zeros((2, 1)), np.zeros((2, 1)))
        xa = pendulum_impulse_reference(ts[:, 0])
        np.testing.assert_allclose(np.c_[qs[:, 0], vs[:, 0]], xa, atol=1e-6)
        np.testing.assert_allclose(np.c_[qs[:, 1], vs[:, 1]], 0.0, atol=1e-12)   # env 1: zero wrenches
        assert np.abs(vs[:, 0]).max() > 0.1



# ---------------------------------------------------------------------------------------------
# constraint path on the device, against physical closed forms (no oracle): normal force = weight at rest,
# Coulomb cone (stick below mu * weight, slide with a = F - mu * weight above), a joint stopped on its bound.
def constraint_closed_forms(api=None):
    r = M.build_robot_table(os.path.join(DATA, "point_mass.urdf"), True)
    r.add_contact_points(["MassBody"])
    mu = 0.8
    for Fx, slides in ((4.0, False), (15.0, True)):
        opt = _opt(dtMax=1e-3, controllerUpdatePeriod=1e-3, odeSolver="runge_kutta_4")
        opt["contacts"].update(model="constraint", friction=mu, transitionEps=1e-6)
        opt["world"]["gravity"] = [Fx, 0.0, -9.81, 0.0, 0.0, 0.0]
        eng = BatchedEngine(r, opt, 2, api_=api)
        q0 = np.tile(r.neutral(), (2, 1))
        eng.start(q0, np.zeros((2, 6)))
        for _ in range(30):
            eng.step(0.01)
        _, q, v, a = eng.get_state()
        fext = eng.get_efforts()[3]
        np.testing.assert_allclose(fext[:, 1, 2], 9.81, rtol=1e-3)                    # normal force = weight
        if slides:
            np.testing.assert_allclose(a[:, 0], Fx - mu * 9.81, rtol=2e-3)
            np.testing.assert_allclose(np.hypot(fext[:, 1, 0], fext[:, 1, 1]), mu * 9.81, rtol=2e-3)
        else:
            assert np.abs(v[:, 0]).max() < 1e-3 and np.abs(a[:, 0]).max() < 1e-2
            np.testing.assert_allclose(fext[:, 1, 0], -Fx, rtol=5e-3)                 # static friction balances the push
        assert np.abs(q[:, 2]).max() < 1e-4 and not eng.get_status().any()
    # pendulum resting on its upper position bound
    rp = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    rp.q_upper[0], rp.q_lower[0] = 0.5, -0.5
    opt = _opt(odeSolver="runge_kutta_4", dtMax=1e-3, controllerUpdatePeriod=1e-3, sensorsUpdatePeriod=1e-3)
    eng = BatchedEngine(rp, opt, 2, api_=api)
    eng.start(np.array([[0.3], [0.45]]), np.zeros((2, 1)))
    for _ in range(300):
        eng.step(0.01)
    _, q, v, _ = eng.get_state()
    assert np.all(np.abs(q - 0.5) < 2e-3) and np.abs(v).max() < 1e-4
    np.testing.assert_allclose(np.abs(eng.get_efforts()[0][:, 0]), 5.0 * 9.81 * np.sin(q[:, 0]), rtol=2e-3)
