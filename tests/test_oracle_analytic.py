"""Pins the CPU oracle with the reference's own analytical tests, re-implemented on
hand-written fixtures:
  1. core/unit/engine_sanity_check.cc:47-165 ......... double-pendulum energy conservation
  2. unit_py/test_simple_pendulum.py:100-141 .......... rotor inertia (armature) + spring vs expm
  3. unit_py/test_simple_pendulum.py:240-267 .......... non-linear pendulum vs scipy dopri5
  4. unit_py/test_double_spring_mass.py:85-130 ........ two prismatic masses vs expm (continuous, discrete)
  5. unit_py/test_simple_mass.py:113-176, :248-344 .... spring-damper contact equilibrium, friction steady state
  6. unit_py/test_simple_mass.py:183-246 .............. contact / force sensor == external force in frame
  7. unit_py/test_simulator.py:26-109 ................. Euler finite difference of v equals a; IMU reads g at rest
  8. unit_py/test_simple_pendulum.py:143-211 .......... SimpleMotor velocity-dependent effort limit
  9. unit_py/test_foot_pendulum.py:25-107 ............. redundant contact constraints at an unstable equilibrium
 10. unit_py/test_dense_pole.py:160-199 ............... joint-bound constraint enable / disable hysteresis
 11. gym_jiminy/unit_py/test_pipeline_control.py ...... PD pipeline: Atlas stands still, target consistency, Mahony filter
The reference binary itself cannot run here, so these analytical pins are what anchors the oracle.
"""
import os

import numpy as np
import pytest
import scipy.integrate
import scipy.linalg

from jiminy_b200 import model as M
from jiminy_b200 import robots as R
from oracle.oracle import OracleBatch

from conftest import DATA

TOL = 1e-7


def _opt(**stepper):
    opt = M.default_engine_options()
    opt["contacts"]["model"] = "spring_damper"
    opt["stepper"].update(stepper)
    return opt


def _pendulum(armature=None):
    r = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    kw = dict(enableVelocityLimit=False, enableEffortLimit=False)
    if armature is not None:
        kw.update(enableArmature=True, armature=armature)
    M.attach_motor(r, "PendulumJoint", "PendulumJoint", **kw)
    return r


def test_armature_spring_vs_expm():
    J, k = 0.1, 500.0
    r = _pendulum(J)
    opt = _opt(tolAbs=TOL * 0.1, tolRel=TOL * 0.1)
    opt["world"]["gravity"] = [0.0] * 6
    o = OracleBatch(r, opt)
    o.set_springs([k], [0.0])
    ts, qs, vs, _ = o.simulate(2.0, [0.1], [0.0])
    I_eq = 5.0 * 1.0 ** 2 + J
    A = np.array([[0.0, 1.0], [-k / I_eq, 0.0]])
    xa = np.stack([scipy.linalg.expm(A * t) @ np.array([0.1, 0.0]) for t in ts])
    np.testing.assert_allclose(np.c_[qs, vs], xa, atol=TOL)


def test_pendulum_vs_scipy_dopri5():
    r = _pendulum()
    o = OracleBatch(r, _opt(tolAbs=1e-10, tolRel=1e-10))
    ts, qs, vs, _ = o.simulate(2.0, [0.3], [0.0])
    g, l = 9.81, 1.0
    # joint axis +Y, mass at +z: q is measured from the upright position
    sol = scipy.integrate.solve_ivp(lambda t, x: [x[1], g / l * np.sin(x[0])], (0.0, 2.0), [0.3, 0.0], method="DOP853",
                                    t_eval=ts, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(qs[:, 0], sol.y[0], atol=TOL)
    np.testing.assert_allclose(vs[:, 0], sol.y[1], atol=TOL)


@pytest.mark.parametrize("period", [0.0, 1e-3])
def test_two_masses_vs_expm(period):
    r = M.build_robot_table(os.path.join(DATA, "linear_two_masses.urdf"), False)
    o = OracleBatch(r, _opt(tolAbs=TOL * 0.1, tolRel=TOL * 0.1, sensorsUpdatePeriod=period, controllerUpdatePeriod=period))
    k, nu, m = np.array([200.0, 20.0]), np.array([0.1, 0.2]), np.array([1.0, 2.5])
    o.set_springs(k, nu)
    Iq = 1.0 / m[1] + 1.0 / m[0]
    A = np.array([[0, 0, 1, 0], [0, 0, 0, 1],
                  [-k[0] / m[0], k[1] / m[0], -nu[0] / m[0], nu[1] / m[0]],
                  [k[0] / m[0], -k[1] * Iq, nu[0] / m[0], -nu[1] * Iq]])
    x0 = np.array([0.1, -0.1, 0.0, 0.0])
    ts, qs, vs, _ = o.simulate(4.0, x0[:2], x0[2:])
    idx = np.linspace(0, len(ts) - 1, 60).astype(int)
    xa = np.stack([scipy.linalg.expm(A * t) @ x0 for t in ts[idx]])
    np.testing.assert_allclose(np.c_[qs, vs][idx], xa, rtol=1e-5, atol=TOL)   # == np.allclose(.., atol=TOL) of the reference test


def test_two_masses_python_controller_callback():
    """Same system driven through the FunctionalController-style callback instead of built-in springs."""
    r = M.build_robot_table(os.path.join(DATA, "linear_two_masses.urdf"), False)
    for j in ("FirstJoint", "SecondJoint"):
        M.attach_motor(r, j, j, enableVelocityLimit=False, enableEffortLimit=False)
    k, nu = np.array([200.0, 20.0]), np.array([0.1, 0.2])
    a, b = OracleBatch(r, _opt(tolAbs=1e-9, tolRel=1e-9)), OracleBatch(r, _opt(tolAbs=1e-9, tolRel=1e-9))
    a.set_callbacks(0, controller=lambda t, q, v, s, out: out.__setitem__(slice(None), -k * q - nu * v))
    b.set_springs(k, nu)
    ta, qa, va, _ = a.simulate(0.5, [0.1, -0.1], [0.0, 0.0])
    tb, qb, vb, _ = b.simulate(0.5, [0.1, -0.1], [0.0, 0.0])
    np.testing.assert_allclose(qa, qb, atol=1e-12)


def _point_mass(**contacts):
    r = M.build_robot_table(os.path.join(DATA, "point_mass.urdf"), True)
    r.add_contact_points(["MassBody"])
    M.attach_sensor(r, "ContactSensor", "MassBody", frame_name="MassBody")
    r.add_frame("Sensor", "MassBody", M.SE3(M.rpy_to_matrix([0.3, -0.2, 0.5]), np.array([0.1, 0.2, -0.05])))
    M.attach_sensor(r, "ForceSensor", "F", frame_name="Sensor")
    opt = _opt(dtMax=1e-5, controllerUpdatePeriod=1e-5)
    opt["contacts"].update(stiffness=1e6, damping=2e3, transitionEps=1e-6, **contacts)
    return r, opt


def test_contact_equilibrium_and_sensors():
    r, opt = _point_mass()
    o = OracleBatch(r, opt)
    q0 = r.neutral()
    q0[2] = 1.0
    assert not o.start(q0, np.zeros(6)).any()
    energies = []
    for _ in range(150):
        assert not o.step(0.01).any()
        energies.append(o.get_extra_terms()[0][0].sum())
    _, q, v, _ = o.get_state()
    weight = 9.81
    np.testing.assert_allclose(-q[0, 2], weight / 1e6, atol=TOL)          # equilibrium depth = weight / k
    fext = o.get_efforts()[3]
    np.testing.assert_allclose(fext[0, 1, 2], weight, atol=TOL)           # f_external on the parent joint
    s = o.get_sensors()[0]
    lay = r.sensor_layout()
    cont = s[lay["ContactSensor"][0]:lay["ContactSensor"][0] + 3]
    np.testing.assert_allclose(cont, [0, 0, weight], atol=TOL)
    # force sensor = same wrench expressed in the (rotated, shifted) sensor frame
    F = s[lay["ForceSensor"][0]:lay["ForceSensor"][0] + 6]
    P = r.frames["Sensor"].placement
    f_expected = P.R.T @ np.array([0, 0, weight])
    t_expected = P.R.T @ np.cross(-P.p, np.array([0, 0, weight]))
    np.testing.assert_allclose(F[:3], f_expected, atol=TOL)
    np.testing.assert_allclose(F[3:], t_expected, atol=TOL)
    # mechanical energy (robot + contact spring) never increases beyond numerical noise while settling
    assert energies[-1] < energies[0]


def test_friction_steady_state_velocity():
    """v_steady = Fx / (mu * weight): pins the un-normalised tangential velocity of engine.cc:3218-3222."""
    r, opt = _point_mass(friction=2.0, transitionVelocity=5e-2)
    Fx = 5.0
    opt["world"]["gravity"] = [Fx, 0.0, -9.81, 0.0, 0.0, 0.0]   # horizontal gravity = constant push on the unit mass
    o = OracleBatch(r, opt)
    q0 = r.neutral()
    assert not o.start(q0, np.zeros(6)).any()
    for _ in range(80):
        assert not o.step(0.01).any()
    _, _, v, a = o.get_state()
    np.testing.assert_allclose(v[0, 0], Fx / (2.0 * 9.81), atol=TOL)
    assert abs(a[0, 0]) < 1e-5


def test_double_pendulum_energy_conservation():
    robot, opt = R.load_robot("double_pendulum")
    opt = R.baseline_options("double_pendulum", opt)
    opt["stepper"].update(sensorsUpdatePeriod=0.0, controllerUpdatePeriod=0.0, dtMax=1e-3)
    o = OracleBatch(robot, opt)
    assert not o.start([0.0, 0.1], [0.0, 0.0]).any()
    e0 = o.get_extra_terms()[0][0].sum()
    drift = 0.0
    for _ in range(300):
        assert not o.step(0.01).any()
        drift = max(drift, abs(o.get_extra_terms()[0][0].sum() - e0))
    assert drift < 1e-6, drift     # fixed-step RK4 at 1 ms: O(dt^4) energy error
    # the reference test: DOPRI with tolAbs = tolRel = 1e-11, continuous then discrete 1 ms, drift < 1e-9
    for period, h in ((0.0, 0.02), (1e-3, 0.02)):
        opt["stepper"].update(odeSolver="runge_kutta_dopri", tolAbs=1e-11, tolRel=1e-11, dtMax=0.02,
                              sensorsUpdatePeriod=period, controllerUpdatePeriod=period)
        o = OracleBatch(robot, opt)
        o.start([0.0, 0.1], [0.0, 0.0])
        e0 = o.get_extra_terms()[0][0].sum()
        for _ in range(150):
            assert not o.step(h).any()
        assert abs(o.get_extra_terms()[0][0].sum() - e0) < 1e-9


def test_euler_finite_difference_and_imu_at_rest():
    r = _pendulum()
    r.add_frame("ImuFrame", "PendulumLink", M.SE3(M.rpy_to_matrix([0.1, 0.2, 0.3]), np.zeros(3)))
    M.attach_sensor(r, "ImuSensor", "imu", frame_name="ImuFrame")
    opt = _opt(odeSolver="euler_explicit", dtMax=1e-3, controllerUpdatePeriod=1e-3, sensorsUpdatePeriod=1e-3)
    o = OracleBatch(r, opt)
    ts, qs, vs, as_ = o.simulate(0.05, [0.2], [0.0])
    fd = np.diff(vs[:, 0]) / np.diff(ts)
    np.testing.assert_allclose(fd[1:], as_[1:-1, 0], atol=1e-9)           # v' - v = dt * a(previous)
    # hanging at rest (stable equilibrium, q = pi): accelerometer measures |g|, gyro zero
    opt2 = _opt(odeSolver="runge_kutta_4", dtMax=1e-3)
    o = OracleBatch(r, opt2)
    o.set_springs([0.0], [50.0])
    o.start([np.pi], [0.0])
    o.step(0.01)
    s = o.get_sensors()[0]
    np.testing.assert_allclose(np.linalg.norm(s[3:6]), 9.81, atol=1e-9)
    np.testing.assert_allclose(s[:3], 0.0, atol=1e-9)


def test_lie_group_integrate_difference_roundtrip():
    r = M.build_robot_table(os.path.join(DATA, "branched_arm.urdf"), True)
    o = OracleBatch(r, _opt())
    rng = np.random.default_rng(0)
    q = r.neutral()
    for _ in range(20):
        v = rng.normal(size=r.nv) * 0.3
        q1 = o.integrate(q, v)
        np.testing.assert_allclose(o.difference(q, q1), v, atol=1e-12)
        np.testing.assert_allclose(np.linalg.norm(q1[3:7]), 1.0, atol=1e-12)
        q = q1


def test_step_scheduler_microsecond_first_step_and_iters():
    """First sub-step of an episode is 1 us (engine.cc:1176); afterwards dtMax steps snapped to the us grid."""
    robot, opt = R.load_robot("anymal")
    opt = R.baseline_options("anymal", opt)
    o = OracleBatch(robot, opt)
    q0 = R.ground_base_height(robot, robot.neutral())
    assert not o.start(q0, np.zeros(robot.nv)).any()
    o.step(0.04)
    assert o.get_iters()[0][0] == 41           # 1 us + 4 x 1 ms + 0.999 ms, then 7 x 5 x 1 ms
    o.step(0.04)
    assert o.get_iters()[0][0] == 81
    np.testing.assert_allclose(o.get_state()[0][0], 0.08, atol=1e-15)


# ---------------------------------------------------------------------------------------------
# 8. unit_py/test_simple_pendulum.py:540-660 .......... impulse forces: breakpoints + impulse-momentum
from analytic_device import IMPULSES, pendulum_impulse_reference  # noqa: E402


@pytest.mark.parametrize("period", [0.0, 1e-3])
def test_pendulum_force_impulse(period):
    r = _pendulum()
    opt = _opt(sensorsUpdatePeriod=period, controllerUpdatePeriod=period)
    opt["world"]["gravity"] = [0.0] * 6
    o = OracleBatch(r, opt)
    fr = r.frames["PendulumLink"]
    for f in IMPULSES:
        o.register_impulse_force(fr.joint, fr.placement.p, f["t"], f["dt"], f["F"])
    ts, qs, vs, _ = o.simulate(1.0, [0.0], [0.0])
    xa = pendulum_impulse_reference(ts)
    np.testing.assert_allclose(np.c_[qs, vs], xa, atol=1e-6)
    # without breakpoints at t and t + dt a 1 us, 1e6 N impulse would be mis-integrated by orders of magnitude
    assert np.abs(vs).max() > 0.1 and abs(qs[-1, 0]) > 0.01


# ---------------------------------------------------------------------------------------------
# contacts.model = "constraint" and joint position bounds: the PGS path (engine.cc:3709-3866).
# Physical closed forms the boxed LCP must reproduce: normal force = weight at rest, Coulomb cone
# (stick below mu * weight, slide with a = (F - mu * weight) / m above), a joint stopped at its bound,
# and the reference's own check that the contact / force sensors equal f_external in the frame
# (test_simple_mass.py:181-246, which runs with both contact models).
def _point_mass_constraint(**contacts):
    r, opt = _point_mass(**contacts)
    opt["contacts"]["model"] = "constraint"
    opt["stepper"].update(dtMax=1e-3, controllerUpdatePeriod=1e-3, odeSolver="runge_kutta_4")
    return r, opt


def test_constraint_contact_rest_and_sensors():
    r, opt = _point_mass_constraint()
    o = OracleBatch(r, opt)
    o.set_callbacks(0, internal_dynamics=lambda t, q, v, s, u: u.__setitem__(slice(3, 6), 1.0))   # spinning mass
    q0 = r.neutral()
    q0[2] = 0.02
    assert not o.start(q0, np.zeros(6)).any()
    lay = r.sensor_layout()
    P = r.frames["Sensor"].placement
    for k in range(100):
        assert not o.step(0.01).any()
        # sensors == f_external[parent joint] expressed in the sensor frames, at every step (falling, impact, rest)
        fext = o.get_efforts()[3][0, 1]
        s = o.get_sensors()[0]
        cont = s[lay["ContactSensor"][0]:lay["ContactSensor"][0] + 3]
        F = s[lay["ForceSensor"][0]:lay["ForceSensor"][0] + 6]
        np.testing.assert_allclose(cont, fext[:3], atol=TOL)
        np.testing.assert_allclose(P.R @ F[:3], fext[:3], atol=TOL)
        np.testing.assert_allclose(P.R @ F[3:] + np.cross(P.p, P.R @ F[:3]), fext[3:], atol=TOL)
    _, q, v, a = o.get_state()
    # at rest on the ground: z ~ 0 (Baumgarte pulls the frame back to the surface), zero linear motion
    assert abs(q[0, 2]) < 1e-5 and np.abs(v[0, :3]).max() < 1e-6 and np.abs(a[0, :3]).max() < 1e-4
    # normal force = weight, expressed in the (spinning) body frame: its norm is the weight
    np.testing.assert_allclose(np.linalg.norm(o.get_efforts()[3][0, 1, :3]), 9.81, rtol=1e-4)
    # torsion = 0: the spin about z is not resisted by the contact
    assert v[0, 5] > 0.5


@pytest.mark.parametrize("Fx,slides", [(4.0, False), (15.0, True)])
def test_constraint_contact_coulomb_cone(Fx, slides):
    mu = 0.8
    r, opt = _point_mass_constraint(friction=mu)
    opt["world"]["gravity"] = [Fx, 0.0, -9.81, 0.0, 0.0, 0.0]   # unit mass: horizontal gravity = constant push
    o = OracleBatch(r, opt)
    assert not o.start(r.neutral(), np.zeros(6)).any()
    for _ in range(30):
        assert not o.step(0.01).any()
    _, q, v, a = o.get_state()
    if slides:
        np.testing.assert_allclose(a[0, 0], Fx - mu * 9.81, rtol=2e-3)      # regularization 1e-3 softens the cone
        np.testing.assert_allclose(v[0, 0], (Fx - mu * 9.81) * 0.3, rtol=5e-3)
    else:
        assert abs(v[0, 0]) < 1e-3 and abs(a[0, 0]) < 1e-2                    # sticks (up to the solver compliance)
    assert abs(q[0, 2]) < 1e-4


@pytest.mark.parametrize("model", ["spring_damper", "constraint"])
def test_joint_bound_stops_pendulum(model):
    """A pendulum falling onto its upper position bound stops there (JointConstraint, lambda >= 0 pushes back
    inside), whatever the contact model: the bounds dynamics always goes through the constraint solver."""
    r = _pendulum()
    r.q_upper[0], r.q_lower[0] = 0.5, -0.5
    opt = _opt(odeSolver="runge_kutta_4", dtMax=1e-3, controllerUpdatePeriod=1e-3, sensorsUpdatePeriod=1e-3)
    opt["contacts"]["model"] = model
    o = OracleBatch(r, opt)
    assert not o.start([0.3], [0.0]).any()
    qmax = 0.0
    for _ in range(300):
        assert not o.step(0.01).any()
        qmax = max(qmax, o.get_state()[1][0, 0])
    _, q, v, a = o.get_state()
    assert 0.5 - 1e-3 < q[0, 0] < 0.5 + 2e-3 and qmax < 0.52     # rests on the bound, small overshoot at impact
    assert abs(v[0, 0]) < 1e-4
    u = o.get_efforts()[0]
    # the bound holds the gravity torque m g l sin(q): reported in u (engine.cc:3770-3788)
    np.testing.assert_allclose(abs(u[0, 0]), 5.0 * 9.81 * np.sin(q[0, 0]), rtol=2e-3)
    assert o.get_status()[0] & 8


def test_atlas_pd_pipeline_stands_still_like_the_reference_test():
    """gym_jiminy/unit_py/test_pipeline_control.py:46-113 (`test_pid_standing`, Atlas): after 9 s of zero target motor
    velocities the velocity targets of the last second are below 1e-9 and every generalised velocity below 1e-3.
    Restated on the oracle with the reference's neutral posture, option file and block arguments: constraint contacts,
    knees and shoulders on their position bounds, safety limits, PD controller, adapter and Mahony filter together."""
    import parity_common as pc
    v_target, v_robot, orc, sc = pc.atlas_pd_standing_on_oracle()
    last = int(round(1.0 / sc.step_dt))
    assert np.all(v_target[-last:] < 1.0e-9)
    assert np.all(v_robot[-last:] < 1.0e-3), v_robot[-last:].max()
    assert v_robot[:5].max() > 0.05                                   # it did settle from somewhere
    q = orc.get_state()[1][0]
    assert abs(q[2] - pc.atlas_reference_neutral(sc.robot)[2]) < 5e-3 and (orc.get_status() & ~8 == 0).all()


def test_pd_controller_targets_like_the_reference_test():
    """gym_jiminy/unit_py/test_pipeline_control.py:258-313 (`test_pd_controller`): Atlas PD pipeline with the acceleration
    limits lifted and random target velocities for 2 s.  The logged targets of the last motor must satisfy: the target
    velocity reaches the commanded one at the end of every adapter period; finite differences of target position /
    velocity reproduce the logged velocity / acceleration (shifted by one controller period, as the reference checks
    them); position and velocity targets stay within the motor limits."""
    import parity_common as pc
    from jiminy_b200 import scenarios
    from jiminy_b200.blocks import pd_adapter
    from oracle.oracle import OracleBatch
    sc = scenarios.make("atlas", 1, seed=0, contact_model="constraint", solver="euler_explicit", dt_max=0.005)
    rob, nm = sc.robot, sc.robot.nmotors
    iq = np.array([rob.idx_q[m.joint] for m in rob.motors])
    v_hw = np.array([m.velocity_limit for m in rob.motors])
    vel = np.minimum(v_hw, pc.ATLAS_PIPELINE["joint_velocity_limit"])
    lower = np.stack([rob.q_lower[iq], -vel, np.full(nm, -1e300)])        # controller._command_state_lower[2] = -inf
    upper = np.stack([rob.q_upper[iq], vel, np.full(nm, 1e300)])
    sf = pc.ATLAS_PIPELINE["safety"]
    table = np.stack([np.full(nm, sf["kp"]), np.full(nm, sf["kd"]), rob.q_lower[iq], rob.q_upper[iq], np.minimum(v_hw, sf["soft_velocity_max"])])
    orc = OracleBatch(rob, sc.options, 1)
    orc.set_pd_controller_full(sc.kp, sc.kd, lower, upper, table)
    orc.set_command(np.zeros((1, nm)))
    assert not orc.start(pc.atlas_reference_neutral(rob)[None, :], np.zeros((1, rob.nv))).any()
    rng = np.random.default_rng(0)
    control_dt = sc.options["stepper"]["controllerUpdatePeriod"]
    update_ratio = int(round(sc.step_dt / control_dt))
    pos, velo, acc, cmd = [], [], [], []
    for _ in range(int(round(2.0 / sc.step_dt))):
        action = 0.2 * rng.uniform(lower[1], upper[1])[None, :]             # 0.2 * action_space.sample()
        state, out = orc.get_pd_controller_state(), np.zeros((1, nm))
        pd_adapter(action.copy(), 1, state, lower, upper, False, np.zeros(nm), sc.step_dt, out)
        orc.set_command(out)
        for _ in range(update_ratio):
            assert not orc.step(control_dt).any()
            s = orc.get_pd_controller_state()[0]
            pos.append(s[0, -1]); velo.append(s[1, -1]); acc.append(s[2, -1]); cmd.append(action[0, -1])
    pos, velo, acc, cmd = map(np.array, (pos, velo, acc, cmd))
    TOLERANCE = 1.0e-6                                                       # the reference's
    np.testing.assert_allclose(velo[update_ratio - 1::update_ratio], cmd[update_ratio - 1::update_ratio], atol=TOLERANCE)
    np.testing.assert_allclose((np.diff(velo) / control_dt)[:-1], acc[1:-1], atol=TOLERANCE)
    np.testing.assert_allclose((np.diff(pos) / control_dt)[:-1], velo[1:-1], atol=TOLERANCE)
    assert np.all((rob.q_lower[iq][-1] <= pos) & (pos <= rob.q_upper[iq][-1])) and np.all(np.abs(velo) <= v_hw[-1])
    assert np.abs(velo).max() > 0.05


def test_mahony_filter_tracks_the_imu_like_the_reference_test():
    """gym_jiminy/unit_py/test_pipeline_control.py:135-190 (`test_mahony_filter_plus_body_observer`, the variant where
    the twist is measured): Atlas PD pipeline with a Mahony filter at kp = ki = 0 and exact initialisation; a constant
    action swings the upper body (back joints at 0.4 / 0.08 / 0.08 rad/s, sign flipped every 50 steps) for 200 steps;
    the roll-pitch-yaw of the estimated IMU orientation must stay within 5e-3 rad of the true orientation of the IMU
    frame (the reference's tolerance; the gyro integration at 5 ms leaves 4.9e-3 here)."""
    import parity_common as pc
    from jiminy_b200 import scenarios, robots as R
    from jiminy_b200.blocks import pd_adapter
    from oracle.oracle import OracleBatch
    sc = scenarios.make("atlas", 1, seed=0, contact_model="constraint", solver="euler_explicit", dt_max=0.005)
    rob, nm = sc.robot, sc.robot.nmotors
    iq = np.array([rob.idx_q[m.joint] for m in rob.motors])
    v_hw = np.array([m.velocity_limit for m in rob.motors])
    vel = np.minimum(v_hw, pc.ATLAS_PIPELINE["joint_velocity_limit"])
    acc = np.full(nm, pc.ATLAS_PIPELINE["joint_acceleration_limit"])
    lower, upper = np.stack([rob.q_lower[iq], -vel, -acc]), np.stack([rob.q_upper[iq], vel, acc])
    sf = pc.ATLAS_PIPELINE["safety"]
    table = np.stack([np.full(nm, sf["kp"]), np.full(nm, sf["kd"]), rob.q_lower[iq], rob.q_upper[iq], np.minimum(v_hw, sf["soft_velocity_max"])])
    orc = OracleBatch(rob, sc.options, 1)
    orc.set_pd_controller_full(sc.kp, sc.kd, lower, upper, table)
    orc.set_mahony_filter(0.0, 0.0)
    orc.set_command(np.zeros((1, nm)))
    assert not orc.start(pc.atlas_reference_neutral(rob)[None, :], np.zeros((1, rob.nv))).any()
    names = [m.name for m in rob.motors]
    action = np.zeros((1, nm))
    for name, value in (("back_bkz", 0.4), ("back_bky", 0.08), ("back_bkx", 0.08)):
        action[0, names.index(name)] = value

    def quat_to_matrix(q):
        x, y, z, w = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])

    def matrix_to_rpy(m):
        return np.array([np.arctan2(m[2, 1], m[2, 2]), -np.arcsin(m[2, 0]), np.arctan2(m[1, 0], m[0, 0])])
    frame, swing = rob.imu_frames[0], 0.0
    for i in range(200):
        a = action * (1 - 2 * ((i // 50) % 2))
        state, out = orc.get_pd_controller_state(), np.zeros((1, nm))
        pd_adapter(a.copy(), 1, state, lower, upper, False, None, sc.step_dt, out)
        orc.set_command(out)
        assert not orc.step(sc.step_dt).any()
        rpy_true = matrix_to_rpy(R.frame_placements(rob, orc.get_state()[1][0], [frame])[frame].R)
        rpy_est = matrix_to_rpy(quat_to_matrix(orc.get_mahony_filter()[0, 0, :4]))
        np.testing.assert_allclose(rpy_true, rpy_est, atol=5e-3)
        swing = max(swing, np.abs(rpy_true).max())
    assert swing > 0.3


def test_motor_velocity_bounds_like_the_reference_test():
    """unit_py/test_simple_pendulum.py:143-211 (`test_velocity_bounds`): constant command against the velocity-dependent
    effort limit of `SimpleMotor` -- the velocity saturates at the limit, the acceleration decays exponentially from the
    velocity where the taper starts."""
    import analytic_device as ad
    r, opt = ad.velocity_bounds_robot()
    o = OracleBatch(r, opt)
    o.set_command(np.array([[50.0]]))
    ts, qs, vs, as_ = o.simulate(4.0, [0.0], [0.0])
    ad.velocity_bounds_criteria(np.asarray(ts).reshape(len(ts), -1)[:, 0], vs[:, 0], as_[:, 0], r.motors[0])


def test_foot_pendulum_holds_its_equilibrium_like_the_reference_test():
    """unit_py/test_foot_pendulum.py:25-107: inverted pendulum on a square foot, `constraint` contacts with four redundant
    contact points, no Baumgarte stabilisation, regularisation 1e-9."""
    import analytic_device as ad
    r, opt = ad.foot_pendulum_robot()
    ad.foot_pendulum_criteria(OracleBatch(r, opt), r)


def test_joint_position_limits_like_the_reference_test():
    """unit_py/test_dense_pole.py:160-199: enable / keep / disable logic of a joint-bound constraint (hysteresis band
    `transitionEps`), all five cases visited."""
    import analytic_device as ad
    r, opt, lim, eps = ad.joint_position_limits_robot()
    ad.joint_position_limits_criteria(OracleBatch(r, opt), lim, eps)


def test_centroidal_terms_against_first_principles():
    """computeExtraTerms (engine.cc:817-832, :890-904) on the oracle, checked against quantities derived independently
    of any spatial algebra: the whole-robot centre of mass from a numpy forward kinematics (sum of m_i * oMi c_i), the
    subtree masses, Newton's law for the centroidal momentum derivative (dhg.linear = M g + sum of the contact forces;
    a robot in free fall has dhg = (M g, 0) exactly) and hg.linear = M * d(com)/dt by finite differences."""
    from jiminy_b200 import scenarios
    sc = scenarios.make("anymal", 2, seed=2)
    rob = sc.robot
    q0 = sc.q0.copy()
    q0[1, 2] += 0.5                      # env 1 starts in the air: free fall, no contact
    v0 = np.zeros_like(sc.v0)
    v0[1, 3:6] = [0.4, -0.3, 0.2]        # ... tumbling
    orc = OracleBatch(rob, sc.options, 2)
    orc.set_pd_controller(sc.kp, sc.kd)
    orc.set_command(sc.target0)
    assert not orc.start(q0, v0).any()
    Mtot = rob.mass
    g = np.array(sc.options["world"]["gravity"][:3])
    coms = []
    for k in range(3):
        orc.step(1e-3)
        _, q, v, _ = orc.get_state()
        ycrb, com, vcom, hg, dhg = orc.get_centroidal()
        for e in range(2):
            oMi = R.forward_kinematics(rob, q[e])
            c = sum(rob.inertia[j, 0] * (oMi[j].R @ rob.inertia[j, 1:4] + oMi[j].p) for j in range(1, rob.njoints)) / Mtot
            np.testing.assert_allclose(com[e, 0], c, rtol=0, atol=1e-13)
            np.testing.assert_allclose(ycrb[e, 1, 0], Mtot, rtol=1e-14)
            np.testing.assert_allclose(hg[e, :3], Mtot * vcom[e, 0], rtol=0, atol=1e-11)
        # free fall: total external force = weight, no moment about the centre of mass
        np.testing.assert_allclose(dhg[1, :3], Mtot * g, rtol=0, atol=1e-9)
        np.testing.assert_allclose(dhg[1, 3:], 0.0, atol=1e-9)
        # standing: weight + contact forces (world frame) = d(hg)/dt
        fext = orc.get_efforts()[3][0]                      # per joint wrench in the joint frame
        oMi = R.forward_kinematics(rob, q[0])
        f_world = sum(oMi[j].R @ fext[j, :3] for j in range(1, rob.njoints))
        np.testing.assert_allclose(dhg[0, :3], Mtot * g + f_world, rtol=0, atol=1e-8 * max(1.0, np.abs(f_world).max()))
        coms.append((orc.get_state()[0][1], com[1, 0].copy(), vcom[1, 0].copy()))
    # finite difference of the free-falling CoM against vcom (world-frame velocity of the centre of mass)
    (t0, c0, _), (t1, c1, w1), (t2, c2, _) = coms
    np.testing.assert_allclose((c2 - c0) / (t2 - t0), w1, rtol=0, atol=1e-5)


# ---------------------------------------------------------------------------------------------------------------
# 12. unit_py/test_simple_pendulum.py:662-750 -- flexibility joint + rotor inertia vs a series-elastic actuator
def _flexible_pendulum(J, k, nu, inertia=1e-5):
    r = _pendulum(J)
    return M.add_flexibility_joints(r, [dict(frameName="PendulumJoint", stiffness=k * np.ones(3), damping=nu * np.ones(3),
                                             inertia=inertia * np.ones(3))])


def test_flexibility_model_like_the_reference_api_test():
    """test_simple_pendulum.py:815-842: `robot.flexibility_joint_indices == [1]`, the mechanical joint moves to index 2."""
    r = _flexible_pendulum(0.1, 1.0, 1.0, inertia=1.0)
    assert r.joint_names == ["universe", "PendulumJointFlexibility", "PendulumJoint"]
    assert [r.joint_index(n) for n in r.flexibility_joint_names] == [1]
    assert (r.nq, r.nv) == (5, 4) and r.parent.tolist() == [0, 0, 1]
    assert r.motors[0].joint == 2 and r.rotor_inertia.tolist() == [1.0, 1.0, 1.0, 0.1]
    np.testing.assert_array_equal(r.placement[2], M.SE3().flat())
    o = OracleBatch(r, _opt())
    ts, qs, vs, _ = o.simulate(0.1, r.neutral(), np.zeros(4))
    assert np.isfinite(qs).all()


def test_flexibility_armature_vs_series_elastic_actuator_like_the_reference_test():
    J, k, nu = 0.1, 20.0, 0.1
    k_control, nu_control = 100.0, 1.0
    r = _flexible_pendulum(J, k, nu)
    opt = _opt(tolAbs=TOL * 0.1, tolRel=TOL * 0.1)
    opt["world"]["gravity"] = [0.0] * 6
    o = OracleBatch(r, opt)
    o.set_callbacks(0, controller=lambda t, q, v, s, out: out.__setitem__(0, -k_control * q[4] - nu_control * v[3]))
    v_init = 0.1
    ts, qs, vs, _ = o.simulate(10.0, [0.0, 0.0, 0.0, 1.0, 0.0], [0.0, v_init, 0.0, 0.0])
    # quaternion -> angle about y; no motion about the other axes
    assert np.abs(qs[:, [0, 2]]).max() < 1e-12 and np.abs(vs[:, [0, 2]]).max() < 1e-12
    flex_angle = 2.0 * np.arctan2(qs[:, 1], qs[:, 3])
    x = np.c_[flex_angle, qs[:, 4], vs[:, 1], vs[:, 3]]
    I = 5.0 * 1.0 ** 2
    A = np.array([[0.0, 0.0, 1.0, 0.0],
                  [0.0, 0.0, 0.0, 1.0],
                  [-k * (1 / I + 1 / J), k_control / J, -nu * (1 / I + 1 / J), nu_control / J],
                  [k / J, -k_control / J, nu / J, -nu_control / J]])
    idx = np.linspace(0, len(ts) - 1, 80).astype(int)
    xa = np.stack([scipy.linalg.expm(A * t) @ x[0] for t in ts[idx]])
    np.testing.assert_allclose(x[idx], xa, atol=1e-4)       # the reference's own tolerance: the flexible element has inertia 1e-5, not 0


def _flexible_arm_urdf(path, n_segments, mass=0.1, inertia=0.001, length=0.01):
    """The procedural arm of unit_py/test_flexible_arm.py:32-79: one motorised revolute joint, then a chain of links held
    by fixed joints."""
    import xml.etree.ElementTree as ET
    robot = ET.Element("robot", name="flexible_arm")
    ET.SubElement(robot, "link", name="base")
    for i in range(n_segments):
        link = ET.SubElement(robot, "link", name=f"link{i}")
        inertial = ET.SubElement(link, "inertial")
        ET.SubElement(inertial, "origin", xyz=f"{length / 2} 0 0", rpy="0 0 0")
        ET.SubElement(inertial, "mass", value=f"{mass}")
        ET.SubElement(inertial, "inertia", ixx="0", ixy="0", ixz="0", iyy="0", iyz="0", izz=f"{inertia}")
    motor = ET.SubElement(robot, "joint", name="base_to_link0", type="revolute")
    ET.SubElement(motor, "parent", link="base")
    ET.SubElement(motor, "child", link="link0")
    ET.SubElement(motor, "origin", xyz="0 0 0", rpy=f"{np.pi / 2} 0 0")
    ET.SubElement(motor, "axis", xyz="0 0 1")
    ET.SubElement(motor, "limit", effort="100.0", lower=f"{-np.pi}", upper=f"{np.pi}", velocity="10.0")
    for i in range(1, n_segments):
        joint = ET.SubElement(robot, "joint", name=f"link{i - 1}_to_link{i}", type="fixed")
        ET.SubElement(joint, "parent", link=f"link{i - 1}")
        ET.SubElement(joint, "child", link=f"link{i}")
        ET.SubElement(joint, "origin", xyz=f"{length} 0 0", rpy="0 0 0")
    ET.ElementTree(robot).write(path)


def test_rigid_vs_flexibility_at_fixed_frames_like_the_reference_test(tmp_path):
    """unit_py/test_flexible_arm.py:178-259: with an extremely large flexibility inertia the arm split at its fixed frames
    moves like the rigid one (1e-5), whatever the order in which the flexibility joints are inserted."""
    n_flex = 12
    urdf = str(tmp_path / "flexible_arm.urdf")
    _flexible_arm_urdf(urdf, n_flex + 1)
    rigid = M.build_robot_table(urdf, False)
    M.attach_motor(rigid, "base_to_link0", "base_to_link0", enableVelocityLimit=False, enableEffortLimit=False)
    opt = _opt(tolAbs=1e-9, tolRel=1e-9)
    t_end = 1.0
    o = OracleBatch(rigid, opt)
    o.simulate(t_end, [0.0], [0.0], log=False)
    q_rigid = o.get_state()[1][0]
    assert abs(q_rigid[0]) > 0.1                    # the arm does fall
    tables = []
    for order in (range(n_flex), range(n_flex)[::-1]):
        flex = M.add_flexibility_joints(rigid, [dict(frameName=f"link{i}_to_link{i + 1}", stiffness=np.zeros(3),
                                                     damping=np.zeros(3), inertia=np.full(3, 1e6)) for i in order])
        tables.append(flex)
        assert flex.njoints == 2 + n_flex and flex.nq == 1 + 4 * n_flex
        np.testing.assert_allclose(flex.mass, rigid.mass, rtol=1e-14)
        of = OracleBatch(flex, opt)
        of.simulate(t_end, flex.neutral(), np.zeros(flex.nv), log=False)
        q = of.get_state()[1][0]
        # get_theoretical_position_from_extended: the mechanical joint's coordinate
        np.testing.assert_allclose(q[flex.idx_q[flex.joint_index("base_to_link0")]], q_rigid[0], atol=1e-5)
    a, b = tables
    assert a.joint_names == b.joint_names and a.parent.tolist() == b.parent.tolist()
    np.testing.assert_allclose(a.inertia, b.inertia, atol=1e-12)
    np.testing.assert_allclose(a.placement, b.placement, atol=1e-12)


# ---------------------------------------------------------------------------------------------------------------
# 13. unit_py/test_simple_pendulum.py:269-333 -- transmission backlash: free play, then one body with the rotor
def _backlash_pendulum(J=1.0, backlash=1.1):
    r = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    M.attach_motor(r, "PendulumJoint", "PendulumJoint", enableVelocityLimit=False, enableEffortLimit=False,
                   enableArmature=True, armature=J, enableBacklash=True, backlash=2 * backlash)
    return M.add_backlash_joints(r)


def _rk_reference(times, x0, dynamics):
    sol = scipy.integrate.solve_ivp(dynamics, (times[0], times[-1]), x0, t_eval=times, method="DOP853", rtol=1e-12, atol=1e-12)
    return sol.y.T


def test_backlash_like_the_reference_test():
    J, B, TAU, g, l, m = 1.0, 1.1, 5.0, 9.81, 1.0, 5.0
    r = _backlash_pendulum(J, B)
    assert r.joint_names == ["universe", "PendulumJoint", "PendulumJointBacklash"] and r.motors[0].joint == 1
    assert (r.q_lower[1], r.q_upper[1]) == (-B, B) and r.inertia[1, 0] == 0.0 and r.inertia[2, 0] == m
    opt = _opt(tolAbs=1e-9, tolRel=1e-9)
    opt["constraints"]["regularization"] = 0.0
    o = OracleBatch(r, opt)
    o.set_command(np.array([[-TAU]]))
    x0 = np.array([0.0, 0.1, 0.0, 0.0])
    ts, qs, vs, _ = o.simulate(5.0, x0[:2], x0[2:])
    x = np.c_[qs, vs]
    # phase 1: inside the backlash the rotor spins up alone and the pendulum falls freely (its angle is q0 + q1)
    # (the fixture's pendulum is the inverted one -- mass above the joint -- so it falls into the backlash earlier than the
    # reference's hanging pendulum, whose impact time is sqrt(2 B J / TAU): the impact is read off the trajectory)
    t_impact = ts[np.argmax(np.abs(qs[:, 1]) >= B - 1e-6)]
    assert 0.2 < t_impact < np.sqrt(B / (TAU / J) * 2)
    t1, t2 = np.searchsorted(ts, [t_impact - 0.02, t_impact + 0.4])

    def free(t, y):
        return np.array([y[2], y[3], -TAU / J, g / l * np.sin(y[0] + y[1]) + TAU / J])
    np.testing.assert_allclose(x[:t1], _rk_reference(ts[:t1], x0, free), atol=TOL)
    # phase 2: on the backlash limit both move as one body, rotor and body inertia summed up
    I_total = m * l ** 2 + J
    G = m * g * l / I_total

    def locked(t, y):
        return np.array([y[2], y[3], G * np.sin(y[0] + y[1]) - TAU / I_total, 0.0])
    # (as long as the transmission stays on its limit: the falling inverted pendulum eventually pulls it off again)
    off = np.nonzero(np.abs(qs[t2:, 1]) < B - 1e-4)[0]
    t3 = t2 + (off[0] if off.size else len(ts) - t2) - 3        # (the multiplier fades out over the last samples before the release)
    assert ts[t3 - 1] - ts[t2] > 0.25
    np.testing.assert_allclose(x[t2:t3], _rk_reference(ts[t2:t3] - ts[t2], x[t2], locked), atol=1e-5)
