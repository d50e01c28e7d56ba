"""Flexibility joints (Engine::computeInternalDynamics, core/src/engine/engine.cc:3367-3391; the spherical joints of
Model::addFlexibilityJointsToExtendedModel, core/src/robot/model.cc:1087-1165) on a jiminy_b200 BatchedEngine -- the
CUDA library in the `-m gpu` suite, the same kernel source under the warp emulator in the CPU suite -- against the
closed form of the reference's own test and against the oracle."""
import os

import numpy as np
import scipy.linalg

from jiminy_b200 import model as M
from jiminy_b200 import robots as R
from jiminy_b200 import scenarios
from jiminy_b200.core import BatchedEngine
from oracle.oracle import OracleBatch

from conftest import DATA
import parity_common as pc


def _opt(**stepper):
    opt = M.default_engine_options()
    opt["contacts"]["model"] = "spring_damper"
    opt["stepper"].update(stepper)
    return opt


def flexible_pendulum(J=0.1, k=20.0, nu=0.1, inertia=1e-5):
    r = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    M.attach_motor(r, "PendulumJoint", "PendulumJoint", enableVelocityLimit=False, enableEffortLimit=False,
                   enableArmature=True, armature=J)
    return M.add_flexibility_joints(r, [dict(frameName="PendulumJoint", stiffness=k * np.ones(3), damping=nu * np.ones(3),
                                             inertia=inertia * np.ones(3))])


def series_elastic_actuator(api=None, device=0):
    """unit_py/test_simple_pendulum.py:662-750 on the device: a flexibility in front of a motorised pendulum with rotor
    inertia behaves like a series-elastic actuator (1e-4: the flexible element has inertia 1e-5, not 0); the same run
    against the oracle, adaptive steps included."""
    J, k, nu, k_control, nu_control = 0.1, 20.0, 0.1, 100.0, 1.0
    r = flexible_pendulum(J, k, nu)
    opt = _opt(odeSolver="runge_kutta_dopri", tolAbs=1e-8, tolRel=1e-8)
    opt["world"]["gravity"] = [0.0] * 6
    v_init = np.array([0.1, -0.05, 0.2])
    n = len(v_init)
    q0 = np.tile([0.0, 0.0, 0.0, 1.0, 0.0], (n, 1))
    v0 = np.zeros((n, 4))
    v0[:, 1] = v_init
    eng = BatchedEngine(r, opt, n, device=device, api_=api)
    orc = OracleBatch(r, opt, n)
    for e in (eng, orc):
        # the reference's controller, command = -k_control q[4] - nu_control v[3], as the built-in joint spring-damper
        (e.set_joint_springs if e is eng else e.set_springs)([0.0, 0.0, 0.0, k_control], [0.0, 0.0, 0.0, nu_control])
    eng.start(q0, v0)
    assert not orc.start(q0, v0).any()
    I = 5.0
    A = np.array([[0.0, 0.0, 1.0, 0.0],
                  [0.0, 0.0, 0.0, 1.0],
                  [-k * (1 / I + 1 / J), k_control / J, -nu * (1 / I + 1 / J), nu_control / J],
                  [k / J, -k_control / J, nu / J, -nu_control / J]])
    worst, log = 0.0, []
    for _ in range(40):
        eng.step(0.05)
        assert not orc.step(0.05).any()
        (t1, q1, v1, a1), (t0, q0_, v0_, a0) = eng.get_state(), orc.get_state()
        np.testing.assert_allclose(t1, t0, rtol=0, atol=1e-15)
        np.testing.assert_array_equal(eng.get_iters()[0], orc.get_iters()[0])        # accepted AND rejected steps agree
        np.testing.assert_array_equal(eng.get_iters()[1], orc.get_iters()[1])
        worst = max(worst, np.abs(q1 - q0_).max(), np.abs(v1 - v0_).max())
        assert np.abs(q1[:, [0, 2]]).max() < 1e-12 and np.abs(v1[:, [0, 2]]).max() < 1e-12   # motion about y only
        log.append((t1.copy(), q1.copy(), v1.copy()))
    assert worst < 1e-10, worst
    pc.compare_extra_terms(eng, orc, 1e-11)
    # closed form (after the run: scipy's expm and the thread emulator of the CPU suite do not mix well in one process)
    for t1, q1, v1 in log[::4]:
        for e in range(n):
            x = np.array([2.0 * np.arctan2(q1[e, 1], q1[e, 3]), q1[e, 4], v1[e, 1], v1[e, 3]])
            xa = scipy.linalg.expm(A * t1[e]) @ np.array([0.0, 0.0, v_init[e], 0.0])
            np.testing.assert_allclose(x, xa, atol=1e-4)
    return worst


def flexible_branched_arm():
    """Every joint model of the path on a branched tree (tests/data/branched_arm.urdf) with flexibilities in front of a
    joint of each branch and at a fixed frame; contacts, IMU, force sensor, encoders."""
    robot = M.build_robot_table(os.path.join(DATA, "branched_arm.urdf"), True)
    robot.add_contact_points(["b_sole", "a_tool"])
    for jn in ("a_shoulder", "a_elbow", "a_spin", "b_hip", "b_slide", "b_skew_slide", "b_ankle_z", "c_spin_skew"):
        M.attach_motor(robot, jn, jn, enableVelocityLimit=(jn == "b_hip"), velocityEffortInvSlope=0.05,
                       enableArmature=True, armature=0.01)
        M.attach_sensor(robot, "EncoderSensor", jn, motor_name=jn)
        M.attach_sensor(robot, "EffortSensor", jn, motor_name=jn)
    M.attach_sensor(robot, "ImuSensor", "imu", frame_name="a_hand")
    M.attach_sensor(robot, "ForceSensor", "sole", frame_name="b_toe")
    M.attach_sensor(robot, "ContactSensor", "sole_c", frame_name="b_sole")
    fixed = [n for n, f in robot.frames.items() if f.kind == "fixed_joint"]
    cfg = [dict(frameName="a_elbow", stiffness=[500.0, 600.0, 700.0], damping=[5.0, 6.0, 7.0], inertia=[0.01, 0.02, 0.03]),
           dict(frameName="b_hip", stiffness=[800.0, 600.0, 700.0], damping=[5.0, 3.0, 7.0], inertia=[0.02, 0.02, 0.01])]
    if fixed:
        cfg.append(dict(frameName=fixed[0], stiffness=[300.0, 300.0, 400.0], damping=[2.0, 3.0, 2.0], inertia=[0.01, 0.01, 0.02]))
    flex = M.add_flexibility_joints(robot, cfg)
    opt = M.default_engine_options()
    opt["contacts"].update(model="spring_damper", stiffness=1e5, damping=5e2)
    opt["stepper"].update(odeSolver="runge_kutta_4", dtMax=5e-4, sensorsUpdatePeriod=2e-3, controllerUpdatePeriod=4e-3)
    return robot, flex, opt


def branched_arm_parity(api=None, device=0, solver="runge_kutta_4", n_steps=3, tol_state=1e-10, tol_sens=1e-8):
    """Single evaluations and a few env-steps of the flexible branched arm, everything `parity_common.compare` checks
    (state, sensors, iteration counters, energies, data.a / data.f, centroidal terms)."""
    robot, flex, opt = flexible_branched_arm()
    opt["stepper"]["odeSolver"] = solver
    rng = np.random.default_rng(5)
    n = 3
    q, v = pc.random_states(flex, n, rng, base_height=0.55)
    cmd = rng.uniform(-5, 5, size=(n, flex.nmotors))
    orc = OracleBatch(flex, opt, n)
    eng = BatchedEngine(flex, opt, n, device=device, api_=api)
    a0, f0, u0 = orc.compute_dynamics(q, v, cmd)
    a1, f1, u1 = eng.compute_dynamics(q, v, cmd)
    np.testing.assert_allclose(a1, a0, rtol=0, atol=1e-12 * max(1.0, np.abs(a0).max()))
    np.testing.assert_allclose(f1, f0, rtol=0, atol=1e-12 * max(1.0, np.abs(f0).max()))
    np.testing.assert_allclose(u1, u0, rtol=0, atol=1e-12 * max(1.0, np.abs(u0).max()))
    assert np.abs(u0[:, flex.idx_v[flex.joint_index("a_elbowFlexibility")]]).max() > 1e-3      # the flexibility does act
    for e in (orc, eng):
        e.set_command(cmd)
    eng.start(q, v)
    assert not orc.start(q, v).any()
    pc.compare(eng, orc, 1e-13, 1e-11)
    for k in range(n_steps):
        eng.step(4e-3)
        assert not orc.step(4e-3).any()
        pc.compare(eng, orc, tol_state, tol_sens)
    u1, u0 = eng.get_efforts()[0], orc.get_efforts()[0]
    np.testing.assert_allclose(u1, u0, rtol=0, atol=1e-9 * max(1.0, np.abs(u0).max()))
    return eng, orc


def flexible_pendulum_on_its_bounds(api=None, device=0, model="spring_damper", n_steps=60):
    """`parity_common.bounds_scenario` with a flexibility in front of the joint: the pendulum is thrown against its
    position bounds, the JointConstraint of the mechanical joint is solved with a spherical joint in the tree (joint-space
    inertia, Cholesky factor and M^-1 J^T over nv = 4), multiplier reported in u."""
    r = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    M.attach_motor(r, "PendulumJoint", "PendulumJoint", enableVelocityLimit=False, enableEffortLimit=False)
    r.q_upper[0], r.q_lower[0] = 0.5, -0.5
    r = M.add_flexibility_joints(r, [dict(frameName="PendulumJoint", stiffness=[400.0, 300.0, 500.0], damping=[2.0, 3.0, 2.5],
                                          inertia=[0.02, 0.03, 0.01])])
    opt = _opt(odeSolver="runge_kutta_4", dtMax=1e-3, controllerUpdatePeriod=1e-3, sensorsUpdatePeriod=1e-3)
    opt["contacts"]["model"] = model
    eng, orc = BatchedEngine(r, opt, 3, device=device, api_=api), OracleBatch(r, opt, 3)
    q0 = np.tile(r.neutral(), (3, 1))
    q0[:, 4] = [0.3, 0.1, -0.45]
    v0 = np.zeros((3, 4))
    v0[:, 3] = [0.0, 1.0, -2.0]
    for x in (eng, orc):
        x.set_command(np.zeros((3, 1)))
    eng.start(q0, v0)
    assert not orc.start(q0, v0).any()
    pc.compare(eng, orc, 1e-13, 1e-11)
    hit = np.zeros(3, dtype=bool)
    for _ in range(n_steps):
        eng.step(0.01)
        assert not orc.step(0.01).any()
        pc.compare(eng, orc, 1e-9, 1e-7)
        np.testing.assert_allclose(eng.get_efforts()[0], orc.get_efforts()[0], rtol=0, atol=1e-7)
        hit |= (eng.get_status() & 8) != 0
    assert hit.all()                                           # every env reached a bound (JB_ENV_JOINT_LIMIT)
    assert np.abs(eng.get_state()[1][:, 4]).max() < 0.5 + 5e-3
    return eng, orc


def flexible_anymal_parity(api=None, device=0, n_env=8, n_steps=2, tol_state=1e-9, tol_sens=1e-7, contact_model=None):
    """ANYmal standing under its PD controller with a flexibility in front of a joint of every leg: the spherical
    records sit inside the legs' private chains, four lanes per env; spring-damper ground, or (`contact_model`) the
    reference's default `constraint` contacts through the generic constraint solver."""
    sc = scenarios.make("anymal_flexible", n_env, seed=3, contact_model=contact_model)
    flex, q0, v0 = sc.robot, sc.q0, sc.v0
    eng = BatchedEngine(flex, sc.options, n_env, device=device, api_=api)
    orc = OracleBatch(flex, sc.options, n_env)
    for e in (eng, orc):
        e.set_pd_controller(sc.kp, sc.kd)
        e.set_command(sc.target0)
    eng.start(q0, v0)
    assert not orc.start(q0, v0).any()
    pc.compare(eng, orc, 1e-13, 1e-11)
    for k in range(n_steps):
        act = sc.sample_targets(k)
        eng.set_command(act)
        orc.set_command(act)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
        pc.compare(eng, orc, tol_state, tol_sens)
    qf = eng.get_state()[1]
    iq = flex.idx_q[flex.joint_index("LF_HFEFlexibility")]
    assert np.abs(qf[:, iq:iq + 3]).max() > 1e-6          # the flexibilities deform under the robot's weight
    return eng, orc


def backlash_pendulum_parity(api=None, device=0, n_steps=120):
    """Transmission backlash (`Robot::initializeExtendedModel`, robot.cc:582-629; `model.add_backlash_joints`): the
    reference's `test_backlash` system -- a motorised pendulum with rotor inertia and a backlash joint behind the motor --
    driven by a constant torque through the free play and onto the limit, device vs oracle.  The extra joint is an
    ordinary bounded joint for the engine: the bound is a JointConstraint of the constraint path."""
    J, B, TAU = 1.0, 1.1, 5.0
    r = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    M.attach_motor(r, "PendulumJoint", "PendulumJoint", enableVelocityLimit=False, enableEffortLimit=False,
                   enableArmature=True, armature=J, enableBacklash=True, backlash=2 * B)
    r = M.add_backlash_joints(r)
    opt = _opt(odeSolver="runge_kutta_4", dtMax=1e-3, controllerUpdatePeriod=1e-3, sensorsUpdatePeriod=1e-3)
    opt["constraints"]["regularization"] = 0.0
    n = 3
    eng, orc = BatchedEngine(r, opt, n, device=device, api_=api), OracleBatch(r, opt, n)
    cmd = np.array([[-TAU], [TAU], [-0.5 * TAU]])
    q0 = np.array([[0.0, 0.1], [0.0, -0.2], [0.1, 0.0]])
    for x in (eng, orc):
        x.set_command(cmd)
    eng.start(q0, np.zeros((n, 2)))
    assert not orc.start(q0, np.zeros((n, 2))).any()
    pc.compare(eng, orc, 1e-13, 1e-11)
    for _ in range(n_steps):
        eng.step(0.01)
        assert not orc.step(0.01).any()
        pc.compare(eng, orc, 1e-9, 1e-7)
    q = eng.get_state()[1]
    assert (np.abs(q[:2, 1]) > B - 1e-2).all() and (eng.get_status()[:2] & 8).all()      # on the limit: the bound has been active
    return eng, orc


def flexible_atlas_trunk_parity(api=None, device=0, n_env=3):
    """Atlas with flexibilities in front of a back joint (a TRUNK joint of the lane plan: the spherical record is walked by
    all the lanes and its articulated inertia all-reduced), a knee and a shoulder, deformed and moving at the start:
    single evaluation and one PD env-step against the oracle."""
    sc = scenarios.make("atlas", n_env, seed=1)
    rigid = sc.robot
    cfg = [dict(frameName=jn, stiffness=[8e3, 9e3, 7e3], damping=[40.0, 30.0, 35.0], inertia=[0.2, 0.3, 0.25])
           for jn in ("back_bky", "l_leg_kny", "r_arm_shx")]
    flex = M.add_flexibility_joints(rigid, cfg)
    q0, v0 = M.extended_state_from_theoretical(flex, rigid, sc.q0, sc.v0)
    rng = np.random.default_rng(0)
    for name in flex.flexibility_joint_names:
        j = flex.joint_index(name)
        iq, iv = flex.idx_q[j], flex.idx_v[j]
        w = rng.normal(size=(n_env, 3)) * 0.05
        ang = np.linalg.norm(w, axis=1, keepdims=True)
        q0[:, iq:iq + 3], q0[:, iq + 3] = np.sin(ang / 2) * w / ang, np.cos(ang / 2)[:, 0]
        v0[:, iv:iv + 3] = rng.normal(size=(n_env, 3)) * 0.3
    eng = BatchedEngine(flex, sc.options, n_env, device=device, api_=api)
    orc = OracleBatch(flex, sc.options, n_env)
    assert "ntrunk=5" in eng.describe()                     # pelvis, the three back joints and the flexibility
    cmd = rng.uniform(-20, 20, size=(n_env, flex.nmotors))
    a0, f0, u0 = orc.compute_dynamics(q0, v0, cmd)
    a1, f1, u1 = eng.compute_dynamics(q0, v0, cmd)
    np.testing.assert_allclose(a1, a0, rtol=0, atol=1e-11 * max(1.0, np.abs(a0).max()))
    np.testing.assert_allclose(u1, u0, rtol=0, atol=1e-11 * max(1.0, np.abs(u0).max()))
    for e in (eng, orc):
        e.set_pd_controller(sc.kp, sc.kd)
        e.set_command(sc.target0)
    eng.start(q0, v0)
    assert not orc.start(q0, v0).any()
    pc.compare(eng, orc, 1e-12, 1e-10)
    act = sc.sample_targets(0)
    eng.set_command(act)
    orc.set_command(act)
    eng.step(sc.step_dt)
    assert not orc.step(sc.step_dt, parallel=True).any()
    pc.compare(eng, orc, 1e-8, 1e-6)
    return eng, orc
