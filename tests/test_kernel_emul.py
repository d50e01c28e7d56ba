"""CPU-only checks of the *kernel source* (jiminy_b200/csrc) against the oracle, by running it under
the warp emulator of tests/emul (threads stand for lanes; shuffles are rendez-vous).  This is test
infrastructure: it proves the lane plan, the trunk all-reduce, the scheduler and the sensor code
before GPU time is spent.  The real parity gate is tests/test_gpu_parity.py (-m gpu)."""
import os

import numpy as np
import pytest

from jiminy_b200 import model as M
from jiminy_b200 import robots as R
from jiminy_b200 import scenarios
from jiminy_b200.core import BatchedEngine, plan_describe
from oracle.oracle import OracleBatch

from conftest import DATA
from emul import emul_api, host_has_fma
import parity_common as pc


@pytest.fixture(scope="module")
def api():
    return emul_api()


def _lanes(n):
    os.environ["JB_LANES"] = str(n)


@pytest.mark.parametrize("name,lanes", [("anymal", 0), ("anymal", 1), ("anymal", 2), ("atlas", 0), ("atlas", 8),
                                        ("cartpole", 0), ("double_pendulum", 0)])
def test_single_rhs_matches_oracle(api, name, lanes):
    """One Engine::computeRobotsDynamics evaluation (FK + contacts + motors + ABA) on random states."""
    _lanes(lanes)
    robot, opt = R.load_robot(name)
    opt = R.baseline_options(name, opt)
    rng = np.random.default_rng(3)
    n = 3
    q, v = pc.random_states(robot, n, rng)
    cmd = rng.uniform(-20, 20, size=(n, max(robot.nmotors, 1)))
    a0, f0, u0 = OracleBatch(robot, opt, n).compute_dynamics(q, v, cmd)
    eng = BatchedEngine(robot, opt, n, api_=api)
    a1, f1, u1 = eng.compute_dynamics(q, v, cmd)
    _lanes(0)
    np.testing.assert_allclose(a1, a0, rtol=0, atol=1e-12 * max(1.0, np.abs(a0).max()))
    np.testing.assert_allclose(f1, f0, rtol=0, atol=1e-12 * max(1.0, np.abs(f0).max()))
    np.testing.assert_allclose(u1, u0, rtol=0, atol=1e-12)


def test_all_joint_models_and_internal_branching(api):
    """Every joint model of the path (RX/RY/RZ/RU, RUB*, PZ/PU, free-flyer) on a branched tree, with a
    contact, an IMU, a force sensor and encoders; lanes = 1, 2 and automatic."""
    robot = M.build_robot_table(os.path.join(DATA, "branched_arm.urdf"), True)
    robot.add_contact_points(["b_sole", "a_tool"])
    for jn in ("a_shoulder", "a_elbow", "a_spin", "b_hip", "b_slide", "b_skew_slide", "b_ankle_z", "c_spin_skew"):
        M.attach_motor(robot, jn, jn, enableVelocityLimit=(jn == "b_hip"), velocityEffortInvSlope=0.05,
                       enableArmature=True, armature=0.01, enableFriction=(jn == "a_elbow"),
                       frictionViscousPositive=-0.1, frictionViscousNegative=-0.2, frictionDryPositive=-0.05,
                       frictionDryNegative=-0.07, frictionDrySlope=3.0)
        M.attach_sensor(robot, "EncoderSensor", jn, motor_name=jn)
        M.attach_sensor(robot, "EffortSensor", jn, motor_name=jn)
    M.attach_sensor(robot, "ImuSensor", "imu", frame_name="a_hand")
    M.attach_sensor(robot, "ForceSensor", "sole", frame_name="b_toe")
    M.attach_sensor(robot, "ContactSensor", "sole_c", frame_name="b_sole")
    opt = M.default_engine_options()
    opt["contacts"].update(model="spring_damper", stiffness=1e5, damping=5e2)
    opt["stepper"].update(odeSolver="runge_kutta_4", dtMax=5e-4, sensorsUpdatePeriod=2e-3, controllerUpdatePeriod=4e-3)
    rng = np.random.default_rng(5)
    n = 2
    q, v = pc.random_states(robot, n, rng, base_height=0.55)
    cmd = rng.uniform(-5, 5, size=(n, robot.nmotors))
    for lanes in (0, 1, 2):
        _lanes(lanes)
        orc = OracleBatch(robot, opt, n)
        eng = BatchedEngine(robot, opt, n, api_=api)
        for e in (orc, eng):
            e.set_command(cmd)
        a0, f0, u0 = orc.compute_dynamics(q, v, cmd)
        a1, f1, u1 = eng.compute_dynamics(q, v, cmd)
        np.testing.assert_allclose(a1, a0, rtol=0, atol=1e-11 * max(1.0, np.abs(a0).max()))
        np.testing.assert_allclose(u1, u0, rtol=0, atol=1e-12)
        assert not orc.start(q, v).any()
        eng.start(q, v)
        pc.compare(eng, orc, 1e-13, 1e-11)
        for _ in range(3):
            eng.step(0.01)
            assert not orc.step(0.01).any()
            pc.compare(eng, orc, 1e-10, 1e-8)
    _lanes(0)


def test_anymal_pd_env_steps(api):
    pc.run_scenario("anymal", 3, 2, api=api)


def test_anymal_torque_mode_euler(api):
    """Zero-order-held effort commands, explicit Euler at 1e-4 (the alternative profile of SURVEY.md 8d)."""
    sc = scenarios.make("anymal", 2, dt_max=1e-4, solver="euler_explicit")
    sc.kp = None
    rng = np.random.default_rng(2)
    sc.target0 = rng.uniform(-10, 10, size=(2, 12))
    eng, orc = pc.make_pair(sc, api)
    eng.step(0.005)
    assert not orc.step(0.005).any()
    pc.compare(eng, orc, 1e-10, 1e-8)


def test_atlas_pd_env_step(api):
    pc.run_scenario("atlas", 2, 1, api=api)


def test_cartpole_and_double_pendulum(api):
    pc.run_scenario("cartpole", 5, 6, api=api, tol_state=1e-13, tol_sens=1e-12)
    pc.run_scenario("double_pendulum", 2, 20, api=api, tol_state=1e-13, tol_sens=1e-12)


def test_masked_restart_and_odd_env_count(api):
    """jb_start with a mask restarts only the selected envs (batched reset); n_env not a multiple of a warp."""
    sc = scenarios.make("anymal", 9)
    eng, orc = pc.make_pair(sc, api)
    eng.step(sc.step_dt)
    orc.step(sc.step_dt, parallel=True)
    mask = np.zeros(9, dtype=np.uint8)
    mask[[1, 8]] = 1
    eng.start(sc.q0, sc.v0, mask=mask)
    orc.start(sc.q0, sc.v0, mask=mask)
    t = eng.get_state()[0]
    assert t[1] == 0.0 and t[8] == 0.0 and t[0] == pytest.approx(0.04)
    pc.compare(eng, orc, 1e-9, 1e-7)
    eng.step(sc.step_dt)
    orc.step(sc.step_dt, parallel=True)
    pc.compare(eng, orc, 1e-9, 1e-7)
    it = eng.get_iters()[0]
    assert it[1] == 41 and it[0] == 81       # restarted envs do the 1 us first step again


def test_status_flags(api):
    """Joint bound violation raises JB_ENV_JOINT_LIMIT on both sides; step before start is a control-flow error."""
    from jiminy_b200.core import BadControlFlow, JB_ENV_JOINT_LIMIT
    robot, opt = R.load_robot("anymal")
    opt = R.baseline_options("anymal", opt)
    eng = BatchedEngine(robot, opt, 1, api_=api)
    with pytest.raises(BadControlFlow):
        eng.step(0.04)
    q = R.ground_base_height(robot, robot.neutral())
    q[robot.idx_q[robot.joint_index("LF_HAA")]] = 0.48      # upper bound 0.49
    v = np.zeros(robot.nv)
    v[robot.idx_v[robot.joint_index("LF_HAA")]] = 3.0
    orc = OracleBatch(robot, opt, 1)
    eng.start(q, v)
    orc.start(q, v)
    eng.set_command(np.zeros((1, robot.nmotors)))
    orc.set_command(np.zeros((1, robot.nmotors)))
    pc.compare(eng, orc, 1e-13, 1e-11)
    # the joint hits its bound inside the first step: the fast kernel hands the env over to the full kernel
    # (constraint path), which keeps it until the bound constraint switches off again
    for _ in range(6):
        eng.step(0.01)
        assert not orc.step(0.01).any()
        pc.compare(eng, orc, 1e-9, 1e-7)
    assert eng.get_status()[0] & JB_ENV_JOINT_LIMIT and orc.get_status()[0] & JB_ENV_JOINT_LIMIT
    bad = q.copy()
    bad[robot.idx_q[robot.joint_index("LF_HAA")]] = 0.6
    with pytest.raises(ValueError):
        eng.start(bad, v)


def test_lane_plans(api):
    for name, lanes, expect in (("anymal", 0, 4), ("atlas", 0, 4), ("cartpole", 0, 1)):
        robot, _ = R.load_robot(name)
        text, joint_lane = plan_describe(robot, lanes, api)
        assert f"lanes={expect}" in text
    robot, _ = R.load_robot("anymal")
    _, jl = plan_describe(robot, 0, api)
    assert jl[1] == -1 and sorted(set(jl[2:].tolist())) == [0, 1, 2, 3]     # root is trunk, one leg per lane
    assert len({int(jl[robot.joint_index(f"LF_{s}")]) for s in ("HAA", "HFE", "KFE")}) == 1


def test_batched_env_reset_step_autoreset(api):
    """BatchedJiminyEnv: gym-style reset/step over N envs with automatic masked restart of finished envs."""
    from jiminy_b200.envs import BatchedJiminyEnv
    sc = scenarios.make("anymal", 4)
    env = BatchedJiminyEnv(sc, api_=api, simulation_duration_max=0.16)
    obs, _ = env.reset()
    assert obs["states"]["agent"]["q"].shape == (4, 19) and obs["measurements"]["ImuSensor"].shape == (4, 6, 1)
    assert obs["measurements"]["EncoderSensor"].shape == (4, 2, 12)
    obs, rew, term, trunc, info = env.step(sc.sample_targets(0))
    assert rew.shape == (4,) and not term.any() and not trunc.any()
    np.testing.assert_allclose(obs["t"], 0.04)
    for k in (1, 2):
        obs, rew, term, trunc, info = env.step(sc.sample_targets(k))
    # settled on the ground after 0.12 s: the foot force sensors see a push of the order of the weight
    fz = np.abs(obs["measurements"]["ForceSensor"][:, :3, :]).sum(axis=(1, 2))
    assert np.all(fz > 0.2 * sc.robot.mass * 9.81)
    obs, rew, term, trunc, info = env.step(sc.sample_targets(3))
    assert trunc.all()                              # duration limit reached -> every env is restarted
    np.testing.assert_allclose(env.engine.get_state()[0], 0.0)
    # the returned observation is the one after the restart; the terminal one is kept aside, untouched by the restart
    np.testing.assert_allclose(obs["t"], 0.0)
    fin = info["final_observation"]
    assert info["_final_observation"].all() and np.allclose(fin["t"], 0.16)
    assert np.abs(fin["measurements"]["ForceSensor"][:, :3, :]).sum() > 0.0
    assert not np.shares_memory(fin["measurements"]["ForceSensor"], obs["measurements"]["ForceSensor"])
    env.close()


@pytest.mark.parametrize("period", [0.0, 1e-3])
def test_dopri_double_pendulum_and_energy(api, period):
    """Adaptive Dormand-Prince on the device: same accepted / rejected step sequence as the oracle
    (iteration counters equal), state within 1e-10, continuous and discrete (1 ms) modes."""
    robot, opt = R.load_robot("double_pendulum")
    opt = R.baseline_options("double_pendulum", opt)
    opt["stepper"].update(odeSolver="runge_kutta_dopri", tolAbs=1e-9, tolRel=1e-9, dtMax=0.02,
                          sensorsUpdatePeriod=period, controllerUpdatePeriod=period)
    eng, orc = BatchedEngine(robot, opt, 2, api_=api), OracleBatch(robot, opt, 2)
    q0, v0 = np.array([[0.0, 0.1], [0.3, -0.2]]), np.array([[0.0, 0.0], [0.5, 0.1]])
    eng.start(q0, v0)
    assert not orc.start(q0, v0).any()
    for _ in range(10):
        eng.step(0.02)
        assert not orc.step(0.02).any()
        pc.compare(eng, orc, 1e-10, 1e-9)
        np.testing.assert_array_equal(eng.get_iters()[1], orc.get_iters()[1])      # same number of rejected steps


def test_dopri_free_flyer_contact_and_unbounded_joint(api):
    """DOPRI with the SE(3) and SO(2) difference operators in the error norm: branched arm (free-flyer,
    continuous joints, prismatic) dropping on the ground with the engine's default tolerances."""
    robot = M.build_robot_table(os.path.join(DATA, "branched_arm.urdf"), True)
    robot.add_contact_points(["b_sole", "a_tool"])
    opt = M.default_engine_options()
    opt["contacts"].update(model="spring_damper", stiffness=1e5, damping=5e2)
    opt["stepper"].update(odeSolver="runge_kutta_dopri", sensorsUpdatePeriod=5e-3, controllerUpdatePeriod=5e-3)
    rng = np.random.default_rng(9)
    q, v = pc.random_states(robot, 2, rng, base_height=0.5)
    for lanes in (0, 1):
        _lanes(lanes)
        eng, orc = BatchedEngine(robot, opt, 2, api_=api), OracleBatch(robot, opt, 2)
        eng.start(q, v)
        assert not orc.start(q, v).any()
        for _ in range(6):
            eng.step(0.01)
            assert not orc.step(0.01).any()
        pc.compare(eng, orc, 1e-8, 1e-6)
        np.testing.assert_array_equal(eng.get_iters()[1], orc.get_iters()[1])
    _lanes(0)


def test_engine_facade_python_controller(api):
    """`Engine` facade: reference method names, in-place RobotState buffers, Python controller called
    back once per controller period (spring-damper law written as a controller, vs the analytic solution
    of test_double_spring_mass.py)."""
    import scipy.linalg
    from jiminy_b200.core import Engine, FunctionalController, BadControlFlow
    robot = M.build_robot_table(os.path.join(DATA, "linear_two_masses.urdf"), False)
    for j in ("FirstJoint", "SecondJoint"):
        M.attach_motor(robot, j, j, enableVelocityLimit=False, enableEffortLimit=False)
    k, nu, m = np.array([200.0, 20.0]), np.array([0.1, 0.2]), np.array([1.0, 2.5])
    calls = []

    def compute_command(t, q, v, sensors, command):
        calls.append(t)
        command[:] = -k * q - nu * v

    engine = Engine(api_=api)
    engine.add_robot(robot, FunctionalController(compute_command))
    opt = engine.get_options()
    opt["stepper"].update(odeSolver="runge_kutta_4", dtMax=1e-4, controllerUpdatePeriod=1e-3, sensorsUpdatePeriod=1e-3)
    engine.set_options(opt)
    with pytest.raises(BadControlFlow):
        engine.step(0.01)
    x0 = np.array([0.1, -0.1, 0.0, 0.0])
    engine.start(x0[:2], x0[2:])
    q_view = engine.robot_states[0].q
    assert engine.is_simulation_running
    for _ in range(20):
        engine.step(0.005)
    assert q_view is engine.robot_states[0].q and engine.stepper_state.t == pytest.approx(0.1)
    assert len(calls) >= 100
    # discrete 1 kHz zero-order-held spring law ~ continuous law up to O(period)
    Iq = 1.0 / m[1] + 1.0 / m[0]
    A = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [-k[0] / m[0], k[1] / m[0], -nu[0] / m[0], nu[1] / m[0]],
                  [k[0] / m[0], -k[1] * Iq, nu[0] / m[0], -nu[1] * Iq]])
    xa = scipy.linalg.expm(A * 0.1) @ x0
    np.testing.assert_allclose(np.r_[q_view, engine.robot_states[0].v], xa, atol=2e-2)
    engine.stop()
    assert not engine.is_simulation_running


def test_device_code_vs_closed_forms(api):
    """The reference's analytical tests on the kernel source itself (no oracle in the loop)."""
    import analytic_device as ad
    ad.armature_spring(api)
    ad.foot_pendulum(api, t_end=0.01)            # (the full second at dtMax = 1e-5 is run on the oracle)
    # (ad.velocity_bounds and ad.joint_position_limits: 15 s / 3 min under the emulator -- in the GPU suite; the oracle is
    # pinned by the same criteria, and the emulator ran both once when they were written)
    ad.two_masses(api, t_end=0.25)               # the emulator is slow: shorter horizons than the GPU suite
    ad.contact_equilibrium_and_friction(api)
    ad.energy_conservation(api)
    ad.force_impulse(api, t_end=0.45)            # covers the first six impulses
    ad.constraint_closed_forms(api)


def test_external_forces_anymal(api):
    pc.external_forces_scenario(api)


@pytest.mark.parametrize("toggle", [None, "JB_NO_STRUCTURED_CONS"])
def test_external_forces_with_constraint_contacts(api, monkeypatch, toggle):
    """Impulse / profile forces while the contact constraints are solved: register-resident quadruped solver, and
    the body-space solver when it is switched off."""
    if toggle:
        monkeypatch.setenv(toggle, "1")
    pc.external_forces_scenario(api, n_env=2, n_steps=2, solver="euler_explicit", dt_max=0.005, contact_model="constraint", tol=1e-8)


def test_external_forces_control_flow(api):
    sc = scenarios.make("cartpole", 2)
    eng = BatchedEngine(sc.robot, sc.options, 2, api_=api)
    with pytest.raises(ValueError):
        eng.register_impulse_force("universe", 0.0, 1e-3, np.zeros(6))
    with pytest.raises(ValueError):
        eng.register_impulse_force("no_such_frame", 0.0, 1e-3, np.zeros(6))
    fr = next(n for n, f in sc.robot.frames.items() if f.joint > 0)
    with pytest.raises(ValueError):
        eng.register_impulse_force(fr, 0.0, 1e-12, np.zeros(6))       # duration below STEPPER_MIN_TIMESTEP
    with pytest.raises(ValueError):
        eng.register_impulse_force(fr, -1.0, 1e-3, np.zeros(6))
    eng.set_command(sc.target0)
    eng.start(sc.q0, sc.v0)
    from jiminy_b200.core import BadControlFlow
    with pytest.raises(BadControlFlow):
        eng.register_impulse_force(fr, 0.0, 1e-3, np.zeros(6))
    with pytest.raises(BadControlFlow):
        eng.remove_all_forces()
    eng.stop()
    assert (eng.get_status() & 16).all()
    eng.register_impulse_force(fr, 0.0, 1e-3, np.zeros(6))
    eng.remove_all_forces()


def test_engine_facade_impulse_forces(api):
    """`Engine.register_impulse_force(robot_name, frame_name, t, dt, F)` + `simulate`, read like
    test_simple_pendulum.py:540-605 (first three forces, 0.25 s)."""
    import analytic_device as ad
    from jiminy_b200.core import Engine, BadControlFlow
    robot = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    engine = Engine(api_=api)
    engine.add_robot(robot)
    opt = engine.get_options()
    opt["world"]["gravity"] = np.zeros(6)
    opt["stepper"].update(sensorsUpdatePeriod=0.0, controllerUpdatePeriod=0.0)
    engine.set_options(opt)
    for f in ad.IMPULSES:
        engine.register_impulse_force("", "PendulumLink", f["t"], f["dt"], np.array(f["F"]))
    with pytest.raises(ValueError):
        engine.register_impulse_force("", "universe", 0.0, 1e-3, np.zeros(6))
    engine.start(np.zeros(1), np.zeros(1))
    with pytest.raises(BadControlFlow):
        engine.register_impulse_force("", "PendulumLink", 0.0, 1e-3, np.zeros(6))
    ts, xs = [], []
    while engine.stepper_state.t < 0.25 - 1e-9:
        engine.step(1e-3)
        ts.append(engine.stepper_state.t)
        xs.append([engine.robot_states[0].q[0], engine.robot_states[0].v[0]])
    xa = ad.pendulum_impulse_reference(np.array(ts))
    np.testing.assert_allclose(np.array(xs), xa, atol=1e-6)
    engine.stop()
    engine.reset(remove_all_forces=True)
    assert engine.impulse_forces == []
    engine.start(np.zeros(1), np.zeros(1))
    engine.step(0.01)
    assert abs(engine.robot_states[0].v[0]) < 1e-14


def test_engine_facade_profile_force_function(api):
    """`Engine.register_profile_force(robot_name, frame_name, force_func, update_period)`: the Python function is sampled at
    its update period (an integration breakpoint) and held in between -- against the oracle driven by hand the same way."""
    from jiminy_b200.core import Engine
    robot = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    engine = Engine(api_=api)
    engine.add_robot(robot)
    opt = engine.get_options()
    opt["stepper"].update(sensorsUpdatePeriod=0.0, controllerUpdatePeriod=0.0, odeSolver="runge_kutta_4", dtMax=1e-3)
    engine.set_options(opt)
    calls = []

    def force(t, q, v, out):
        calls.append(t)
        out[:] = [30.0 * np.sin(40.0 * t), 0.0, 5.0 * q[0], 0.0, -2.0 * v[0], 0.0]
    with pytest.raises(NotImplementedError):
        engine.register_profile_force("", "PendulumLink", force, 0.0)
    engine.register_profile_force("", "PendulumLink", force, 2e-3)
    fr = robot.frames["PendulumLink"]
    orc = OracleBatch(robot, opt, 1)
    slot = orc.register_profile_force(fr.joint, fr.placement.p, 2e-3)
    q0, v0 = np.array([[0.2]]), np.array([[-0.5]])
    w = np.zeros(6)
    force(0.0, q0[0], v0[0], w)
    orc.set_profile_force(slot, w[None, :])
    assert not orc.start(q0, v0).any()
    engine.start(q0[0], v0[0])
    for k in range(10):
        engine.step(5e-3)                      # not a multiple of the force period: breakpoints fall inside the steps
        t = 5e-3 * k
        while t < 5e-3 * (k + 1) - 1e-12:      # the oracle, stopped by hand at every multiple of 2 ms
            _, q, v, _ = orc.get_state()
            if abs(t / 2e-3 - round(t / 2e-3)) < 1e-9:
                force(t, q[0], v[0], w)
                orc.set_profile_force(slot, w[None, :])
            h = min(5e-3 * (k + 1), (np.floor(t / 2e-3 + 1e-9) + 1.0) * 2e-3) - t
            assert not orc.step(h).any()
            t += h
        _, q, v, a = orc.get_state()
        np.testing.assert_allclose([engine.robot_states[0].q[0], engine.robot_states[0].v[0]], [q[0, 0], v[0, 0]], rtol=0, atol=1e-10)
    assert abs(engine.robot_states[0].v[0] + 0.5) > 1e-2 and len(calls) > 40
    engine.stop()
    engine.remove_all_forces()
    assert engine._profile_forces == []


def test_engine_facade_telemetry_log(api, tmp_path):
    """`Engine.log_data` / `Engine.write_log` (binary format of the reference): one line at start and per step."""
    from jiminy_b200.core import Engine, BadControlFlow
    from jiminy_b200 import telemetry as T
    robot = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    engine = Engine(api_=api)
    engine.add_robot(robot)
    with pytest.raises(BadControlFlow):
        engine.write_log(str(tmp_path / "none.data"))
    opt = engine.get_options()
    opt["stepper"].update(sensorsUpdatePeriod=0.0, controllerUpdatePeriod=0.0)
    opt["telemetry"]["enableEnergy"] = True
    engine.set_options(opt)
    engine.start(np.array([0.3]), np.zeros(1))
    qs = [engine.robot_states[0].q[0]]
    for _ in range(5):
        engine.step(1e-3)
        qs.append(engine.robot_states[0].q[0])
    path = str(tmp_path / "pendulum.data")
    engine.write_log(path)
    log = T.read_log(path)
    np.testing.assert_allclose(log["variables"]["Global.Time"], 1e-3 * np.arange(6), atol=1e-12)
    np.testing.assert_array_equal(log["variables"]["currentPositionPendulum"], np.array(qs))
    e = log["variables"]["energy"]
    assert np.abs(e - e[0]).max() < 1e-8                      # RK4 at 1 ms conserves the pendulum's energy
    assert engine.log_data["variables"].keys() == log["variables"].keys()
    with pytest.raises(NotImplementedError):
        engine.write_log(path, format="hdf5")


@pytest.mark.parametrize("model", ["spring_damper", "constraint"])
def test_joint_bounds_constraint_path(api, model):
    pc.bounds_scenario(api, DATA, model)


def test_start_on_joint_bounds(api):
    pc.start_on_bounds_scenario(api, DATA)


def test_constraint_contact_point_mass(api):
    pc.point_mass_constraint_scenario(api, DATA)
    pc.point_mass_constraint_scenario(api, DATA, n_steps=15, torsion=0.05)
    pc.point_mass_constraint_scenario(api, DATA, n_steps=20, solver="runge_kutta_dopri")     # adaptive steps + PGS
    pc.point_mass_constraint_scenario(api, DATA, n_steps=20, impulse=True)                   # external forces + PGS


def test_constraint_contact_anymal(api):
    # (the structured solver exchanges through shuffles, which the thread emulator pays dearly: small case here,
    # 40 envs x 3 steps in the GPU suite)
    eng, orc, sc = pc.robot_constraint_scenario("anymal", 3, 1, api, seed=2)
    assert (eng.get_state()[1][:, 2] > 0.4).all()      # still standing


def test_constraint_contact_atlas_rhs(api):
    """Atlas (12 contact points, 30 bounded joints, 78 constraint rows at most): start + a short step."""
    sc = scenarios.make("atlas", 1, seed=1)
    sc.options["contacts"]["model"] = "constraint"
    eng, orc = pc.make_pair(sc, api)
    pc.compare(eng, orc, 1e-12, 1e-8)
    eng.step(0.005)
    assert not orc.step(0.005).any()
    pc.compare(eng, orc, 1e-9, 1e-6)
    # the reference's own Atlas settings (atlas_options.toml, pipeline_benchmark.py): explicit Euler at 5 ms
    pc.robot_constraint_scenario("atlas", 1, 1, api, seed=2, solver="euler_explicit", dt_max=0.005)


@pytest.mark.parametrize("robot,toggle", [("atlas", None), ("atlas", "JB_NO_BODY_CONS"), ("atlas", "JB_NO_BLOCK_CONS"),
                                          ("anymal", "JB_NO_STRUCTURED_CONS"), ("atlas", "torsion")])
def test_constraint_solver_variants(api, monkeypatch, robot, toggle):
    """Every device formulation of the constraint solve against the oracle: body-space contact solver (default for
    Atlas; for ANYmal once the register-resident quadruped solver is switched off), lane-block solver, dense generic."""
    torsion = 0.05 if toggle == "torsion" else None      # torsional friction block of the sweep
    if toggle and torsion is None:
        monkeypatch.setenv(toggle, "1")
    eng, orc, sc = pc.robot_constraint_scenario(robot, 2, 1, api, seed=3, torsion=torsion, solver="euler_explicit", dt_max=0.005)
    want = {None: "body-space", "JB_NO_BODY_CONS": "lane-block", "JB_NO_BLOCK_CONS": "generic", "JB_NO_STRUCTURED_CONS": "body-space", "torsion": "body-space"}[toggle]
    assert want in eng.describe() and (toggle != "JB_NO_BODY_CONS" or "body-space" not in eng.describe())


@pytest.mark.parametrize("lanes,toggle,solver", [(0, None, "body-space"), (0, "JB_NO_BODY_CONS", "lane-block"), (1, None, "generic")])
def test_constraint_contacts_all_joint_models(api, monkeypatch, lanes, toggle, solver):
    """The branched arm (free-flyer, revolute about arbitrary axes, unbounded revolute, prismatic incl. a skewed axis)
    lying on two contact frames with `contacts.model = "constraint"`, torsion on: Jacobians of every joint model
    through each formulation of the constraint solve."""
    if toggle:
        monkeypatch.setenv(toggle, "1")
    robot = M.build_robot_table(os.path.join(DATA, "branched_arm.urdf"), True)
    robot.add_contact_points(["b_sole", "a_tool"])
    for jn in ("a_shoulder", "a_elbow", "a_spin", "b_hip", "b_slide", "b_skew_slide", "b_ankle_z", "c_spin_skew"):
        M.attach_motor(robot, jn, jn, enableArmature=True, armature=0.01)
        M.attach_sensor(robot, "EncoderSensor", jn, motor_name=jn)
    M.attach_sensor(robot, "ForceSensor", "sole", frame_name="b_toe")
    M.attach_sensor(robot, "ContactSensor", "sole_c", frame_name="b_sole")
    opt = M.default_engine_options()
    opt["contacts"].update(model="constraint", friction=0.7, torsion=0.02)
    opt["stepper"].update(odeSolver="runge_kutta_4", dtMax=1e-3, sensorsUpdatePeriod=2e-3, controllerUpdatePeriod=4e-3)
    rng = np.random.default_rng(5)
    n = 3
    q, v = pc.random_states(robot, n, rng, base_height=0.12)
    cmd = rng.uniform(-3, 3, size=(n, robot.nmotors))
    _lanes(lanes)
    try:
        eng, orc = BatchedEngine(robot, opt, n, api_=api), OracleBatch(robot, opt, n)
        assert solver in eng.describe()
        for e in (eng, orc):
            e.set_command(cmd)
        eng.start(q, 0.3 * v)
        assert not orc.start(q, 0.3 * v).any()
        f_max = 0.0
        for _ in range(25):
            eng.step(0.004)
            assert not orc.step(0.004).any()
            pc.compare(eng, orc, 1e-9, 1e-7)
            f_ref = orc.get_efforts()[3]
            np.testing.assert_allclose(eng.get_efforts()[3], f_ref, rtol=0, atol=1e-8 * max(1.0, np.abs(f_ref).max()))
            f_max = max(f_max, np.abs(f_ref).max())
        assert f_max > 50.0          # the contact constraints did carry the robot
    finally:
        _lanes(0)


@pytest.mark.parametrize("robot", ["atlas", "anymal"])
def test_dopri_with_constraint_contacts(api, robot):
    """Adaptive Dormand-Prince steps (error control, rejected steps) over the body-space contact solver."""
    eng, orc, sc = pc.robot_constraint_scenario(robot, 1, 1, api, seed=2, solver="runge_kutta_dopri")
    assert "body-space" in eng.describe()
    it, it_failed = eng.get_iters()
    assert it[0] > 20


@pytest.mark.parametrize("robot", ["atlas", "anymal"])
def test_masked_restart_with_constraint_contacts(api, robot):
    pc.masked_restart_constraint_scenario(api, robot)


def test_gpu_like_rounding_stays_within_the_gpu_tolerances():
    """The kernel source built with contracted multiply-adds (what nvcc does for the device): the deviation from the
    oracle, which is built without contraction, must stay far inside the tolerances of the `-m gpu` suite -- in
    particular no Gauss-Seidel stopping decision may be so close to its threshold that rounding flips it visibly."""
    from emul import host_has_fma
    if not host_has_fma():
        pytest.skip("host CPU without FMA")
    api_fma = emul_api(fma=True)
    pc.atlas_bounds_and_contacts_scenario(api_fma, n_env=2, n_steps=4, tol_state=1e-10, tol_sens=1e-8)
    pc.robot_constraint_scenario("anymal", 4, 1, api_fma, seed=2, tol_state=1e-10, tol_sens=1e-8)
    pc.pd_adapter_scenario(api_fma, n_env=4, n_steps=2, order=0)


def test_atlas_pd_standing_first_steps(api):
    """The reference's Atlas PD-standing test (tests/test_oracle_analytic.py, 9 s on the oracle): its first 0.4 s on the
    device path against the oracle -- neutral posture with knees and shoulders on their bounds, full block pipeline."""
    pc.atlas_pd_standing_on_device(api, 0.4)


def test_restart_is_exactly_repeatable(api):
    pc.atlas_repeatability_scenario(api, steps=(0, 2, 4, 0))


def test_atlas_bounds_and_contacts_together(api):
    pc.atlas_bounds_and_contacts_scenario(api)


@pytest.mark.parametrize("drop_a_foot,solver", [(False, "lane-block"), (True, "body-space")])
def test_constraint_contact_on_trunk_body(api, drop_a_foot, solver):
    """A contact frame on the floating base -- a trunk joint, replicated on every lane -- next to the feet: five contact
    bodies go to the lane-block solver, four (one foot removed) to the body-space solver."""
    sc = scenarios.make("anymal", 2, seed=4, solver="euler_explicit", dt_max=0.005)
    sc.options["contacts"]["model"] = "constraint"
    rob = sc.robot
    base = [n for n, f in rob.frames.items() if f.joint == 1 and f.kind == "body"][0]
    z0 = float(sc.q0[:, 2].min())
    rob.add_frame("belly_contact", base, M.SE3(np.eye(3), np.array([0.05, 0.02, -(z0 + 0.002)])))     # 2 mm into the ground
    if drop_a_foot:
        rob.remove_contact_points([rob.contact_frame_names[0]])
    rob.add_contact_points(["belly_contact"])
    eng, orc = pc.make_pair(sc, api)
    assert solver in eng.describe() and "quadruped" not in eng.describe()
    pc.compare(eng, orc, 1e-12, 1e-9)
    for k in range(2):
        act = sc.sample_targets(k)
        eng.set_command(act)
        orc.set_command(act)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt).any()
        pc.compare(eng, orc, 1e-9, 1e-7)


@pytest.mark.parametrize("safety", [False, True])
def test_pd_controller_block(api, safety):
    pc.pd_block_scenario(api, safety=safety)


@pytest.mark.parametrize("order,instantaneous", [(0, False), (1, True)])
def test_pd_adapter_pipeline(api, order, instantaneous):
    pc.pd_adapter_scenario(api, order=order, instantaneous=instantaneous)


def test_pd_control_pipeline_env(api):
    """`PDControlBatchedEnv` = MotorSafetyLimit + PDController + PDAdapter(order 1) + MahonyFilter with the arguments of
    `AtlasPDControlJiminyEnv` (atlas.py:239-295), on ANYmal: derived bounds, observation layout, and the trajectories
    against the oracle driven by the same adapter function."""
    from jiminy_b200.envs import PDControlBatchedEnv, flatten_observation
    from jiminy_b200.blocks import pd_adapter
    from jiminy_b200._ctypes_abi import safety_table
    sc = scenarios.make("anymal", 3, seed=12)
    env = PDControlBatchedEnv(sc, joint_velocity_limit=4.0, joint_acceleration_limit=30.0, order=1, mahony=(0.75, 0.057),
                              safety=dict(kp=50.0, kd=0.15, soft_position_margin=0.0, soft_velocity_max=4.0), api_=api)
    rob, nm = sc.robot, sc.robot.nmotors
    v_hw = np.array([m.velocity_limit for m in rob.motors])
    np.testing.assert_array_equal(env.command_state_upper[1], np.minimum(v_hw, 4.0))
    np.testing.assert_array_equal(env.command_state_upper[2], np.full(nm, 30.0))
    np.testing.assert_array_equal(env.action_high, env.command_state_upper[1])
    iq = np.array([rob.idx_q[m.joint] for m in rob.motors])
    np.testing.assert_array_equal(env.command_state_lower[0], rob.q_lower[iq])
    # inferred acceleration bound (joint_acceleration_limit=None): bang-bang on velocity or effort, whichever binds
    env2 = PDControlBatchedEnv(sc, api_=api)
    eff = np.array([m.effort_limit for m in rob.motors])
    expect = np.minimum(2.0 * v_hw / sc.step_dt, eff / (sc.kp * sc.step_dt * np.maximum(sc.step_dt, sc.kd)))
    np.testing.assert_allclose(env2.command_state_upper[2], expect, rtol=1e-15)
    env2.close()
    # the oracle with the same blocks
    orc = OracleBatch(rob, sc.options, sc.n_env)
    table = np.stack([np.full(nm, 50.0), np.full(nm, 0.15), rob.q_lower[iq], rob.q_upper[iq], np.minimum(v_hw, 4.0)])
    orc.set_pd_controller_full(sc.kp, sc.kd, env.command_state_lower, env.command_state_upper, table)
    orc.set_mahony_filter(0.75, 0.057)
    orc.set_command(np.zeros((sc.n_env, nm)))
    obs, _ = env.reset()
    assert not orc.start(sc.q0, sc.v0).any()
    assert set(obs) == {"t", "states", "measurements", "features"} and obs["states"]["pd_controller"].shape == (3, 2, nm)
    assert obs["features"]["mahony_filter"].shape == (3, 4, 1)
    rng = np.random.default_rng(5)
    for k in range(2):
        act = rng.uniform(-0.3, 0.3, size=(sc.n_env, nm))
        obs, reward, terminated, truncated, info = env.step(act)
        st, out = orc.get_pd_controller_state(), np.zeros((sc.n_env, nm))
        pd_adapter(act.copy(), 1, st, env.command_state_lower, env.command_state_upper, False, None, sc.step_dt, out)
        orc.set_command(out)
        assert not orc.step(sc.step_dt, parallel=True).any()
        pc.compare(env.engine, orc, 1e-9, 1e-7)
        np.testing.assert_allclose(obs["states"]["pd_controller"], orc.get_pd_controller_state()[:, :2], atol=1e-10)
        np.testing.assert_allclose(obs["features"]["mahony_filter"][:, :, 0], orc.get_mahony_filter()[:, 0, :4], atol=1e-10)
        assert not terminated.any() and not truncated.any() and (reward == 1.0).all()
    keys = [("states", "pd_controller"), ("measurements", "EncoderSensor"), ("features", "mahony_filter")]
    low = {keys[0]: env.command_state_lower[:2]}
    high = {keys[0]: env.command_state_upper[:2]}
    flat = flatten_observation(obs, keys, low, high)
    assert flat.shape == (3, 2 * nm + 2 * nm + 4)
    assert np.abs(flat[:, :2 * nm]).max() <= 1.0 + 1e-12            # normalised by the command-state bounds
    np.testing.assert_array_equal(flat[:, 2 * nm:4 * nm], obs["measurements"]["EncoderSensor"].reshape(3, -1))   # unbounded: untouched
    env.close()


@pytest.mark.parametrize("in_kernel", [True, False])
def test_joint_bounds_on_the_hot_path_and_by_handoff(api, monkeypatch, in_kernel):
    """ANYmal envs driven through their hip bounds: solved inside the hot-path evaluation (default for the quadruped
    signature), or -- JB_NO_FAST_BOUNDS=1, the path every other robot takes -- aborted and replayed by the full body."""
    if not in_kernel:
        monkeypatch.setenv("JB_NO_FAST_BOUNDS", "1")
    pc.bounds_handoff_scenario(api, n_env=9, n_steps=4)


def test_long_horizon_resynchronised_gpu_like_rounding():
    """Short version of the `-m gpu` long-horizon test on the emulator build that contracts multiply-adds like nvcc."""
    if not host_has_fma():
        pytest.skip("host CPU without FMA")
    fma = emul_api(fma=True)
    resync, free = pc.resync_long_horizon_scenario("anymal", 4, 25, api=fma, tol_rel=1e-11)
    assert free.max() < 1e-10
    pc.resync_long_horizon_scenario("atlas", 2, 4, api=fma, tol_rel=1e-11, free_running=False)


@pytest.mark.parametrize("in_kernel", [True, False])
def test_handoff_with_stateful_blocks(api, monkeypatch, in_kernel):
    if not in_kernel:
        monkeypatch.setenv("JB_NO_FAST_BOUNDS", "1")
    pc.stateful_handoff_scenario(api, n_env=6, n_steps=6)


def test_mahony_filter_observer(api):
    pc.mahony_scenario(api, "anymal")
    pc.mahony_scenario(api, "atlas", n_env=1, n_steps=1)


def test_compute_dynamics_leaves_the_running_state_alone(api):
    """`compute_dynamics` on arbitrary states (out-of-bounds joints included) between two steps must not change the
    running envs: neither their held command nor their constraint state."""
    sc = scenarios.make("anymal", 3, seed=9)
    runs = []
    for probe in (False, True):
        eng = BatchedEngine(sc.robot, sc.options, 3, api_=api)
        eng.set_pd_controller(sc.kp, sc.kd)
        eng.set_command(sc.target0)
        eng.start(sc.q0, sc.v0)
        for k in range(2):
            eng.set_command(sc.sample_targets(k))
            if probe:
                rng = np.random.default_rng(k)
                q, v = pc.random_states(sc.robot, 3, rng)
                q[:, 7] = sc.robot.q_upper[7] + 0.2          # a hip joint beyond its upper bound
                eng.compute_dynamics(q, v, rng.uniform(-30, 30, size=(3, sc.robot.nmotors)))
            eng.step(sc.step_dt)
        runs.append((eng.get_state(), eng.get_status(), eng.get_constraints()[0].copy()))
    for x, y in zip(runs[0][0], runs[1][0]):
        np.testing.assert_array_equal(x, y)
    np.testing.assert_array_equal(runs[0][1], runs[1][1])
    np.testing.assert_array_equal(runs[0][2], runs[1][2])


def test_state_views_are_stable_and_current(api):
    """`jb_state_ptrs` / `BatchedEngine.state_views`: the arrays are created once (same memory for the life of the batch)
    and hold, after a synchronise, what the getters return -- start, steps, and a masked restart included."""
    sc = scenarios.make("anymal", 5, seed=4)
    eng = BatchedEngine(sc.robot, sc.options, 5, api_=api)
    eng.set_pd_controller(sc.kp, sc.kd)
    eng.set_command(sc.target0)
    eng.start(sc.q0, sc.v0)
    views = eng.state_views()                      # enabled on a running batch: filled right away
    addr = {k: a.__array_interface__["data"][0] for k, a in views.items()}

    def check():
        eng.synchronize()
        t, q, v, a = eng.get_state()
        for key, ref in (("t", t), ("q", q), ("v", v), ("a", a), ("sensors", eng.get_sensors())):
            np.testing.assert_array_equal(views[key], ref)
        assert eng.state_views() is views
        assert {k: a.__array_interface__["data"][0] for k, a in views.items()} == addr
    check()
    for k in range(3):
        eng.set_command(sc.sample_targets(k))
        eng.step(sc.step_dt)
        check()
    mask = np.array([1, 0, 0, 1, 0], dtype=bool)
    eng.start(sc.q0, sc.v0, mask=mask)
    check()
    assert (views["t"][mask] == 0.0).all() and (views["t"][~mask] > 0.0).all()


def test_model_variants_match_per_variant_oracles(api):
    """Model randomisation (`jb_set_model_variants`): three draws of `biased_robot` (mass, centre of mass, inertia, joint
    placement) over four groups of envs; every group follows the oracle built on ITS variant, centroidal terms included."""
    from jiminy_b200 import model as M
    n = 26                                         # ANYmal: 8 envs per group -> 4 groups, the last one partial
    sc = scenarios.make("anymal", n, seed=6)
    rng = np.random.default_rng(11)
    robots = [sc.robot] + [M.biased_robot(sc.robot, rng, mass_std=0.05, com_std=0.05, inertia_std=0.05, relative_position_std=0.002)
                           for _ in range(2)]
    assert abs(robots[1].mass - sc.robot.mass) > 1e-3 and not np.array_equal(robots[1].placement, sc.robot.placement)
    np.testing.assert_array_equal(robots[1].placement[:, :9], sc.robot.placement[:, :9])      # rotations untouched
    eng = BatchedEngine(sc.robot, sc.options, n, api_=api)
    assert eng.envs_per_group == 8
    vog = np.array([1, 0, 2, 1], dtype=np.int32)
    eng.set_model_variants(robots, vog)
    with pytest.raises(ValueError):
        eng.set_model_variants(robots, vog[:2])
    eng.set_pd_controller(sc.kp, sc.kd)
    eng.set_command(sc.target0)
    eng.start(sc.q0, sc.v0)
    for k in range(2):
        eng.set_command(sc.sample_targets(k))
        eng.step(sc.step_dt)
    t, q, v, a = eng.get_state()
    cen = eng.get_centroidal()
    worst = 0.0
    for g, var in enumerate(vog):
        rows = slice(8 * g, min(8 * g + 8, n))
        m = rows.stop - rows.start
        orc = OracleBatch(robots[var], sc.options, m)
        orc.set_pd_controller(sc.kp, sc.kd)
        orc.set_command(sc.target0[rows])
        orc.start(sc.q0[rows], sc.v0[rows])
        for k in range(2):
            orc.set_command(sc.sample_targets(k)[rows])
            orc.step(sc.step_dt)
        to, qo, vo, ao = orc.get_state()
        np.testing.assert_allclose(q[rows], qo, rtol=0, atol=1e-9)
        np.testing.assert_allclose(v[rows], vo, rtol=0, atol=1e-8)
        np.testing.assert_allclose(a[rows], ao, rtol=1e-7, atol=1e-6)
        for x, y in zip(cen, orc.get_centroidal()):
            np.testing.assert_allclose(np.asarray(x)[rows], y, rtol=1e-8, atol=1e-8)
        worst = max(worst, float(np.abs(q[rows] - qo).max()))
    # the variants really differ: group 0 (variant 1) against the unbiased model
    orc0 = OracleBatch(sc.robot, sc.options, 8)
    orc0.set_pd_controller(sc.kp, sc.kd); orc0.set_command(sc.target0[:8]); orc0.start(sc.q0[:8], sc.v0[:8])
    for k in range(2):
        orc0.set_command(sc.sample_targets(k)[:8]); orc0.step(sc.step_dt)
    assert np.abs(orc0.get_state()[1] - q[:8]).max() > 1e4 * max(worst, 1e-13)
    # a variant with another tree is refused
    other = scenarios.make("atlas", 1).robot
    with pytest.raises(Exception):
        eng.set_model_variants([other], np.zeros(4, dtype=np.int32))


# ---- flexibility joints (Engine::computeInternalDynamics, engine.cc:3367-3391; spherical joints in the lane plan)
def test_flexibility_branched_arm_matches_oracle(api):
    import flexibility_common as fc
    for lanes in (0, 1):
        _lanes(lanes)
        try:
            fc.branched_arm_parity(api)
        finally:
            _lanes(0)


def test_flexibility_dopri_matches_oracle(api):
    import flexibility_common as fc
    fc.branched_arm_parity(api, solver="runge_kutta_dopri")


def test_flexibility_anymal_matches_oracle(api):
    import flexibility_common as fc
    fc.flexible_anymal_parity(api, n_env=8, n_steps=1)


@pytest.mark.parametrize("model", ["spring_damper", "constraint"])
def test_flexibility_joint_bounds_constraint_path(api, model):
    import flexibility_common as fc
    fc.flexible_pendulum_on_its_bounds(api, model=model)


def test_flexibility_constraint_contacts_anymal(api):
    import flexibility_common as fc
    fc.flexible_anymal_parity(api, n_env=3, n_steps=1, contact_model="constraint", tol_state=1e-8, tol_sens=1e-6)


def test_flexibility_series_elastic_actuator_like_the_reference_test(api):
    import flexibility_common as fc
    assert fc.series_elastic_actuator(api) < 1e-11


def test_flexibility_engine_facade_like_the_reference_api_test(api):
    """unit_py/test_simple_pendulum.py:815-842 through the single-env `Engine` facade: the flexibility API works, the
    indices survive a simulation."""
    import flexibility_common as fc
    from jiminy_b200.core import Engine
    th = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    M.attach_motor(th, "PendulumJoint", "PendulumJoint", enableVelocityLimit=False, enableEffortLimit=False)
    robot = M.Robot(th)
    model_options = robot.get_model_options()
    model_options["dynamics"]["enableFlexibility"] = True
    model_options["dynamics"]["flexibilityConfig"] = [{"frameName": "PendulumJoint", "stiffness": np.ones(3),
                                                       "damping": np.ones(3), "inertia": np.ones(3)}]
    robot.set_model_options(model_options)
    assert robot.is_flexibility_enabled and robot.flexibility_joint_indices == [1]
    engine = Engine(api_=api)
    engine.add_robot(robot)
    engine.simulate(0.1, np.array([0.0, 0.0, 0.0, 1.0, 0.0]), np.zeros(4))
    assert engine.stepper_state.t >= 0.1 - 1e-9 and np.isfinite(engine.stepper_state.q).all()
    assert abs(np.linalg.norm(engine.stepper_state.q[:4]) - 1.0) < 1e-9          # the quaternion stays on the sphere
    assert robot.flexibility_joint_indices == [1]


def test_backlash_joints_match_oracle(api):
    import flexibility_common as fc
    fc.backlash_pendulum_parity(api)


def test_flexibility_batched_env(api):
    """The gym-style batched env over the flexible ANYmal (`scenarios.make("anymal_flexible")`): extended state in the
    observation, restart of the finished envs with undeformed flexibilities."""
    from jiminy_b200.envs import BatchedJiminyEnv
    sc = scenarios.make("anymal_flexible", 3)
    assert sc.robot.is_flexibility_enabled and sc.robot.nq == 19 + 4 * 4
    env = BatchedJiminyEnv(sc, api_=api, simulation_duration_max=0.08)
    obs, _ = env.reset()
    assert obs["states"]["agent"]["q"].shape == (3, 35) and obs["states"]["agent"]["v"].shape == (3, 30)
    obs, rew, term, trunc, info = env.step(sc.sample_targets(0))
    assert not term.any() and not trunc.any()
    iq = sc.robot.idx_q[sc.robot.joint_index("LF_HFEFlexibility")]
    assert np.abs(obs["states"]["agent"]["q"][:, iq:iq + 3]).max() > 1e-7      # deformed under load
    obs, rew, term, trunc, info = env.step(sc.sample_targets(1))
    assert trunc.all()                                                          # duration limit -> restarted ...
    np.testing.assert_allclose(obs["states"]["agent"]["q"][:, iq:iq + 4], [[0.0, 0.0, 0.0, 1.0]] * 3)   # ... undeformed
    env.close()


def test_batched_rollout_telemetry_logs(api, tmp_path):
    """One reference-format log per recorded env of a batched rollout (`telemetry.BatchTelemetryRecorder`), read back with the
    restated reference reader: times, states and sensors of the chosen envs, the batch read once per snapshot."""
    from jiminy_b200 import telemetry as T
    sc = scenarios.make("anymal", 5, seed=2)
    opt = dict(sc.options)
    opt["telemetry"] = {"enableConfiguration": True, "enableVelocity": True, "enableAcceleration": True, "enableEnergy": True}
    eng = BatchedEngine(sc.robot, opt, sc.n_env, api_=api)
    eng.set_pd_controller(sc.kp, sc.kd)
    eng.set_command(sc.target0)
    eng.start(sc.q0, sc.v0)
    rec = T.BatchTelemetryRecorder(eng, envs=[0, 3, 4], options=opt)
    rec.snapshot()
    states = [eng.get_state()]
    for k in range(2):
        eng.set_command(sc.sample_targets(k))
        eng.step(sc.step_dt)
        rec.snapshot()
        states.append(eng.get_state())
    paths = rec.write_logs(str(tmp_path), prefix="rollout")
    assert [os.path.basename(p) for p in paths] == ["rollout_000000.data", "rollout_000003.data", "rollout_000004.data"]
    for e, path in zip((0, 3, 4), paths):
        var = T.read_log(path)["variables"]
        np.testing.assert_allclose(var["Global.Time"], [0.0, 0.04, 0.08], rtol=0, atol=1e-10)
        for k, (t, q, v, a) in enumerate(states):
            assert var["currentFreeflyerPositionTransZ"][k] == q[e, 2]
            assert var["currentVelocityRH_KFE"][k] == v[e, sc.robot.idx_v[sc.robot.joint_index("RH_KFE")]]
        assert np.isfinite(var["energy"]).all() and f"ImuSensor.{sc.robot.imu_names[0]}.GyroX" in var
    with pytest.raises(ValueError):
        T.BatchTelemetryRecorder(eng, envs=[7])


def test_centroidal_terms_of_a_massless_subtree_are_finite(api):
    """A flexibility at a fixed frame whose bodies are massless leaves a joint with a massless subtree: its centre of mass
    is the joint origin (InertiaTpl::__pequ__ divides by max(mass, eps)), not 0 / 0; everything else matches the oracle."""
    import flexibility_common as fc
    robot, _, opt = fc.flexible_branched_arm()
    flex = M.add_flexibility_joints(robot, [dict(frameName="b_sole_fixed", stiffness=[300.0] * 3, damping=[2.0] * 3, inertia=[0.01] * 3)])
    j = flex.joint_index("b_sole_fixed")
    assert flex.inertia[j, 0] == 0.0
    rng = np.random.default_rng(1)
    q, v = pc.random_states(flex, 2, rng, base_height=0.55)
    eng, orc = BatchedEngine(flex, opt, 2, api_=api), OracleBatch(flex, opt, 2)
    cmd = np.zeros((2, flex.nmotors))
    for x in (eng, orc):
        x.set_command(cmd)
    eng.start(q, v)
    assert not orc.start(q, v).any()
    eng.step(4e-3)
    orc.step(4e-3)
    ycrb, com, vcom, hg, dhg = [np.asarray(x) for x in eng.get_centroidal()]
    yo, co, vo, hgo, dhgo = [np.asarray(x) for x in orc.get_centroidal()]
    assert np.isfinite(ycrb).all() and np.isfinite(com).all()
    np.testing.assert_array_equal(com[:, j], 0.0)
    keep = np.arange(flex.njoints) != j          # (the reference keeps the meaningless lever of its inertia bookkeeping there)
    np.testing.assert_allclose(ycrb[:, keep], yo[:, keep], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(com[:, keep], co[:, keep], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(hg, hgo, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(dhg, dhgo, rtol=1e-8, atol=1e-7)


def test_flexibility_on_a_trunk_joint_of_atlas(api):
    import flexibility_common as fc
    fc.flexible_atlas_trunk_parity(api)
