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
            elif t == 14:   # spherical (flexibility) joint: a moderate rotation
                w = rng.normal(size=3) * 0.2
                ang = np.linalg.norm(w)
                q[i, iq:iq + 3] = np.sin(ang / 2) * w / ang
                q[i, iq + 3] = np.cos(ang / 2)
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


def rel_state_error(q1, v1, q0, v0):
    """Per-env relative deviation of (q, v): max |x1 - x0| / max(1, max |x0|), the larger of the two blocks."""
    eq = np.abs(q1 - q0).max(axis=1) / np.maximum(1.0, np.abs(q0).max(axis=1))
    ev = np.abs(v1 - v0).max(axis=1) / np.maximum(1.0, np.abs(v0).max(axis=1))
    return np.maximum(eq, ev)


def resync_long_horizon_scenario(name, n_env, n_steps, api=None, tol_rel=1e-10, free_running=True, seed=21, **kw):
    """Long horizons without the chaotic amplification: after every env-step the device is handed the ORACLE's state
    (jb_set_stepper_state: q, v, a, scheduler scalars, iteration counters, held command), so that each step is compared
    from identical inputs -- `resync[k]` is the error the device path adds in ONE env-step, at every point of a long
    trajectory.  A second, free-running device engine records how a rounding-level difference grows when nothing is
    re-synchronised (`free[k]`, a property of the dynamics: stiff contacts amplify any perturbation).
    Returns (resync [n_steps, n_env], free [n_steps, n_env])."""
    sc = scenarios.make(name, n_env, seed=seed, **kw)
    eng, orc = make_pair(sc, api)
    free = None
    if free_running:
        free = BatchedEngine(sc.robot, sc.options, sc.n_env, api_=api)
        if sc.kp is not None:
            free.set_pd_controller(sc.kp, sc.kd)
        free.set_command(sc.target0)
        free.start(sc.q0, sc.v0)
    resync, growth = [], []
    for k in range(n_steps):
        act = sc.sample_targets(k)
        for x in (eng, orc) + ((free,) if free is not None else ()):
            x.set_command(act)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
        (t1, q1, v1, a1), (t0, q0, v0, a0) = eng.get_state(), orc.get_state()
        e = rel_state_error(q1, v1, q0, v0)
        resync.append(e)
        assert e.max() <= tol_rel, f"{name}: env-step {k}: one-step relative deviation {e.max():.3e} > {tol_rel:.1e} (env {int(e.argmax())})"
        np.testing.assert_allclose(t1, t0, rtol=0, atol=1e-15)
        np.testing.assert_array_equal(eng.get_iters()[0], orc.get_iters()[0])
        np.testing.assert_array_equal(eng.get_status(), orc.get_status())
        s1, s0 = eng.get_sensors(), orc.get_sensors()
        if s0.size:
            np.testing.assert_allclose(s1, s0, rtol=0, atol=1e4 * tol_rel * max(1.0, np.abs(s0).max()))
        if free is not None:
            free.step(sc.step_dt)
            _, qf, vf, _ = free.get_state()
            growth.append(rel_state_error(qf, vf, q0, v0))
        sched, held = orc.get_stepper_state()
        it, itf = orc.get_iters()
        eng.set_stepper_state(sched, q0, v0, a0, it, itf, held)
    return np.array(resync), np.array(growth)


def compare_extra_terms(eng, orc, tol=1e-9):
    """computeExtraTerms outputs: energies, joint spatial accelerations, joint internal wrenches."""
    e1, a1, f1 = eng.get_extra_terms()
    e0, a0, f0 = orc.get_extra_terms()
    np.testing.assert_allclose(e1, e0, rtol=0, atol=tol * max(1.0, np.abs(e0).max()))
    np.testing.assert_allclose(a1, a0, rtol=0, atol=tol * max(1.0, np.abs(a0).max()))
    np.testing.assert_allclose(f1, f0, rtol=0, atol=tol * max(1.0, np.abs(f0).max()))
    # subtree inertias / centres of mass / their velocities, centroidal momentum and its derivative (engine.cc:817-832, :890-904)
    for x1, x0 in zip(eng.get_centroidal(), orc.get_centroidal()):
        np.testing.assert_allclose(x1, x0, rtol=0, atol=tol * max(1.0, np.abs(x0).max()))


def _force_pair(sc, api):
    eng = BatchedEngine(sc.robot, sc.options, sc.n_env, api_=api)
    orc = OracleBatch(sc.robot, sc.options, sc.n_env)
    if sc.kp is not None:
        eng.set_pd_controller(sc.kp, sc.kd)
        orc.set_pd_controller(sc.kp, sc.kd)
    return eng, orc


def external_forces_scenario(api, n_env=6, n_steps=3, solver="runge_kutta_4", tol=1e-9, contact_model=None, dt_max=None):
    """Impulse forces with per-env schedules on the base (trunk joint) and on a shank (private joint, off-origin
    frame), a sampled profile force (finite update period) and a continuous one: Engine::computeExternalForces
    + the breakpoint handling of Engine::step, against the oracle."""
    sc = scenarios.make("anymal", n_env, seed=3, solver=solver, contact_model=contact_model, dt_max=dt_max)
    eng, orc = _force_pair(sc, api)
    rng = np.random.default_rng(5)
    rob = sc.robot
    base = rob.frames["base"] if "base" in rob.frames else rob.frames["root_joint"]
    shank_name = next(n for n in rob.frames if "SHANK" in n.upper())
    shank = rob.frames[shank_name]
    imp = [
        ((base.joint, base.placement.p), rng.uniform(0.0, 0.05, n_env), rng.uniform(2e-3, 2e-2, n_env), rng.normal(size=(n_env, 6)) * 200.0),
        ((base.joint, base.placement.p), np.zeros(n_env), np.full(n_env, 7e-3), rng.normal(size=(n_env, 6)) * 100.0),
        ((shank.joint, shank.placement.p + [0.0, 0.0, -0.1]), rng.uniform(0.03, 0.09, n_env), rng.uniform(1e-3, 3e-2, n_env), rng.normal(size=(n_env, 6)) * 50.0),
    ]
    for fr, t, dt, F in imp:
        k1 = eng.register_impulse_force(fr, t, dt, F)
        k0 = orc.register_impulse_force(fr[0], fr[1], t, dt, F)
        assert k1 == k0
    s_eng = [eng.register_profile_force((base.joint, base.placement.p), 0.01), eng.register_profile_force(shank_name, 0.0)]
    s_orc = [orc.register_profile_force(base.joint, base.placement.p, 0.01), orc.register_profile_force(shank.joint, shank.placement.p, 0.0)]
    assert s_eng == s_orc
    for x in (eng, orc):
        x.set_command(sc.target0)
    w0, w1 = rng.normal(size=(n_env, 6)) * 30.0, rng.normal(size=(n_env, 6)) * 20.0
    for x, sl in ((eng, s_eng), (orc, s_orc)):
        x.set_profile_force(sl[0], w0)
        x.set_profile_force(sl[1], w1)
    eng.start(sc.q0, sc.v0)
    assert not orc.start(sc.q0, sc.v0).any()
    compare(eng, orc, 1e-13, 1e-12)
    np.testing.assert_allclose(eng.get_efforts()[3], orc.get_efforts()[3], rtol=0, atol=1e-9)
    for k in range(n_steps):
        act = sc.sample_targets(k)
        w0, w1 = rng.normal(size=(n_env, 6)) * 30.0, rng.normal(size=(n_env, 6)) * 20.0
        for x, sl in ((eng, s_eng), (orc, s_orc)):
            x.set_command(act)
            x.set_profile_force(sl[0], w0)
            x.set_profile_force(sl[1], w1)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
        compare(eng, orc, tol, 100 * tol)
        f0 = orc.get_efforts()[3]
        np.testing.assert_allclose(eng.get_efforts()[3], f0, rtol=0, atol=1e-8 * max(1.0, np.abs(f0).max()))
    # the forces did something: the same run without them ends elsewhere
    q_with = eng.get_state()[1]
    eng.stop()
    eng.remove_all_forces()
    eng.set_command(sc.target0)
    eng.start(sc.q0, sc.v0)
    for k in range(n_steps):
        eng.set_command(sc.sample_targets(k))
        eng.step(sc.step_dt)
    assert np.abs(eng.get_state()[1] - q_with).max() > 1e-4
    return eng, orc


# ---------------------------------------------------------------------------------------------
# constraint path (joint position bounds, contacts.model = "constraint") against the oracle
def _cons_opt(**stepper):
    from jiminy_b200 import model as M
    opt = M.default_engine_options()
    opt["contacts"]["model"] = "spring_damper"
    opt["stepper"].update(stepper)
    return opt


def bounds_scenario(api, data_dir, model="spring_damper", n_steps=60):
    """A pendulum thrown against its position bounds: JointConstraint enable / disable hysteresis, PGS with a
    single boxed multiplier, multiplier reported in u."""
    import os
    from jiminy_b200 import model as M
    r = M.build_robot_table(os.path.join(data_dir, "simple_pendulum.urdf"), False)
    M.attach_motor(r, "PendulumJoint", "PendulumJoint", enableVelocityLimit=False, enableEffortLimit=False)
    r.q_upper[0], r.q_lower[0] = 0.5, -0.5
    opt = _cons_opt(odeSolver="runge_kutta_4", dtMax=1e-3, controllerUpdatePeriod=1e-3, sensorsUpdatePeriod=1e-3)
    opt["contacts"]["model"] = model
    eng, orc = BatchedEngine(r, opt, 3, api_=api), OracleBatch(r, opt, 3)
    q0, v0 = np.array([[0.3], [0.1], [-0.45]]), np.array([[0.0], [1.0], [-2.0]])
    for x in (eng, orc):
        x.set_command(np.zeros((3, 1)))
    eng.start(q0, v0)
    assert not orc.start(q0, v0).any()
    compare(eng, orc, 1e-13, 1e-11)
    for _ in range(n_steps):
        eng.step(0.01)
        assert not orc.step(0.01).any()
        compare(eng, orc, 1e-9, 1e-7)
        np.testing.assert_allclose(eng.get_efforts()[0], orc.get_efforts()[0], rtol=0, atol=1e-7)
    q = eng.get_state()[1]
    assert np.all(np.abs(np.abs(q) - 0.5) < 2e-3) and (eng.get_status() & 8).all()
    return eng, orc


def start_on_bounds_scenario(api, data_dir):
    """`Engine::start` with joint-bound constraints enabled (constraint contact model: every bound starts enabled): the
    INIT_ITERATIONS loop rebuilds u from a uInternal that carries the multipliers of the previous iteration
    (engine.cc:1452-1461, :3770-3788), so the initial acceleration of a joint resting on its bound is the fixed point of
    that feedback, not the plain regularised solve."""
    import os
    from jiminy_b200 import model as M
    r = M.build_robot_table(os.path.join(data_dir, "simple_pendulum.urdf"), False)
    M.attach_motor(r, "PendulumJoint", "PendulumJoint", enableVelocityLimit=False, enableEffortLimit=False)
    r.q_upper[0], r.q_lower[0] = 0.5, -0.5
    opt = _cons_opt(odeSolver="runge_kutta_4", dtMax=1e-3, controllerUpdatePeriod=1e-3, sensorsUpdatePeriod=1e-3)
    opt["contacts"]["model"] = "constraint"
    eng, orc = BatchedEngine(r, opt, 4, api_=api), OracleBatch(r, opt, 4)
    q0, v0 = np.array([[0.5], [-0.5], [0.3], [0.5]]), np.array([[0.0], [0.0], [0.0], [-1.0]])
    for x in (eng, orc):
        x.set_command(np.zeros((4, 1)))
    eng.start(q0, v0)
    assert not orc.start(q0, v0).any()
    a1, a0 = eng.get_state()[3], orc.get_state()[3]
    np.testing.assert_allclose(a1, a0, rtol=1e-12, atol=1e-13)
    assert abs(a0[1, 0]) < 1e-4 and abs(a0[2, 0]) > 1.0          # held on its lower bound / free inside the bounds
    for _ in range(5):
        eng.step(0.002)
        assert not orc.step(0.002).any()
        compare(eng, orc, 1e-11, 1e-9)
    return eng, orc


def point_mass_constraint_scenario(api, data_dir, n_steps=40, torsion=0.0, solver="runge_kutta_4", impulse=False):
    """A free-flying mass on the ground with the constraint contact model: resting, sliding (Coulomb cone) and
    spinning envs; contact / force sensors and f_external included in the comparison."""
    import os
    from jiminy_b200 import model as M
    r = M.build_robot_table(os.path.join(data_dir, "point_mass.urdf"), True)
    r.add_contact_points(["MassBody"])
    M.attach_sensor(r, "ContactSensor", "MassBody", frame_name="MassBody")
    r.add_frame("Sensor", "MassBody", M.SE3(M.rpy_to_matrix([0.3, -0.2, 0.5]), np.array([0.1, 0.2, -0.05])))
    M.attach_sensor(r, "ForceSensor", "F", frame_name="Sensor")
    opt = _cons_opt(dtMax=1e-3, controllerUpdatePeriod=1e-3, odeSolver=solver)
    if solver == "runge_kutta_dopri":
        opt["stepper"].update(dtMax=5e-3, tolAbs=1e-7, tolRel=1e-6)
    opt["contacts"].update(model="constraint", friction=0.8, transitionEps=1e-6, torsion=torsion)
    opt["world"]["gravity"] = [4.0, 1.0, -9.81, 0, 0, 0]
    n = 3
    eng, orc = BatchedEngine(r, opt, n, api_=api), OracleBatch(r, opt, n)
    if impulse:   # a push and a lift while in contact: external forces and constraints in the same evaluation
        fr = r.frames["MassBody"]
        for x in (eng, orc):
            args = ((fr.joint, fr.placement.p),) if x is eng else (fr.joint, fr.placement.p)
            x.register_impulse_force(*args, [0.05, 0.02, 0.11], [0.05, 0.1, 0.02], [[20.0, 0, 0, 0, 0, 0], [0, 0, 30.0, 0, 0, 0], [0, -15.0, 5.0, 0, 0, 1.0]])
    q0 = np.tile(r.neutral(), (n, 1))
    q0[:, 2] = [0.0, 0.02, -1e-4]
    v0 = np.zeros((n, 6))
    v0[1, :3] = [0.3, 0.0, 0.0]
    v0[2, 3:] = [0.5, 0.2, 1.0]
    eng.start(q0, v0)
    assert not orc.start(q0, v0).any()
    compare(eng, orc, 1e-13, 1e-10)
    for _ in range(n_steps):
        eng.step(0.01)
        assert not orc.step(0.01).any()
        compare(eng, orc, 1e-9, 1e-7)
        np.testing.assert_allclose(eng.get_efforts()[3], orc.get_efforts()[3], rtol=0, atol=1e-7)
    return eng, orc


def robot_constraint_scenario(name, n_env, n_steps, api=None, tol_state=1e-8, tol_sens=1e-6, torsion=None, **kw):
    """A BASELINE robot with contacts.model = "constraint" (the default of the reference's option files)."""
    sc = scenarios.make(name, n_env, **kw)
    sc.options["contacts"]["model"] = "constraint"
    if torsion is not None:
        sc.options["contacts"]["torsion"] = torsion
    eng, orc = make_pair(sc, api)
    compare(eng, orc, 1e-12, 1e-9)
    for k in range(n_steps):
        act = sc.sample_targets(k)
        eng.set_command(act)
        orc.set_command(act)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
        compare(eng, orc, tol_state, tol_sens)
    return eng, orc, sc


def pd_block_scenario(api, name="anymal", n_env=4, n_steps=3, safety=False):
    """gym_jiminy's PDController block (integrate_zoh + pd_controller, optional MotorSafetyLimit) on the device
    against the oracle's restatement, which is itself pinned by golden vectors of the reference's own functions
    (tests/test_golden_controller_blocks.py).  Actions = target motor accelerations."""
    sc = scenarios.make(name, n_env, seed=6)
    rob = sc.robot
    nm = rob.nmotors
    iq = np.array([rob.idx_q[m.joint] for m in rob.motors])
    vlim = np.array([m.velocity_limit for m in rob.motors])
    # gentle target motion (the stiff spring-damper ground does not survive flailing legs at RK4 1 ms)
    lower = np.stack([rob.q_lower[iq] + 0.05, np.full(nm, -0.6), np.full(nm, -15.0)])
    upper = np.stack([rob.q_upper[iq] - 0.05, np.full(nm, 0.6), np.full(nm, 15.0)])
    # MotorSafetyLimit: kp, kd, soft position bounds, and a soft velocity limit below the motors' own (soft_velocity_max)
    sf = np.stack([np.full(nm, 20.0), np.full(nm, 0.5), rob.q_lower[iq] + 0.02, rob.q_upper[iq] - 0.02,
                   np.minimum(vlim, 4.0)]) if safety else None
    eng, orc = BatchedEngine(rob, sc.options, n_env, api_=api), OracleBatch(rob, sc.options, n_env)
    rng = np.random.default_rng(11)
    act = rng.uniform(-10.0, 10.0, size=(n_env, nm))
    for x in (eng, orc):
        x.set_pd_controller_full(sc.kp, sc.kd, lower, upper, sf)
        x.set_command(act)
    eng.start(sc.q0, sc.v0)
    assert not orc.start(sc.q0, sc.v0).any()
    compare(eng, orc, 1e-13, 1e-11)
    for k in range(n_steps):
        act = rng.uniform(-25.0, 25.0, size=(n_env, nm))     # beyond the acceleration bound: exercises the clipping
        eng.set_command(act)
        orc.set_command(act)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
        compare(eng, orc, 1e-9, 1e-7)
        np.testing.assert_allclose(eng.get_efforts()[1], orc.get_efforts()[1], rtol=0, atol=1e-7)   # motor efforts
    return eng, orc


def masked_restart_constraint_scenario(api, name, n_env=5, tol_state=1e-9, tol_sens=1e-7):
    """Masked `start` (a vectorised env restarting some of its envs) while contact constraints are enabled: the
    restarted envs get fresh constraint state (enabled set, multipliers, reference placements), the others keep theirs."""
    kw = dict(solver="euler_explicit", dt_max=0.005, contact_model="constraint")
    sc, sc2 = scenarios.make(name, n_env, seed=7, **kw), scenarios.make(name, n_env, seed=8, **kw)
    eng, orc = make_pair(sc, api)
    for k in range(2):
        act = sc.sample_targets(k)
        eng.set_command(act)
        orc.set_command(act)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
    mask = (np.arange(n_env) % 2 == 0).astype(np.uint8)
    eng.start(sc2.q0, sc2.v0, mask=mask)
    assert not orc.start(sc2.q0, sc2.v0, mask=mask).any()
    compare(eng, orc, tol_state, tol_sens)
    for k in range(2):
        act = sc.sample_targets(2 + k)
        eng.set_command(act)
        orc.set_command(act)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
        compare(eng, orc, tol_state, tol_sens)
    t = eng.get_state()[0]
    np.testing.assert_allclose(t, np.where(mask, 2, 4) * sc.step_dt, atol=1e-12)
    return eng, orc


def atlas_bounds_and_contacts_scenario(api, n_env=2, n_steps=6, tol_state=1e-8, tol_sens=1e-6):
    """Atlas on `constraint` contacts whose elbows are driven past their position bounds: contact frames and joint
    bounds are enabled together (lane-block solver, rows of both kinds on a robot with a four-joint trunk) and the
    solve switches between the body-space and the lane-block formulation as the bounds come and go."""
    sc = scenarios.make("atlas", n_env, seed=5, solver="euler_explicit", dt_max=0.005)
    sc.options["contacts"]["model"] = "constraint"
    rob = sc.robot
    eng, orc = make_pair(sc, api)
    iq = np.array([rob.idx_q[m.joint] for m in rob.motors])
    sel = [k for k, m in enumerate(rob.motors) if "elx" in m.name or "ely" in m.name]
    hit = False
    for k in range(n_steps):
        act = sc.sample_targets(k)
        act[:, sel] = rob.q_upper[iq][sel] + 0.4
        eng.set_command(act)
        orc.set_command(act)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
        compare(eng, orc, tol_state, tol_sens)
        hit = hit or bool((eng.get_status() & 8).any())
        c1, c0 = eng.get_constraints(), orc.get_constraints()          # is_enabled / lambda of every constraint
        np.testing.assert_array_equal(c1[0], c0[0])
        np.testing.assert_array_equal(c1[2], c0[2])
        scale = max(1.0, np.abs(c0[3]).max())
        np.testing.assert_allclose(c1[1], c0[1], rtol=0, atol=1e-5 * scale)
        np.testing.assert_allclose(c1[3], c0[3], rtol=0, atol=1e-5 * scale)
    assert hit and not (eng.get_status() & ~8).any() and c0[0].any() and c0[2].any()
    return eng, orc


def atlas_reference_neutral(robot):
    """`AtlasJiminyEnv._neutral` (atlas.py:147-166) clipped to the joint limits as `_sample_state` does
    (generic.py:1300-1335; two shoulder angles exceed the rounded URDF limits by 2e-7), base lifted so that the feet
    touch the ground, zero velocity.  The knees and the shoulders sit exactly on position bounds."""
    import json
    import os
    from jiminy_b200 import robots as R
    with open(os.path.join(os.path.dirname(os.path.abspath(R.__file__)), "robots", "atlas.json")) as fh:
        q = np.array(json.load(fh)["meta"]["neutral"])
    q[7:] = np.clip(q[7:], robot.q_lower[7:], robot.q_upper[7:])
    return R.ground_base_height(robot, q)


ATLAS_PIPELINE = dict(joint_velocity_limit=4.0, joint_acceleration_limit=30.0, order=1, mahony=(0.75, 0.057),
                      safety=dict(kp=50.0, kd=0.15, soft_position_margin=0.0, soft_velocity_max=4.0))    # atlas.py:28-37, :239-295


def atlas_pd_standing_on_oracle(t_end=9.0):
    """The reference's acceptance test of the PD pipeline (gym_jiminy/unit_py/test_pipeline_control.py:46-113):
    `AtlasPDControlJiminyEnv` in evaluation mode, zero target motor velocities for 9 s.  Returns per env-step the
    largest |target velocity| and the largest |generalised velocity|, and the oracle."""
    from jiminy_b200.blocks import pd_adapter
    sc = scenarios.make("atlas", 1, seed=0, contact_model="constraint", solver="euler_explicit", dt_max=0.005)
    rob, nm = sc.robot, sc.robot.nmotors
    iq = np.array([rob.idx_q[m.joint] for m in rob.motors])
    v_hw = np.array([m.velocity_limit for m in rob.motors])
    vel = np.minimum(v_hw, ATLAS_PIPELINE["joint_velocity_limit"])
    acc = np.full(nm, ATLAS_PIPELINE["joint_acceleration_limit"])
    lower, upper = np.stack([rob.q_lower[iq], -vel, -acc]), np.stack([rob.q_upper[iq], vel, acc])
    sf = ATLAS_PIPELINE["safety"]
    table = np.stack([np.full(nm, sf["kp"]), np.full(nm, sf["kd"]), rob.q_lower[iq], rob.q_upper[iq], np.minimum(v_hw, sf["soft_velocity_max"])])
    orc = OracleBatch(rob, sc.options, 1)
    orc.set_pd_controller_full(sc.kp, sc.kd, lower, upper, table)
    orc.set_mahony_filter(*ATLAS_PIPELINE["mahony"])
    orc.set_command(np.zeros((1, nm)))
    q0 = atlas_reference_neutral(rob)[None, :]
    assert not orc.start(q0, np.zeros((1, rob.nv))).any()
    action, deadband = np.zeros((1, nm)), np.zeros(nm)       # evaluation mode: dead band enabled, 0 wide
    v_target, v_robot = [], []
    for _ in range(int(round(t_end / sc.step_dt))):
        st, out = orc.get_pd_controller_state(), np.zeros((1, nm))
        pd_adapter(action.copy(), 1, st, lower, upper, False, deadband, sc.step_dt, out)
        orc.set_command(out)
        assert not orc.step(sc.step_dt).any()
        v_target.append(np.abs(orc.get_pd_controller_state()[0, 1]).max())
        v_robot.append(np.abs(orc.get_state()[2]).max())
    return np.array(v_target), np.array(v_robot), orc, sc


def atlas_pd_standing_on_device(api, t_end, tol_state=1e-8):
    """Same run through `PDControlBatchedEnv` (evaluation mode) on the device path; compared with the oracle at the end."""
    from jiminy_b200.envs import PDControlBatchedEnv
    v_target, v_robot, orc, sc = atlas_pd_standing_on_oracle(t_end)
    sc.q0, sc.v0 = atlas_reference_neutral(sc.robot)[None, :], np.zeros((1, sc.robot.nv))
    env = PDControlBatchedEnv(sc, training=False, api_=api, **ATLAS_PIPELINE)
    env.reset()
    v_dev = []
    for _ in range(len(v_robot)):
        obs, reward, terminated, truncated, info = env.step(np.zeros((1, sc.robot.nmotors)))
        assert not terminated.any() and not truncated.any()
        v_dev.append(np.abs(obs["states"]["agent"]["v"]).max())
    (_, q1, v1, _), (_, q0, v0, _) = env.engine.get_state(), orc.get_state()
    np.testing.assert_allclose(q1, q0, rtol=0, atol=tol_state)
    np.testing.assert_allclose(v1, v0, rtol=0, atol=10 * tol_state)
    np.testing.assert_allclose(env.engine.get_pd_controller_state(), orc.get_pd_controller_state(), rtol=0, atol=tol_state)
    assert (env.engine.get_status() & 8).all()          # bound constraints were active (knees / shoulders on their bounds)
    env.close()
    return np.array(v_dev), v_robot, sc


def atlas_repeatability_scenario(api, steps=(0, 5, 20, 10, 0), n_env=2):
    """gym_jiminy/unit_py/test_pipeline_control.py:315-330 (`test_repeatability`): restarting from the same state after
    any number of steps must give exactly the same initial acceleration -- nothing of the previous episode (constraint
    multipliers and enabled set, controller targets, filter state, solver workspace) may survive `start`."""
    from jiminy_b200.envs import PDControlBatchedEnv
    sc = scenarios.make("atlas", n_env, seed=0, contact_model="constraint", solver="euler_explicit", dt_max=0.005)
    sc.q0 = np.tile(atlas_reference_neutral(sc.robot), (n_env, 1))
    sc.q0[1:, 7:] = np.clip(sc.q0[1:, 7:] + 0.01, sc.robot.q_lower[7:], sc.robot.q_upper[7:])    # a second, different env
    sc.v0 = np.zeros((n_env, sc.robot.nv))
    env = PDControlBatchedEnv(sc, training=False, api_=api, **ATLAS_PIPELINE)
    a_prev = None
    for n in steps:
        env.engine.set_command(np.zeros((n_env, sc.robot.nmotors)))
        env.engine.start(sc.q0, sc.v0)
        a = env.engine.get_state()[3].copy()
        s = env.engine.get_sensors().copy()
        if a_prev is None:
            a_prev, s_prev = a, s
        np.testing.assert_array_equal(a, a_prev)
        np.testing.assert_array_equal(s, s_prev)
        env._started = True
        for _ in range(n):
            env.step(np.zeros((n_env, sc.robot.nmotors)))
    env.close()


def pd_adapter_scenario(api, name="anymal", n_env=3, n_steps=3, order=0, instantaneous=False):
    """`PDAdapter` -> `PDController` pipeline (the `*-pid` envs of gym_jiminy): the host-side adapter of
    jiminy_b200/blocks.py drives the device block through the command-state getter / setter; the same adapter function
    drives the oracle's block.  Actions = target motor positions (order 0) or velocities (order 1)."""
    from jiminy_b200.blocks import PDAdapter, pd_adapter
    sc = scenarios.make(name, n_env, seed=9)
    rob = sc.robot
    nm = rob.nmotors
    iq = np.array([rob.idx_q[m.joint] for m in rob.motors])
    lower = np.stack([rob.q_lower[iq] + 0.05, np.full(nm, -0.8), np.full(nm, -20.0)])
    upper = np.stack([rob.q_upper[iq] - 0.05, np.full(nm, 0.8), np.full(nm, 20.0)])
    eng, orc = BatchedEngine(rob, sc.options, n_env, api_=api), OracleBatch(rob, sc.options, n_env)
    for x in (eng, orc):
        x.set_pd_controller_full(sc.kp, sc.kd, lower, upper, None)
        x.set_command(np.zeros((n_env, nm)))
    eng.start(sc.q0, sc.v0)
    assert not orc.start(sc.q0, sc.v0).any()
    np.testing.assert_allclose(eng.get_pd_controller_state(), orc.get_pd_controller_state(), rtol=0, atol=1e-13)
    deadband = np.full(nm, 0.02)
    adapter = PDAdapter(eng, lower, upper, order=order, is_instantaneous=instantaneous, velocity_deadband=deadband, step_dt=sc.step_dt)
    rng = np.random.default_rng(13)
    for k in range(n_steps):
        if order == 0:
            act = sc.target0 + rng.uniform(-0.03, 0.03, size=(n_env, nm))
        else:
            act = rng.uniform(-0.3, 0.3, size=(n_env, nm))
        adapter.apply(act)
        st, out = orc.get_pd_controller_state(), np.zeros((n_env, nm))
        pd_adapter(act.copy(), order, st, lower, upper, instantaneous, deadband, sc.step_dt, out)
        if instantaneous:
            orc.set_pd_controller_state(st)
        orc.set_command(out)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
        compare(eng, orc, 1e-9, 1e-7)
        np.testing.assert_allclose(eng.get_pd_controller_state(), orc.get_pd_controller_state(), rtol=0, atol=1e-10)
    return eng, orc


def bounds_handoff_scenario(api, n_env=9, n_steps=3, tol_state=1e-8):
    """ANYmal envs of which every third is driven into its hip position bounds (PD targets beyond the limits): inside
    one warp some envs stay on the fast kernel while others abort and are redone by the full kernel with their
    joint-bound constraints -- all of them must match the oracle."""
    sc = scenarios.make("anymal", n_env, seed=8)
    rob = sc.robot
    eng, orc = make_pair(sc, api)
    compare(eng, orc, 1e-13, 1e-12)
    iq = np.array([rob.idx_q[m.joint] for m in rob.motors])
    haa = [k for k, m in enumerate(rob.motors) if "HAA" in m.name]
    hit_any = False
    for k in range(n_steps):
        act = sc.sample_targets(k)
        for j in haa:
            act[::3, j] = rob.q_upper[iq[j]] + 0.3      # beyond the upper bound
        eng.set_command(act)
        orc.set_command(act)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
        compare(eng, orc, tol_state, 1e-6)
        np.testing.assert_allclose(eng.get_efforts()[0], orc.get_efforts()[0], rtol=0, atol=1e-5 * max(1.0, np.abs(orc.get_efforts()[0]).max()))
        hit_any = hit_any or bool((eng.get_status() & 8).any())
    st = eng.get_status()
    assert hit_any and (st[::3] & 8).all() and not (st[1::3] & 8).any()     # only the driven envs touched their bounds
    return eng, orc


def stateful_handoff_scenario(api, n_env=6, n_steps=6, tol_state=1e-8):
    """Hand-off from the hot-path body to the full body with the STATEFUL device blocks enabled (PDController targets
    integrated by integrate_zoh, MahonyFilter): every third env drives its hip targets beyond the joint bounds, so its
    step is aborted midway and replayed from the top -- the replay must not see targets / filter states that the
    aborted pass already advanced."""
    sc = scenarios.make("anymal", n_env, seed=8)
    rob = sc.robot
    nm = rob.nmotors
    iq = np.array([rob.idx_q[m.joint] for m in rob.motors])
    haa = [k for k, m in enumerate(rob.motors) if "HAA" in m.name]
    lower = np.stack([rob.q_lower[iq] - 0.5, np.full(nm, -4.0), np.full(nm, -40.0)])
    upper = np.stack([rob.q_upper[iq] + 0.5, np.full(nm, 4.0), np.full(nm, 40.0)])     # targets may leave the joint bounds
    eng, orc = BatchedEngine(rob, sc.options, n_env, api_=api), OracleBatch(rob, sc.options, n_env)
    rng = np.random.default_rng(12)
    act = np.zeros((n_env, nm))
    for x in (eng, orc):
        x.set_pd_controller_full(sc.kp, sc.kd, lower, upper, None)
        x.set_mahony_filter(1.0, 0.1)
        x.set_command(act)
    eng.start(sc.q0, sc.v0)
    assert not orc.start(sc.q0, sc.v0).any()
    compare(eng, orc, 1e-13, 1e-11)
    hit_any = False
    for k in range(n_steps):
        act = rng.uniform(-3.0, 3.0, size=(n_env, nm))
        for j in haa:
            act[::3, j] = 40.0        # the hip targets accelerate towards (and beyond) the upper bound
        eng.set_command(act)
        orc.set_command(act)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
        compare(eng, orc, tol_state, 1e-6)
        np.testing.assert_allclose(eng.get_pd_controller_state(), orc.get_pd_controller_state(), rtol=0, atol=1e-9)
        np.testing.assert_allclose(eng.get_mahony_filter(), orc.get_mahony_filter(), rtol=0, atol=1e-8)
        hit_any = hit_any or bool((eng.get_status() & 8).any())
    st = eng.get_status()
    assert hit_any and (st[::3] & 8).all() and not (st[1::3] & 8).any()
    return eng, orc


def mahony_scenario(api, name="anymal", n_env=3, n_steps=3):
    """Device-side MahonyFilter observer against the oracle's (pinned by golden vectors of the reference's own
    `mahony_filter` / `matrices_to_quat`): exact initialisation at start, one iteration per sensor refresh."""
    sc = scenarios.make(name, n_env, seed=5)
    eng, orc = BatchedEngine(sc.robot, sc.options, n_env, api_=api), OracleBatch(sc.robot, sc.options, n_env)
    for x in (eng, orc):
        x.set_pd_controller(sc.kp, sc.kd)
        x.set_mahony_filter(1.0, 0.1)
        x.set_command(sc.target0)
    v0 = sc.v0.copy()
    v0[:, 3:6] = [0.3, -0.2, 0.5]          # spinning base: the filter has something to track
    eng.start(sc.q0, v0)
    assert not orc.start(sc.q0, v0).any()
    np.testing.assert_allclose(eng.get_mahony_filter(), orc.get_mahony_filter(), rtol=0, atol=1e-14)
    for k in range(n_steps):
        act = sc.sample_targets(k)
        eng.set_command(act)
        orc.set_command(act)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
        compare(eng, orc, 1e-9, 1e-7)
        np.testing.assert_allclose(eng.get_mahony_filter(), orc.get_mahony_filter(), rtol=0, atol=1e-8)
    m = eng.get_mahony_filter()
    assert np.all(np.abs(np.linalg.norm(m[:, :, :4], axis=2) - 1.0) < 1e-6) and np.abs(m[:, :, 7:]).max() > 1e-3
    return eng, orc
