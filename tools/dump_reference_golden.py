#!/usr/bin/env python
"""Golden trajectories of the REAL reference (jiminy 1.8.12), for `tests/test_reference_golden.py`.

Run on a machine where `import jiminy_py` works (pip install jiminy_py==1.8.12), from the repo root:

    python tools/dump_reference_golden.py --data /path/to/jiminy/data --out tests/golden

It reproduces, through the reference's own public API (`jiminy_py.simulator.Simulator.build`,
`jiminy.FunctionalController`, `Engine.start` / `Engine.step`; core/src/engine/engine.cc:952-1533, :1724-2417), the
very scenarios `jiminy_b200.scenarios.make` defines for the BASELINE configs -- same URDF / hardware files, same
engine options (taken from jiminy_b200/robots/<name>.json, i.e. what the loader of this repo derived from the
reference's TOML files, plus `robots.baseline_options`), same initial state (env 0 of the seeded batch), same PD law,
same per-env-step targets -- and records after every `Engine.step(step_dt)`: t, q, v, a, the motor efforts and the
sensor matrix of each type.  The file also carries the reference's own joint / motor / sensor / contact ordering so
that the "bit-identical joint indexing" part of the north star is checked against the real thing, not against
SURVEY.md's Appendix C.

STATUS: this script has never been executed in the build container (jiminy cannot be installed there: no Eigen /
Boost / Pinocchio, no network).  The consumer test skips while no `tests/golden/reference_*.npz` exists, and parity
against the reference binary stays UNPINNED until someone runs this once and commits the files.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {   # name -> (urdf relative to --data, has_freeflyer, number of env-steps recorded)
    "double_pendulum": ("toys_models/double_pendulum/double_pendulum.urdf", False, 250),
    "cartpole": ("toys_models/cartpole/cartpole.urdf", False, 100),
    "anymal": ("quadrupedal_robots/anymal/anymal.urdf", True, 250),
    "atlas": ("bipedal_robots/atlas/atlas.urdf", True, 50),
    # ANYmal with `dynamics.enableFlexibility` (scenarios.FLEXIBLE_ANYMAL_CONFIG): pins the flexibility joints
    # (Engine::computeInternalDynamics, engine.cc:3367-3391) and the joint order of the extended model
    "anymal_flexible": ("quadrupedal_robots/anymal/anymal.urdf", True, 50),
}


def _set_nested(dst, src):
    """Overwrite the leaves of the reference's option dict that our option dict also has (same camelCase keys)."""
    for k, v in src.items():
        if k not in dst:
            continue
        if isinstance(v, dict):
            _set_nested(dst[k], v)
        else:
            dst[k] = type(dst[k])(v) if not isinstance(dst[k], (list, np.ndarray)) else np.asarray(v, dtype=np.float64)


def dump(name: str, data_dir: str, out_dir: str) -> str:
    import jiminy_py.core as jiminy
    from jiminy_py.simulator import Simulator

    from jiminy_b200 import scenarios

    urdf_rel, has_freeflyer, n_steps = CONFIGS[name]
    sc = scenarios.make(name, 1, seed=0)
    sim = Simulator.build(os.path.join(data_dir, urdf_rel), has_freeflyer=has_freeflyer, config_path="")
    robot, engine = sim.robot, sim.engine
    if name == "atlas":   # the env's contact-point clean-up (gym_jiminy/envs/atlas.py:95-111), needed for equal contact sets
        from gym_jiminy.envs.atlas import _cleanup_contact_points
        _cleanup_contact_points(robot)
    if name == "anymal_flexible":
        model_options = robot.get_model_options()
        model_options["dynamics"]["enableFlexibility"] = True
        model_options["dynamics"]["flexibilityConfig"] = [
            {k: (np.asarray(v, dtype=np.float64) if k != "frameName" else v) for k, v in cfg.items()}
            for cfg in scenarios.FLEXIBLE_ANYMAL_CONFIG]
        robot.set_model_options(model_options)
    opts = engine.get_options()
    _set_nested(opts, sc.options)
    engine.set_options(opts)

    motors = list(robot.motors)
    mq = np.array([m.joint_position_index for m in motors])
    mv = np.array([m.joint_velocity_index for m in motors])
    red = np.array([m.get_options()["mechanicalReduction"] for m in motors])
    lim = np.array([m.effort_limit for m in motors])
    target = sc.target0[0].copy()

    def compute_command(t, q, v, sensor_measurements, command):
        if sc.kp is None:
            command[:] = target
            return
        tau = sc.kp * ((target - q[mq] * red) + sc.kd * (0.0 - v[mv] * red))
        command[:] = np.clip(tau, -lim, lim)

    robot.controller = jiminy.FunctionalController(compute_command, None)
    engine.start(sc.q0[0], sc.v0[0])
    rec = {k: [] for k in ("t", "q", "v", "a", "u_motor")}
    sensors = {}

    def snap():
        st, rs = engine.stepper_state, engine.robot_states[0]
        rec["t"].append(st.t)
        rec["q"].append(np.array(rs.q)); rec["v"].append(np.array(rs.v)); rec["a"].append(np.array(rs.a))
        rec["u_motor"].append(np.array(rs.u_motor))
        for stype, tree in robot.sensor_measurements.items() if hasattr(robot.sensor_measurements, "items") else ():
            sensors.setdefault(stype, []).append(np.array(tree))

    snap()
    for k in range(n_steps):
        target[:] = sc.sample_targets(k)[0]
        engine.step(sc.step_dt)
        snap()
    engine.stop()

    pin_model = robot.pinocchio_model
    meta = {
        "jiminy_version": getattr(jiminy, "__version__", "?"),
        "name": name, "step_dt": sc.step_dt, "n_steps": n_steps, "seed": 0,
        "joint_names": list(pin_model.names),
        "idx_q": [int(pin_model.joints[i].idx_q) for i in range(pin_model.njoints)],
        "idx_v": [int(pin_model.joints[i].idx_v) for i in range(pin_model.njoints)],
        "motor_names": [m.name for m in motors],
        "contact_frame_names": list(robot.contact_frame_names),
        "sensor_names": {stype: [s.name for s in robot.sensors[stype]] for stype in robot.sensors},
        "options": json.loads(json.dumps(sc.options, default=lambda x: np.asarray(x).tolist())),
    }
    path = os.path.join(out_dir, f"reference_{name}.npz")
    np.savez_compressed(path, meta=json.dumps(meta), **{k: np.stack(v) for k, v in rec.items()},
                        **{f"sensor_{k}": np.stack(v) for k, v in sensors.items()})
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", required=True, help="the reference's data/ directory")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--only", nargs="*", default=list(CONFIGS))
    args = ap.parse_args()
    for name in args.only:
        print("wrote", dump(name, args.data, args.out))


if __name__ == "__main__":
    main()
