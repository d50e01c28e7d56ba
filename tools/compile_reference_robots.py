#!/usr/bin/env python
"""Compile the robots of the BASELINE configs from the reference's data files into the flat tables
shipped under `jiminy_b200/robots/*.json`.

Inputs (read-only, only available in the build container): `/root/reference/data/**` URDF,
`*_hardware.toml`, `*_options.toml`.  Output: `RobotTable` + engine options per robot.  Run again
whenever `jiminy_b200/model.py` changes:   python tools/compile_reference_robots.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from jiminy_b200 import model as M  # noqa: E402
from jiminy_b200 import robots as R  # noqa: E402

DATA = os.environ.get("JIMINY_REFERENCE_DATA", "/root/reference/data")


def compile_legged(name, subdir, urdf, neutral_fn=None, cleanup=False):
    base = os.path.join(DATA, subdir)
    robot = M.build_robot_table(os.path.join(base, urdf), has_freeflyer=True)
    M.load_hardware_description_file(robot, os.path.join(base, urdf.replace(".urdf", "_hardware.toml")),
                                     avoid_instable_collisions=True)
    opt = R.simulator_build_options()
    M.load_options_file(opt, os.path.join(base, urdf.replace(".urdf", "_options.toml")))
    if cleanup:
        q = robot.neutral() if neutral_fn is None else neutral_fn(robot)
        R.cleanup_contact_points(robot, q)
    return robot, opt


def atlas_neutral(robot):
    # AtlasJiminyEnv._neutral (python/gym_jiminy/envs/gym_jiminy/envs/atlas.py:147-166)
    hip = 0.2  # NEUTRAL_SAGITTAL_HIP_ANGLE (atlas.py:29)
    q = robot.neutral()

    def iq(n):
        return int(robot.idx_q[robot.joint_index(n)])
    q[iq("back_bky")] = hip
    q[iq("l_arm_elx")] = hip
    q[iq("l_arm_shx")] = -np.pi / 2.0
    q[iq("l_arm_shz")] = np.pi / 4.0
    q[iq("l_arm_ely")] = np.pi / 4.0 + np.pi / 2.0
    q[iq("r_arm_elx")] = -hip
    q[iq("r_arm_shx")] = np.pi / 2.0
    q[iq("r_arm_shz")] = -np.pi / 4.0
    q[iq("r_arm_ely")] = np.pi / 4.0 + np.pi / 2.0
    return q


def main():
    meta = {"generator": "tools/compile_reference_robots.py", "source": "duburcqa/jiminy data/ @ v1.8.12",
            "joint_order": "alphabetical"}
    # -- config 3: ANYmal
    robot, opt = compile_legged("anymal", "quadrupedal_robots/anymal", "anymal.urdf")
    print(R.save_robot("anymal", robot, opt, meta), robot.njoints, robot.nq, robot.nv, robot.contact_frame_names)
    # -- config 4: Atlas
    robot, opt = compile_legged("atlas", "bipedal_robots/atlas", "atlas.urdf", atlas_neutral, cleanup=True)
    meta_atlas = dict(meta, neutral=atlas_neutral(robot).tolist())
    print(R.save_robot("atlas", robot, opt, meta_atlas), robot.njoints, robot.nq, robot.nv,
          len(robot.contact_frame_names), robot.contact_frame_names)
    # -- config 2: cartpole (python/gym_jiminy/envs/gym_jiminy/envs/cartpole.py:108-147)
    robot = M.build_robot_table(os.path.join(DATA, "toys_models/cartpole/cartpole.urdf"), has_freeflyer=False)
    M.attach_motor(robot, "slider_to_cart", "slider_to_cart", enableVelocityLimit=False)
    M.attach_sensor(robot, "EncoderSensor", "slider", joint_name="slider_to_cart")
    M.attach_sensor(robot, "EncoderSensor", "pole", joint_name="cart_to_pole")
    opt = M.default_engine_options()  # Simulator(robot) without build(): engine defaults
    opt["stepper"]["odeSolver"] = "euler_explicit"
    opt["stepper"]["dtMax"] = 0.02
    print(R.save_robot("cartpole", robot, opt, meta), robot.njoints, robot.nq, robot.nv)
    # -- config 1: double pendulum (core/examples/double_pendulum/double_pendulum.cc:68-126)
    robot = M.build_robot_table(os.path.join(DATA, "toys_models/double_pendulum/double_pendulum.urdf"),
                                has_freeflyer=False)
    M.attach_motor(robot, "SecondPendulumJoint", "SecondPendulumJoint")
    opt = M.default_engine_options()
    opt["contacts"].update(model="spring_damper", stiffness=1.0e6, damping=2000.0, friction=5.0,
                           transitionEps=0.001, transitionVelocity=0.01)
    opt["stepper"].update(odeSolver="runge_kutta_dopri", tolRel=1.0e-5, tolAbs=1.0e-4, dtMax=3.0e-3,
                          sensorsUpdatePeriod=1.0e-3, controllerUpdatePeriod=1.0e-3)
    print(R.save_robot("double_pendulum", robot, opt, meta), robot.njoints, robot.nq, robot.nv)


if __name__ == "__main__":
    main()
