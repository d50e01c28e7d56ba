"""ctypes mirror of the POD structs declared in `include/jiminy_b200.h` (JbModelDesc, JbOptions,
JbSensorLayout) and the marshalling of a `RobotTable` / options dict into them.

Pure data plumbing: no physics, no device code.  Used by `jiminy_b200.core` (the product binding)
and by `oracle/oracle.py` (the test oracle binding) so both consume byte-identical model tables.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List

import numpy as np

from .model import RobotTable

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)
c_int64_p = C.POINTER(C.c_int64)


class JbModelDesc(C.Structure):
    _fields_ = [
        ("njoints", C.c_int32), ("nq", C.c_int32), ("nv", C.c_int32),
        ("joint_type", c_int32_p), ("parent", c_int32_p), ("idx_q", c_int32_p), ("idx_v", c_int32_p),
        ("placement", c_double_p), ("axis", c_double_p), ("inertia", c_double_p),
        ("rotor_inertia", c_double_p), ("q_lower", c_double_p), ("q_upper", c_double_p),
        ("nmotors", C.c_int32), ("motor_joint", c_int32_p), ("motor_flags", c_int32_p),
        ("motor_params", c_double_p),
        ("ncontacts", C.c_int32), ("contact_joint", c_int32_p), ("contact_placement", c_double_p),
        ("nimu", C.c_int32), ("imu_joint", c_int32_p), ("imu_placement", c_double_p),
        ("nforce", C.c_int32), ("force_joint", c_int32_p), ("force_placement", c_double_p),
        ("nencoder", C.c_int32), ("encoder_joint", c_int32_p), ("encoder_reduction", c_double_p),
        ("neffort", C.c_int32), ("effort_motor", c_int32_p),
        ("ncontact_sensor", C.c_int32), ("contact_sensor_index", c_int32_p),
        ("flexibility", c_double_p),
    ]


class JbOptions(C.Structure):
    _fields_ = [
        ("ode_solver", C.c_int32), ("successive_iter_failed_max", C.c_int32),
        ("iter_max", C.c_int32), ("contact_model", C.c_int32),
        ("tol_abs", C.c_double), ("tol_rel", C.c_double), ("dt_max", C.c_double),
        ("dt_restore_threshold_rel", C.c_double),
        ("sensors_update_period", C.c_double), ("controller_update_period", C.c_double),
        ("contact_stiffness", C.c_double), ("contact_damping", C.c_double),
        ("contact_friction", C.c_double), ("contact_transition_eps", C.c_double),
        ("contact_transition_velocity", C.c_double), ("gravity", C.c_double * 6),
        ("contact_torsion", C.c_double), ("contact_stabilization_freq", C.c_double),
        ("constraint_regularization", C.c_double),
    ]


class JbSensorLayout(C.Structure):
    _fields_ = [("imu_offset", C.c_int32), ("force_offset", C.c_int32), ("encoder_offset", C.c_int32),
                ("effort_offset", C.c_int32), ("contact_offset", C.c_int32), ("width", C.c_int32)]


class JbStateViews(C.Structure):
    _fields_ = [("t", C.POINTER(C.c_double)), ("qv", C.POINTER(C.c_double)), ("a", C.POINTER(C.c_double)),
                ("sensors", C.POINTER(C.c_double)), ("n_env", C.c_int32), ("nq", C.c_int32), ("nv", C.c_int32),
                ("width", C.c_int32)]


SOLVERS = {"euler_explicit": 0, "runge_kutta_4": 1, "runge_kutta_dopri": 2}
CONTACT_MODELS = {"spring_damper": 0, "constraint": 1}


def _d(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def dptr(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


def iptr(a: np.ndarray):
    return a.ctypes.data_as(c_int32_p)


class ModelDescHolder:
    """Owns the numpy arrays a JbModelDesc points to (keeps them alive for the call)."""

    def __init__(self, robot: RobotTable):
        self.arrays: List[np.ndarray] = []
        d = JbModelDesc()
        d.njoints, d.nq, d.nv = robot.njoints, robot.nq, robot.nv

        def keep(a: np.ndarray) -> np.ndarray:
            if a.size == 0:  # never hand a NULL/dangling pointer to C for empty tables
                a = np.zeros(1, dtype=a.dtype)
            self.arrays.append(a)
            return a

        d.joint_type = iptr(keep(_i(robot.joint_type)))
        d.parent = iptr(keep(_i(robot.parent)))
        d.idx_q = iptr(keep(_i(robot.idx_q)))
        d.idx_v = iptr(keep(_i(robot.idx_v)))
        d.placement = dptr(keep(_d(robot.placement)))
        d.axis = dptr(keep(_d(robot.axis)))
        d.inertia = dptr(keep(_d(robot.inertia)))
        d.rotor_inertia = dptr(keep(_d(robot.rotor_inertia)))
        d.q_lower = dptr(keep(_d(robot.q_lower)))
        d.q_upper = dptr(keep(_d(robot.q_upper)))

        d.nmotors = robot.nmotors
        mp = np.zeros((max(robot.nmotors, 1), 10))
        mf = np.zeros(max(robot.nmotors, 1), dtype=np.int32)
        for k, m in enumerate(robot.motors):
            mp[k] = [m.reduction, m.effort_limit, m.velocity_limit, m.velocity_effort_inv_slope,
                     m.friction_viscous_positive, m.friction_viscous_negative,
                     m.friction_dry_positive, m.friction_dry_negative, m.friction_dry_slope, 0.0]
            mf[k] = (1 if m.enable_effort_limit else 0) | (2 if m.enable_velocity_limit else 0) | \
                    (4 if m.enable_friction else 0)
        d.motor_joint = iptr(keep(_i([m.joint for m in robot.motors])))
        d.motor_flags = iptr(keep(mf))
        d.motor_params = dptr(keep(mp))

        def frames(names):
            joints = _i([robot.frames[n].joint for n in names])
            plc = _d([robot.frames[n].placement.flat() for n in names]).reshape(-1)
            return iptr(keep(joints)), dptr(keep(plc))

        d.ncontacts = len(robot.contact_frame_names)
        d.contact_joint, d.contact_placement = frames(robot.contact_frame_names)
        d.nimu = len(robot.imu_names)
        d.imu_joint, d.imu_placement = frames(robot.imu_frames)
        d.nforce = len(robot.force_names)
        d.force_joint, d.force_placement = frames(robot.force_frames)
        d.nencoder = len(robot.encoder_names)
        d.encoder_joint = iptr(keep(_i(robot.encoder_joints)))
        d.encoder_reduction = dptr(keep(_d(robot.encoder_reduction)))
        d.neffort = len(robot.effort_names)
        d.effort_motor = iptr(keep(_i(robot.effort_motors)))
        d.ncontact_sensor = len(robot.contact_sensor_names)
        d.contact_sensor_index = iptr(keep(_i(robot.contact_sensor_index)))
        flex = getattr(robot, "flexibility", None)
        if flex is not None and len(flex):
            d.flexibility = dptr(keep(_d(flex).reshape(-1)))
        self.desc = d


def make_options(opt: Dict[str, Any]) -> JbOptions:
    """Engine option dict (reference layout, `Engine.get_options()`) -> JbOptions."""
    st, ct, world = opt["stepper"], opt["contacts"], opt["world"]
    if ct["model"] not in CONTACT_MODELS:
        raise ValueError(f"Requested contact model '{ct['model']}' not available.")
    if opt.get("constraints", {}).get("solver", "PGS") != "PGS":
        raise ValueError("Requested constraint solver not available.")
    o = JbOptions()
    o.contact_model = CONTACT_MODELS[ct["model"]]
    o.contact_torsion = float(ct.get("torsion", 0.0))
    o.contact_stabilization_freq = float(ct.get("stabilizationFreq", 20.0))
    o.constraint_regularization = float(opt.get("constraints", {}).get("regularization", 1e-3))
    o.ode_solver = SOLVERS[st["odeSolver"]]
    o.successive_iter_failed_max = int(st["successiveIterFailedMax"])
    o.iter_max = int(st.get("iterMax", 0))
    o.tol_abs, o.tol_rel = float(st["tolAbs"]), float(st["tolRel"])
    o.dt_max = float(st["dtMax"])
    o.dt_restore_threshold_rel = float(st["dtRestoreThresholdRel"])
    o.sensors_update_period = float(st["sensorsUpdatePeriod"])
    o.controller_update_period = float(st["controllerUpdatePeriod"])
    o.contact_stiffness, o.contact_damping = float(ct["stiffness"]), float(ct["damping"])
    o.contact_friction = float(ct["friction"])
    o.contact_transition_eps = float(ct["transitionEps"])
    o.contact_transition_velocity = float(ct["transitionVelocity"])
    for k in range(6):
        o.gravity[k] = float(world["gravity"][k])
    return o


def safety_table(safety, robot) -> "np.ndarray | None":
    """`MotorSafetyLimit` parameters as the [5, nmotors] table of jb_set_pd_controller_full: kp, kd, soft lower /
    upper motor position, velocity limit.  Four rows are accepted for the common case soft_velocity_max = inf: the
    fifth is then the motors' own velocity limit (motor_safety_limit.py:172-175)."""
    if safety is None:
        return None
    nm = robot.nmotors
    sf = np.asarray(safety, dtype=np.float64)
    if sf.shape == (4, nm):
        sf = np.concatenate([sf, np.array([[m.velocity_limit for m in robot.motors]], dtype=np.float64)], axis=0)
    return np.ascontiguousarray(sf.reshape(5, nm))
