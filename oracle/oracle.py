"""TEST INFRASTRUCTURE -- Python binding of the CPU oracle (`oracle/liboracle.so`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.  The product (`jiminy_b200/`) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from typing import Callable, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from jiminy_b200._ctypes_abi import (JbModelDesc, JbOptions, JbSensorLayout, ModelDescHolder,  # noqa: E402
                                     c_double_p, c_int32_p, c_int64_p, c_uint8_p, dptr, make_options, safety_table)
from jiminy_b200.model import RobotTable  # noqa: E402

_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("engine.cpp", "constraints.cpp", "controllers.cpp", "capi.cpp", "engine.hpp", "spatial.hpp")]
    srcs.append(os.path.join(_HERE, "..", "include", "jiminy_b200.h"))
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B", "liboracle.so"], check=True)
    return _LIB_PATH


_lib = None

CONTROLLER_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_double, c_double_p, c_double_p, c_double_p, c_double_p)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(JbModelDesc), C.POINTER(JbOptions), C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_sensor_width.argtypes = [C.c_void_p]
        L.orc_sensor_layout.argtypes = [C.c_void_p, C.POINTER(JbSensorLayout)]
        L.orc_set_options.argtypes = [C.c_void_p, C.POINTER(JbOptions)]
        L.orc_set_callbacks.argtypes = [C.c_void_p, C.c_int, CONTROLLER_FN, CONTROLLER_FN, C.c_void_p]
        L.orc_set_springs.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.orc_set_pd.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.orc_integrate_zoh.argtypes = [c_double_p] * 3 + [C.c_int, C.c_double]
        L.orc_pd_controller.argtypes = [c_double_p] * 7 + [C.c_int, C.c_double, c_double_p]
        L.orc_apply_safety_limits.argtypes = [c_double_p] * 9 + [C.c_int, c_double_p]
        L.orc_set_pd_full.argtypes = [C.c_void_p] + [c_double_p] * 5
        L.orc_mahony_filter.argtypes = [c_double_p] * 5 + [C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_matrix_to_quat.argtypes = [c_double_p, c_double_p]
        L.orc_set_mahony.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_get_mahony.argtypes = [C.c_void_p, c_double_p]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_stop.argtypes = [C.c_void_p]
        L.orc_register_impulse_force.argtypes = [C.c_void_p, C.c_int] + [c_double_p] * 4
        L.orc_set_impulse_force.argtypes = [C.c_void_p, C.c_int, c_uint8_p] + [c_double_p] * 3
        L.orc_register_profile_force.argtypes = [C.c_void_p, C.c_int, c_double_p, C.c_double]
        L.orc_set_profile_force.argtypes = [C.c_void_p, C.c_int, c_double_p]
        L.orc_remove_all_forces.argtypes = [C.c_void_p]
        L.orc_start.argtypes = [C.c_void_p, c_uint8_p, c_double_p, c_double_p, c_int32_p]
        L.orc_set_command.argtypes = [C.c_void_p, c_double_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_double, C.c_int, c_int32_p]
        L.orc_get_state.argtypes = [C.c_void_p] + [c_double_p] * 4
        L.orc_get_efforts.argtypes = [C.c_void_p] + [c_double_p] * 4
        L.orc_get_stepper_state.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.orc_get_sensors.argtypes = [C.c_void_p, c_double_p]
        L.orc_get_sensor_data.argtypes = [C.c_void_p, c_double_p]
        L.orc_set_sensor_options.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, C.c_double, C.c_double, C.c_int]
        L.orc_set_seeds.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_random_draws.argtypes = [C.c_uint32, C.c_int, C.c_int, c_double_p]
        L.orc_get_extra_terms.argtypes = [C.c_void_p] + [c_double_p] * 3
        L.orc_get_status.argtypes = [C.c_void_p, c_int32_p]
        L.orc_get_centroidal.argtypes = [C.c_void_p] + [c_double_p] * 5
        L.orc_get_iters.argtypes = [C.c_void_p, c_int64_p, c_int64_p]
        L.orc_rhs_count.argtypes = [C.c_void_p]
        L.orc_rhs_count.restype = C.c_int64
        L.orc_compute_dynamics.argtypes = [C.c_void_p] + [c_double_p] * 6
        L.orc_integrate.argtypes = [C.c_void_p] + [c_double_p] * 3
        L.orc_difference.argtypes = [C.c_void_p] + [c_double_p] * 3
        _lib = L
    return _lib


SENSOR_TYPES = ("ImuSensor", "ForceSensor", "EncoderSensor", "EffortSensor", "ContactSensor")


def random_draws(seed: int, kind: str, n: int) -> np.ndarray:
    """n draws of the restated generator PCG32(seed_seq{seed}): kind = 'normal' | 'uniform' | 'raw'."""
    out = np.zeros(n)
    lib().orc_random_draws(int(seed), {"normal": 0, "uniform": 1, "raw": 2}[kind], int(n), dptr(out))
    return out


class OracleBatch:
    """N independent single-robot engines (restated `jiminy::Engine`) stepped on the host CPU."""

    def __init__(self, robot: RobotTable, options: dict, n_env: int = 1):
        self.robot, self.n = robot, int(n_env)
        self._holder = ModelDescHolder(robot)
        self._opt = make_options(options)
        self._h = lib().orc_create(C.byref(self._holder.desc), C.byref(self._opt), self.n)
        self.nq, self.nv, self.nm, self.nj = robot.nq, robot.nv, robot.nmotors, robot.njoints
        self.width = lib().orc_sensor_width(self._h)
        self._cbs = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    @staticmethod
    def max_threads() -> int:
        return lib().orc_max_threads()

    @staticmethod
    def use_all_cores() -> int:
        """OpenMP thread count = the cores this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it
        to 1 for its workers)."""
        n = OracleBatch.usable_cores()["threads"]
        lib().orc_set_threads(int(n))
        return lib().orc_max_threads()

    @staticmethod
    def usable_cores() -> dict:
        """What this process may actually use: the affinity mask AND the cgroup CPU quota (a container that shows 128
        CPUs in its affinity mask may be throttled to a handful by `cpu.max`).  `threads` = min of the two."""
        import math
        import os
        aff = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        quota = None
        try:   # cgroup v2, then v1
            txt = open("/sys/fs/cgroup/cpu.max").read().split()
            if txt[0] != "max":
                quota = float(txt[0]) / float(txt[1])
        except Exception:
            try:
                q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
                per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    quota = q / per
            except Exception:
                pass
        threads = aff if quota is None else max(1, min(aff, int(math.ceil(quota))))
        return {"affinity": aff, "cgroup_quota_cpus": quota, "threads": threads}

    def set_options(self, options: dict) -> None:
        self._opt = make_options(options)
        lib().orc_set_options(self._h, C.byref(self._opt))

    def set_springs(self, k, d) -> None:
        k, d = np.ascontiguousarray(k, dtype=np.float64), np.ascontiguousarray(d, dtype=np.float64)
        lib().orc_set_springs(self._h, dptr(k), dptr(d))

    def set_pd_controller(self, kp, kd) -> None:
        if kp is None:
            lib().orc_set_pd(self._h, None, None)
            return
        kp = np.ascontiguousarray(np.broadcast_to(kp, (self.nm,)), dtype=np.float64)
        kd = np.ascontiguousarray(np.broadcast_to(kd, (self.nm,)), dtype=np.float64)
        lib().orc_set_pd(self._h, dptr(kp), dptr(kd))

    # ---- external forces: same calling convention as jiminy_b200.core.BatchedEngine
    def stop(self) -> None:
        lib().orc_stop(self._h)

    def _per_env(self, x, shape):
        return np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (self.n,) + shape))

    def register_impulse_force(self, joint: int, p, t, dt, force) -> int:
        p = np.ascontiguousarray(p, dtype=np.float64)
        t, dt, force = self._per_env(t, ()), self._per_env(dt, ()), self._per_env(force, (6,))
        rc = lib().orc_register_impulse_force(self._h, int(joint), dptr(p), dptr(t), dptr(dt), dptr(force))
        if rc < 0:
            raise ValueError(f"register_impulse_force failed ({rc})")
        return rc

    def set_impulse_force(self, k: int, t, dt, force, mask=None) -> None:
        t, dt, force = self._per_env(t, ()), self._per_env(dt, ()), self._per_env(force, (6,))
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        rc = lib().orc_set_impulse_force(self._h, int(k), None if m is None else m.ctypes.data_as(c_uint8_p),
                                         dptr(t), dptr(dt), dptr(force))
        if rc < 0:
            raise ValueError(f"set_impulse_force failed ({rc})")

    def register_profile_force(self, joint: int, p, update_period: float = 0.0) -> int:
        p = np.ascontiguousarray(p, dtype=np.float64)
        rc = lib().orc_register_profile_force(self._h, int(joint), dptr(p), float(update_period))
        if rc < 0:
            raise ValueError(f"register_profile_force failed ({rc})")
        return rc

    def set_profile_force(self, slot: int, force) -> None:
        force = self._per_env(force, (6,))
        if lib().orc_set_profile_force(self._h, int(slot), dptr(force)) < 0:
            raise ValueError("set_profile_force failed")

    def remove_all_forces(self) -> None:
        lib().orc_remove_all_forces(self._h)

    def set_pd_controller_full(self, kp, kd, lower, upper, safety=None) -> None:
        """gym_jiminy `PDController` block (+ `MotorSafetyLimit` when `safety` = [kp, kd, soft_lower, soft_upper])."""
        nm = self.nm
        kp = np.ascontiguousarray(np.broadcast_to(kp, (nm,)), dtype=np.float64)
        kd = np.ascontiguousarray(np.broadcast_to(kd, (nm,)), dtype=np.float64)
        lower = np.ascontiguousarray(lower, dtype=np.float64).reshape(3, nm)
        upper = np.ascontiguousarray(upper, dtype=np.float64).reshape(3, nm)
        sf = safety_table(safety, self.robot)
        lib().orc_set_pd_full(self._h, dptr(kp), dptr(kd), dptr(lower), dptr(upper), None if sf is None else dptr(sf))

    def get_constraints(self):
        nj, nc = self.robot.njoints, max(len(self.robot.contact_frame_names), 1)
        je, jl = np.zeros((self.n, nj), dtype=np.uint8), np.zeros((self.n, nj))
        ce, cl = np.zeros((self.n, nc), dtype=np.uint8), np.zeros((self.n, nc, 4))
        lib().orc_get_constraints.argtypes = [C.c_void_p, c_uint8_p, c_double_p, c_uint8_p, c_double_p]
        lib().orc_get_constraints(self._h, je.ctypes.data_as(c_uint8_p), dptr(jl), ce.ctypes.data_as(c_uint8_p), dptr(cl))
        n = len(self.robot.contact_frame_names)
        return je.astype(bool), jl, ce[:, :n].astype(bool), cl[:, :n]

    def get_pd_controller_state(self) -> np.ndarray:
        out = np.zeros((self.n, 3, self.nm))
        lib().orc_get_pd_state.argtypes = [C.c_void_p, c_double_p]
        lib().orc_get_pd_state(self._h, dptr(out))
        return out

    def set_pd_controller_state(self, state) -> None:
        state = np.ascontiguousarray(state, dtype=np.float64).reshape(self.n, 3, self.nm)
        lib().orc_set_pd_state.argtypes = [C.c_void_p, c_double_p]
        lib().orc_set_pd_state(self._h, dptr(state))

    def set_mahony_filter(self, kp: Optional[float] = 1.0, ki: float = 0.1) -> None:
        lib().orc_set_mahony(self._h, -1.0 if kp is None else float(kp), float(ki))

    def get_mahony_filter(self) -> np.ndarray:
        """[n_env, nimu, 10]: quaternion estimate (x, y, z, w), gyro bias estimate, unbiased angular velocity."""
        out = np.zeros((self.n, max(self.robot.sensor_layout()["ImuSensor"][2], 0), 10))
        if out.size:
            lib().orc_get_mahony(self._h, dptr(out))
        return out

    def set_callbacks(self, env: int, controller: Optional[Callable] = None,
                      internal_dynamics: Optional[Callable] = None) -> None:
        """`controller(t, q, v, sensors, out)` / `internal_dynamics(t, q, v, sensors, out)` write `out`
        in place, like `jiminy.FunctionalController` (controller_functor.h:15-20)."""
        nq, nv, nm, w = self.nq, self.nv, self.nm, self.width

        def wrap(fn, nout):
            if fn is None:
                return C.cast(None, CONTROLLER_FN)

            def tramp(_ctx, t, q, v, s, out):
                sens = np.ctypeslib.as_array(s, (w,)) if (w and s) else np.zeros(0)
                fn(t, np.ctypeslib.as_array(q, (nq,)), np.ctypeslib.as_array(v, (nv,)), sens,
                   np.ctypeslib.as_array(out, (nout,)))
            return CONTROLLER_FN(tramp)
        c, d = wrap(controller, nm), wrap(internal_dynamics, nv)
        self._cbs.append((c, d))
        lib().orc_set_callbacks(self._h, env, c, d, None)

    def start(self, q0, v0, mask=None) -> np.ndarray:
        q0 = np.ascontiguousarray(np.broadcast_to(q0, (self.n, self.nq)), dtype=np.float64)
        v0 = np.ascontiguousarray(np.broadcast_to(v0, (self.n, self.nv)), dtype=np.float64)
        rc = np.zeros(self.n, dtype=np.int32)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().orc_start(self._h, None if m is None else m.ctypes.data_as(c_uint8_p), dptr(q0), dptr(v0),
                        rc.ctypes.data_as(c_int32_p))
        return rc

    def set_command(self, cmd) -> None:
        cmd = np.ascontiguousarray(np.broadcast_to(cmd, (self.n, self.nm)), dtype=np.float64)
        if self.nm:
            lib().orc_set_command(self._h, dptr(cmd))

    def step(self, step_dt: float, parallel: bool = False) -> np.ndarray:
        rc = np.zeros(self.n, dtype=np.int32)
        lib().orc_step(self._h, float(step_dt), int(parallel), rc.ctypes.data_as(c_int32_p))
        return rc

    def get_state(self):
        t = np.zeros(self.n)
        q, v, a = np.zeros((self.n, self.nq)), np.zeros((self.n, self.nv)), np.zeros((self.n, self.nv))
        lib().orc_get_state(self._h, dptr(t), dptr(q), dptr(v), dptr(a))
        return t, q, v, a

    def get_stepper_state(self):
        sched, cmd = np.zeros((self.n, 6)), np.zeros((self.n, max(self.nm, 1)))
        lib().orc_get_stepper_state(self._h, dptr(sched), dptr(cmd))
        return sched, cmd[:, :self.nm]

    def get_efforts(self):
        u, um = np.zeros((self.n, self.nv)), np.zeros((self.n, max(self.nm, 1)))
        cmd, fext = np.zeros((self.n, max(self.nm, 1))), np.zeros((self.n, self.nj, 6))
        lib().orc_get_efforts(self._h, dptr(u), dptr(um), dptr(cmd), dptr(fext))
        return u, um[:, :self.nm], cmd[:, :self.nm], fext

    def get_sensors(self) -> np.ndarray:
        out = np.zeros((self.n, max(self.width, 1)))
        if self.width:
            lib().orc_get_sensors(self._h, dptr(out))
        return out[:, :self.width]

    def get_sensor_data(self) -> np.ndarray:
        """True values behind the measurements (what `sensor.data` holds in the reference)."""
        out = np.zeros((self.n, max(self.width, 1)))
        if self.width:
            lib().orc_get_sensor_data(self._h, dptr(out))
        return out[:, :self.width]

    def set_sensor_options(self, sensor_type: str, index: int, noise_std=None, bias=None, delay: float = 0.0,
                           jitter: float = 0.0, delay_interpolation_order: int = 1) -> None:
        t = SENSOR_TYPES.index(sensor_type)
        ns = None if noise_std is None else np.ascontiguousarray(noise_std, dtype=np.float64)
        bs = None if bias is None else np.ascontiguousarray(bias, dtype=np.float64)
        lib().orc_set_sensor_options(self._h, t, int(index), None if ns is None else dptr(ns), None if bs is None else dptr(bs),
                                     float(delay), float(jitter), int(delay_interpolation_order))

    def set_seeds(self, seeds) -> None:
        s = np.ascontiguousarray(seeds, dtype=np.uint32)
        assert s.shape == (self.n,)
        lib().orc_set_seeds(self._h, s.ctypes.data_as(C.POINTER(C.c_uint32)))

    def get_extra_terms(self):
        e, ja, jf = np.zeros((self.n, 2)), np.zeros((self.n, self.nj, 6)), np.zeros((self.n, self.nj, 6))
        lib().orc_get_extra_terms(self._h, dptr(e), dptr(ja), dptr(jf))
        return e, ja, jf

    def get_centroidal(self):
        """(Ycrb [n, njoints, 10], com [n, njoints, 3], vcom [n, njoints, 3], hg [n, 6], dhg [n, 6])."""
        y, c, vc = np.zeros((self.n, self.nj, 10)), np.zeros((self.n, self.nj, 3)), np.zeros((self.n, self.nj, 3))
        hg, dhg = np.zeros((self.n, 6)), np.zeros((self.n, 6))
        lib().orc_get_centroidal(self._h, dptr(y), dptr(c), dptr(vc), dptr(hg), dptr(dhg))
        return y, c, vc, hg, dhg

    def get_status(self) -> np.ndarray:
        s = np.zeros(self.n, dtype=np.int32)
        lib().orc_get_status(self._h, s.ctypes.data_as(c_int32_p))
        return s

    def get_iters(self):
        it, itf = np.zeros(self.n, dtype=np.int64), np.zeros(self.n, dtype=np.int64)
        lib().orc_get_iters(self._h, it.ctypes.data_as(c_int64_p), itf.ctypes.data_as(c_int64_p))
        return it, itf

    def rhs_count(self) -> int:
        return int(lib().orc_rhs_count(self._h))

    def compute_dynamics(self, q, v, cmd):
        q = np.ascontiguousarray(np.broadcast_to(q, (self.n, self.nq)), dtype=np.float64)
        v = np.ascontiguousarray(np.broadcast_to(v, (self.n, self.nv)), dtype=np.float64)
        cmd = np.ascontiguousarray(np.broadcast_to(cmd, (self.n, max(self.nm, 1))), dtype=np.float64)
        a, fext, u = np.zeros((self.n, self.nv)), np.zeros((self.n, self.nj, 6)), np.zeros((self.n, self.nv))
        lib().orc_compute_dynamics(self._h, dptr(q), dptr(v), dptr(cmd), dptr(a), dptr(fext), dptr(u))
        return a, fext, u

    def integrate(self, q, v):
        q, v = np.ascontiguousarray(q, dtype=np.float64), np.ascontiguousarray(v, dtype=np.float64)
        out = np.zeros(self.nq)
        lib().orc_integrate(self._h, dptr(q), dptr(v), dptr(out))
        return out

    def difference(self, q0, q1):
        q0, q1 = np.ascontiguousarray(q0, dtype=np.float64), np.ascontiguousarray(q1, dtype=np.float64)
        out = np.zeros(self.nv)
        lib().orc_difference(self._h, dptr(q0), dptr(q1), dptr(out))
        return out

    def simulate(self, t_end: float, q0, v0, step_dt: Optional[float] = None, log: bool = True):
        """`Engine::simulate` (engine.cc:1614-1699): start, then step(min(period or dtMax, tEnd - t))."""
        rc = self.start(q0, v0)
        if rc.any():
            raise ValueError(f"start failed: {rc}")
        st = self._opt
        period = min([p for p in (st.sensors_update_period, st.controller_update_period) if p > 2.3e-16],
                     default=np.inf)
        ts, qs, vs, as_ = [], [], [], []

        def snap():
            t, q, v, a = self.get_state()
            ts.append(t[0]); qs.append(q[0].copy()); vs.append(v[0].copy()); as_.append(a[0].copy())
        if log:
            snap()
        while True:
            t = self.get_state()[0][0]
            if t_end - t < 1e-6:
                break
            h = min(period if np.isfinite(period) else st.dt_max, t_end - t) if step_dt is None else \
                min(step_dt, t_end - t)
            rc = self.step(h)
            if rc.any():
                raise RuntimeError(f"step failed rc={rc} status={self.get_status()}")
            if log:
                snap()
        return np.array(ts), np.array(qs), np.array(vs), np.array(as_)
