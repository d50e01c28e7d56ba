"""Telemetry log in the reference's binary format, for rollouts produced by the batched engine.

Layout written by `TelemetryRecorder::flushSnapshot / writeLog` and `TelemetryData::formatHeader`
(core/src/telemetry/telemetry_recorder.cc:121-171, core/src/telemetry/telemetry_data.cc:39-116) and parsed back by
`TelemetryRecorder::readLog / parseLogDataRaw` (telemetry_recorder.cc:173-445):

    int32  version (= 1, little endian)
    "StartConstants\\0"
    "StartLine" <name> "=" <value> "\\0"            for every constant, `Global.TIME_UNIT` last of the user ones,
    "StartLineNumIntEntries=<n_int + 1>\\0"         then the two counters the reader relies on
    "StartLineNumFloatEntries=<n_float>\\0"
    "StartColumns\\0" "Global.Time\\0" <integer names \\0 ...> <float names \\0 ...> "StartData\\0"
    then one line per snapshot:  "StartLine" int64 round(t / time_unit)  int64[n_int]  float64[n_float]

Variable names follow `Model::refreshProxies` (core/src/robot/model.cc:1313-1365: `current[Freeflyer]Position<Joint><suffix>`,
joint name without its "Joint" suffix, suffixes of core/src/utilities/pinocchio.cc:162-205), `Robot` motor commands
(`currentCommand<motor>`, robot.cc:256), `energy`, and sensors as `<SensorType>.<sensor>.<field>`
(abstract_sensor.hxx:275-278, basic_sensors.cc:66-67,194,284-285,394,546); everything is prefixed by `<robot name>.` when
the robot has a name (engine.cc:595-632).  Which groups are logged follows the `telemetry.enable*` options
(engine.h:330-339: configuration, velocity, acceleration on by default).

What cannot be reproduced: the `robot` constant is a Boost.Serialization archive of the C++ robot (engine.cc:1505), so
`build_robot_from_log` of the reference needs the robot passed explicitly; every variable reads back by name.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import model as M

TELEMETRY_VERSION = 1                       # telemetry/fwd.h:10
START_CONSTANTS, START_COLUMNS, START_LINE_TOKEN, START_DATA = b"StartConstants", b"StartColumns", b"StartLine", b"StartData"
NUM_INTS, NUM_FLOATS = "NumIntEntries", "NumFloatEntries"
GLOBAL_TIME, TIME_UNIT = "Global.Time", "Global.TIME_UNIT"      # constants.h:14-16
CONSTANT_DELIMITER, FIELDNAME_DELIMITER = "=", "."
STEPPER_MIN_TIMESTEP = 1e-10                # constants.h:18 = Engine::getTelemetryTimeUnit (engine.cc:2902-2905)

_POSITION_SUFFIXES = {"free": ("TransX", "TransY", "TransZ", "QuatX", "QuatY", "QuatZ", "QuatW"), "unbounded": ("Cos", "Sin"), "1dof": ("",),
                      "spherical": ("QuatX", "QuatY", "QuatZ", "QuatW")}      # utilities/pinocchio.cc:155-205
_VELOCITY_SUFFIXES = {"free": ("LinX", "LinY", "LinZ", "AngX", "AngY", "AngZ"), "unbounded": ("",), "1dof": ("",),
                      "spherical": ("AngX", "AngY", "AngZ")}
SENSOR_FIELDS = {"ImuSensor": ("GyroX", "GyroY", "GyroZ", "AccelX", "AccelY", "AccelZ"),
                 "ForceSensor": ("FX", "FY", "FZ", "MX", "MY", "MZ"), "EncoderSensor": ("Q", "V"),
                 "EffortSensor": ("U",), "ContactSensor": ("FX", "FY", "FZ")}


def _json_default(x):
    return x.tolist() if isinstance(x, np.ndarray) else float(x)


def _circumfix(name: str, prefix: str) -> str:
    """`addCircumfix(name, prefix, {}, ".")` (core/src/utilities/helpers.cc:144-158)."""
    return f"{prefix}{FIELDNAME_DELIMITER}{name}" if prefix else name


def _joint_class(joint_type: int) -> str:
    if joint_type == M.JB_JOINT_FREEFLYER:
        return "free"
    if joint_type == M.JB_JOINT_SPHERICAL:
        return "spherical"
    if joint_type in (M.JB_JOINT_RUBX, M.JB_JOINT_RUBY, M.JB_JOINT_RUBZ, M.JB_JOINT_RUBU):
        return "unbounded"
    return "1dof"


def log_fieldnames(robot: M.RobotTable) -> Dict[str, List[str]]:
    """Position / velocity / acceleration / effort / command fieldnames of a robot, unprefixed."""
    out: Dict[str, List[str]] = {"position": [], "velocity": [], "acceleration": [], "effort": [], "command": []}
    for j in range(1, robot.njoints):
        name = robot.joint_names[j]
        short = name[:-5] if len(name) > 5 and name.endswith("Joint") else name      # removeSuffix(name, "Joint")
        cls = _joint_class(int(robot.joint_type[j]))
        prefix = "current"
        if cls == "free":
            prefix, short = "currentFreeflyer", ""
        out["position"] += [f"{prefix}Position{short}{s}" for s in _POSITION_SUFFIXES[cls]]
        for s in _VELOCITY_SUFFIXES[cls]:
            out["velocity"].append(f"{prefix}Velocity{short}{s}")
            out["acceleration"].append(f"{prefix}Acceleration{short}{s}")
            out["effort"].append(f"{prefix}Effort{short}{s}")
    out["command"] = [f"currentCommand{m.name}" for m in robot.motors]
    return out


def sensor_fieldnames(robot: M.RobotTable) -> Tuple[List[str], np.ndarray]:
    """(names, index into the flattened sensor row) of every sensor value, sensor by sensor."""
    lay = robot.sensor_layout()
    per_type = {"ImuSensor": robot.imu_names, "ForceSensor": robot.force_names, "EncoderSensor": robot.encoder_names,
                "EffortSensor": robot.effort_names, "ContactSensor": robot.contact_sensor_names}
    names, index = [], []
    for stype in sorted(per_type):                    # the robot keeps its sensors in a map keyed by type name
        off, nf, ns = lay[stype]
        for s, sname in enumerate(per_type[stype]):
            for f, fname in enumerate(SENSOR_FIELDS[stype]):
                names.append(f"{stype}{FIELDNAME_DELIMITER}{sname}{FIELDNAME_DELIMITER}{fname}")
                index.append(off + f * ns + s)        # the sensor row is field-major per type
    return names, np.asarray(index, dtype=np.int64)


class TelemetryRecorder:
    """Accumulates snapshots of ONE env and writes them as a reference-format binary log."""

    def __init__(self, robot: M.RobotTable, options: Optional[dict] = None, constants: Optional[Dict[str, str]] = None,
                 time_unit: float = STEPPER_MIN_TIMESTEP, robot_name: str = ""):
        tel = dict(M.default_engine_options().get("telemetry", {}))
        if options is not None:
            tel.update(options.get("telemetry", {}))
        self.robot, self.time_unit = robot, float(time_unit)
        fn = log_fieldnames(robot)
        pre = robot_name            # the engine-side name of the robot ("" for the single robot of a Simulator)
        self._groups: List[Tuple[str, List[str]]] = []
        for key, flag in (("position", "enableConfiguration"), ("velocity", "enableVelocity"), ("acceleration", "enableAcceleration"),
                          ("effort", "enableEffort"), ("command", "enableCommand")):
            if tel.get(flag, key in ("position", "velocity", "acceleration")):
                self._groups.append((key, [_circumfix(n, pre) for n in fn[key]]))
        if tel.get("enableEnergy", False):
            self._groups.append(("energy", [_circumfix("energy", pre)]))
        snames, self._sensor_index = sensor_fieldnames(robot)
        if snames:
            self._groups.append(("sensors", [_circumfix(n, pre) for n in snames]))
        self.integer_names: List[str] = []
        self.float_names: List[str] = [n for _, names in self._groups for n in names]
        self.constants: List[Tuple[str, str]] = list((constants or {}).items())
        if options is not None:
            self.constants.append(("options", json.dumps(options, separators=(",", ":"), default=_json_default)))
        # `TelemetryRecorder::initialize` (telemetry_recorder.cc:29-35): scientific notation, 10 digits
        self.constants.append((TIME_UNIT, f"{self.time_unit:.10e}"))
        self._times: List[int] = []
        self._rows: List[np.ndarray] = []

    # ---- recording
    def append(self, t: float, q, v, a, sensors=None, u=None, command=None, energy=None) -> None:
        """One snapshot (`TelemetryRecorder::flushSnapshot`)."""
        src = {"position": q, "velocity": v, "acceleration": a, "effort": u, "command": command,
               "energy": None if energy is None else [energy],
               "sensors": None if sensors is None else np.asarray(sensors, dtype=np.float64)[self._sensor_index]}
        row = []
        for key, names in self._groups:
            val = src[key]
            if val is None:
                raise ValueError(f"telemetry group '{key}' is enabled but no value was given")
            val = np.asarray(val, dtype=np.float64).ravel()
            if val.size != len(names):
                raise ValueError(f"telemetry group '{key}': expected {len(names)} values, got {val.size}")
            row.append(val)
        self._times.append(int(round(float(t) / self.time_unit)))
        self._rows.append(np.concatenate(row) if row else np.zeros(0))

    def snapshot(self, engine, env: int = 0) -> None:
        """Pull env `env` of a `BatchedEngine` (state, efforts, sensors, energy) and append it."""
        t, q, v, a = engine.get_state()
        keys = {k for k, _ in self._groups}
        u = cmd = energy = sensors = None
        if keys & {"effort", "command"}:
            uu, _, cc, _ = engine.get_efforts()
            u, cmd = uu[env], cc[env]
        if "energy" in keys:
            energy = float(engine.get_extra_terms()[0][env].sum())
        if "sensors" in keys:
            sensors = engine.get_sensors()[env]
        self.append(float(t[env]), q[env], v[env], a[env], sensors=sensors, u=u, command=cmd, energy=energy)

    # ---- output
    def header(self) -> bytes:
        """`TelemetryData::formatHeader` (telemetry_data.cc:39-116)."""
        out = bytearray(struct.pack("<i", TELEMETRY_VERSION))

        def line(*parts) -> None:
            for p in parts:
                out.extend(p if isinstance(p, bytes) else str(p).encode())
            out.append(0)
        line(START_CONSTANTS)
        for name, value in self.constants:
            line(START_LINE_TOKEN, name, CONSTANT_DELIMITER, value)
        line(START_LINE_TOKEN, NUM_INTS, CONSTANT_DELIMITER, len(self.integer_names) + 1)
        line(START_LINE_TOKEN, NUM_FLOATS, CONSTANT_DELIMITER, len(self.float_names))
        line(START_COLUMNS)
        line(GLOBAL_TIME)
        for n in self.integer_names + self.float_names:
            line(n)
        line(START_DATA)
        return bytes(out)

    def to_bytes(self) -> bytes:
        out = bytearray(self.header())
        for t, row in zip(self._times, self._rows):
            out.extend(START_LINE_TOKEN)
            out.extend(struct.pack("<q", t))
            out.extend(np.ascontiguousarray(row, dtype="<f8").tobytes())
        return bytes(out)

    def write_log(self, path: str) -> None:
        """`Engine.write_log(path, format="binary")`."""
        with open(path, "wb") as f:
            f.write(self.to_bytes())

    @property
    def log_data(self) -> dict:
        """`Engine.log_data`-like dict: constants, times (s) and one array per variable."""
        return read_log_bytes(self.to_bytes())


class BatchTelemetryRecorder:
    """Telemetry of a batched rollout: one reference-format log per recorded env of a `BatchedEngine`, so that
    `jiminy_py.log` / `plot` / `viewer.replay` keep working on the rollouts of the batched path.  The batch is
    read once per snapshot (state, efforts, sensors, energies: one device-to-host copy each, whatever the number of
    recorded envs), not once per env."""

    def __init__(self, engine, envs: Optional[Sequence[int]] = None, options: Optional[dict] = None,
                 constants: Optional[Dict[str, str]] = None):
        self.engine = engine
        self.envs = list(range(engine.n_env)) if envs is None else [int(e) for e in envs]
        for e in self.envs:
            if not 0 <= e < engine.n_env:
                raise ValueError(f"env index {e} out of range")
        self.recorders = {e: TelemetryRecorder(engine.robot, options, constants) for e in self.envs}

    def snapshot(self) -> None:
        """Append the current state of every recorded env (call after `start` and after every `step`)."""
        eng = self.engine
        t, q, v, a = eng.get_state()
        keys = {k for k, _ in next(iter(self.recorders.values()))._groups} if self.recorders else set()
        uu = cc = energies = sensors = None
        if keys & {"effort", "command"}:
            uu, _, cc, _ = eng.get_efforts()
        if "energy" in keys:
            energies = eng.get_extra_terms()[0].sum(axis=1)
        if "sensors" in keys:
            sensors = eng.get_sensors()
        for e, rec in self.recorders.items():
            rec.append(float(t[e]), q[e], v[e], a[e], sensors=None if sensors is None else sensors[e],
                       u=None if uu is None else uu[e], command=None if cc is None else cc[e],
                       energy=None if energies is None else float(energies[e]))

    def write_logs(self, directory: str, prefix: str = "env") -> List[str]:
        """One `<prefix>_<env>.data` per recorded env (`Engine.write_log(path, format="binary")`); returns the paths."""
        import os
        os.makedirs(directory, exist_ok=True)
        paths = []
        for e, rec in self.recorders.items():
            path = os.path.join(directory, f"{prefix}_{e:06d}.data")
            rec.write_log(path)
            paths.append(path)
        return paths

    def log_data(self, env: int) -> dict:
        return self.recorders[int(env)].log_data


def read_log_bytes(buf: bytes) -> dict:
    """The reference's reader restated (`TelemetryRecorder::readLog` + `parseLogDataRaw`, telemetry_recorder.cc:173-445):
    version check, constants up to `StartColumns`, the two counters taken from the LAST two constants, variable names
    up to `StartData`, then fixed-size lines that must start with the line token."""
    (version,) = struct.unpack_from("<i", buf, 0)
    if version != TELEMETRY_VERSION:
        raise RuntimeError("Log telemetry version not supported. Impossible to read log.")
    pos = 4

    def cstring(p: int) -> Tuple[bytes, int]:
        e = buf.index(b"\0", p)
        return buf[p:e], e + 1
    tok, pos = cstring(pos)
    if tok != START_CONSTANTS:
        raise RuntimeError("Invalid log file.")
    constants: List[Tuple[str, str]] = []
    while True:
        tok, pos = cstring(pos)
        if tok == START_COLUMNS:
            break
        if not tok.startswith(START_LINE_TOKEN):
            raise RuntimeError("Invalid log file.")
        key, _, value = tok[len(START_LINE_TOKEN):].decode().partition(CONSTANT_DELIMITER)
        constants.append((key, value))
    if len(constants) < 2 or constants[-2][0] != NUM_INTS or constants[-1][0] != NUM_FLOATS:
        raise RuntimeError("Invalid log file.")
    n_int, n_float = int(constants[-2][1]) - 1, int(constants[-1][1])      # Global.Time is counted with the integers
    names: List[str] = []
    while True:
        tok, pos = cstring(pos)
        if tok == START_DATA:
            break
        names.append(tok.decode())
    if len(names) != 1 + n_int + n_float or names[0] != GLOBAL_TIME:
        raise RuntimeError("Invalid log file.")
    time_unit = STEPPER_MIN_TIMESTEP
    for k, val in constants:
        if k == TIME_UNIT:
            time_unit = float(val)
            break
    line = len(START_LINE_TOKEN) + 8 + 8 * n_int + 8 * n_float
    n_lines = (len(buf) - pos) // line
    times = np.zeros(n_lines, dtype=np.int64)
    ints, floats = np.zeros((n_int, n_lines), dtype=np.int64), np.zeros((n_float, n_lines))
    k = 0
    while k < n_lines:
        p = pos + k * line
        if buf[p:p + 1] != START_LINE_TOKEN[:1]:      # a pre-allocated chunk may not be full
            break
        p += len(START_LINE_TOKEN)
        times[k] = struct.unpack_from("<q", buf, p)[0]
        ints[:, k] = np.frombuffer(buf, dtype="<i8", count=n_int, offset=p + 8)
        floats[:, k] = np.frombuffer(buf, dtype="<f8", count=n_float, offset=p + 8 + 8 * n_int)
        k += 1
    variables = {GLOBAL_TIME: times[:k] * time_unit}
    for i, n in enumerate(names[1:1 + n_int]):
        variables[n] = ints[i, :k]
    for i, n in enumerate(names[1 + n_int:]):
        variables[n] = floats[i, :k]
    return {"version": version, "time_unit": time_unit, "constants": dict(constants[:-2]), "variables": variables,
            "variable_names": names, "times_raw": times[:k]}


def read_log(path: str) -> dict:
    with open(path, "rb") as f:
        return read_log_bytes(f.read())
