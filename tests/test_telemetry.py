"""Binary telemetry log in the reference's format: the header against the byte layout of
`TelemetryData::formatHeader` (core/src/telemetry/telemetry_data.cc:39-116) written out by hand, the field names of
`Model::refreshProxies` / the sensors, and a round trip through the restated reader (`readLog` + `parseLogDataRaw`,
telemetry_recorder.cc:173-445)."""
import os
import struct

import numpy as np
import pytest

from jiminy_b200 import model as M
from jiminy_b200 import robots as R
from jiminy_b200 import telemetry as T

from conftest import DATA


def test_header_bytes_of_a_pendulum():
    robot = M.build_robot_table(os.path.join(DATA, "simple_pendulum.urdf"), False)
    rec = T.TelemetryRecorder(robot, constants={"urdf_file": "simple_pendulum.urdf"})
    expected = struct.pack("<i", 1) + b"StartConstants\0" + \
        b"StartLineurdf_file=simple_pendulum.urdf\0" + b"StartLineGlobal.TIME_UNIT=1.0000000000e-10\0" + \
        b"StartLineNumIntEntries=1\0" + b"StartLineNumFloatEntries=3\0" + b"StartColumns\0" + b"Global.Time\0" + \
        b"currentPositionPendulum\0" + b"currentVelocityPendulum\0" + b"currentAccelerationPendulum\0" + b"StartData\0"
    assert rec.header() == expected
    rec.append(0.0, [0.1], [0.2], [0.3])
    rec.append(1e-3, [1.1], [1.2], [1.3])
    blob = rec.to_bytes()
    line = b"StartLine" + struct.pack("<q", 10_000_000) + struct.pack("<3d", 1.1, 1.2, 1.3)
    assert blob.endswith(line) and len(blob) == len(expected) + 2 * len(line)


def test_fieldnames_and_round_trip_anymal(tmp_path):
    robot, opt = R.load_robot("anymal")
    names = T.log_fieldnames(robot)
    assert names["position"][:7] == ["currentFreeflyerPosition" + s for s in ("TransX", "TransY", "TransZ", "QuatX", "QuatY", "QuatZ", "QuatW")]
    assert names["velocity"][:6] == ["currentFreeflyerVelocity" + s for s in ("LinX", "LinY", "LinZ", "AngX", "AngY", "AngZ")]
    assert len(names["position"]) == robot.nq and len(names["velocity"]) == robot.nv == len(names["acceleration"])
    assert "currentPositionLF_HAA" in names["position"] and "currentAccelerationRH_KFE" in names["acceleration"]
    assert names["command"] == ["currentCommand" + m.name for m in robot.motors]
    opt = dict(opt)
    opt["telemetry"] = {"enableConfiguration": True, "enableVelocity": True, "enableAcceleration": False,
                        "enableCommand": True, "enableEnergy": True}
    rec = T.TelemetryRecorder(robot, opt)
    assert not any("Acceleration" in n for n in rec.float_names)
    width = robot.sensor_layout()["width"][0]
    rng = np.random.default_rng(0)
    rows = []
    for k in range(5):
        q, v, a = rng.normal(size=robot.nq), rng.normal(size=robot.nv), rng.normal(size=robot.nv)
        sens, cmd, e = rng.normal(size=width), rng.normal(size=robot.nmotors), rng.normal()
        rec.append(0.04 * k, q, v, a, sensors=sens, command=cmd, energy=e)
        rows.append((q, v, sens, cmd, e))
    path = str(tmp_path / "log.data")
    rec.write_log(path)
    log = T.read_log(path)
    assert log["version"] == 1 and log["constants"]["Global.TIME_UNIT"] == "1.0000000000e-10" and "options" in log["constants"]
    var = log["variables"]
    np.testing.assert_allclose(var["Global.Time"], 0.04 * np.arange(5), rtol=0, atol=1e-10)
    off, nf, ns = robot.sensor_layout()["EncoderSensor"]
    for k, (q, v, sens, cmd, e) in enumerate(rows):
        assert var["currentFreeflyerPositionQuatW"][k] == q[6]
        assert var["currentVelocityLF_HAA"][k] == v[robot.idx_v[robot.joint_index("LF_HAA")]]
        assert var["energy"][k] == e
        assert var["currentCommand" + robot.motors[3].name][k] == cmd[3]
        s = 2
        assert var[f"EncoderSensor.{robot.encoder_names[s]}.V"][k] == sens[off + 1 * ns + s]      # field-major sensor row
        ioff, _, ins = robot.sensor_layout()["ImuSensor"]
        assert var[f"ImuSensor.{robot.imu_names[0]}.AccelZ"][k] == sens[ioff + 5 * ins]
    # a truncated pre-allocated chunk is tolerated by the reader, a wrong version is not
    assert len(T.read_log_bytes(rec.to_bytes() + b"\0" * 400)["variables"]["Global.Time"]) == 5
    with pytest.raises(RuntimeError):
        T.read_log_bytes(struct.pack("<i", 2) + rec.to_bytes()[4:])
    with pytest.raises(ValueError):
        rec.append(1.0, np.zeros(robot.nq), np.zeros(robot.nv), np.zeros(robot.nv))               # enabled groups need values


def test_robot_name_prefix_and_unbounded_joints():
    robot, _ = R.load_robot("cartpole")
    rec = T.TelemetryRecorder(robot, robot_name="cart")
    cos_sin = [n for n in rec.float_names if n.endswith(("Cos", "Sin"))]
    assert len(cos_sin) == 2 and all(n.startswith("cart.currentPosition") for n in cos_sin)
    assert any(n.startswith("cart.EncoderSensor.") and n.endswith(".Q") for n in rec.float_names)
