"""Scenario shared by the CPU (emulator) and GPU suites: the sensor measurement pipeline (delay ring, noise, bias) of the
device against the oracle's restatement, on ANYmal with options on sensors of every type."""
import numpy as np

from jiminy_b200 import scenarios
from jiminy_b200.core import BatchedEngine
from oracle.oracle import OracleBatch


def configure(x, robot):
    """Same options on both sides: every mechanism at least once, mixed per sensor like a real robot description."""
    x.set_sensor_options("ImuSensor", 0, noise_std=[0.01, 0.01, 0.01, 0.1, 0.1, 0.1], bias=[0.002, -0.001, 0.0, 0.05, 0.0, -0.02],
                         delay=0.0025, jitter=0.0, delay_interpolation_order=1)
    n_enc = len(robot.encoder_names)
    for k in range(n_enc):
        if k % 3 == 0:
            x.set_sensor_options("EncoderSensor", k, noise_std=[1e-3, 1e-2], delay=0.004, jitter=0.002, delay_interpolation_order=0)
        elif k % 3 == 1:
            x.set_sensor_options("EncoderSensor", k, bias=[0.01, 0.0], delay=0.0055, delay_interpolation_order=1)
    x.set_sensor_options("EffortSensor", 1, noise_std=[0.5])
    x.set_sensor_options("ForceSensor", 2, noise_std=[1.0] * 6, bias=[0.0, 0.0, 3.0, 0.0, 0.0, 0.0], delay=0.001)
    if robot.contact_sensor_names:
        x.set_sensor_options("ContactSensor", 0, delay=0.012, delay_interpolation_order=1)


def pipeline_scenario(api, n_env=3, n_steps=3, tol=1e-9):
    sc = scenarios.make("anymal", n_env, seed=13)
    eng, orc = BatchedEngine(sc.robot, sc.options, n_env, api_=api), OracleBatch(sc.robot, sc.options, n_env)
    seeds = np.arange(n_env, dtype=np.uint32) * 7919 + 3
    for x in (eng, orc):
        x.set_pd_controller(sc.kp, sc.kd)
        x.set_mahony_filter(1.0, 0.1)
        configure(x, sc.robot)
        x.set_seeds(seeds)
        x.set_command(sc.target0)
    eng.start(sc.q0, sc.v0)
    assert not orc.start(sc.q0, sc.v0).any()
    hist = []
    for k in range(-1, n_steps):
        if k >= 0:
            act = sc.sample_targets(k)
            eng.set_command(act)
            orc.set_command(act)
            eng.step(sc.step_dt)
            assert not orc.step(sc.step_dt, parallel=True).any()
        m1, m0 = eng.get_sensors(), orc.get_sensors()
        d1, d0 = eng.get_sensor_data(), orc.get_sensor_data()
        scale = max(1.0, np.abs(d0).max())
        np.testing.assert_allclose(d1, d0, rtol=0, atol=1e3 * tol * scale)          # true values: the physics parity
        # measurements: identical noise draws (the generators are integer state machines), the rest is the physics parity
        np.testing.assert_allclose(m1 - d1, m0 - d0, rtol=0, atol=1e3 * tol * scale)
        np.testing.assert_allclose(eng.get_mahony_filter(), orc.get_mahony_filter(), rtol=0, atol=1e-7)
        hist.append((m0.copy(), d0.copy()))
    # the options did something: noise on the IMU, bias on an encoder, the delayed contact sensor lags
    m, d = hist[-1]
    assert np.abs(m - d).max() > 1e-3
    # envs differ in their noise (different seeds), same env reproduces it after a restart with the same seed
    lay = sc.robot.sensor_layout()
    imu = slice(lay["ImuSensor"][0], lay["ImuSensor"][0] + 6)
    assert np.abs((m - d)[0, imu] - (m - d)[1, imu]).max() > 1e-4
    mask = np.zeros(n_env, dtype=np.uint8)
    mask[0] = 1
    eng.start(sc.q0, sc.v0, mask=mask)
    orc.start(sc.q0, sc.v0, mask=mask)
    np.testing.assert_allclose(eng.get_sensors()[0] - eng.get_sensor_data()[0], orc.get_sensors()[0] - orc.get_sensor_data()[0],
                               rtol=0, atol=1e3 * tol * max(1.0, np.abs(orc.get_sensor_data()).max()))
    return eng, orc
