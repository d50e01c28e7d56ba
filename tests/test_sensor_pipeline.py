"""Sensor measurement pipeline (delay, jitter, white noise, bias; abstract_sensor.hxx:305-522, abstract_sensor.cc:71-85):
the oracle's restatement against closed forms and distribution moments, then the device code (under the emulator) against
the oracle."""
import numpy as np
import pytest

from jiminy_b200 import scenarios
from oracle import oracle as orc_mod
from oracle.oracle import OracleBatch

from emul import emul_api
import sensor_pipeline_common as spc


def test_generators_have_the_right_distributions():
    x = orc_mod.random_draws(42, "normal", 400000)
    assert abs(x.mean()) < 5e-3 and abs(x.std() - 1.0) < 5e-3
    assert abs(((x - x.mean()) ** 4).mean() / x.var() ** 2 - 3.0) < 0.05          # kurtosis of a Gaussian
    assert 4.0 < np.abs(x).max() < 6.5                                             # the ziggurat tail is sampled
    u = orc_mod.random_draws(42, "uniform", 400000)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 2e-3 and abs(u.var() - 1.0 / 12.0) < 1e-3
    r = orc_mod.random_draws(1, "raw", 200000)
    assert r.max() < 2 ** 32 and abs(r.mean() / 2 ** 32 - 0.5) < 5e-3
    # same seed -> same stream, other seed -> other stream
    np.testing.assert_array_equal(orc_mod.random_draws(7, "raw", 16), orc_mod.random_draws(7, "raw", 16))
    assert (orc_mod.random_draws(7, "raw", 16) != orc_mod.random_draws(8, "raw", 16)).any()


def test_delay_bias_and_noise_closed_forms_on_the_oracle():
    """Encoders of the cartpole-free ANYmal: a pure delay with zero-order hold returns the sample `delay` ago, linear
    interpolation the interpolated one, bias shifts, noise has the requested standard deviation."""
    sc = scenarios.make("anymal", 1, seed=2)
    orc = OracleBatch(sc.robot, sc.options, 1)
    orc.set_pd_controller(sc.kp, sc.kd)
    P = sc.options["stepper"]["sensorsUpdatePeriod"]
    orc.set_sensor_options("EncoderSensor", 0, delay=3 * P, delay_interpolation_order=0)
    orc.set_sensor_options("EncoderSensor", 1, delay=2.5 * P, delay_interpolation_order=1)
    orc.set_sensor_options("EncoderSensor", 2, bias=[0.25, -0.5])
    orc.set_sensor_options("EncoderSensor", 3, noise_std=[0.02, 0.3])
    orc.set_command(sc.target0)
    assert not orc.start(sc.q0, sc.v0).any()
    lay = sc.robot.sensor_layout()
    off, _, ns = lay["EncoderSensor"]
    true_hist, meas_hist = [orc.get_sensor_data()[0].copy()], [orc.get_sensors()[0].copy()]
    rng = np.random.default_rng(0)
    for k in range(400):
        orc.set_command(sc.target0 + rng.uniform(-0.05, 0.05, size=sc.target0.shape))
        assert not orc.step(P).any()
        true_hist.append(orc.get_sensor_data()[0].copy())
        meas_hist.append(orc.get_sensors()[0].copy())
    T, Mz = np.array(true_hist), np.array(meas_hist)
    pos = lambda s: off + s          # noqa: E731  (field 0 = position, field 1 = velocity at off + ns + s)
    k = np.arange(10, 401)
    np.testing.assert_allclose(Mz[k, pos(0)], T[k - 3, pos(0)], rtol=0, atol=1e-15)                                   # hold, 3 samples ago
    np.testing.assert_allclose(Mz[k, pos(1)], 0.5 * (T[k - 3, pos(1)] + T[k - 2, pos(1)]), rtol=0, atol=1e-12)        # halfway between two samples
    np.testing.assert_allclose(Mz[k, pos(2)] - T[k, pos(2)], 0.25, rtol=0, atol=1e-15)
    np.testing.assert_allclose(Mz[k, off + ns + 2] - T[k, off + ns + 2], -0.5, rtol=0, atol=1e-13)
    e0, e1 = Mz[k, pos(3)] - T[k, pos(3)], Mz[k, off + ns + 3] - T[k, off + ns + 3]
    assert abs(e0.std() - 0.02) < 0.004 and abs(e1.std() - 0.3) < 0.06 and abs(e0.mean()) < 0.005
    # before the buffer is old enough the oldest real sample is returned (abstract_sensor.hxx:408-424)
    np.testing.assert_allclose(Mz[1, pos(0)], T[0, pos(0)], rtol=0, atol=1e-15)
    # untouched sensors read their true value
    np.testing.assert_array_equal(Mz[:, pos(5)], T[:, pos(5)])


def test_device_pipeline_matches_the_oracle():
    spc.pipeline_scenario(emul_api(), n_env=3, n_steps=2)


def test_hand_off_replays_the_pipeline_state(monkeypatch):
    """An env aborted by the hot-path body and replayed by the full body draws the same noise as the oracle."""
    import parity_common as pc
    monkeypatch.setenv("JB_NO_FAST_BOUNDS", "1")      # (the quadruped signature would solve the bounds on the hot path)
    api = emul_api()
    sc = scenarios.make("anymal", 6, seed=8, flagged_fraction=1.0 / 3.0)
    from jiminy_b200.core import BatchedEngine
    eng, orc = BatchedEngine(sc.robot, sc.options, 6, api_=api), OracleBatch(sc.robot, sc.options, 6)
    for x in (eng, orc):
        x.set_pd_controller(sc.kp, sc.kd)
        x.set_sensor_options("ImuSensor", 0, noise_std=[0.01] * 6, delay=0.003)
        x.set_seeds(np.arange(6, dtype=np.uint32))
        x.set_command(sc.target0)
    eng.start(sc.q0, sc.v0)
    assert not orc.start(sc.q0, sc.v0).any()
    for k in range(3):
        act = sc.sample_targets(k)
        eng.set_command(act); orc.set_command(act)
        eng.step(sc.step_dt)
        assert not orc.step(sc.step_dt, parallel=True).any()
        np.testing.assert_allclose(eng.get_sensors() - eng.get_sensor_data(), orc.get_sensors() - orc.get_sensor_data(), rtol=0, atol=1e-6)
    assert (eng.get_status()[::3] & 8).all()
