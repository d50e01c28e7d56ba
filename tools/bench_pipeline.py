"""End-to-end timing of the reference's only published benchmark, restated over the batched engine.

`python/gym_jiminy/examples/pipeline_benchmark.py` steps `AtlasPDControlJiminyEnv` (MotorSafetyLimit + PDController +
PDAdapter(order=1) + MahonyFilter, atlas.py:239-295) behind FilterObservation / NormalizeObservation / FlattenObservation
with a constant action, 100 000 times: 27.4 s = 3.65 k env-steps/s on one CPU thread.  Here the same blocks
with the same arguments run for N lockstep envs: `PDControlBatchedEnv.step(action)` (host adapter, H2D of the target
accelerations, step kernel with controller / safety limits / observer inside, D2H of state, sensors, controller state and
filter state) followed by `flatten_observation` of the same three observation leaves.  Wall-clock, everything included.

    python tools/bench_pipeline.py [--n-env 4096] [--steps 10] [--warmup 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# atlas.py:28-37, :77-78
MOTOR_POSITION_MARGIN, MOTOR_VELOCITY_SAFE_GAIN, MOTOR_VELOCITY_MAX, MOTOR_ACCELERATION_MAX = 0.02, 0.15, 4.0, 30.0
MAHONY_KP, MAHONY_KI = 0.75, 0.057
KEYS = [("states", "pd_controller"), ("measurements", "EncoderSensor"), ("features", "mahony_filter")]


def make_env(n_env: int, api_=None, robot: str = "atlas"):
    from jiminy_b200 import scenarios
    from jiminy_b200.envs import PDControlBatchedEnv
    sc = scenarios.make(robot, n_env, seed=0, contact_model="constraint", solver="euler_explicit", dt_max=0.005)
    return PDControlBatchedEnv(
        sc, joint_position_margin=0.0, joint_velocity_limit=MOTOR_VELOCITY_MAX, joint_acceleration_limit=MOTOR_ACCELERATION_MAX,
        safety=dict(kp=1.0 / MOTOR_POSITION_MARGIN, kd=MOTOR_VELOCITY_SAFE_GAIN, soft_position_margin=0.0, soft_velocity_max=MOTOR_VELOCITY_MAX),
        order=1, mahony=(MAHONY_KP, MAHONY_KI), api_=api_)


def run(n_env: int, steps: int, warmup: int, api_=None, robot: str = "atlas") -> dict:
    from jiminy_b200.envs import flatten_observation
    env = make_env(n_env, api_, robot)
    low = {KEYS[0]: env.command_state_lower[:2]}
    high = {KEYS[0]: env.command_state_upper[:2]}
    obs, _ = env.reset()
    action = np.zeros((n_env, env.robot.nmotors))          # `env.action` after reset: target velocities = 0
    flat = flatten_observation(obs, KEYS, low, high)
    for _ in range(warmup):
        obs, *_ = env.step(action)
        flat = flatten_observation(obs, KEYS, low, high)
    t0 = time.perf_counter()
    n_done = 0
    for _ in range(steps):
        obs, reward, terminated, truncated, info = env.step(action)
        flat = flatten_observation(obs, KEYS, low, high)
        n_done += int((terminated | truncated).sum())
    dt = time.perf_counter() - t0
    status = env.engine.get_status()
    out = {"metric": "env_steps_per_sec", "unit": "env-steps/s", "value": n_env * steps / dt, "ms_per_step": 1e3 * dt / steps,
           "n_env": n_env, "steps": steps, "warmup": warmup, "timing": "wall clock around env.step + flatten_observation",
           "config": {"workload": f"{robot} PD-control pipeline (MotorSafetyLimit + PDController + PDAdapter(order=1) + MahonyFilter), "
                                  f"{n_env} envs, step_dt {env.step_dt}, constant action",
                      "lane_plan": env.engine.describe(), "observation_width": int(flat.shape[1])},
           "published_reference": {"value": 100000 / 27.4, "unit": "env-steps/s", "source": "pipeline_benchmark.py:46, one CPU thread"},
           "envs_restarted": n_done, "envs_flagged": int(((status & ~8) != 0).sum()),
           "base_height_min": float(obs["states"]["agent"]["q"][:, 2].min())}
    env.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-env", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    print(json.dumps(run(a.n_env, a.steps, a.warmup)))
