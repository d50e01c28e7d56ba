"""Batched counterpart of `gym_jiminy.common.envs.BaseJiminyEnv` / `WalkerJiminyEnv` for the
accelerated path: `reset()` / `step(action)` over N lockstep envs, one kernel launch per step.

Mirrors the reference flow (python/gym_jiminy/common/gym_jiminy/common/envs/generic.py:521-880):
`reset` samples an initial state (`_sample_state`: neutral posture + perturbation, feet on the
ground, :1300-1335) and starts the engine (:673-690); `step` copies the action into the controller
buffer (:806), advances the engine by `step_dt` (:810), refreshes the observation (:834),
evaluates termination (:846-875; `WalkerJiminyEnv.has_terminated`: base height under a threshold,
locomotion.py:380-400) and truncation (numerical failure -> the reference raises and the env
truncates, generic.py:809-817).  Terminated / truncated envs are restarted with a masked `start`,
which is what a vectorised gym env does between steps.

The observation is the reference's `{"t", "states": {"agent": {"q", "v"}}, "measurements": {...}}`
nested dict with a leading env axis; sensor matrices are `[n_env, n_fields, n_sensors]` views of
the flat sensor row.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from . import core, scenarios
from .model import RobotTable

SENSOR_FIELDS = {"ImuSensor": 6, "ForceSensor": 6, "EncoderSensor": 2, "EffortSensor": 1, "ContactSensor": 3}


class BatchedJiminyEnv:
    def __init__(self, scenario: scenarios.Scenario, device: int = 0, height_threshold_ratio: float = 0.5,
                 simulation_duration_max: float = 20.0, api_: Optional[core.Api] = None):
        self.sc = scenario
        self.robot: RobotTable = scenario.robot
        self.n_env, self.step_dt = scenario.n_env, scenario.step_dt
        self.engine = core.BatchedEngine(self.robot, scenario.options, self.n_env, device=device, api_=api_)
        if scenario.kp is not None:
            self.engine.set_pd_controller(scenario.kp, scenario.kd)
        self.simulation_duration_max = simulation_duration_max
        self._height_min = height_threshold_ratio * float(np.mean(scenario.q0[:, 2])) if self.robot.has_freeflyer else None
        self._layout = self.robot.sensor_layout()
        self._sens = np.zeros((self.n_env, max(self.engine.width, 1)))
        self.num_steps = np.zeros(self.n_env, dtype=np.int64)
        self._rng = np.random.default_rng(scenario.seed)
        lim = np.array([m.effort_limit for m in self.robot.motors]) if self.robot.nmotors else np.zeros(0)
        # action space bounds: motor effort limits (generic.py:344-361) or, in PD mode, joint position bounds
        if scenario.kp is None:
            self.action_low, self.action_high = -lim, lim
        else:
            iq = np.array([self.robot.idx_q[m.joint] for m in self.robot.motors])
            self.action_low, self.action_high = self.robot.q_lower[iq], self.robot.q_upper[iq]
        self._started = False

    # ------------------------------------------------------------------ helpers
    def _observation(self) -> Dict[str, Any]:
        t, q, v, _ = self.engine.get_state()
        self.engine.get_sensors(self._sens)
        meas = {}
        for name, nf in SENSOR_FIELDS.items():
            off, _, ns = self._layout[name]
            if ns:
                meas[name] = self._sens[:, off:off + nf * ns].reshape(self.n_env, nf, ns)
        return {"t": t, "states": {"agent": {"q": q, "v": v}}, "measurements": meas}

    def _sample_state(self, n: int) -> Tuple[np.ndarray, np.ndarray]:
        """Fresh draws from the scenario's initial-state distribution (perturbed posture, feet on ground)."""
        sc = scenarios.make(self.sc.name, n, seed=int(self._rng.integers(0, 2 ** 31 - 1)))
        return sc.q0, sc.v0

    # ------------------------------------------------------------------ gym API
    def reset(self, mask: Optional[np.ndarray] = None) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        if mask is None or not self._started:
            q0, v0 = (self.sc.q0, self.sc.v0) if not self._started else self._sample_state(self.n_env)
            self.engine.set_command(self.sc.target0)
            self.engine.start(q0, v0)
            self.num_steps[:] = 0
            self._started = True
        elif mask.any():
            q0, v0 = self._sample_state(self.n_env)
            self.engine.start(q0, v0, mask=mask)
            self.num_steps[mask.astype(bool)] = 0
        return self._observation(), {}

    def step(self, action: np.ndarray):
        """action: [n_env, nmotors] efforts (or position targets in PD mode).  Returns the gymnasium
        5-tuple with per-env arrays; terminated / truncated envs are restarted before returning."""
        if not self._started:
            raise core.BadControlFlow("No simulation running. Please call `reset` before `step`.")
        action = np.clip(np.asarray(action, dtype=np.float64), self.action_low, self.action_high)
        self.engine.set_command(action)
        self.engine.step(self.step_dt)
        obs = self._observation()
        self.num_steps += 1
        status = self.engine.get_status()
        q = obs["states"]["agent"]["q"]
        terminated = np.zeros(self.n_env, dtype=bool)
        if self._height_min is not None:
            terminated |= q[:, 2] < self._height_min
        truncated = (status != 0) | (self.num_steps * self.step_dt >= self.simulation_duration_max)
        reward = np.where(terminated, 0.0, 1.0)          # SurviveReward
        info = {"status": status}
        done = terminated | truncated
        if done.any():
            self.reset(mask=done.astype(np.uint8))
        return obs, reward, terminated, truncated, info

    def close(self) -> None:
        self.engine.close()
