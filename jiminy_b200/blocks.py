"""Host side of gym_jiminy's controller blocks for batched rollouts.

The `PDController` block itself runs inside the step kernel (`BatchedEngine.set_pd_controller_full`).  The `PDAdapter`
block in front of it (python/gym_jiminy/common/gym_jiminy/common/blocks/proportional_derivative_controller.py:166-262,
class :538-660) is evaluated once per env-step on the action, so it stays on the host: it turns "the target motor
position (order 0) or velocity (order 1) wanted at the end of the step" into the target motor acceleration the
`PDController` holds during the step, reading -- and in its instantaneous mode updating -- the controller's command
state.  `pd_adapter` below restates the reference function for a batch of envs; golden vectors of the reference's own
function pin it (tests/golden/pd_adapter.npz, tools/make_golden_controller_blocks.py).
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def pd_adapter(action: np.ndarray, order: int, command_state: np.ndarray, command_state_lower: np.ndarray,
               command_state_upper: np.ndarray, is_instantaneous: bool, motors_velocity_deadband: Optional[np.ndarray],
               step_dt: float, out: np.ndarray) -> None:
    """`pd_adapter` (proportional_derivative_controller.py:166-262) for arrays with a leading env axis:
    action, out [n_env, nmotors]; command_state [n_env, 3, nmotors] (updated in place when `is_instantaneous`);
    bounds [3, nmotors]; deadband [nmotors] or None."""
    if abs(step_dt) < 1e-9:
        return
    if order not in (0, 1):
        raise ValueError("Derivative order of the action out-of-bounds.")
    lo, hi = np.asarray(command_state_lower), np.asarray(command_state_upper)
    if is_instantaneous:
        if order == 0:
            velocity = (action - command_state[:, 0]) / step_dt
            velocity = np.minimum(np.maximum(velocity, lo[1]), hi[1])
            if motors_velocity_deadband is not None:
                velocity[np.abs(velocity) < motors_velocity_deadband] = 0.0
            command_state[:, 0] += velocity * step_dt
            command_state[:, 1] = 0.0
        else:
            if motors_velocity_deadband is not None:
                action = action * (np.abs(action) > motors_velocity_deadband)
            acceleration = (action - command_state[:, 1]) / step_dt
            acceleration = np.minimum(np.maximum(acceleration, lo[2]), hi[2])
            command_state[:, 1] += acceleration * step_dt
        out[:] = 0.0
    else:
        velocity = (action - command_state[:, 0]) / step_dt if order == 0 else np.array(action, dtype=np.float64)
        velocity = np.minimum(np.maximum(velocity, lo[1]), hi[1])
        if motors_velocity_deadband is not None:
            velocity[np.abs(velocity) < motors_velocity_deadband] = 0.0
        out[:] = (velocity - command_state[:, 1]) / step_dt


class PDAdapter:
    """`PDAdapter` block (proportional_derivative_controller.py:538-660) in front of a `BatchedEngine` whose
    `PDController` block is enabled: `apply(action)` uploads the target accelerations of the coming env-step."""

    def __init__(self, engine, state_lower, state_upper, order: int = 1, is_instantaneous: bool = False,
                 velocity_deadband: Optional[np.ndarray] = None, step_dt: Optional[float] = None):
        if order not in (0, 1):
            raise ValueError("Derivative order of the action out-of-bounds.")      # :590-592
        self.engine, self.order, self.is_instantaneous = engine, int(order), bool(is_instantaneous)
        self.lower = np.ascontiguousarray(state_lower, dtype=np.float64).reshape(3, engine.nm)
        self.upper = np.ascontiguousarray(state_upper, dtype=np.float64).reshape(3, engine.nm)
        self.deadband = None if velocity_deadband is None else np.broadcast_to(np.asarray(velocity_deadband, dtype=np.float64), (engine.nm,)).copy()
        self.step_dt = step_dt
        self._out = np.zeros((engine.n_env, engine.nm))

    def apply(self, action: np.ndarray, step_dt: Optional[float] = None) -> np.ndarray:
        dt = float(step_dt if step_dt is not None else self.step_dt)
        state = self.engine.get_pd_controller_state()
        self._out[:] = 0.0
        pd_adapter(np.asarray(action, dtype=np.float64).reshape(self._out.shape), self.order, state, self.lower, self.upper,
                   self.is_instantaneous, self.deadband, dt, self._out)
        if self.is_instantaneous:
            self.engine.set_pd_controller_state(state)
        self.engine.set_command(self._out)
        return self._out
