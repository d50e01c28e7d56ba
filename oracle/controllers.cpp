// TEST INFRASTRUCTURE -- CPU oracle, not product code.  See oracle/README.md.
//
// gym_jiminy's low-level controller blocks, restated from
//   python/gym_jiminy/common/gym_jiminy/common/blocks/proportional_derivative_controller.py:23-165
//       integrate_zoh (:24-98), pd_controller (:104-165)
//   python/gym_jiminy/common/gym_jiminy/common/blocks/motor_safety_limit.py:21-86  apply_safety_limits
// PARITY: PINNED against the reference's own code -- tests/golden/controller_blocks.npz holds input / output
// vectors produced by running these very reference functions (tools/make_golden_controller_blocks.py).
#include <algorithm>
#include <cmath>

#include "engine.hpp"

namespace orc {

// state, state_min, state_max: [3][n] (position, velocity, acceleration), row-major
void integrate_zoh(double* state, const double* state_min, const double* state_max, int n, double dt) {
    if (std::fabs(dt) < 1e-9) return;
    double* position = state; double* velocity = state + n; double* acceleration = state + 2 * n;
    for (int i = 0; i < n; ++i) {
        const double position_min = state_min[i], acceleration_min = state_min[2 * n + i];
        const double position_max = state_max[i], acceleration_max = state_max[2 * n + i];
        double velocity_min = state_min[n + i], velocity_max = state_max[n + i];
        acceleration[i] = std::min(std::max(acceleration[i], acceleration_min), acceleration_max);
        const double velocity_prev = velocity[i];
        velocity[i] += acceleration[i] * dt;
        velocity[i] = std::min(std::max(velocity[i], velocity_min), velocity_max);
        // slow down early enough not to violate the acceleration limit when hitting the position bounds
        const double horizon = std::max(static_cast<double>(static_cast<long long>(std::fabs(velocity_prev) / acceleration_max / dt)) * dt, dt);
        double position_min_delta = position_min - position[i];
        double position_max_delta = position_max - position[i];
        if (horizon > dt) {
            const double drift = 0.5 * (horizon * (horizon - dt)) * acceleration_max;
            position_min_delta -= drift;
            position_max_delta += drift;
        }
        velocity_min = position_min_delta / horizon;
        velocity_max = position_max_delta / horizon;
        velocity[i] = std::min(std::max(velocity[i], velocity_min), velocity_max);
        // the velocity after hitting the bounds must be cancellable in a single step
        if (std::fabs(velocity[i]) > dt * acceleration_max) {
            velocity_min = -std::max(position_min_delta / velocity[i], dt) * acceleration_max;
            velocity_max = std::max(position_max_delta / velocity[i], dt) * acceleration_max;
            velocity[i] = std::min(std::max(velocity[i], velocity_min), velocity_max);
        }
        acceleration[i] = (velocity[i] - velocity_prev) / dt;
        position[i] += dt * velocity[i];
    }
}

// encoder_data: [2][n]; out: [n]
void pd_controller(const double* encoder_data, double* command_state, const double* lower, const double* upper,
                   const double* kp, const double* kd, const double* effort_limit, int n, double control_dt, double* out) {
    integrate_zoh(command_state, lower, upper, n, control_dt);
    for (int i = 0; i < n; ++i) {
        const double q_error = command_state[i] - encoder_data[i];
        const double v_error = command_state[n + i] - encoder_data[n + i];
        const double tau = kp[i] * (q_error + kd[i] * v_error);
        out[i] = std::min(std::max(tau, -effort_limit[i]), effort_limit[i]);
    }
}

void apply_safety_limits(const double* command, const double* q, const double* v, const double* kp, const double* kd,
                         const double* soft_lower, const double* soft_upper, const double* velocity_limit,
                         const double* effort_limit, int n, double* out) {
    for (int i = 0; i < n; ++i) {
        const double safe_velocity_lower = velocity_limit[i] * std::min(std::max(-kp[i] * (q[i] - soft_lower[i]), -1.0), 1.0);
        const double safe_velocity_upper = velocity_limit[i] * std::min(std::max(-kp[i] * (q[i] - soft_upper[i]), -1.0), 1.0);
        const double safe_effort_lower = effort_limit[i] * std::min(std::max(-kd[i] * (v[i] - safe_velocity_lower), -1.0), 1.0);
        const double safe_effort_upper = effort_limit[i] * std::min(std::max(-kd[i] * (v[i] - safe_velocity_upper), -1.0), 1.0);
        out[i] = std::min(std::max(command[i], safe_effort_lower), safe_effort_upper);
    }
}

}  // namespace orc

extern "C" {
void orc_integrate_zoh(double* state, const double* lo, const double* hi, int n, double dt) { orc::integrate_zoh(state, lo, hi, n, dt); }
void orc_pd_controller(const double* enc, double* state, const double* lo, const double* hi, const double* kp, const double* kd,
                       const double* elim, int n, double dt, double* out) { orc::pd_controller(enc, state, lo, hi, kp, kd, elim, n, dt, out); }
void orc_apply_safety_limits(const double* cmd, const double* q, const double* v, const double* kp, const double* kd, const double* lo,
                             const double* hi, const double* vlim, const double* elim, int n, double* out) {
    orc::apply_safety_limits(cmd, q, v, kp, kd, lo, hi, vlim, elim, n, out);
}
}
