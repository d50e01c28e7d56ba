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
#include <vector>

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

// gym_jiminy `mahony_filter` (blocks/mahony_filter.py:28-101), M IMUs at once: q [4][M] (x, y, z, w), omega / gyro / acc /
// bias_hat [3][M].  The early return looks at all the IMUs together, like the vectorised reference.
void mahony_filter(double* q, double* omega, const double* gyro, const double* acc, double* bias_hat, int M, double kp, double ki,
                   double dt) {
    const double EARTH_SURFACE_GRAVITY = 9.81;
    std::vector<double> cf(3 * M), omega_mes(3 * M);
    bool still = true;
    for (int i = 0; i < M; ++i) {
        const double q_x = q[i], q_y = q[M + i], q_z = q[2 * M + i], q_w = q[3 * M + i];
        // compute_tilt_from_quat (utils/math.py:1046-1060): R(q)^T e_z
        const double v_x = 2 * (q_x * q_z - q_y * q_w), v_y = 2 * (q_y * q_z + q_w * q_x), v_z = 1 - 2 * (q_x * q_x + q_y * q_y);
        for (int k = 0; k < 3; ++k) omega[k * M + i] = gyro[k * M + i] - bias_hat[k * M + i];
        const double ax = acc[i] / EARTH_SURFACE_GRAVITY, ay = acc[M + i] / EARTH_SURFACE_GRAVITY, az = acc[2 * M + i] / EARTH_SURFACE_GRAVITY;
        omega_mes[i] = ay * v_z - az * v_y;
        omega_mes[M + i] = az * v_x - ax * v_z;
        omega_mes[2 * M + i] = ax * v_y - ay * v_x;
        for (int k = 0; k < 3; ++k) {
            cf[k * M + i] = omega[k * M + i] + kp * omega_mes[k * M + i];
            if (!(std::fabs(cf[k * M + i]) < 1e-6)) still = false;
        }
    }
    if (still) return;
    for (int i = 0; i < M; ++i) {
        double theta = std::sqrt(cf[i] * cf[i] + cf[M + i] * cf[M + i] + cf[2 * M + i] * cf[2 * M + i]);
        const double a_x = cf[i] / theta, a_y = cf[M + i] / theta, a_z = cf[2 * M + i] / theta;
        theta *= dt / 2;
        const double sn = std::sin(theta), p_w = std::cos(theta);
        const double p_x = a_x * sn, p_y = a_y * sn, p_z = a_z * sn;
        const double q_x = q[i], q_y = q[M + i], q_z = q[2 * M + i], q_w = q[3 * M + i];
        double n_x = q_x * p_w + q_w * p_x - q_z * p_y + q_y * p_z;
        double n_y = q_y * p_w + q_z * p_x + q_w * p_y - q_x * p_z;
        double n_z = q_z * p_w - q_y * p_x + q_x * p_y + q_w * p_z;
        double n_w = q_w * p_w - q_x * p_x - q_y * p_y - q_z * p_z;
        const double scale = (3.0 - (n_x * n_x + n_y * n_y + n_z * n_z + n_w * n_w)) / 2;   // first-order normalisation
        q[i] = n_x * scale; q[M + i] = n_y * scale; q[2 * M + i] = n_z * scale; q[3 * M + i] = n_w * scale;
        for (int k = 0; k < 3; ++k) bias_hat[k * M + i] -= ki * dt * omega_mes[k * M + i];
    }
}

// gym_jiminy `matrices_to_quat` (utils/math.py:307-350) for one row-major rotation matrix -> (x, y, z, w)
void matrix_to_quat_ref(const double* R, double* out) {
    double t;
    if (R[8] < 0) {
        if (R[0] > R[4]) { t = 1 + R[0] - R[4] - R[8]; out[0] = t; out[1] = R[3] + R[1]; out[2] = R[2] + R[6]; out[3] = R[7] - R[5]; }
        else { t = 1 - R[0] + R[4] - R[8]; out[0] = R[3] + R[1]; out[1] = t; out[2] = R[7] + R[5]; out[3] = R[2] - R[6]; }
    } else {
        if (R[0] < -R[4]) { t = 1 - R[0] - R[4] + R[8]; out[0] = R[2] + R[6]; out[1] = R[7] + R[5]; out[2] = t; out[3] = R[3] - R[1]; }
        else { t = 1 + R[0] + R[4] + R[8]; out[0] = R[7] - R[5]; out[1] = R[2] - R[6]; out[2] = R[3] - R[1]; out[3] = t; }
    }
    const double d = 2 * std::sqrt(t);
    for (int k = 0; k < 4; ++k) out[k] /= d;
}

}  // namespace orc

extern "C" {
void orc_mahony_filter(double* q, double* omega, const double* gyro, const double* acc, double* bias, int M, double kp, double ki, double dt) {
    orc::mahony_filter(q, omega, gyro, acc, bias, M, kp, ki, dt);
}
void orc_matrix_to_quat(const double* R, double* out) { orc::matrix_to_quat_ref(R, out); }
void orc_integrate_zoh(double* state, const double* lo, const double* hi, int n, double dt) { orc::integrate_zoh(state, lo, hi, n, dt); }
void orc_pd_controller(const double* enc, double* state, const double* lo, const double* hi, const double* kp, const double* kd,
                       const double* elim, int n, double dt, double* out) { orc::pd_controller(enc, state, lo, hi, kp, kd, elim, n, dt, out); }
void orc_apply_safety_limits(const double* cmd, const double* q, const double* v, const double* kp, const double* kd, const double* lo,
                             const double* hi, const double* vlim, const double* elim, int n, double* out) {
    orc::apply_safety_limits(cmd, q, v, kp, kd, lo, hi, vlim, elim, n, out);
}
}
