// TEST INFRASTRUCTURE -- CPU oracle, not product code.  See oracle/README.md.
//
// C entry points over orc::Engine for ctypes (tests/, __graft_entry__.smoke, bench.py cpu_baseline).
// A "batch" is a plain array of independent single-robot engines, stepped sequentially or with an
// OpenMP `parallel for` over envs (the reference engine itself has no threading; N reference
// processes on N cores is what this mirrors).
#include <cstring>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "engine.hpp"

using orc::Engine;

struct OrcBatch {
    std::vector<std::unique_ptr<Engine>> envs;
    int nq, nv, nmotors, njoints, width;
};

extern "C" {

void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

OrcBatch* orc_create(const JbModelDesc* model, const JbOptions* opt, int n_env) {
    auto* b = new OrcBatch;
    for (int i = 0; i < n_env; ++i) b->envs.emplace_back(new Engine(*model, *opt));
    b->nq = model->nq; b->nv = model->nv; b->nmotors = model->nmotors; b->njoints = model->njoints;
    b->width = b->envs[0]->model.layout.width;
    return b;
}
void orc_destroy(OrcBatch* b) { delete b; }
int orc_sensor_width(OrcBatch* b) { return b->width; }
void orc_sensor_layout(OrcBatch* b, JbSensorLayout* out) { *out = b->envs[0]->model.layout; }

void orc_set_options(OrcBatch* b, const JbOptions* opt) {
    for (auto& e : b->envs) e->set_options(*opt);
}
void orc_set_callbacks(OrcBatch* b, int env, orc::ControllerFn c, orc::InternalDynFn d, void* ctx) {
    b->envs[env]->controller = c; b->envs[env]->internalDyn = d; b->envs[env]->ctx = ctx;
}
void orc_set_springs(OrcBatch* b, const double* k, const double* d) {
    for (auto& e : b->envs) { e->spring_k.assign(k, k + b->nv); e->spring_d.assign(d, d + b->nv); }
}

// ---- external forces (Engine::register_impulse_force / register_profile_force / remove_all_forces / stop)
void orc_stop(OrcBatch* b) { for (auto& e : b->envs) e->stop(); }
int orc_register_impulse_force(OrcBatch* b, int joint, const double* p, const double* t, const double* dt, const double* F) {
    int rc = 0;
    for (size_t i = 0; i < b->envs.size(); ++i) {
        const int r = b->envs[i]->registerImpulseForce(joint, p, t[i], dt[i], F + 6 * i);
        if (r < 0) rc = r;
    }
    return rc < 0 ? rc : static_cast<int>(b->envs[0]->impulseForces.size()) - 1;
}
// rewrite impulse `k` of the envs selected by `mask` (what re-registering at an episode reset does)
int orc_set_impulse_force(OrcBatch* b, int k, const uint8_t* mask, const double* t, const double* dt, const double* F) {
    for (size_t i = 0; i < b->envs.size(); ++i) {
        if (mask && !mask[i]) continue;
        Engine& e = *b->envs[i];
        if (k < 0 || k >= static_cast<int>(e.impulseForces.size())) return JB_ERR_INVALID_ARGUMENT;
        if (dt[i] < orc::STEPPER_MIN_TIMESTEP || t[i] < 0.0) return JB_ERR_INVALID_ARGUMENT;
        auto& f = e.impulseForces[k];
        f.t = t[i]; f.dt = dt[i];
        f.F = orc::Force{orc::V3(F[6 * i], F[6 * i + 1], F[6 * i + 2]), orc::V3(F[6 * i + 3], F[6 * i + 4], F[6 * i + 5])};
        e.impulseForceBreakpoints.clear();
        for (const auto& g : e.impulseForces) { e.impulseForceBreakpoints.push_back(g.t); e.impulseForceBreakpoints.push_back(g.t + g.dt); }
        std::sort(e.impulseForceBreakpoints.begin(), e.impulseForceBreakpoints.end());
        e.impulseForceBreakpoints.erase(std::unique(e.impulseForceBreakpoints.begin(), e.impulseForceBreakpoints.end()),
                                        e.impulseForceBreakpoints.end());
    }
    return JB_OK;
}
int orc_register_profile_force(OrcBatch* b, int joint, const double* p, double update_period) {
    int rc = 0;
    for (auto& e : b->envs) rc = e->registerProfileForce(joint, p, update_period);
    return rc;
}
int orc_set_profile_force(OrcBatch* b, int slot, const double* F) {
    for (size_t i = 0; i < b->envs.size(); ++i) {
        Engine& e = *b->envs[i];
        if (slot < 0 || slot >= static_cast<int>(e.profileForces.size())) return JB_ERR_INVALID_ARGUMENT;
        e.profileForces[slot].pending =
            orc::Force{orc::V3(F[6 * i], F[6 * i + 1], F[6 * i + 2]), orc::V3(F[6 * i + 3], F[6 * i + 4], F[6 * i + 5])};
    }
    return JB_OK;
}
void orc_remove_all_forces(OrcBatch* b) { for (auto& e : b->envs) e->removeAllForces(); }

// returns the number of envs whose start failed; rc[i] receives the per-env return code
int orc_start(OrcBatch* b, const uint8_t* mask, const double* q0, const double* v0, int* rc) {
    int bad = 0;
    for (size_t i = 0; i < b->envs.size(); ++i) {
        if (mask && !mask[i]) { if (rc) rc[i] = 0; continue; }
        int r = b->envs[i]->start(q0 + i * b->nq, v0 + i * b->nv);
        if (rc) rc[i] = r;
        bad += (r != 0);
    }
    return bad;
}
void orc_set_command(OrcBatch* b, const double* cmd) {
    for (size_t i = 0; i < b->envs.size(); ++i) {
        Engine& e = *b->envs[i];
        if (e.pdf_enabled) e.pdf_action.assign(cmd + i * b->nmotors, cmd + (i + 1) * b->nmotors);
        else if (e.pd_enabled) e.pd_target.assign(cmd + i * b->nmotors, cmd + (i + 1) * b->nmotors);
        else std::memcpy(e.state.command.data(), cmd + i * b->nmotors, sizeof(double) * b->nmotors);
    }
}
void orc_set_pd(OrcBatch* b, const double* kp, const double* kd) {
    for (auto& e : b->envs) {
        e->pd_enabled = kp != nullptr;
        if (kp) { e->pd_kp.assign(kp, kp + b->nmotors); e->pd_kd.assign(kd, kd + b->nmotors); e->pd_target.assign(b->nmotors, 0.0); }
    }
}
// PDController block with optional MotorSafetyLimit: lower / upper [3][nmotors], safety [5][nmotors] or null
void orc_set_pd_full(OrcBatch* b, const double* kp, const double* kd, const double* lower, const double* upper, const double* safety) {
    const int nm = b->nmotors;
    for (auto& e : b->envs) {
        e->pdf_enabled = kp != nullptr;
        if (!kp) continue;
        e->pd_enabled = false;
        e->pdf_kp.assign(kp, kp + nm); e->pdf_kd.assign(kd, kd + nm);
        e->pdf_lower.assign(lower, lower + 3 * nm); e->pdf_upper.assign(upper, upper + 3 * nm);
        e->pdf_state.assign(3 * nm, 0.0); e->pdf_action.assign(nm, 0.0);
        e->pdf_safety = safety != nullptr;
        if (safety) {
            e->pdf_skp.assign(safety, safety + nm); e->pdf_skd.assign(safety + nm, safety + 2 * nm);
            e->pdf_slo.assign(safety + 2 * nm, safety + 3 * nm); e->pdf_shi.assign(safety + 3 * nm, safety + 4 * nm);
            e->pdf_svlim.assign(safety + 4 * nm, safety + 5 * nm);
        }
    }
}
// the block's command state, [n_env][3][nmotors] (what the PDAdapter block reads / writes)
void orc_get_pd_state(OrcBatch* b, double* out) {
    const size_t n = 3 * static_cast<size_t>(b->nmotors);
    for (size_t i = 0; i < b->envs.size(); ++i) std::copy(b->envs[i]->pdf_state.begin(), b->envs[i]->pdf_state.begin() + n, out + i * n);
}
void orc_set_pd_state(OrcBatch* b, const double* in) {
    const size_t n = 3 * static_cast<size_t>(b->nmotors);
    for (size_t i = 0; i < b->envs.size(); ++i) b->envs[i]->pdf_state.assign(in + i * n, in + (i + 1) * n);
}
// is_enabled / lambda of the bound constraints (by joint) and of the contact constraints (by contact frame)
void orc_get_constraints(OrcBatch* b, uint8_t* joint_enabled, double* joint_lambda, uint8_t* contact_enabled, double* contact_lambda) {
    for (size_t e = 0; e < b->envs.size(); ++e) {
        const auto& E = *b->envs[e];
        const int nj = E.model.njoints, ncc = E.model.ncontacts;
        for (int j = 0; j < nj; ++j) { if (joint_enabled) joint_enabled[e * nj + j] = 0; if (joint_lambda) joint_lambda[e * nj + j] = 0.0; }
        for (const auto& c : E.constraints) {
            if (c.kind == 0) {
                if (joint_enabled) joint_enabled[e * nj + c.joint] = c.enabled;
                if (joint_lambda) joint_lambda[e * nj + c.joint] = c.lambda[0];
            } else {
                if (contact_enabled) contact_enabled[e * ncc + c.contact] = c.enabled;
                if (contact_lambda) for (int k = 0; k < 4; ++k) contact_lambda[(e * ncc + c.contact) * 4 + k] = c.lambda[k];
            }
        }
    }
}
void orc_set_mahony(OrcBatch* b, double kp, double ki) {
    for (auto& e : b->envs) { e->mahony_enabled = kp >= 0.0; e->mahony_kp = kp; e->mahony_ki = ki; }
}
// out: [n_env][nimu][10] = quaternion (x, y, z, w), gyro bias estimate (3), unbiased angular velocity (3)
void orc_get_mahony(OrcBatch* b, double* out) {
    for (size_t i = 0; i < b->envs.size(); ++i) {
        const Engine& e = *b->envs[i];
        const int M = e.model.nimu;
        for (int k = 0; k < M; ++k) {
            double* o = out + (i * M + k) * 10;
            for (int c = 0; c < 4; ++c) o[c] = e.mahony_q.empty() ? 0.0 : e.mahony_q[c * M + k];
            for (int c = 0; c < 3; ++c) { o[4 + c] = e.mahony_bias.empty() ? 0.0 : e.mahony_bias[c * M + k]; o[7 + c] = e.mahony_omega.empty() ? 0.0 : e.mahony_omega[c * M + k]; }
        }
    }
}
int orc_step(OrcBatch* b, double step_dt, int parallel, int* rc) {
    const int n = static_cast<int>(b->envs.size());
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad) if (parallel)
    for (int i = 0; i < n; ++i) {
        int r = b->envs[i]->step(step_dt);
        if (rc) rc[i] = r;
        bad += (r != 0);
    }
    return bad;
}
void orc_get_state(OrcBatch* b, double* t, double* q, double* v, double* a) {
    for (size_t i = 0; i < b->envs.size(); ++i) {
        Engine& e = *b->envs[i];
        if (t) t[i] = e.t;
        if (q) std::memcpy(q + i * b->nq, e.q.data(), sizeof(double) * b->nq);
        if (v) std::memcpy(v + i * b->nv, e.v.data(), sizeof(double) * b->nv);
        if (a) std::memcpy(a + i * b->nv, e.a.data(), sizeof(double) * b->nv);
    }
}
// StepperState fields besides (t, q, v, a) and the command held since the last controller update
void orc_get_stepper_state(OrcBatch* b, double* sched, double* command_held) {
    for (size_t i = 0; i < b->envs.size(); ++i) {
        Engine& e = *b->envs[i];
        if (sched) { double* s = sched + 6 * i; s[0] = e.t; s[1] = e.dt; s[2] = e.dtLargest; s[3] = e.dtLargestPrev; s[4] = e.tError; s[5] = e.tPrev; }
        if (command_held && b->nmotors) std::memcpy(command_held + i * b->nmotors, e.state.command.data(), sizeof(double) * b->nmotors);
    }
}
void orc_get_efforts(OrcBatch* b, double* u, double* u_motor, double* command, double* fext) {
    for (size_t i = 0; i < b->envs.size(); ++i) {
        Engine& e = *b->envs[i];
        if (u) std::memcpy(u + i * b->nv, e.state.u.data(), sizeof(double) * b->nv);
        if (u_motor) std::memcpy(u_motor + i * b->nmotors, e.state.uMotor.data(), sizeof(double) * b->nmotors);
        if (command) std::memcpy(command + i * b->nmotors, e.state.command.data(), sizeof(double) * b->nmotors);
        if (fext)
            for (int j = 0; j < b->njoints; ++j) orc::to6(e.state.fExternal[j], fext + (i * b->njoints + j) * 6);
    }
}
void orc_get_sensors(OrcBatch* b, double* out) {
    for (size_t i = 0; i < b->envs.size(); ++i)
        std::memcpy(out + i * b->width, b->envs[i]->sensorOutput().data(), sizeof(double) * b->width);
}
// true values behind the measurements (AbstractSensorTpl::data())
void orc_get_sensor_data(OrcBatch* b, double* out) {
    for (size_t i = 0; i < b->envs.size(); ++i)
        std::memcpy(out + i * b->width, b->envs[i]->sensors.data(), sizeof(double) * b->width);
}
void orc_set_sensor_options(OrcBatch* b, int type, int index, const double* noise_std, const double* bias, double delay, double jitter, int order) {
    for (auto& e : b->envs) e->setSensorOptions(type, index, noise_std, bias, delay, jitter, static_cast<uint32_t>(order));
}
void orc_set_seeds(OrcBatch* b, const uint32_t* seeds) {
    for (size_t i = 0; i < b->envs.size(); ++i) b->envs[i]->engineSeed = seeds[i];
}
// raw generator access for the distribution tests: n draws of normal(0, 1) / uniform01 / raw 32-bit words from PCG32(seed_seq{seed})
void orc_random_draws(uint32_t seed, int kind, int n, double* out) {
    std::seed_seq seq{seed};
    orc::PCG32 g = orc::pcg32_from_seed_seq(seq);
    for (int i = 0; i < n; ++i) out[i] = kind == 0 ? static_cast<double>(orc::normal(g)) : (kind == 1 ? static_cast<double>(orc::uniform01(g)) : static_cast<double>(g()));
}
void orc_get_extra_terms(OrcBatch* b, double* energy, double* joint_a, double* joint_f) {
    for (size_t i = 0; i < b->envs.size(); ++i) {
        Engine& e = *b->envs[i];
        if (energy) { energy[2 * i] = e.data.kinetic_energy; energy[2 * i + 1] = e.data.potential_energy; }
        for (int j = 0; j < b->njoints; ++j) {
            if (joint_a) orc::to6(e.data.a[j], joint_a + (i * b->njoints + j) * 6);
            if (joint_f) orc::to6(e.data.f[j], joint_f + (i * b->njoints + j) * 6);
        }
    }
}
// ycrb [n][njoints][10] (mass, lever, inertia about the subtree CoM), com / vcom [n][njoints][3], hg / dhg [n][6]
void orc_get_centroidal(OrcBatch* b, double* ycrb, double* com, double* vcom, double* hg, double* dhg) {
    const int nj = b->njoints;
    for (size_t i = 0; i < b->envs.size(); ++i) {
        const Engine& e = *b->envs[i];
        for (int j = 0; j < nj; ++j) {
            if (ycrb) {
                double* o = ycrb + (i * nj + j) * 10;
                const orc::Inertia& Y = e.data.Ycrb[j];
                o[0] = Y.mass; o[1] = Y.c.x; o[2] = Y.c.y; o[3] = Y.c.z;
                for (int k = 0; k < 6; ++k) o[4 + k] = Y.I[k];
            }
            if (com) { double* o = com + (i * nj + j) * 3; o[0] = e.data.com[j].x; o[1] = e.data.com[j].y; o[2] = e.data.com[j].z; }
            if (vcom) { double* o = vcom + (i * nj + j) * 3; o[0] = e.data.vcom[j].x; o[1] = e.data.vcom[j].y; o[2] = e.data.vcom[j].z; }
        }
        if (hg) { double* o = hg + 6 * i; const orc::Force& f = e.data.hg; o[0] = f.lin.x; o[1] = f.lin.y; o[2] = f.lin.z; o[3] = f.ang.x; o[4] = f.ang.y; o[5] = f.ang.z; }
        if (dhg) { double* o = dhg + 6 * i; const orc::Force& f = e.data.dhg; o[0] = f.lin.x; o[1] = f.lin.y; o[2] = f.lin.z; o[3] = f.ang.x; o[4] = f.ang.y; o[5] = f.ang.z; }
    }
}
void orc_get_status(OrcBatch* b, int32_t* status) {
    for (size_t i = 0; i < b->envs.size(); ++i) status[i] = b->envs[i]->status;
}
void orc_get_iters(OrcBatch* b, int64_t* iter, int64_t* iter_failed) {
    for (size_t i = 0; i < b->envs.size(); ++i) {
        if (iter) iter[i] = b->envs[i]->iter;
        if (iter_failed) iter_failed[i] = b->envs[i]->iterFailed;
    }
}
int64_t orc_pgs_iterations(OrcBatch* b) {
    int64_t n = 0;
    for (auto& e : b->envs) n += e->pgsIterations;
    return n;
}
// diagnostics: iteration count of every PGS solve of one env since the history was switched on
void orc_pgs_history_enable(OrcBatch* b, int32_t on) {
    for (auto& e : b->envs) { e->keepPgsHistory = on != 0; e->pgsHistory.clear(); }
}
int64_t orc_pgs_history(OrcBatch* b, int32_t env, int32_t* out, int64_t cap) {
    const auto& h = b->envs[env]->pgsHistory;
    const int64_t n = std::min<int64_t>(cap, static_cast<int64_t>(h.size()));
    for (int64_t i = 0; i < n; ++i) out[i] = h[i];
    return static_cast<int64_t>(h.size());
}
int64_t orc_rhs_count(OrcBatch* b) {
    int64_t s = 0;
    for (auto& e : b->envs) s += e->rhs_count;
    return s;
}

// One evaluation of Engine::computeRobotsDynamics on an arbitrary (q, v, command), per env, on a
// scratch engine state: FK -> contacts -> motors -> ABA.  Used for per-RHS parity.
void orc_compute_dynamics(OrcBatch* b, const double* q, const double* v, const double* cmd, double* a,
                          double* fext, double* u) {
    for (size_t i = 0; i < b->envs.size(); ++i) {
        Engine& e = *b->envs[i];
        const bool was = e.running;
        e.running = true;
        std::memcpy(e.state.command.data(), cmd + i * b->nmotors, sizeof(double) * b->nmotors);
        std::vector<double> zeros(b->nv, 0.0), aout(b->nv);
        e.statePrev.a = zeros;
        e.computeRobotsDynamics(0.0, q + i * b->nq, v + i * b->nv, aout, false);
        std::memcpy(a + i * b->nv, aout.data(), sizeof(double) * b->nv);
        if (u) std::memcpy(u + i * b->nv, e.state.u.data(), sizeof(double) * b->nv);
        if (fext)
            for (int j = 0; j < b->njoints; ++j) orc::to6(e.state.fExternal[j], fext + (i * b->njoints + j) * 6);
        e.running = was;
    }
}

// Lie-group helpers exposed for unit tests (pinocchio::integrate / difference)
void orc_integrate(OrcBatch* b, const double* q, const double* v, double* out) { b->envs[0]->integrate(q, v, out); }
void orc_difference(OrcBatch* b, const double* q0, const double* q1, double* out) { b->envs[0]->difference(q0, q1, out); }

}  // extern "C"
