// The env-step kernel: per-env scheduler of Engine::step (core/src/engine/engine.cc:1724-2417)
// around the device steppers, plus Engine::start (engine.cc:952-1533) and a single-RHS mode.
#pragma once
#include "jb_device.cuh"

namespace jb {

// q / v component of a record <-> SoA global state
JB_DI void load_record_state(const Ctx& c, const KParams* P, const RecInt* ri, int base, const double* __restrict__ q,
                             const double* __restrict__ v, const double* __restrict__ a, size_t stride, size_t col) {
    if (ri->kind == REC_FREE) {
#pragma unroll
        for (int k = 0; k < 7; ++k) SMF(c, base + RF_Q + k) = q[(ri->idx_q + k) * stride + col];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            SMF(c, base + RF_V + k) = v[(ri->idx_v + k) * stride + col];
            SMF(c, base + RF_A + k) = a ? a[(ri->idx_v + k) * stride + col] : 0.0;
        }
    } else {
        SMF(c, base + R1_Q) = q[ri->idx_q * stride + col];
        SMF(c, base + R1_Q + 1) = (ri->kind == REC_REVU) ? q[(ri->idx_q + 1) * stride + col] : 0.0;
        SMF(c, base + R1_V) = v[ri->idx_v * stride + col];
        SMF(c, base + R1_A) = a ? a[ri->idx_v * stride + col] : 0.0;
    }
}

// Loads from env-major (AoS) arrays: element (env, k) at p[env * width + k]
JB_DI void load_record_state_aos(const Ctx& c, const KParams* P, const RecInt* ri, int base, const double* __restrict__ q,
                                 const double* __restrict__ v, int env) {
    const double* qe = q + static_cast<size_t>(env) * P->nq;
    const double* ve = v + static_cast<size_t>(env) * P->nv;
    if (ri->kind == REC_FREE) {
#pragma unroll
        for (int k = 0; k < 7; ++k) SMF(c, base + RF_Q + k) = qe[ri->idx_q + k];
#pragma unroll
        for (int k = 0; k < 6; ++k) { SMF(c, base + RF_V + k) = ve[ri->idx_v + k]; SMF(c, base + RF_A + k) = 0.0; }
    } else {
        SMF(c, base + R1_Q) = qe[ri->idx_q];
        SMF(c, base + R1_Q + 1) = (ri->kind == REC_REVU) ? qe[ri->idx_q + 1] : 0.0;
        SMF(c, base + R1_V) = ve[ri->idx_v];
        SMF(c, base + R1_A) = 0.0;
    }
}

// pinocchio::normalize on the record (Engine::start, engine.cc:1042-1043)
JB_DI void normalize_record(const Ctx& c, const RecInt* ri, int base) {
    if (ri->kind == REC_FREE) {
        double n2 = 0.0;
#pragma unroll
        for (int k = 3; k < 7; ++k) n2 += SMF(c, base + RF_Q + k) * SMF(c, base + RF_Q + k);
        const double n = sqrt(n2);
#pragma unroll
        for (int k = 3; k < 7; ++k) SMF(c, base + RF_Q + k) /= n;
    } else if (ri->kind == REC_REVU) {
        const double n = sqrt(SMF(c, base + R1_Q) * SMF(c, base + R1_Q) + SMF(c, base + R1_Q + 1) * SMF(c, base + R1_Q + 1));
        SMF(c, base + R1_Q) /= n; SMF(c, base + R1_Q + 1) /= n;
    }
}

// Device-side controller block: gym_jiminy.common.blocks.pd_controller
// (python/gym_jiminy/common/gym_jiminy/common/blocks/proportional_derivative_controller.py:101-165)
// for a zero-order-held position target and zero target velocity,
//     tau = clip(kp * ((q_des - q_enc) + kd * (0 - v_enc)), +-effort_limit),
// evaluated on the motor-side encoder data of the accepted state at every controller breakpoint.
// The action buffer (`command`) holds the targets; the torque goes to the CMD field of the record.
__device__ __noinline__ void update_pd_commands(const Ctx c, const KParams* P) {
    const int L = P->L;
    for (int r = 0; r < P->nrec; ++r) {
        const RecInt* ri = P->rint + (r * L + c.sub);
        if (ri->kind == REC_PAD || ri->kind == REC_FREE || ri->motor < 0) continue;
        const RecDbl* rd = P->rdbl + (r * L + c.sub);
        const int base = P->rec_off[r];
        const double red = rd->motor[0], lim = rd->motor[1];
        const double pos = (ri->kind == REC_REVU) ? atan2(SMF(c, base + R1_Q + 1), SMF(c, base + R1_Q)) : SMF(c, base + R1_Q);
        const double q_enc = pos * red, v_enc = SMF(c, base + R1_V) * red;
        const double target = P->command[static_cast<size_t>(c.env) * P->nmotors + ri->motor];
        const double tau = P->pd_gains[ri->motor] * ((target - q_enc) + P->pd_gains[P->nmotors + ri->motor] * (0.0 - v_enc));
        SMF(c, base + R1_CMD) = fmin(fmax(tau, -lim), lim);
    }
}

JB_DI bool period_hit(double t, double period) {
    // `dtNext < SIMULATION_MIN_TIMESTEP || period - dtNext < STEPPER_MIN_TIMESTEP` (engine.cc:1924-1927, :2388-2395)
    const double dtNext = period - fmod(t, period);
    return dtNext < SIMULATION_MIN_TIMESTEP || period - dtNext < STEPPER_MIN_TIMESTEP;
}

__device__ __noinline__ void store_outputs(const Ctx c, const KParams* P) {
    if (!c.valid) return;
    const int L = P->L;
    const size_t N = P->n_pad, col = c.env;
    for (int r = 0; r < P->nrec; ++r) {
        const RecInt* ri = P->rint + (r * L + c.sub);
        if (ri->kind == REC_PAD || !ri->owner) continue;
        const RecDbl* rd = P->rdbl + (r * L + c.sub);
        const int base = P->rec_off[r];
        double* qv = P->qv_out ? P->qv_out + col * (P->nq + P->nv) : nullptr;
        if (ri->kind == REC_FREE) {
#pragma unroll
            for (int k = 0; k < 7; ++k) { const double x = SMF(c, base + RF_Q + k); P->q[(ri->idx_q + k) * N + col] = x; if (qv) qv[ri->idx_q + k] = x; }
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const double x = SMF(c, base + RF_V + k);
                P->v[(ri->idx_v + k) * N + col] = x; if (qv) qv[P->nq + ri->idx_v + k] = x;
                P->a[(ri->idx_v + k) * N + col] = SMF(c, base + RF_A + k);
                if (P->eff_u) P->eff_u[col * P->nv + ri->idx_v + k] = 0.0;
            }
        } else {
            const double q0 = SMF(c, base + R1_Q), q1 = SMF(c, base + R1_Q + 1), vv = SMF(c, base + R1_V);
            P->q[ri->idx_q * N + col] = q0; if (qv) qv[ri->idx_q] = q0;
            if (ri->kind == REC_REVU) { P->q[(ri->idx_q + 1) * N + col] = q1; if (qv) qv[ri->idx_q + 1] = q1; }
            P->v[ri->idx_v * N + col] = vv; if (qv) qv[P->nq + ri->idx_v] = vv;
            P->a[ri->idx_v * N + col] = SMF(c, base + R1_A);
            // RobotState.u = uInternal + uCustom + uTransmission (engine.cc:3694-3702), rebuilt from the
            // accepted state because the backward sweep reuses the U field for `data.u`
            if (P->eff_u) {
                double u = 0.0;
                if (P->springs != nullptr && ri->kind != REC_REVU) u = -P->springs[ri->idx_v] * q0 - P->springs[P->nv + ri->idx_v] * vv;
                if (ri->motor >= 0) {
                    double uM, uT;
                    motor_effort(rd, ri->motor_flags, SMF(c, base + R1_CMD), vv, uM, uT);
                    u += uT;
                }
                P->eff_u[col * P->nv + ri->idx_v] = u;
            }
            if (P->eff_umotor && ri->motor >= 0) P->eff_umotor[col * P->nmotors + ri->motor] = SMF(c, base + R1_UMOTOR);
        }
        if (P->eff_fext) {
            Mot fext = mzero();
            for (int k = 0; k < ri->ncontact; ++k) {
                const int cs = ri->contact0 + k;
                const ContactSlot* ct = P->cslots + (cs * L + c.sub);
                const int co = P->cslot_off + CSLOT_SIZE * cs;
                const V3 Fl = mk(SMF(c, co), SMF(c, co + 1), SMF(c, co + 2));
                fext.l = fext.l + Fl; fext.a = fext.a + cross(ld3(ct->placement + 9), Fl);
            }
            double* o = P->eff_fext + (col * P->njoints + ri->joint) * 6;
            o[0] = fext.l.x; o[1] = fext.l.y; o[2] = fext.l.z; o[3] = fext.a.x; o[4] = fext.a.y; o[5] = fext.a.z;
        }
    }
}

// MODE_DYNAMICS outputs: a, fext, u (all AoS)
__device__ __noinline__ void store_dynamics(const Ctx c, const KParams* P) {
    if (!c.valid) return;
    const int L = P->L;
    const size_t col = c.env;
    for (int r = 0; r < P->nrec; ++r) {
        const RecInt* ri = P->rint + (r * L + c.sub);
        if (ri->kind == REC_PAD || !ri->owner) continue;
        const RecDbl* rd = P->rdbl + (r * L + c.sub);
        const int base = P->rec_off[r];
        if (ri->kind == REC_FREE) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                P->a_out[col * P->nv + ri->idx_v + k] = SMF(c, base + RF_A + k);
                if (P->u_out) P->u_out[col * P->nv + ri->idx_v + k] = 0.0;
            }
        } else {
            P->a_out[col * P->nv + ri->idx_v] = SMF(c, base + R1_A);
            if (P->u_out) {
                double u = 0.0;
                const double q0 = SMF(c, base + R1_QS), vv = SMF(c, base + R1_VS);
                if (P->springs != nullptr && ri->kind != REC_REVU) u = -P->springs[ri->idx_v] * q0 - P->springs[P->nv + ri->idx_v] * vv;
                if (ri->motor >= 0) { double uM, uT; motor_effort(rd, ri->motor_flags, SMF(c, base + R1_CMD), vv, uM, uT); u += uT; }
                P->u_out[col * P->nv + ri->idx_v] = u;
            }
        }
        if (P->fext_out) {
            Mot fext = mzero();
            for (int k = 0; k < ri->ncontact; ++k) {
                const int cs = ri->contact0 + k;
                const ContactSlot* ct = P->cslots + (cs * L + c.sub);
                const int co = P->cslot_off + CSLOT_SIZE * cs;
                const V3 Fl = mk(SMF(c, co), SMF(c, co + 1), SMF(c, co + 2));
                fext.l = fext.l + Fl; fext.a = fext.a + cross(ld3(ct->placement + 9), Fl);
            }
            double* o = P->fext_out + (col * P->njoints + ri->joint) * 6;
            o[0] = fext.l.x; o[1] = fext.l.y; o[2] = fext.l.z; o[3] = fext.a.x; o[4] = fext.a.y; o[5] = fext.a.z;
        }
    }
}

__global__ void __launch_bounds__(32) env_step_kernel(const __grid_constant__ KParams Pk) {
    const KParams* P = &Pk;
#ifdef JB_HOST_EMUL
    double* smem = emul_smem;
#else
    extern __shared__ double smem[];
#endif
    Ctx c;
    c.lane = threadIdx.x & 31;
    c.sm = smem + c.lane;
    const int L = P->L;
    c.sub = c.lane % L;
    const int epw = 32 / L;
    const int env_raw = blockIdx.x * epw + c.lane / L;
    c.valid = env_raw < P->n_env;
    c.env = c.valid ? env_raw : (P->n_env - 1);
    c.gmask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << (c.lane - c.sub));
    const size_t N = P->n_pad, col = c.env;
    const int mode = P->mode;
    int status = P->status[c.env];

    const bool masked_out = (mode == MODE_START) && P->mask != nullptr && P->mask[c.env] == 0;
    if (masked_out) return;   // whole env (all its lanes) leaves: group masks keep the others safe
    if (mode == MODE_STEP && (status & (JB_ENV_NOT_STARTED | JB_ENV_NAN | JB_ENV_ITER_FAILED | JB_ENV_DT_UNDERFLOW))) return;

    // ---------------- load state into the lane records
    for (int r = 0; r < P->nrec; ++r) {
        const RecInt* ri = P->rint + (r * L + c.sub);
        if (ri->kind == REC_PAD) continue;
        const int base = P->rec_off[r];
        if (mode == MODE_STEP) load_record_state(c, P, ri, base, P->q, P->v, P->a, N, col);
        else {
            load_record_state_aos(c, P, ri, base, P->q_in, P->v_in, c.env);
            if (mode == MODE_START) normalize_record(c, ri, base);
        }
        if (ri->kind != REC_FREE) {
            // torque command: the action itself, or (PD mode) the torque held since the last breakpoint
            const double* cmd_src = (P->pd_gains != nullptr && mode == MODE_STEP) ? P->cmd_torque : P->command;
            SMF(c, base + R1_CMD) = (ri->motor >= 0) ? cmd_src[col * P->nmotors + ri->motor] : 0.0;
            SMF(c, base + R1_UMOTOR) = 0.0;
        }
    }
    for (int k = 0; k < CSLOT_SIZE * P->ncslot; ++k) SMF(c, P->cslot_off + k) = 0.0;
    for (int k = 0; k < IMUSLOT_SIZE * P->nimuslot; ++k) SMF(c, P->imu_off + k) = 0.0;

    if (mode == MODE_DYNAMICS) {
        stage_from_accepted(c, P);
        int st = 0;
        rhs(c, P, false, &st);
        store_dynamics(c, P);
        return;
    }

    double t, dt, dtLargest, dtLargestPrev, tError, tPrev;
    long long iter, iterFailed;
    if (mode == MODE_START) {
        // Engine::start (engine.cc:952-1533): stepperState_.reset(SIMULATION_MIN_TIMESTEP, ...), forward
        // kinematics, initial contact-force guard, then the INIT_ITERATIONS fixed point, which for a
        // zero-order-held command converges to one evaluation of the dynamics.
        status = JB_ENV_OK;
        t = 0.0; tPrev = 0.0; tError = 0.0;
        dt = SIMULATION_MIN_TIMESTEP; dtLargest = dt; dtLargestPrev = dt;
        iter = 0; iterFailed = 0;
        if (P->pd_gains != nullptr) update_pd_commands(c, P);
        stage_from_accepted(c, P);
        rhs(c, P, false, &status);
        // forceMax > 1e5 guard (engine.cc:1310-1346)
        double fmax2 = 0.0;
        for (int k = 0; k < P->ncslot; ++k) {
            const int co = P->cslot_off + CSLOT_SIZE * k;
            const double fx = SMF(c, co), fy = SMF(c, co + 1), fz = SMF(c, co + 2);
            fmax2 = fmax(fmax2, fx * fx + fy * fy + fz * fz);
        }
        for (int o = 1; o < L; o <<= 1) fmax2 = fmax(fmax2, __shfl_xor_sync(c.gmask, fmax2, o));
        if (fmax2 > 1e10) status |= JB_ENV_CONTACT_FORCE | JB_ENV_NOT_STARTED;
        bool bad = accel_has_nan(c, P);
        bad = __any_sync(c.gmask, bad);
        if (bad) status |= JB_ENV_NAN;
        write_sensors(c, P);
    } else {
        t = P->sched[SCH_T * N + col]; dt = P->sched[SCH_DT * N + col];
        dtLargest = P->sched[SCH_DTLARGEST * N + col]; dtLargestPrev = P->sched[SCH_DTLARGESTPREV * N + col];
        tError = P->sched[SCH_TERROR * N + col]; tPrev = P->sched[SCH_TPREV * N + col];
        iter = P->iters[col]; iterFailed = P->iters[N + col];

        // ------------- Engine::step (engine.cc:1724-2417)
        const JbOptions& opt = P->opt;
        double stepSize = P->step_dt;
        if (stepSize < D_EPS) {
            if (opt.controller_update_period > D_EPS) stepSize = opt.controller_update_period;
            else if (opt.sensors_update_period > D_EPS) stepSize = opt.sensors_update_period;
            else stepSize = opt.dt_max;
        }
        // Kahan-compensated end time (engine.cc:1793-1795)
        const double stepSizeCorrected = stepSize - tError;
        const double tEnd = t + stepSizeCorrected;
        tError = (tEnd - t) - stepSizeCorrected;
        const double supd = P->stepper_update_period;
        const bool finitePeriod = supd < 1e300;
        bool hasDynamicsChanged = false;
        bool failed = false;
        // The cached contact forces of the last evaluation are a pure function of the accepted
        // state: rebuild them (and the IMU captures) once, so that `up_to_date` evaluations and
        // sensor refreshes see what the reference keeps in RobotData between calls.
        // Shared memory does not persist between launches: the first evaluation of this launch is a
        // full one (it rebuilds the cached contact forces, a pure function of the accepted state).
        bool need_refresh = true;

        while (tEnd - t >= STEPPER_MIN_TIMESTEP && !failed) {
            double tNext = t;
            if (finitePeriod && opt.controller_update_period > D_EPS) {
                if (period_hit(t, opt.controller_update_period)) {
                    // computeCommand (engine.cc:1920-1940): zero-order hold of the action, or the PD block
                    if (P->pd_gains != nullptr) update_pd_commands(c, P);
                    hasDynamicsChanged = true;
                }
            }
            if (!finitePeriod && hasDynamicsChanged) {
                stage_from_accepted(c, P);
                rhs(c, P, !need_refresh, &status);
                need_refresh = false;
                hasDynamicsChanged = false;
            }
            if (finitePeriod) {
                double dtNextGlobal;
                const double dtNextUpdatePeriod = supd - fmod(t, supd);
                if (dtNextUpdatePeriod < SIMULATION_MIN_TIMESTEP) dtNextGlobal = dtNextUpdatePeriod + supd;
                else dtNextGlobal = dtNextUpdatePeriod;
                if (tEnd - t - STEPPER_MIN_TIMESTEP < dtNextGlobal) dtNextGlobal = tEnd - t;
                tNext += dtNextGlobal;
                while (tNext - t > STEPPER_MIN_TIMESTEP) {
                    if (hasDynamicsChanged) {
                        // FSAL repair: same state, cached contact forces, new command (engine.cc:2032-2037)
                        stage_from_accepted(c, P);
                        rhs(c, P, !need_refresh, &status);
                        need_refresh = false;
                        hasDynamicsChanged = false;
                    }
                    if (dt < STEPPER_MIN_TIMESTEP) break;
                    // successiveIterTooLarge == 0 always holds for the fixed-step steppers
                    const double dtResidualThr = fmin(fmax(0.1 * dt, STEPPER_MIN_TIMESTEP), SIMULATION_MIN_TIMESTEP);
                    if (tNext - t < dt || tNext - t < dt + dtResidualThr) dt = tNext - t;
                    if (dt > SIMULATION_MIN_TIMESTEP) {
                        const double dtResidual = fmod(dt, SIMULATION_MIN_TIMESTEP);
                        if (dtResidual > STEPPER_MIN_TIMESTEP && dtResidual < SIMULATION_MIN_TIMESTEP - STEPPER_MIN_TIMESTEP &&
                            dt - dtResidual > STEPPER_MIN_TIMESTEP)
                            dt -= dtResidual;
                    }
                    const bool isBreakpointReached = (dtLargest > dt);
                    dtLargest = dt;
                    // stepper_->tryStep (abstract_stepper.cc:16-62)
                    const double t_next = t + dtLargest;
                    if (opt.ode_solver == JB_SOLVER_EULER_EXPLICIT) step_euler(c, P, dtLargest, &status);
                    else step_rk4(c, P, dtLargest, &status);
                    need_refresh = false;
                    dtLargest = D_INF;
                    bool bad = accel_has_nan(c, P);
                    bad = __any_sync(c.gmask, bad);
                    if (bad) { status |= JB_ENV_NAN; failed = true; ++iterFailed; break; }
                    t = t_next;
                    ++iter;
                    if (isBreakpointReached) {
                        const double thr = dtLargestPrev * opt.dt_restore_threshold_rel;
                        if (dt < dtLargest && dtLargest < thr) dtLargest = dtLargestPrev;
                    }
                    tPrev = t;
                    dtLargestPrev = dtLargest;
                    dt = fmin(dtLargest, opt.dt_max);
                }
            } else {
                dt = fmin(dt, tEnd - t);
                const bool isBreakpointReached = (dtLargest > dt);
                dtLargest = dt;
                const double t_next = t + dtLargest;
                if (opt.ode_solver == JB_SOLVER_EULER_EXPLICIT) step_euler(c, P, dtLargest, &status);
                else step_rk4(c, P, dtLargest, &status);
                need_refresh = false;
                dtLargest = D_INF;
                bool bad = accel_has_nan(c, P);
                bad = __any_sync(c.gmask, bad);
                if (bad) { status |= JB_ENV_NAN; failed = true; ++iterFailed; break; }
                t = t_next;
                ++iter;
                if (isBreakpointReached) {
                    const double thr = dtLargestPrev * opt.dt_restore_threshold_rel;
                    if (dt < dtLargest && dtLargest < thr) dtLargest = dtLargestPrev;
                }
                tPrev = t;
                dtLargestPrev = dtLargest;
                dt = fmin(dtLargest, opt.dt_max);
            }
            if (dt < STEPPER_MIN_TIMESTEP) { status |= JB_ENV_DT_UNDERFLOW; failed = true; break; }
            // sensors refresh (engine.cc:2386-2410)
            const double sp = opt.sensors_update_period;
            bool mustUpdateSensors = sp < D_EPS;
            if (!mustUpdateSensors) mustUpdateSensors = period_hit(t, sp);
            if (mustUpdateSensors && !failed) write_sensors(c, P);
        }
        if (!failed) t = tEnd;
    }

    // ---------------- store
    store_outputs(c, P);
    if (P->pd_gains != nullptr && c.valid) {
        for (int r = 0; r < P->nrec; ++r) {
            const RecInt* ri = P->rint + (r * L + c.sub);
            if (ri->kind == REC_PAD || ri->kind == REC_FREE || ri->motor < 0 || !ri->owner) continue;
            P->cmd_torque[col * P->nmotors + ri->motor] = SMF(c, P->rec_off[r] + R1_CMD);
        }
    }
    if (c.valid && c.sub == 0) {
        P->sched[SCH_T * N + col] = t; P->sched[SCH_DT * N + col] = dt;
        P->sched[SCH_DTLARGEST * N + col] = dtLargest; P->sched[SCH_DTLARGESTPREV * N + col] = dtLargestPrev;
        P->sched[SCH_TERROR * N + col] = tError; P->sched[SCH_TPREV * N + col] = tPrev;
        P->iters[col] = iter; P->iters[N + col] = iterFailed;
    }
    // status bits can be raised by any lane of the env
    for (int o = 1; o < L; o <<= 1) status |= __shfl_xor_sync(c.gmask, status, o);
    if (c.valid && c.sub == 0) P->status[c.env] = status;
}

}  // namespace jb
