// jiminy_b200 -- the constraint path of the step: joint position bounds and contacts.model == "constraint".
//
// What Engine::computeAcceleration does once a kinematic constraint is enabled
// (core/src/engine/engine.cc:3709-3866): instead of plain ABA it solves the boxed forward dynamics
//     M ddq + nle = u + J^T lambda ,   J ddq + gamma = 0 on the active set, lambda in its box / cone
// with a projected Gauss-Seidel sweep (core/src/solver/constraint_solvers.cc:107-448) over the
// constraint Jacobians of JointConstraint (core/src/constraints/joint_constraint.cc:141-163) and
// FrameConstraint (core/src/constraints/frame_constraint.cc:103-183) with Baumgarte stabilisation.
//
// Device formulation.  The unconstrained acceleration M^-1 (u - nle + J_ext^T f_ext) is what the ABA sweeps
// of rhs_impl already produce, so they run first and this file only adds the correction:
//   1. sub-lane 0 of the env walks the whole tree in joint order (records of the other lanes are read
//      straight out of shared memory): world placements, velocities, drift accelerations, composite
//      inertias, the joint-space inertia M with rotor inertia (CRBA, pinocchio_overload_algorithms.h:99-124)
//      and its dense Cholesky factor;
//   2. the L lanes share the enabled constraints: rows of J, drift gamma, rows of L^-1 J^T, then rows of
//      A = J M^-1 J^T (+ regularisation), b = -gamma - J ddq_free, warm start;
//   3. sub-lane 0 runs the PGS sweep (or the exact equality solve of the first start iteration), then
//      ddq = ddq_free + M^-1 J^T lambda is written back into the records, the contact wrenches into the
//      contact slots, and every lane refreshes its spatial accelerations.
// Large, rarely used matrices live in a per-env global-memory workspace (L2 resident), not in shared memory,
// so the occupancy of the common ABA path is unchanged.  Per-constraint state (enabled, multipliers, reference
// placements) persists in global memory across evaluations and launches, like the reference's constraint
// objects do.
#pragma once

constexpr double CONS_MIN_REGULARIZER = 1.0e-11;   // constraint_solvers.cc:15
constexpr double CONS_RELAX_MIN = 0.01, CONS_RELAX_MAX = 1.0;
constexpr int CONS_PGS_MAX_ITER = 100;             // engine.cc:62
constexpr int CONS_RELAX_MIN_ITER = 20, CONS_RELAX_MAX_ITER = 30;
// the per-lane count of enabled constraints (shared-memory field cons_off) weighs a joint bound CONS_BOUND_UNIT and a
// contact frame 1, so that "contacts only" is one comparison
constexpr double CONS_BOUND_UNIT = 1024.0;

// ---- addressing -----------------------------------------------------------------------------------
// this env's own row of the per-env global tables (padding envs of the last warp get their own rows, < n_pad)
#define CONS_COL(c) (static_cast<size_t>(blockIdx.x) * (32 / KP->L) + (c).lane / KP->L)
#define CST(off) (KP->cstate[CONS_COL(c) * KP->cs_total + (off)])
// inside the solver the row base is hoisted once (cw / cs): the index arithmetic above costs more than the load
#define CW_ROW(c) (static_cast<size_t>(jb_cw_slot) * (32 / KP->L) + (c).lane / KP->L)
#define CWK(off) (KP->cwork[CW_ROW(c) * KP->cw_total + (off)])
JB_DI int cs_joint(int k) { return CS_JOINT0 + CS_JOINT_SIZE * k; }
JB_DI int cs_contact(int k) { return CS_JOINT0 + CS_JOINT_SIZE * KP->n_jc + CS_CONTACT_SIZE * k; }
// workspace layout (doubles per env)
struct CwLayout { int OM, VV, AD, YC, MM, JJ, YY, AA, AL, GA, BB, LA, YV, YP, AC, DD, TT, total; };
JB_HD CwLayout cw_layout(int njoints, int nv, int m_max) {
    CwLayout w; int o = 0;
    w.OM = o; o += 12 * njoints;
    w.VV = o; o += 6 * njoints;
    w.AD = o; o += 6 * njoints;
    w.YC = o; o += 21 * njoints;
    w.MM = o; o += nv * nv;
    w.JJ = o; o += m_max * nv;
    w.YY = o; o += m_max * nv;
    w.AA = o; o += m_max * m_max;
    w.AL = o; o += m_max * m_max;
    w.GA = o; o += m_max; w.BB = o; o += m_max; w.LA = o; o += m_max; w.YV = o; o += m_max; w.YP = o; o += m_max; w.AC = o; o += m_max;
    w.DD = o; o += nv; w.TT = o; o += nv;
    w.total = o;
    return w;
}

// a record of another lane of the same env, seen from this lane
JB_DI const double* rec_of(const Ctx& c, const JointMap& jm) { return jb_smem + KP->rec_off[jm.rec] * 32 + (c.lane - c.sub + jm.sub); }
JB_DI double* rec_of_mut(const Ctx& c, const JointMap& jm, int s) { return jb_smem + KP->rec_off[jm.rec] * 32 + (c.lane - c.sub + s); }
JB_DI Xf ld_xf32(const double* p) {
    Xf M;
#pragma unroll
    for (int k = 0; k < 9; ++k) M.R[k] = p[k * 32];
    M.p = mk(p[9 * 32], p[10 * 32], p[11 * 32]);
    return M;
}
JB_DI Mot ld_mot32(const double* p) { Mot m; m.l = mk(p[0], p[32], p[64]); m.a = mk(p[96], p[128], p[160]); return m; }
JB_DI Mot motion_act(const Xf& M, Mot m) { Mot r; r.a = rmul(M.R, m.a); r.l = rmul(M.R, m.l) + cross(M.p, r.a); return r; }
// motion subspace column d of a joint, in the joint frame
JB_DI Mot subspace_col(int kind, V3 ax, int d) {
    Mot s = mzero();
    if (kind == REC_FREE) {
        if (d < 3) s.l = mk(d == 0, d == 1, d == 2); else s.a = mk(d == 3, d == 4, d == 5);
    } else if (kind == REC_SPH) s.a = mk(d == 0, d == 1, d == 2);   // (`ax` holds the rotor inertias of a spherical record)
    else if (kind == REC_PRISM) s.l = ax;
    else s.a = ax;
    return s;
}
JB_DI double mdot(Mot a, Mot b) { return dot(a.l, b.l) + dot(a.a, b.a); }

// pinocchio::log3 (explog.hpp)
JB_DI V3 cons_log3(const double* R) {
    const double PI = 3.14159265358979323846;
    const double tr = R[0] + R[4] + R[8];
    double theta;
    if (tr >= 3.0) theta = 0.0;
    else if (tr <= -1.0) theta = PI;
    else theta = acos((tr - 1.0) / 2.0);
    if (theta >= PI - 1e-2) {
        const double cphi = -(tr - 1.0) / 2.0;
        const double beta = theta * theta / (1.0 + cphi);
        const V3 tmp = mk((R[0] + cphi) * beta, (R[4] + cphi) * beta, (R[8] + cphi) * beta);
        return mk((R[7] > R[5] ? 1.0 : -1.0) * (tmp.x > 0.0 ? sqrt(tmp.x) : 0.0),
                  (R[2] > R[6] ? 1.0 : -1.0) * (tmp.y > 0.0 ? sqrt(tmp.y) : 0.0),
                  (R[3] > R[1] ? 1.0 : -1.0) * (tmp.z > 0.0 ? sqrt(tmp.z) : 0.0));
    }
    // TaylorSeriesExpansion<double>::precision<3>() = eps^(1/4)
    const double t = ((theta > 1.220703125e-4) ? theta / sin(theta) : 1.0) / 2.0;
    return mk(t * (R[7] - R[5]), t * (R[2] - R[6]), t * (R[3] - R[1]));
}

// ---- per-constraint state updates, called from the forward sweep of rhs_impl ------------------------
// Model::resetConstraints + the start-time configuration of Engine::start (model.cc:1026-1045,
// engine.cc:1268-1309): everything disabled and zeroed, then -- with the constraint contact model -- every
// bound and contact constraint enabled (the first computeAllTerms disables those that are not active).
__device__ __noinline__ void cons_reset(const Ctx c) {
    const int L = KP->L;
    double count = 0.0;
    const bool cm = KP->opt.contact_model == JB_CONTACT_CONSTRAINT;
    for (int r = 0; r < KP->nrec; ++r) {
        const RecInt* ri = KP->rint + (r * L + c.sub);
        if (ri->kind == REC_PAD || !ri->owner) continue;
        if (!rec_is_big(ri->kind)) {
            const int k = KP->jc_of_joint[ri->joint];
            if (k >= 0) {
                const double* rp = jb_smem + KP->rec_off[r] * 32 + c.lane;
                const bool on = cm && ri->has_limit;   // unbounded joints are disabled right away (engine.cc:3297-3310)
                CST(cs_joint(k) + 0) = on ? 1.0 : 0.0;
                if (cm) CST(cs_joint(k) + 1) = 0.0;   // setRotationDir(false)
                CST(cs_joint(k) + 2) = RP(R1_Q);
                CST(cs_joint(k) + 3) = 0.0;
                if (on) count += CONS_BOUND_UNIT;
            }
        }
        for (int q = 0; q < ri->ncontact; ++q) {
            const ContactSlot* ct = KP->cslots + ((ri->contact0 + q) * L + c.sub);
            if (ct->contact < 0) continue;
            const int o = cs_contact(ct->contact);
            for (int e = 0; e < CS_CONTACT_SIZE; ++e) CST(o + e) = 0.0;
            CST(o) = cm ? 1.0 : 0.0;
            if (cm) count += 1.0;
        }
    }
    if (c.sub == 0) CST(CS_SOLVE_FAILED) = 0.0;
    SMF(c, KP->cons_off) = count;
}

// number of enabled constraints this lane owns, from the persistent state (kernel entry)
__device__ __noinline__ void cons_load_count(const Ctx c) {
    const int L = KP->L;
    double count = 0.0;
    for (int r = 0; r < KP->nrec; ++r) {
        const RecInt* ri = KP->rint + (r * L + c.sub);
        if (ri->kind == REC_PAD || !ri->owner) continue;
        if (!rec_is_big(ri->kind)) {
            const int k = KP->jc_of_joint[ri->joint];
            if (k >= 0 && CST(cs_joint(k)) != 0.0) count += CONS_BOUND_UNIT;
        }
        for (int q = 0; q < ri->ncontact; ++q) {
            const ContactSlot* ct = KP->cslots + ((ri->contact0 + q) * L + c.sub);
            if (ct->contact >= 0 && CST(cs_contact(ct->contact)) != 0.0) count += 1.0;
        }
    }
    SMF(c, KP->cons_off) = count;
}

// computePositionLimitsForcesAlgo (engine.cc:3253-3338) for the bounded joints this lane owns, at the stage state
__device__ __noinline__ void cons_update_bounds(const Ctx c, int* status) {
    const int L = KP->L;
    const double eps = KP->opt.contact_transition_eps;
    for (int r = 0; r < KP->nrec; ++r) {
        const RecInt* ri = KP->rint + (r * L + c.sub);
        if (ri->kind == REC_PAD || rec_is_big(ri->kind) || !ri->owner || !ri->has_limit) continue;
        const int k = KP->jc_of_joint[ri->joint];
        if (k < 0) continue;
        const RecDbl* rd = JB_RDBL + (r * L + c.sub);
        const double q = SMF(c, KP->rec_off[r] + R1_QS), lo = rd->q_lo, hi = rd->q_hi;
        const int o = cs_joint(k);
        const bool was = CST(o) != 0.0;
        if (hi < q || q < lo) {
            CST(o + 2) = fmin(fmax(q, lo), hi);
            CST(o + 1) = (hi < q) ? 1.0 : 0.0;
            CST(o) = 1.0;
            if (!was) SMF(c, KP->cons_off) += CONS_BOUND_UNIT;
            *status |= JB_ENV_JOINT_LIMIT;
        } else if (lo + eps < q && q < hi - eps) {
            if (was) { CST(o) = 0.0; CST(o + 3) = 0.0; SMF(c, KP->cons_off) -= CONS_BOUND_UNIT; }
        }
    }
}

// computeContactDynamicsAtFrame, contacts.model == "constraint" (engine.cc:3133-3194): enable below the
// ground, disable above transitionEps, and keep the reference placement on the ground surface.
__device__ __noinline__ void cons_update_contact(const Ctx c, int contact, const Xf oM, const double* placement, bool owner) {
    if (!owner) return;
    Xf P;
#pragma unroll
    for (int k = 0; k < 9; ++k) P.R[k] = placement[k];
    P.p = ld3(placement + 9);
    const V3 pos = oM.p + rmul(oM.R, P.p);
    const double depth = pos.z;   // flat ground, n = z
    const int o = cs_contact(contact);
    bool on = CST(o) != 0.0;
    if (depth < 0.0) { if (!on) { CST(o) = 1.0; SMF(c, KP->cons_off) += 1.0; on = true; } }
    else if (depth > KP->opt.contact_transition_eps) {
        if (on) { CST(o) = 0.0; for (int e = 1; e <= 4; ++e) CST(o + e) = 0.0; SMF(c, KP->cons_off) -= 1.0; on = false; }
    }
    if (on) {
        double Rf[9];
        mat3mul(oM.R, P.R, Rf);
#pragma unroll
        for (int k = 0; k < 9; ++k) CST(o + 5 + k) = Rf[k];
        CST(o + 14) = pos.x; CST(o + 15) = pos.y; CST(o + 16) = pos.z - depth;
    }
}

// ---- dense helpers on the workspace ------------------------------------------------------------------
// from here on `cw` is the hoisted base of this env's workspace row
#undef CWK
#define CWK(off) (cw[(off)])
// in-place lower Cholesky of the n x n matrix at `off` (row stride ld); false if not positive definite
JB_DI bool cw_llt(double* const cw, int off, int n, int ld) {
    for (int j = 0; j < n; ++j) {
        double s = CWK(off + j * ld + j);
        for (int k = 0; k < j; ++k) { const double l = CWK(off + j * ld + k); s -= l * l; }
        if (!(s > 0.0)) return false;
        const double d = sqrt(s);
        CWK(off + j * ld + j) = d;
        for (int i = j + 1; i < n; ++i) {
            double t = CWK(off + i * ld + j);
            for (int k = 0; k < j; ++k) t -= CWK(off + i * ld + k) * CWK(off + j * ld + k);
            CWK(off + i * ld + j) = t / d;
        }
    }
    return true;
}
JB_DI void cw_forward(double* const cw, int Loff, int n, int ld, int x) {   // L y = x, in place
    for (int i = 0; i < n; ++i) {
        double s = CWK(x + i);
        for (int k = 0; k < i; ++k) s -= CWK(Loff + i * ld + k) * CWK(x + k);
        CWK(x + i) = s / CWK(Loff + i * ld + i);
    }
}
JB_DI void cw_backward(double* const cw, int Loff, int n, int ld, int x) {  // L^T y = x, in place
    for (int i = n - 1; i >= 0; --i) {
        double s = CWK(x + i);
        for (int k = i + 1; k < n; ++k) s -= CWK(Loff + k * ld + i) * CWK(x + k);
        CWK(x + i) = s / CWK(Loff + i * ld + i);
    }
}

// PGSSolver::ProjectedGaussSeidelIter + Solver (constraint_solvers.cc:107-318).  Constraint order: joint
// bounds, then contact frames (ConstraintTree::foreach, model.h:43-46).
JB_DI bool cons_pgs(const Ctx& c, double* const cw, const CwLayout& w, int m, int n_active) {
    const int ld = KP->m_max;
    const JbOptions& opt = KP->opt;
    for (int k = 0; k < m; ++k) CWK(w.YV + k) = 0.0;
    auto residual = [&](int k) {
        double s = 0.0;
        for (int r = 0; r < m; ++r) s += CWK(w.AA + k * ld + r) * CWK(w.LA + r);   // A.col(k).dot(x) == row k (A is symmetric)
        return CWK(w.BB + k) - s;
    };
    for (int iter = 0; iter < CONS_PGS_MAX_ITER; ++iter) {
        for (int k = 0; k < m; ++k) CWK(w.YP + k) = CWK(w.YV + k);
        const double ratio = (static_cast<double>(CONS_PGS_MAX_ITER - CONS_RELAX_MIN_ITER) - iter) /
                             (CONS_PGS_MAX_ITER - CONS_RELAX_MIN_ITER - CONS_RELAX_MAX_ITER);
        double wr = CONS_RELAX_MAX;
        if (ratio < 1.0) {
            wr = CONS_RELAX_MIN;
            if (ratio > 0.0) wr += (CONS_RELAX_MAX - CONS_RELAX_MIN) * (ratio * ratio);
        }
        for (int pass = 0; pass < 3; ++pass) {
            for (int a = 0; a < n_active; ++a) {
                // active list entry: first row * 2 + kind (written by constrained_solve)
                const int code = static_cast<int>(CWK(w.AC + a));
                const bool is_joint = (code & 1) == 0;
                const int start = code >> 1;
                if (is_joint) {
                    if (pass != 0) continue;
                    const double y = residual(start);
                    CWK(w.YV + start) = y;
                    double e = CWK(w.LA + start) + wr * y / CWK(w.AA + start * ld + start);
                    CWK(w.LA + start) = fmax(e, 0.0);
                    continue;
                }
                if (pass == 0) {          // normal force: lambda_z >= 0
                    const int i0 = start + 2;
                    const double y = residual(i0);
                    CWK(w.YV + i0) = y;
                    const double e = CWK(w.LA + i0) + wr * y / CWK(w.AA + i0 * ld + i0);
                    CWK(w.LA + i0) = fmax(e, 0.0);
                } else if (pass == 1) {   // torsional friction |lambda_3| <= torsion * lambda_z
                    const int i0 = start + 3;
                    if (opt.contact_torsion < D_EPS) { CWK(w.LA + i0) *= 0.0; continue; }
                    const double y = residual(i0);
                    CWK(w.YV + i0) = y;
                    const double e = CWK(w.LA + i0) + wr * y / CWK(w.AA + i0 * ld + i0);
                    const double thr = opt.contact_torsion * CWK(w.LA + start + 2);
                    CWK(w.LA + i0) = fmin(fmax(e, -thr), thr);
                } else {                  // Coulomb cone |(lambda_x, lambda_y)| <= friction * lambda_z
                    const int i0 = start, i1 = start + 1;
                    if (opt.contact_friction < D_EPS) { CWK(w.LA + i0) *= 0.0; CWK(w.LA + i1) *= 0.0; continue; }
                    const double y0 = residual(i0), y1 = residual(i1);
                    CWK(w.YV + i0) = y0; CWK(w.YV + i1) = y1;
                    const double A_max = fmax(CWK(w.AA + i0 * ld + i0), CWK(w.AA + i1 * ld + i1));
                    const double iA_max = 1.0 / A_max;     // (one division instead of two on the critical path)
                    double e0 = CWK(w.LA + i0) + wr * y0 * iA_max;
                    double e1 = CWK(w.LA + i1) + wr * y1 * iA_max;
                    const double thr = opt.contact_friction * CWK(w.LA + start + 2);
                    const double sq = e0 * e0 + e1 * e1;
                    { const double scale = sq > thr * thr ? thr * rsqrt(sq) : 1.0; e0 *= scale; e1 *= scale; }   // (thr / sqrt(sq), branch-free)
                    CWK(w.LA + i0) = e0; CWK(w.LA + i1) = e1;
                }
            }
        }
        double ymax = 0.0;
        for (int k = 0; k < m; ++k) ymax = fmax(ymax, fabs(CWK(w.YV + k)));
        const double tol = opt.tol_abs + opt.tol_rel * ymax + D_EPS;
        bool converged = true;
        for (int k = 0; k < m; ++k) if (!(fabs(CWK(w.YV + k) - CWK(w.YP + k)) < tol)) { converged = false; break; }
        if (converged) return true;
    }
    return false;
}

// spatial accelerations (gravity-free frame) from the joint accelerations now in the records: what the
// IMU slots and the pools hold after the third ABA sweep
__device__ __noinline__ void cons_refresh_accelerations(const Ctx c) {
    const int L = KP->L;
    const JbOptions& opt = KP->opt;
    Mot agc = mzero();
    for (int r = 0; r < KP->nrec; ++r) {
        const RecInt* ri = KP->rint + (r * L + c.sub);
        const int kind = ri->kind;
        if (kind == REC_PAD) continue;
        const RecDbl* rd = JB_RDBL + (r * L + c.sub);
        const int base = KP->rec_off[r];
        double* const rp = jb_smem + base * 32 + c.lane;
        if (ri->parent_rec < 0) {
            agc.l = mk(-opt.gravity[0], -opt.gravity[1], -opt.gravity[2]);
            agc.a = mk(-opt.gravity[3], -opt.gravity[4], -opt.gravity[5]);
        } else if (!ri->carry_in) agc = sm_load_mot(c, KP->pool_off + POOL_SIZE * ri->parent_pool);
        Xf li; sm_load_xf(c, base, li);
        Mot ag;
        if (kind == REC_FREE) ag = motion_act_inv(li, agc) + sm_load_mot(c, base + RF_A);
        else if (kind == REC_SPH) {
            ag = sm_load_mot(c, base + KP->sph_off + RS_BIAS) + motion_act_inv(li, agc);
            ag.a = ag.a + mk(RP(RF_A + 3), RP(RF_A + 4), RP(RF_A + 5));
        } else {
            ag = sm_load_mot(c, base + R1_BIAS) + motion_act_inv(li, agc);
            const V3 ax = ld3(rd->axis);
            const double ddq = RP(R1_A);
            if (kind == REC_PRISM) ag.l = ag.l + ddq * ax; else ag.a = ag.a + ddq * ax;
        }
        if (ri->pool >= 0) sm_store_mot(c, KP->pool_off + POOL_SIZE * ri->pool, ag);
        if (ri->imu_slot >= 0) sm_store_mot(c, KP->imu_off + IMUSLOT_SIZE * ri->imu_slot + 6, ag);
        agc = ag;
    }
    __syncwarp(c.gmask);
}

// ---- the solve -----------------------------------------------------------------------------------------
// Called by all lanes of the env right after the ABA sweeps (records hold liMi, bias and the unconstrained
// accelerations).  Returns false when the PGS sweep did not converge.
__device__ __noinline__ bool constrained_solve(const Ctx c, int* status) {
    const int L = KP->L, nv = KP->nv, nj = KP->njoints, ld = KP->m_max;
    const CwLayout w = cw_layout(nj, nv, ld);
    const JbOptions& opt = KP->opt;
    double* const cw = KP->cwork + CW_ROW(c) * KP->cw_total;
    __syncwarp(c.gmask);
    // ---------------- 1. tree quantities, joint-space inertia and its Cholesky factor (sub-lane 0)
    if (c.sub == 0) {
        for (int j = 1; j < nj; ++j) {
            const JointMap jm = KP->jmap[j];
            const double* rq = rec_of(c, jm);
            const RecDbl* rd = JB_RDBL + (jm.rec * L + jm.sub);
            const Xf li = ld_xf32(rq);
            const V3 ax = ld3(rd->axis);
            Mot vJ = mzero();
            if (rec_is_big(jm.kind)) vJ = ld_mot32(rq + RF_VS * 32);   // (spherical: linear half held at zero)
            else if (jm.kind == REC_PRISM) vJ.l = rq[R1_VS * 32] * ax;
            else vJ.a = rq[R1_VS * 32] * ax;
            Xf oM; Mot v, aD;
            if (jm.parent == 0) { oM = li; v = vJ; aD = mzero(); }
            else {
                Xf oMp; Mot vp, aDp;
#pragma unroll
                for (int k = 0; k < 9; ++k) oMp.R[k] = CWK(w.OM + 12 * jm.parent + k);
                oMp.p = mk(CWK(w.OM + 12 * jm.parent + 9), CWK(w.OM + 12 * jm.parent + 10), CWK(w.OM + 12 * jm.parent + 11));
                vp.l = mk(CWK(w.VV + 6 * jm.parent), CWK(w.VV + 6 * jm.parent + 1), CWK(w.VV + 6 * jm.parent + 2));
                vp.a = mk(CWK(w.VV + 6 * jm.parent + 3), CWK(w.VV + 6 * jm.parent + 4), CWK(w.VV + 6 * jm.parent + 5));
                aDp.l = mk(CWK(w.AD + 6 * jm.parent), CWK(w.AD + 6 * jm.parent + 1), CWK(w.AD + 6 * jm.parent + 2));
                aDp.a = mk(CWK(w.AD + 6 * jm.parent + 3), CWK(w.AD + 6 * jm.parent + 4), CWK(w.AD + 6 * jm.parent + 5));
                mat3mul(oMp.R, li.R, oM.R);
                oM.p = oMp.p + rmul(oMp.R, li.p);
                v = motion_act_inv(li, vp) + vJ;
                aD = motion_cross(v, vJ) + motion_act_inv(li, aDp);   // Model::computeConstraints (model.cc:1255-1268)
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) CWK(w.OM + 12 * j + k) = oM.R[k];
            CWK(w.OM + 12 * j + 9) = oM.p.x; CWK(w.OM + 12 * j + 10) = oM.p.y; CWK(w.OM + 12 * j + 11) = oM.p.z;
            CWK(w.VV + 6 * j) = v.l.x; CWK(w.VV + 6 * j + 1) = v.l.y; CWK(w.VV + 6 * j + 2) = v.l.z;
            CWK(w.VV + 6 * j + 3) = v.a.x; CWK(w.VV + 6 * j + 4) = v.a.y; CWK(w.VV + 6 * j + 5) = v.a.z;
            CWK(w.AD + 6 * j) = aD.l.x; CWK(w.AD + 6 * j + 1) = aD.l.y; CWK(w.AD + 6 * j + 2) = aD.l.z;
            CWK(w.AD + 6 * j + 3) = aD.a.x; CWK(w.AD + 6 * j + 4) = aD.a.y; CWK(w.AD + 6 * j + 5) = aD.a.z;
            SymY Y;
            inertia_to_sym(rd->inertia[0], ld3(rd->inertia + 1), rd->inertia + 4, Y);
#pragma unroll
            for (int k = 0; k < 6; ++k) { CWK(w.YC + 21 * j + k) = Y.A[k]; CWK(w.YC + 21 * j + 15 + k) = Y.D[k]; }
#pragma unroll
            for (int k = 0; k < 9; ++k) CWK(w.YC + 21 * j + 6 + k) = Y.B[k];
            // unconstrained accelerations
            if (jm.kind == REC_FREE) { for (int d = 0; d < 6; ++d) CWK(w.DD + jm.idx_v + d) = rq[(RF_A + d) * 32]; }
            else if (jm.kind == REC_SPH) { for (int d = 0; d < 3; ++d) CWK(w.DD + jm.idx_v + d) = rq[(RF_A + 3 + d) * 32]; }
            else CWK(w.DD + jm.idx_v) = rq[R1_A * 32];
        }
        for (int k = 0; k < nv * nv; ++k) CWK(w.MM + k) = 0.0;
        // composite-rigid-body algorithm (pinocchio_overload::crba, overload.h:99-124)
        for (int j = nj - 1; j > 0; --j) {
            const JointMap jm = KP->jmap[j];
            const RecDbl* rd = JB_RDBL + (jm.rec * L + jm.sub);
            const V3 ax = ld3(rd->axis);
            SymY Y;
#pragma unroll
            for (int k = 0; k < 6; ++k) { Y.A[k] = CWK(w.YC + 21 * j + k); Y.D[k] = CWK(w.YC + 21 * j + 15 + k); }
#pragma unroll
            for (int k = 0; k < 9; ++k) Y.B[k] = CWK(w.YC + 21 * j + 6 + k);
            for (int d = 0; d < jm.nvj; ++d) {
                const Mot F = sym_mul_motion(Y, subspace_col(jm.kind, ax, d));
                for (int e = 0; e < jm.nvj; ++e) CWK(w.MM + (jm.idx_v + e) * nv + jm.idx_v + d) = mdot(subspace_col(jm.kind, ax, e), F);
                Mot G = F;
                int jj = j;
                while (KP->jmap[jj].parent > 0) {
                    G = force_act(ld_xf32(rec_of(c, KP->jmap[jj])), G);
                    jj = KP->jmap[jj].parent;
                    const JointMap jp = KP->jmap[jj];
                    const V3 axp = ld3((JB_RDBL + (jp.rec * L + jp.sub))->axis);
                    for (int e = 0; e < jp.nvj; ++e) {
                        const double val = mdot(subspace_col(jp.kind, axp, e), G);
                        CWK(w.MM + (jp.idx_v + e) * nv + jm.idx_v + d) = val;
                        CWK(w.MM + (jm.idx_v + d) * nv + jp.idx_v + e) = val;
                    }
                }
            }
            if (jm.kind == REC_SPH) {
                CWK(w.MM + jm.idx_v * nv + jm.idx_v) += ax.x;
                CWK(w.MM + (jm.idx_v + 1) * nv + jm.idx_v + 1) += ax.y;
                CWK(w.MM + (jm.idx_v + 2) * nv + jm.idx_v + 2) += ax.z;
            } else if (jm.kind != REC_FREE) CWK(w.MM + jm.idx_v * nv + jm.idx_v) += rd->armature;
            if (jm.parent > 0) {
                SymY T;
                sym_transform(ld_xf32(rec_of(c, jm)), Y, T);
                const int p = jm.parent;
#pragma unroll
                for (int k = 0; k < 6; ++k) { CWK(w.YC + 21 * p + k) += T.A[k]; CWK(w.YC + 21 * p + 15 + k) += T.D[k]; }
#pragma unroll
                for (int k = 0; k < 9; ++k) CWK(w.YC + 21 * p + 6 + k) += T.B[k];
            }
        }
        if (!cw_llt(cw, w.MM, nv, nv)) *status |= JB_ENV_NAN;
    }
    __syncwarp(c.gmask);
    // ---------------- 2. rows of J, drift, rows of L^-1 J^T: constraints dealt round-robin to the lanes
    int m = 0, n_active = 0;
    {
        int count = 0;
        for (int k = 0; k < KP->n_jc + KP->n_cc; ++k) {
            const bool is_joint = k < KP->n_jc;
            const int o = is_joint ? cs_joint(k) : cs_contact(k - KP->n_jc);
            if (CST(o) == 0.0) continue;
            const int start = m, dim = is_joint ? 1 : 4;
            m += dim;
            if (c.sub == 0) CWK(w.AC + n_active) = static_cast<double>(2 * start + (is_joint ? 0 : 1));   // list for the sweep
            ++n_active;
            const bool mine = (count++ % L) == c.sub;
            if (!mine) continue;
            for (int r = 0; r < dim; ++r) for (int e = 0; e < nv; ++e) CWK(w.JJ + (start + r) * nv + e) = 0.0;
            if (is_joint) {
                // JointConstraint::computeJacobianAndDrift (joint_constraint.cc:141-163)
                const JointMap jm = KP->jmap[KP->jc_joint[k]];
                const double* rq = rec_of(c, jm);
                const double sgn = CST(o + 1) != 0.0 ? -1.0 : 1.0;
                const double omega = 2.0 * 3.14159265358979323846 * opt.contact_stabilization_freq;
                CWK(w.JJ + start * nv + jm.idx_v) = sgn;
                CWK(w.GA + start) = sgn * (omega * omega * (rq[R1_QS * 32] - CST(o + 2)) + 2.0 * omega * rq[R1_VS * 32]);
                CWK(w.LA + start) = CST(o + 3);
            } else {
                // FrameConstraint::computeJacobianAndDrift (frame_constraint.cc:103-183), flat ground: local frame = world axes
                const ContactMap cm = KP->cmap[k - KP->n_jc];
                const int j = cm.joint;
                Xf oM, P;
#pragma unroll
                for (int e = 0; e < 9; ++e) { oM.R[e] = CWK(w.OM + 12 * j + e); P.R[e] = cm.placement[e]; }
                oM.p = mk(CWK(w.OM + 12 * j + 9), CWK(w.OM + 12 * j + 10), CWK(w.OM + 12 * j + 11));
                P.p = ld3(cm.placement + 9);
                double Rf[9];
                mat3mul(oM.R, P.R, Rf);
                const V3 pf = oM.p + rmul(oM.R, P.p);
                for (int jj = j; jj > 0; jj = KP->jmap[jj].parent) {
                    const JointMap jp = KP->jmap[jj];
                    const V3 axp = ld3((JB_RDBL + (jp.rec * L + jp.sub))->axis);
                    Xf oMj;
#pragma unroll
                    for (int e = 0; e < 9; ++e) oMj.R[e] = CWK(w.OM + 12 * jj + e);
                    oMj.p = mk(CWK(w.OM + 12 * jj + 9), CWK(w.OM + 12 * jj + 10), CWK(w.OM + 12 * jj + 11));
                    for (int d = 0; d < jp.nvj; ++d) {
                        const Mot Jw = motion_act(oMj, subspace_col(jp.kind, axp, d));   // world-frame Jacobian column
                        const V3 lin = Jw.l - cross(pf, Jw.a);                           // transformLocal.actInv, R = 1
                        const int col = jp.idx_v + d;
                        CWK(w.JJ + (start + 0) * nv + col) = lin.x;
                        CWK(w.JJ + (start + 1) * nv + col) = lin.y;
                        CWK(w.JJ + (start + 2) * nv + col) = lin.z;
                        CWK(w.JJ + (start + 3) * nv + col) = Jw.a.z;
                    }
                }
                Mot vj, aDj;
                vj.l = mk(CWK(w.VV + 6 * j), CWK(w.VV + 6 * j + 1), CWK(w.VV + 6 * j + 2));
                vj.a = mk(CWK(w.VV + 6 * j + 3), CWK(w.VV + 6 * j + 4), CWK(w.VV + 6 * j + 5));
                aDj.l = mk(CWK(w.AD + 6 * j), CWK(w.AD + 6 * j + 1), CWK(w.AD + 6 * j + 2));
                aDj.a = mk(CWK(w.AD + 6 * j + 3), CWK(w.AD + 6 * j + 4), CWK(w.AD + 6 * j + 5));
                const Mot vLoc = motion_act_inv(P, vj), aLoc = motion_act_inv(P, aDj);
                const V3 vl = rmul(Rf, vLoc.l), va = rmul(Rf, vLoc.a);
                V3 dl = rmul(Rf, aLoc.l) + cross(va, vl), da = rmul(Rf, aLoc.a);
                const double omega = 2.0 * 3.14159265358979323846 * opt.contact_stabilization_freq;
                const double kp = omega * omega, kd = 2.0 * omega;
                double RrT[9], Rref[9];
#pragma unroll
                for (int e = 0; e < 9; ++e) Rref[e] = CST(o + 5 + e);
                // framePose.R * transformRef.R^T
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b2 = 0; b2 < 3; ++b2) RrT[3 * a + b2] = Rf[3 * a] * Rref[3 * b2] + Rf[3 * a + 1] * Rref[3 * b2 + 1] + Rf[3 * a + 2] * Rref[3 * b2 + 2];
                const V3 dpos = pf - mk(CST(o + 14), CST(o + 15), CST(o + 16));
                const V3 drot = cons_log3(RrT);
                dl = dl + kp * dpos + kd * vl;
                da = da + kp * drot + kd * va;
                CWK(w.GA + start) = dl.x; CWK(w.GA + start + 1) = dl.y; CWK(w.GA + start + 2) = dl.z; CWK(w.GA + start + 3) = da.z;
                for (int e = 0; e < 4; ++e) CWK(w.LA + start + e) = CST(o + 1 + e);
            }
            for (int r = start; r < start + dim; ++r) {
                for (int e = 0; e < nv; ++e) CWK(w.YY + r * nv + e) = CWK(w.JJ + r * nv + e);
                cw_forward(cw, w.MM, nv, nv, w.YY + r * nv);
            }
        }
    }
    __syncwarp(c.gmask);
    // ---------------- 3. rows of A = J M^-1 J^T + regularisation, b = -gamma - J ddq_free
    {
        int count = 0, row = 0;
        for (int k = 0; k < KP->n_jc + KP->n_cc; ++k) {
            const bool is_joint = k < KP->n_jc;
            const int o = is_joint ? cs_joint(k) : cs_contact(k - KP->n_jc);
            if (CST(o) == 0.0) continue;
            const int start = row, dim = is_joint ? 1 : 4;
            row += dim;
            if ((count++ % L) != c.sub) continue;
            for (int r = start; r < start + dim; ++r) {
                // lower triangle only, mirrored (the owner of row r also writes column r of the rows above it)
                for (int q = 0; q <= r; ++q) {
                    double s = 0.0;
                    for (int e = 0; e < nv; ++e) s += CWK(w.YY + r * nv + e) * CWK(w.YY + q * nv + e);
                    if (q == r) s += fmax(s * opt.constraint_regularization, CONS_MIN_REGULARIZER);
                    CWK(w.AA + r * ld + q) = s;
                    CWK(w.AA + q * ld + r) = s;
                }
                double s = 0.0;
                for (int e = 0; e < nv; ++e) s += CWK(w.JJ + r * nv + e) * CWK(w.DD + e);
                CWK(w.BB + r) = -CWK(w.GA + r) - s;
            }
        }
    }
    __syncwarp(c.gmask);
    // ---------------- 4. multipliers, accelerations, contact wrenches (sub-lane 0)
    bool ok = true;
    if (c.sub == 0) {
        if (c.flags & CTX_IGNORE_BOUNDS) {
            // solveJMinvJtv (overload.h:539-551): exact equality solve
            for (int r = 0; r < m; ++r) for (int q = 0; q < m; ++q) CWK(w.AL + r * ld + q) = CWK(w.AA + r * ld + q);
            if (cw_llt(cw, w.AL, m, ld)) {
                for (int r = 0; r < m; ++r) CWK(w.LA + r) = CWK(w.BB + r);
                cw_forward(cw, w.AL, m, ld, w.LA);
                cw_backward(cw, w.AL, m, ld, w.LA);
            }
        } else ok = cons_pgs(c, cw, w, m, n_active);
        // ddq = ddq_free + M^-1 J^T lambda
        for (int e = 0; e < nv; ++e) {
            double s = 0.0;
            for (int r = 0; r < m; ++r) s += CWK(w.JJ + r * nv + e) * CWK(w.LA + r);
            CWK(w.TT + e) = s;
        }
        cw_forward(cw, w.MM, nv, nv, w.TT);
        cw_backward(cw, w.MM, nv, nv, w.TT);
        for (int j = 1; j < nj; ++j) {
            const JointMap jm = KP->jmap[j];
            for (int s = (jm.trunk ? 0 : jm.sub); s < (jm.trunk ? L : jm.sub + 1); ++s) {
                double* rq = rec_of_mut(c, jm, s);
                if (jm.kind == REC_FREE) { for (int d = 0; d < 6; ++d) rq[(RF_A + d) * 32] = CWK(w.DD + jm.idx_v + d) + CWK(w.TT + jm.idx_v + d); }
                else if (jm.kind == REC_SPH) { for (int d = 0; d < 3; ++d) rq[(RF_A + 3 + d) * 32] = CWK(w.DD + jm.idx_v + d) + CWK(w.TT + jm.idx_v + d); }
                else rq[R1_A * 32] = CWK(w.DD + jm.idx_v) + CWK(w.TT + jm.idx_v);
            }
        }
        // multipliers back into the constraints; contact wrenches in the parent joint frame (engine.cc:3790-3822)
        int row = 0;
        for (int k = 0; k < KP->n_jc + KP->n_cc; ++k) {
            const bool is_joint = k < KP->n_jc;
            const int o = is_joint ? cs_joint(k) : cs_contact(k - KP->n_jc);
            if (CST(o) == 0.0) continue;
            if (is_joint) { CST(o + 3) = CWK(w.LA + row); row += 1; continue; }
            for (int e = 0; e < 4; ++e) CST(o + 1 + e) = CWK(w.LA + row + e);
            const ContactMap cm = KP->cmap[k - KP->n_jc];
            double Rj[9];
#pragma unroll
            for (int e = 0; e < 9; ++e) Rj[e] = CWK(w.OM + 12 * cm.joint + e);
            const V3 Fl = rtmul(Rj, mk(CWK(w.LA + row), CWK(w.LA + row + 1), CWK(w.LA + row + 2)));
            const V3 Tl = rtmul(Rj, mk(0.0, 0.0, CWK(w.LA + row + 3)));
            for (int s = (cm.trunk ? 0 : cm.sub); s < (cm.trunk ? L : cm.sub + 1); ++s) {
                double* cp = jb_smem + (KP->cslot_off + CSLOT_SIZE * cm.cslot) * 32 + (c.lane - c.sub + s);
                CO(0) = Fl.x; CO(1) = Fl.y; CO(2) = Fl.z; CO(3) = Tl.x; CO(4) = Tl.y; CO(5) = Tl.z;
            }
            row += 4;
        }
        CST(CS_SOLVE_FAILED) = ok ? 0.0 : CST(CS_SOLVE_FAILED) + 1.0;
    }
    __syncwarp(c.gmask);
    cons_refresh_accelerations(c);
    return ok;
}
