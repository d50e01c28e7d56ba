// jiminy_b200 -- lane-block constraint solve for any lane plan with L > 1: the general form of
// jb_constraints_quadruped.cuh, with its matrices in a per-lane global-memory workspace instead of registers.
//
// Same boxed LCP as jb_constraints.cuh (Engine::computeAcceleration with enabled constraints,
// core/src/engine/engine.cc:3709-3866; PGSSolver, core/src/solver/constraint_solvers.cc:107-448), but the
// joint-space inertia is only ever formed block-wise.  Order the dofs (trunk | lane 0 | ... | lane L-1): lanes
// couple only through the trunk, so with  W_l = M_ll^-1 M_lt  and the Schur complement
// S = M_tt - sum_l M_tl W_l  (n_t x n_t),
//       M^-1 = blockdiag(0, M_ll^-1) + [1; -W] S^-1 [1, -W^T]
//       A_rs = [lane(r) = lane(s)] J_l,r M_ll^-1 J_l,s^T + g_r . S^-1 g_s ,   g_r = J_t,r - J_l,r W_lane(r).
// Every lane builds, in its own row of the workspace, the blocks of its private joints (composite inertias, M_ll
// and its Cholesky factor, M_tl, W), the rows of the constraints it owns (bounds of its joints, its contact
// frames; constraints on trunk joints belong to sub-lane 0), x = M_ll^-1 J_l^T, g, h = S^-1 g and its local block
// of A.  What the lanes of an env exchange, with shuffles: the trunk's composite inertias and S (all-reduces in a
// fixed order, so every lane holds bit-identical copies), and inside the Gauss-Seidel sweep the n_t-vector
// z = sum_r g_r lambda_r.  The sweep visits the constraints in the reference's order (bounds in joint order, then
// contact frames; normal / torsion / friction blocks breadth-first); the lane owning a constraint updates it and
// broadcasts the change of z.
//
// The equality solve of the first start iteration (dense A) stays with jb_constraints.cuh.
#pragma once

constexpr int LB_MAX_NT = 16;   // trunk dofs held in registers during the sweep

// per-lane workspace layout (doubles)
struct LbLayout { int KI, YC, PT, ML, MT, WW, MTT, SS, JL, JT, XL, GG, HH, AL, BB, LA, YV, YP, AD, RG, GA, AC, TT, total; };
JB_HD LbLayout lb_layout(int nrec, int ntrunk, int nl, int nt, int ml, int nact) {
    LbLayout w; int o = 0;
    const int nl1 = nl + 1, nt1 = nt + 1, ml1 = ml + 1;
    w.KI = o; o += 24 * nrec;          // per record: oM (12), v (6), drift acceleration (6)
    w.YC = o; o += 21 * nrec;          // composite inertia of each record's subtree
    w.PT = o; o += 21 * (ntrunk + 1);  // private children's contribution to each trunk record
    w.ML = o; o += nl1 * nl1;          // M_ll, then its Cholesky factor
    w.MT = o; o += nt1 * nl1;          // M_tl  [nt][nl]
    w.WW = o; o += nl1 * nt1;          // W = M_ll^-1 M_lt  [nl][nt]
    w.MTT = o; o += nt1 * nt1;
    w.SS = o; o += nt1 * nt1;          // S, then its Cholesky factor
    w.JL = o; o += ml1 * nl1; w.JT = o; o += ml1 * nt1; w.XL = o; o += ml1 * nl1;
    w.GG = o; o += ml1 * nt1; w.HH = o; o += ml1 * nt1;
    w.AL = o; o += ml1 * ml1;
    w.BB = o; o += ml1; w.LA = o; o += ml1; w.YV = o; o += ml1; w.YP = o; o += ml1; w.AD = o; o += ml1; w.RG = o; o += ml1; w.GA = o; o += ml1;
    w.AC = o; o += nact + 1;           // active constraints in sweep order: owner, first local row, kind
    w.TT = o; o += nl1 + nt1;
    w.total = o;
    return w;
}

#define LBW(off) (lw[(off)])
JB_DI double lb_sum_lanes(const Ctx& c, double x) {   // sum over the lanes of the env, same order on every lane
    const int l0 = c.lane - c.sub;
    double s = __shfl_sync(c.gmask, x, l0);
    for (int k = 1; k < KP->L; ++k) s += __shfl_sync(c.gmask, x, l0 + k);
    return s;
}
// Small dense factorisations and solves run on a copy in local memory: it is cached write-back in L1, whereas a
// value just stored to the global workspace is re-read from L2 (stores do not allocate in L1), which would put an
// L2 round trip on every link of these dependent chains.
constexpr int LB_LOCAL_N = 12;
__device__ __noinline__ bool lb_llt(double* const lw, int off, int n, int ld) {
    if (n > LB_LOCAL_N) {
        for (int j = 0; j < n; ++j) {
            double s = LBW(off + j * ld + j);
            for (int k = 0; k < j; ++k) { const double l = LBW(off + j * ld + k); s -= l * l; }
            if (!(s > 0.0)) return false;
            const double d = sqrt(s);
            LBW(off + j * ld + j) = d;
            for (int i = j + 1; i < n; ++i) {
                double t = LBW(off + i * ld + j);
                for (int k = 0; k < j; ++k) t -= LBW(off + i * ld + k) * LBW(off + j * ld + k);
                LBW(off + i * ld + j) = t / d;
            }
        }
        return true;
    }
    double A[LB_LOCAL_N * LB_LOCAL_N];
    for (int i = 0; i < n; ++i) for (int k = 0; k <= i; ++k) A[i * LB_LOCAL_N + k] = LBW(off + i * ld + k);
    for (int j = 0; j < n; ++j) {
        double s = A[j * LB_LOCAL_N + j];
        for (int k = 0; k < j; ++k) { const double l = A[j * LB_LOCAL_N + k]; s -= l * l; }
        if (!(s > 0.0)) return false;
        const double d = sqrt(s);
        A[j * LB_LOCAL_N + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double t = A[i * LB_LOCAL_N + j];
            for (int k = 0; k < j; ++k) t -= A[i * LB_LOCAL_N + k] * A[j * LB_LOCAL_N + k];
            A[i * LB_LOCAL_N + j] = t / d;
        }
    }
    for (int i = 0; i < n; ++i) for (int k = 0; k <= i; ++k) LBW(off + i * ld + k) = A[i * LB_LOCAL_N + k];
    return true;
}
__device__ __noinline__ void lb_solve(double* const lw, int Loff, int n, int ld, int x) {   // (L L^T) y = x, in place
    if (n > 2 * LB_LOCAL_N) {
        for (int i = 0; i < n; ++i) {
            double s = LBW(x + i);
            for (int k = 0; k < i; ++k) s -= LBW(Loff + i * ld + k) * LBW(x + k);
            LBW(x + i) = s / LBW(Loff + i * ld + i);
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = LBW(x + i);
            for (int k = i + 1; k < n; ++k) s -= LBW(Loff + k * ld + i) * LBW(x + k);
            LBW(x + i) = s / LBW(Loff + i * ld + i);
        }
        return;
    }
    double y[2 * LB_LOCAL_N];
    for (int i = 0; i < n; ++i) y[i] = LBW(x + i);
    for (int i = 0; i < n; ++i) {
        double s = y[i];
        for (int k = 0; k < i; ++k) s -= LBW(Loff + i * ld + k) * y[k];
        y[i] = s / LBW(Loff + i * ld + i);
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < n; ++k) s -= LBW(Loff + k * ld + i) * y[k];
        y[i] = s / LBW(Loff + i * ld + i);
    }
    for (int i = 0; i < n; ++i) LBW(x + i) = y[i];
}
JB_DI Xf lb_load_xf(const double* p) {
    Xf M;
#pragma unroll
    for (int k = 0; k < 9; ++k) M.R[k] = p[k];
    M.p = mk(p[9], p[10], p[11]);
    return M;
}
JB_DI Mot lb_load_mot(const double* p) { Mot m; m.l = mk(p[0], p[1], p[2]); m.a = mk(p[3], p[4], p[5]); return m; }
JB_DI void lb_load_sym(const double* p, SymY& Y) {
#pragma unroll
    for (int k = 0; k < 6; ++k) { Y.A[k] = p[k]; Y.D[k] = p[15 + k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) Y.B[k] = p[6 + k];
}
JB_DI void lb_add_sym(double* p, const SymY& Y) {
#pragma unroll
    for (int k = 0; k < 6; ++k) { p[k] += Y.A[k]; p[15 + k] += Y.D[k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) p[6 + k] += Y.B[k];
}

JB_DI int lb_ndof(const RecInt* rint, int r, int L) { const int k = rint[r * L].kind; return k == REC_PAD ? 0 : (k == REC_FREE ? 6 : 1); }
JB_DI Xf lb_limi(const Ctx& c, int r) { Xf li; sm_load_xf(c, KP->rec_off[r], li); return li; }

// Steps 1-4, shared with jb_constraints_bodies.cuh: kinematics of this lane's records, composite inertias, the
// inertia blocks and their factors.  `rint`, `rdbl`, `dof0` are this lane's columns of the tables.
__device__ __noinline__ void lb_prepare(const Ctx c, const LbLayout w, double* const lw, int* status) {
    const int L = KP->L, nrec = KP->nrec, ntrunk = KP->ntrunk, nt = KP->lb_nt, nl = KP->lb_nl;
    const RecInt* const rint = KP->rint + c.sub;
    const RecDbl* const rdbl = JB_RDBL + c.sub;
    const int32_t* const dof0 = KP->lb_dof0 + c.sub;
    const int my_nl = KP->lb_nl_of[c.sub];
    auto ndof = [&](int r) { return lb_ndof(rint, r, L); };
    auto li_of = [&](int r) { return lb_limi(c, r); };
    // ---------------- 1. world placements, velocities, drift accelerations, own inertias (every record of this lane)
    for (int r = 0; r < nrec; ++r) {
        const RecInt* ri = rint + r * L;
        if (ri->kind == REC_PAD) continue;
        const RecDbl* rd = rdbl + r * L;
        const double* rp = jb_smem + KP->rec_off[r] * 32 + c.lane;
        const Xf li = li_of(r);
        const V3 ax = ld3(rd->axis);
        Mot vJ = mzero();
        if (ri->kind == REC_FREE) vJ = sm_load_mot(c, KP->rec_off[r] + RF_VS);
        else if (ri->kind == REC_PRISM) vJ.l = RP(R1_VS) * ax;
        else vJ.a = RP(R1_VS) * ax;
        Xf oM; Mot v, aD;
        if (ri->parent_rec < 0) { oM = li; v = vJ; aD = mzero(); }
        else {
            const double* pk = lw + w.KI + 24 * ri->parent_rec;
            const Xf oMp = lb_load_xf(pk);
            mat3mul(oMp.R, li.R, oM.R);
            oM.p = oMp.p + rmul(oMp.R, li.p);
            v = motion_act_inv(li, lb_load_mot(pk + 12)) + vJ;
            aD = motion_cross(v, vJ) + motion_act_inv(li, lb_load_mot(pk + 18));   // Model::computeConstraints (model.cc:1255-1268)
        }
        double* k = lw + w.KI + 24 * r;
#pragma unroll
        for (int e = 0; e < 9; ++e) k[e] = oM.R[e];
        k[9] = oM.p.x; k[10] = oM.p.y; k[11] = oM.p.z;
        k[12] = v.l.x; k[13] = v.l.y; k[14] = v.l.z; k[15] = v.a.x; k[16] = v.a.y; k[17] = v.a.z;
        k[18] = aD.l.x; k[19] = aD.l.y; k[20] = aD.l.z; k[21] = aD.a.x; k[22] = aD.a.y; k[23] = aD.a.z;
        SymY Y;
        inertia_to_sym(rd->inertia[0], ld3(rd->inertia + 1), rd->inertia + 4, Y);
        double* y = lw + w.YC + 21 * r;
        for (int e = 0; e < 21; ++e) y[e] = 0.0;
        lb_add_sym(y, Y);
    }
    // ---------------- 2. composite inertias: private subtrees, then the trunk (contributions all-reduced)
    for (int e = 0; e < 21 * ntrunk; ++e) LBW(w.PT + e) = 0.0;
    for (int r = nrec - 1; r >= ntrunk; --r) {
        const RecInt* ri = rint + r * L;
        if (ri->kind == REC_PAD || ri->parent_rec < 0) continue;
        SymY Y, T;
        lb_load_sym(lw + w.YC + 21 * r, Y);
        sym_transform(li_of(r), Y, T);
        lb_add_sym(lw + (ri->parent_rec >= ntrunk ? w.YC : w.PT) + 21 * ri->parent_rec, T);
    }
    for (int e = 0; e < 21 * ntrunk; ++e) LBW(w.PT + e) = lb_sum_lanes(c, LBW(w.PT + e));
    for (int r = ntrunk - 1; r >= 0; --r) {
        double* y = lw + w.YC + 21 * r;
        for (int e = 0; e < 21; ++e) y[e] += LBW(w.PT + 21 * r + e);
        const RecInt* ri = rint + r * L;
        if (ri->parent_rec < 0) continue;
        SymY Y, T;
        lb_load_sym(y, Y);
        sym_transform(li_of(r), Y, T);
        lb_add_sym(lw + w.YC + 21 * ri->parent_rec, T);
    }
    // ---------------- 3. inertia blocks (CRBA with rotor inertia, pinocchio_overload_algorithms.h:99-124):
    //                     M_ll and M_tl of this lane, M_tt replicated
    for (int e = 0; e < nl * nl; ++e) LBW(w.ML + e) = 0.0;
    for (int e = 0; e < nt * nl; ++e) LBW(w.MT + e) = 0.0;
    for (int e = 0; e < nt * nt; ++e) LBW(w.MTT + e) = 0.0;
    for (int r = 0; r < nrec; ++r) {
        const int kind = rint[r * L].kind;
        if (kind == REC_PAD) continue;
        const bool trunk_r = r < ntrunk;
        SymY Y;
        lb_load_sym(lw + w.YC + 21 * r, Y);
        const V3 ax = ld3(rdbl[r * L].axis);
        const int nd = ndof(r), i0 = dof0[r * L];
        for (int d = 0; d < nd; ++d) {
            const Mot F = sym_mul_motion(Y, subspace_col(kind, ax, d));
            const int id = i0 + d;
            for (int e = 0; e < nd; ++e) {
                const double val = mdot(subspace_col(kind, ax, e), F);
                if (trunk_r) LBW(w.MTT + (i0 + e) * nt + id) = val; else LBW(w.ML + (i0 + e) * nl + id) = val;
            }
            Mot G = F;
            int j = r;
            while (rint[j * L].parent_rec >= 0) {
                G = force_act(li_of(j), G);
                j = rint[j * L].parent_rec;
                const V3 axj = ld3(rdbl[j * L].axis);
                const int kj = rint[j * L].kind, ndj = ndof(j), j0 = dof0[j * L];
                for (int e = 0; e < ndj; ++e) {
                    const double val = mdot(subspace_col(kj, axj, e), G);
                    if (trunk_r) { LBW(w.MTT + (j0 + e) * nt + id) = val; LBW(w.MTT + id * nt + j0 + e) = val; }
                    else if (j >= ntrunk) { LBW(w.ML + (j0 + e) * nl + id) = val; LBW(w.ML + id * nl + j0 + e) = val; }
                    else LBW(w.MT + (j0 + e) * nl + id) = val;
                }
            }
        }
        if (kind != REC_FREE) {
            if (trunk_r) LBW(w.MTT + i0 * nt + i0) += rdbl[r * L].armature; else LBW(w.ML + i0 * nl + i0) += rdbl[r * L].armature;
        }
    }
    // ---------------- 4. Cholesky factor of M_ll, W = M_ll^-1 M_lt, Schur complement of the trunk and its factor
    if (!lb_llt(lw, w.ML, my_nl, nl)) *status |= JB_ENV_NAN;
    for (int t = 0; t < nt; ++t) {
        for (int i = 0; i < my_nl; ++i) LBW(w.TT + i) = LBW(w.MT + t * nl + i);
        lb_solve(lw, w.ML, my_nl, nl, w.TT);
        for (int i = 0; i < my_nl; ++i) LBW(w.WW + i * nt + t) = LBW(w.TT + i);
    }
    for (int t1 = 0; t1 < nt; ++t1)
        for (int t2 = 0; t2 <= t1; ++t2) {
            double s = 0.0;
            for (int i = 0; i < my_nl; ++i) s += LBW(w.MT + t1 * nl + i) * LBW(w.WW + i * nt + t2);
            const double val = LBW(w.MTT + t1 * nt + t2) - lb_sum_lanes(c, s);
            LBW(w.SS + t1 * nt + t2) = val; LBW(w.SS + t2 * nt + t1) = val;
        }
    if (!lb_llt(lw, w.SS, nt, nt)) *status |= JB_ENV_NAN;
}

// Called by all lanes of the env after the ABA sweeps.  Returns false when the sweep did not converge.
__device__ __noinline__ bool cons_solve_blocks(const Ctx c, int* status) {
    const int L = KP->L, nrec = KP->nrec, ntrunk = KP->ntrunk, nt = KP->lb_nt, nl = KP->lb_nl, ml = KP->lb_ml;
    const int n_cons = KP->n_jc + KP->n_cc;
    const JbOptions& opt = KP->opt;
    const LbLayout w = lb_layout(nrec, ntrunk, nl, nt, ml, n_cons);
    double* const lw = KP->lwork + (CW_ROW(c) * L + c.sub) * static_cast<size_t>(KP->lw_total);
    const RecInt* const rint = KP->rint + c.sub;
    const RecDbl* const rdbl = JB_RDBL + c.sub;
    const int32_t* const dof0 = KP->lb_dof0 + c.sub;
    const int my_nl = KP->lb_nl_of[c.sub];
    const int lane0 = c.lane - c.sub;
    const double omega = 2.0 * 3.14159265358979323846 * opt.contact_stabilization_freq;
    const double kp = omega * omega, kd = 2.0 * omega;
    __syncwarp(c.gmask);
    auto ndof = [&](int r) { return lb_ndof(rint, r, L); };
    lb_prepare(c, w, lw, status);
    // ---------------- 5. sweep list (replicated on every lane) and the rows this lane owns
    int n_act = 0, my_rows = 0;
    {
        int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < n_cons; ++k) {
            const bool is_joint = k < KP->n_jc;
            const int o = is_joint ? cs_joint(k) : cs_contact(k - KP->n_jc);
            if (CST(o) == 0.0) continue;
            int owner;
            if (is_joint) { const JointMap jm = KP->jmap[KP->jc_joint[k]]; owner = jm.trunk ? 0 : jm.sub; }
            else { const ContactMap cm = KP->cmap[k - KP->n_jc]; owner = cm.trunk ? 0 : cm.sub; }
            const int dim = is_joint ? 1 : 4;
            int start = 0;
#pragma unroll
            for (int s = 0; s < 8; ++s) if (s == owner) { start = cnt[s]; cnt[s] += dim; }
            LBW(w.AC + n_act) = static_cast<double>((owner << 20) | (start << 1) | (is_joint ? 0 : 1));
            ++n_act;
            if (owner != c.sub) continue;
            my_rows = start + dim;
            for (int r = start; r < start + dim; ++r) {
                for (int e = 0; e < nl; ++e) LBW(w.JL + r * nl + e) = 0.0;
                for (int e = 0; e < nt; ++e) LBW(w.JT + r * nt + e) = 0.0;
            }
            if (is_joint) {
                // JointConstraint::computeJacobianAndDrift (joint_constraint.cc:141-163)
                const JointMap jm = KP->jmap[KP->jc_joint[k]];
                const double* rq = jb_smem + KP->rec_off[jm.rec] * 32 + c.lane;
                const double sgn = CST(o + 1) != 0.0 ? -1.0 : 1.0;
                const int id = dof0[jm.rec * L];
                if (jm.trunk) LBW(w.JT + start * nt + id) = sgn; else LBW(w.JL + start * nl + id) = sgn;
                LBW(w.GA + start) = sgn * (kp * (rq[R1_QS * 32] - CST(o + 2)) + kd * rq[R1_VS * 32]);
                LBW(w.LA + start) = CST(o + 3);
            } else {
                // FrameConstraint::computeJacobianAndDrift (frame_constraint.cc:103-183), flat ground: local frame = world axes
                const ContactMap cm = KP->cmap[k - KP->n_jc];
                const int rj = KP->jmap[cm.joint].rec;
                const double* kj = lw + w.KI + 24 * rj;
                const Xf oM = lb_load_xf(kj);
                Xf P;
#pragma unroll
                for (int e = 0; e < 9; ++e) P.R[e] = cm.placement[e];
                P.p = ld3(cm.placement + 9);
                double Rf[9];
                mat3mul(oM.R, P.R, Rf);
                const V3 pf = oM.p + rmul(oM.R, P.p);
                for (int j = rj; j >= 0; j = rint[j * L].parent_rec) {
                    const Xf oMj = lb_load_xf(lw + w.KI + 24 * j);
                    const V3 axj = ld3(rdbl[j * L].axis);
                    const int kj2 = rint[j * L].kind, ndj = ndof(j), j0 = dof0[j * L];
                    for (int d = 0; d < ndj; ++d) {
                        const Mot Jw = motion_act(oMj, subspace_col(kj2, axj, d));   // world-frame Jacobian column
                        const V3 lin = Jw.l - cross(pf, Jw.a);                      // transformLocal.actInv, R = 1
                        const int id = j0 + d;
                        if (j >= ntrunk) {
                            LBW(w.JL + (start + 0) * nl + id) = lin.x; LBW(w.JL + (start + 1) * nl + id) = lin.y;
                            LBW(w.JL + (start + 2) * nl + id) = lin.z; LBW(w.JL + (start + 3) * nl + id) = Jw.a.z;
                        } else {
                            LBW(w.JT + (start + 0) * nt + id) = lin.x; LBW(w.JT + (start + 1) * nt + id) = lin.y;
                            LBW(w.JT + (start + 2) * nt + id) = lin.z; LBW(w.JT + (start + 3) * nt + id) = Jw.a.z;
                        }
                    }
                }
                const Mot vLoc = motion_act_inv(P, lb_load_mot(kj + 12)), aLoc = motion_act_inv(P, lb_load_mot(kj + 18));
                const V3 vl = rmul(Rf, vLoc.l), va = rmul(Rf, vLoc.a);
                V3 dl = rmul(Rf, aLoc.l) + cross(va, vl), da = rmul(Rf, aLoc.a);
                double RrT[9], Rref[9];
#pragma unroll
                for (int e = 0; e < 9; ++e) Rref[e] = CST(o + 5 + e);
                // framePose.R * transformRef.R^T
#pragma unroll
                for (int a2 = 0; a2 < 3; ++a2)
#pragma unroll
                    for (int b2 = 0; b2 < 3; ++b2) RrT[3 * a2 + b2] = Rf[3 * a2] * Rref[3 * b2] + Rf[3 * a2 + 1] * Rref[3 * b2 + 1] + Rf[3 * a2 + 2] * Rref[3 * b2 + 2];
                dl = dl + kp * (pf - mk(CST(o + 14), CST(o + 15), CST(o + 16))) + kd * vl;
                da = da + kp * cons_log3(RrT) + kd * va;
                LBW(w.GA + start) = dl.x; LBW(w.GA + start + 1) = dl.y; LBW(w.GA + start + 2) = dl.z; LBW(w.GA + start + 3) = da.z;
                for (int e = 0; e < 4; ++e) LBW(w.LA + start + e) = CST(o + 1 + e);
            }
        }
    }
    // ---------------- 6. per row: x = M_ll^-1 J_l^T, g, h = S^-1 g, b = -gamma - J ddq_free; local block of A
    for (int r = 0; r < my_rows; ++r) {
        for (int i = 0; i < my_nl; ++i) LBW(w.XL + r * nl + i) = LBW(w.JL + r * nl + i);
        lb_solve(lw, w.ML, my_nl, nl, w.XL + r * nl);
        for (int t = 0; t < nt; ++t) {
            double s = LBW(w.JT + r * nt + t);
            for (int i = 0; i < my_nl; ++i) s -= LBW(w.JL + r * nl + i) * LBW(w.WW + i * nt + t);
            LBW(w.GG + r * nt + t) = s;
            LBW(w.HH + r * nt + t) = s;
        }
        lb_solve(lw, w.SS, nt, nt, w.HH + r * nt);
        double jd = 0.0;
        for (int q = 0; q < nrec; ++q) {
            const int kq = rint[q * L].kind;
            if (kq == REC_PAD) continue;
            const double* rq = jb_smem + KP->rec_off[q] * 32 + c.lane;
            const int ndq = ndof(q), q0 = dof0[q * L];
            for (int d = 0; d < ndq; ++d) {
                const double acc = (kq == REC_FREE) ? rq[(RF_A + d) * 32] : rq[R1_A * 32];
                jd += (q < ntrunk ? LBW(w.JT + r * nt + q0 + d) : LBW(w.JL + r * nl + q0 + d)) * acc;
            }
        }
        LBW(w.BB + r) = -LBW(w.GA + r) - jd;
        LBW(w.YV + r) = 0.0;
    }
    for (int r = 0; r < my_rows; ++r) {
        for (int q = 0; q <= r; ++q) {
            double s = 0.0;
            for (int i = 0; i < my_nl; ++i) s += LBW(w.JL + r * nl + i) * LBW(w.XL + q * nl + i);
            LBW(w.AL + r * ml + q) = s; LBW(w.AL + q * ml + r) = s;
        }
        double a0 = LBW(w.AL + r * ml + r);
        for (int t = 0; t < nt; ++t) a0 += LBW(w.GG + r * nt + t) * LBW(w.HH + r * nt + t);
        const double reg = fmax(a0 * opt.constraint_regularization, CONS_MIN_REGULARIZER);
        LBW(w.AD + r) = a0 + reg; LBW(w.RG + r) = reg;
    }
    double z[LB_MAX_NT];   // z = sum over all rows of g_r lambda_r, identical on every lane
#pragma unroll
    for (int t = 0; t < LB_MAX_NT; ++t) {
        z[t] = 0.0;
        if (t < nt) {
            double s = 0.0;
            for (int r = 0; r < my_rows; ++r) s += LBW(w.GG + r * nt + t) * LBW(w.LA + r);
            z[t] = lb_sum_lanes(c, s);
        }
    }
    // ---------------- 7. projected Gauss-Seidel sweep (constraint_solvers.cc:107-318)
    auto residual = [&](int k) {
        double s = LBW(w.RG + k) * LBW(w.LA + k);
        for (int r = 0; r < my_rows; ++r) s += LBW(w.AL + k * ml + r) * LBW(w.LA + r);
#pragma unroll
        for (int t = 0; t < LB_MAX_NT; ++t) if (t < nt) s += LBW(w.HH + k * nt + t) * z[t];
        return LBW(w.BB + k) - s;
    };
    bool ok = false;
    for (int iter = 0; iter < CONS_PGS_MAX_ITER && !ok; ++iter) {
        for (int r = 0; r < my_rows; ++r) LBW(w.YP + r) = LBW(w.YV + r);
        const double ratio = (static_cast<double>(CONS_PGS_MAX_ITER - CONS_RELAX_MIN_ITER) - iter) /
                             (CONS_PGS_MAX_ITER - CONS_RELAX_MIN_ITER - CONS_RELAX_MAX_ITER);
        double wr = CONS_RELAX_MAX;
        if (ratio < 1.0) {
            wr = CONS_RELAX_MIN;
            if (ratio > 0.0) wr += (CONS_RELAX_MAX - CONS_RELAX_MIN) * (ratio * ratio);
        }
        for (int pass = 0; pass < 3; ++pass) {
            for (int a = 0; a < n_act; ++a) {
                const int code = static_cast<int>(LBW(w.AC + a));
                const int owner = code >> 20, start = (code >> 1) & 0x7ffff;
                const bool is_joint = (code & 1) == 0;
                if (is_joint && pass != 0) continue;
                double d0 = 0.0, d1 = 0.0;
                int r0 = 0, r1 = -1;
                if (owner == c.sub) {
                    if (is_joint || pass == 0) {            // bound / normal force: lambda >= 0
                        r0 = is_joint ? start : start + 2;
                        const double y = residual(r0);
                        LBW(w.YV + r0) = y;
                        const double e = fmax(LBW(w.LA + r0) + wr * y / LBW(w.AD + r0), 0.0);
                        d0 = e - LBW(w.LA + r0);
                        LBW(w.LA + r0) = e;
                    } else if (pass == 1) {                 // torsional friction |lambda_3| <= torsion * lambda_z
                        r0 = start + 3;
                        double e;
                        if (opt.contact_torsion < D_EPS) e = LBW(w.LA + r0) * 0.0;
                        else {
                            const double y = residual(r0);
                            LBW(w.YV + r0) = y;
                            const double thr = opt.contact_torsion * LBW(w.LA + start + 2);
                            e = fmin(fmax(LBW(w.LA + r0) + wr * y / LBW(w.AD + r0), -thr), thr);
                        }
                        d0 = e - LBW(w.LA + r0);
                        LBW(w.LA + r0) = e;
                    } else {                                // Coulomb cone |(lambda_x, lambda_y)| <= friction * lambda_z
                        r0 = start; r1 = start + 1;
                        double e0, e1;
                        if (opt.contact_friction < D_EPS) { e0 = LBW(w.LA + r0) * 0.0; e1 = LBW(w.LA + r1) * 0.0; }
                        else {
                            const double y0 = residual(r0), y1 = residual(r1);
                            LBW(w.YV + r0) = y0; LBW(w.YV + r1) = y1;
                            const double A_max = fmax(LBW(w.AD + r0), LBW(w.AD + r1));
                            const double iA_max = 1.0 / A_max;     // (one division instead of two on the critical path)
                            e0 = LBW(w.LA + r0) + wr * y0 * iA_max;
                            e1 = LBW(w.LA + r1) + wr * y1 * iA_max;
                            const double thr = opt.contact_friction * LBW(w.LA + start + 2);
                            const double sq = e0 * e0 + e1 * e1;
                            { const double scale = sq > thr * thr ? thr * rsqrt(sq) : 1.0; e0 *= scale; e1 *= scale; }   // (thr / sqrt(sq), branch-free)
                        }
                        d0 = e0 - LBW(w.LA + r0); d1 = e1 - LBW(w.LA + r1);
                        LBW(w.LA + r0) = e0; LBW(w.LA + r1) = e1;
                    }
                }
                // the owner's change of z reaches every lane of the env
                if (!__any_sync(c.gmask, d0 != 0.0 || d1 != 0.0)) continue;
#pragma unroll
                for (int t = 0; t < LB_MAX_NT; ++t) {
                    if (t < nt) {
                        double dz = 0.0;
                        if (owner == c.sub) {
                            dz = LBW(w.GG + r0 * nt + t) * d0;
                            if (r1 >= 0) dz += LBW(w.GG + r1 * nt + t) * d1;
                        }
                        z[t] += __shfl_sync(c.gmask, dz, lane0 + owner);
                    }
                }
            }
        }
        double ymax = 0.0;
        for (int r = 0; r < my_rows; ++r) ymax = fmax(ymax, fabs(LBW(w.YV + r)));
        for (int o2 = 1; o2 < L; o2 <<= 1) ymax = fmax(ymax, __shfl_xor_sync(c.gmask, ymax, o2));
        const double tol = opt.tol_abs + opt.tol_rel * ymax + D_EPS;
        bool conv = true;
        for (int r = 0; r < my_rows; ++r) conv = conv && (fabs(LBW(w.YV + r) - LBW(w.YP + r)) < tol);
        ok = __all_sync(c.gmask, conv);
    }
    // ---------------- 8. accelerations: ddq_t += S^-1 z ; ddq_l += sum_r x_r lambda_r - W S^-1 z
#pragma unroll
    for (int t = 0; t < LB_MAX_NT; ++t) if (t < nt) LBW(w.TT + t) = z[t];
    lb_solve(lw, w.SS, nt, nt, w.TT);
    for (int q = 0; q < nrec; ++q) {
        const int kq = rint[q * L].kind;
        if (kq == REC_PAD) continue;
        double* rq = jb_smem + KP->rec_off[q] * 32 + c.lane;
        const int ndq = ndof(q), q0 = dof0[q * L];
        for (int d = 0; d < ndq; ++d) {
            const int id = q0 + d;
            double x;
            if (q < ntrunk) x = LBW(w.TT + id);
            else {
                x = 0.0;
                for (int r = 0; r < my_rows; ++r) x += LBW(w.XL + r * nl + id) * LBW(w.LA + r);
                for (int t = 0; t < nt; ++t) x -= LBW(w.WW + id * nt + t) * LBW(w.TT + t);
            }
            if (kq == REC_FREE) rq[(RF_A + d) * 32] += x; else rq[R1_A * 32] += x;
        }
    }
    // multipliers back into the constraints; contact wrenches in the parent joint frame (engine.cc:3790-3822)
    for (int a = 0, k = 0; k < n_cons; ++k) {
        const bool is_joint = k < KP->n_jc;
        const int o = is_joint ? cs_joint(k) : cs_contact(k - KP->n_jc);
        if (CST(o) == 0.0) continue;
        const int code = static_cast<int>(LBW(w.AC + a));
        ++a;
        const int owner = code >> 20, start = (code >> 1) & 0x7ffff;
        if (is_joint) {
            if (owner == c.sub) CST(o + 3) = LBW(w.LA + start);
            continue;
        }
        const ContactMap cm = KP->cmap[k - KP->n_jc];
        double l4[4] = {0.0, 0.0, 0.0, 0.0};
        if (owner == c.sub)
            for (int e = 0; e < 4; ++e) { l4[e] = LBW(w.LA + start + e); CST(o + 1 + e) = l4[e]; }
        if (cm.trunk)   // a contact frame on a trunk joint is replicated in every lane's slot list
            for (int e = 0; e < 4; ++e) l4[e] = __shfl_sync(c.gmask, l4[e], lane0 + owner);
        if (owner == c.sub || cm.trunk) {
            const Xf oM = lb_load_xf(lw + w.KI + 24 * KP->jmap[cm.joint].rec);
            const V3 Fl = rtmul(oM.R, mk(l4[0], l4[1], l4[2]));
            const V3 Tl = rtmul(oM.R, mk(0.0, 0.0, l4[3]));
            double* cp = jb_smem + (KP->cslot_off + CSLOT_SIZE * cm.cslot) * 32 + c.lane;
            CO(0) = Fl.x; CO(1) = Fl.y; CO(2) = Fl.z; CO(3) = Tl.x; CO(4) = Tl.y; CO(5) = Tl.z;
        }
    }
    if (c.sub == 0) CST(CS_SOLVE_FAILED) = ok ? 0.0 : CST(CS_SOLVE_FAILED) + 1.0;
    __syncwarp(c.gmask);
    cons_refresh_accelerations(c);
    return ok;
}
#undef LBW
