// jiminy_b200 -- contact constraints solved in body space: the constraint path when only contact frames are
// enabled (a standing / walking legged robot away from its joint bounds) and they sit on a few bodies.
//
// Same LCP and the same projected Gauss-Seidel sweep as jb_constraints.cuh (PGSSolver,
// core/src/solver/constraint_solvers.cc:107-318; FrameConstraint, core/src/constraints/frame_constraint.cc:
// 103-183), but the Delassus matrix A = J M^-1 J^T is never formed.  Every contact frame i on body b has
//       J_i = E_i J_b ,   E_i = [ 1  -[r_i]x ; 0 0 0 0 0 1 ]   (4 x 6),
// with J_b the 6 x nv Jacobian of the body at its origin o_b in world axes and r_i = p_i - o_b, so
//       A_ij = E_i Omega_bb' E_j^T ,    Omega_bb' = J_b M^-1 J_b'^T    (6 x 6 per pair of contact bodies).
// Omega has 6 n_b rows (12 for a biped with many contact points per foot, against 4 n_contacts rows of A) and is
// assembled from the lane blocks of jb_constraints_blocks.cuh:  Omega_bb' = [same lane] J_b,l M_ll^-1 J_b',l^T +
// G_b S^-1 G_b'^T,  G_b = J_b,t - J_b,l W.  The sweep keeps  a = Omega F,  F_b = sum_i E_i^T lambda_i  (the
// acceleration of each contact body due to the multipliers): the residual of a row is b_i - E_i a_b - reg lambda_i,
// and a change of lambda_i updates a with one 6-column product.  All lanes of the env run the sweep redundantly on
// private copies of lambda (bit-identical), and share the update of a: each lane owns every L-th row, a is
// double-buffered and one __syncwarp per update orders the exchange.  No shuffles.
// Where the sweep's state lives matters more than its flop count: with one or a few warps per SM every dependent
// access to the global workspace is an exposed L2 round trip (a value just stored is not in L1).  So the mutable
// state of the sweep sits on chip -- lambda, y, y_prev and F in local memory (write-back in L1), a in a small
// shared-memory region spread over the env's lanes -- and only write-once data (Omega, levers, right-hand sides)
// is read from the workspace.
#pragma once

constexpr int BD_MAX_BODIES = 4, BD_MAX_CONTACTS = 16;
struct BdLayout { int OM, AV, GS, HS, PB, total; };
constexpr int BD_PB = 16;   // per contact, shared: r (3), b (4), reciprocals of the regularised diagonal of A (4: 1/max(xx, yy), -, 1/zz, 1/tt), regularisation (4)
JB_HD BdLayout bd_layout(int nb, int nt, int ncc) {
    BdLayout s; int o = 0;
    const int D = 6 * nb, nt1 = nt + 1;
    s.OM = o; o += D * D;
    s.AV = o; o += 0;
    s.GS = o; o += D * nt1;
    s.HS = o; o += D * nt1;
    s.PB = o; o += BD_PB * (ncc + 1);
    s.total = o;
    return s;
}
// per-lane additions, appended to the lane-block layout: per owned body J_l (6 x nl), J_t (6 x nt), X = rows of
// M_ll^-1 J_l^T (6 x nl), c = J ddq_free (6)
struct BdLane { int CB, cb_stride, JBL, JBT, XB, CF, PL, FV, total; };
JB_HD BdLane bd_lane_layout(int base, int ncar, int nl, int nt, int ncc, int nb) {
    BdLane w; int o = base;
    const int nl1 = nl + 1, nt1 = nt + 1;
    w.JBL = 0; w.JBT = 6 * nl1; w.XB = w.JBT + 6 * nt1; w.CF = w.XB + 6 * nl1; w.cb_stride = w.CF + 6;
    w.CB = o; o += w.cb_stride * (ncar + 1);
    w.PL = o; w.FV = o;
    w.total = o;
    return w;
}

#define LBW(off) (lw[(off)])
#define SHW(off) (sh[(off)])
__device__ __noinline__ bool cons_solve_bodies(const Ctx c, int* status) {
    const int L = KP->L, nrec = KP->nrec, ntrunk = KP->ntrunk, nt = KP->lb_nt, nl = KP->lb_nl;
    const int n_cc = KP->n_cc, nb = KP->bd_n, D = 6 * nb;
    const JbOptions& opt = KP->opt;
    const LbLayout w = lb_layout(nrec, ntrunk, nl, nt, KP->lb_ml, KP->n_jc + n_cc);
    const BdLane wl = bd_lane_layout(w.total, KP->bd_ncar, nl, nt, n_cc, nb);
    const BdLayout ws = bd_layout(nb, nt, n_cc);
    double* const lw = KP->lwork + (CW_ROW(c) * L + c.sub) * static_cast<size_t>(KP->lw_total);
    double* const sh = KP->cwork + CW_ROW(c) * static_cast<size_t>(KP->cw_total);
    const RecInt* const rint = KP->rint + c.sub;
    const RecDbl* const rdbl = JB_RDBL + c.sub;
    const int32_t* const dof0 = KP->lb_dof0 + c.sub;
    const int my_nl = KP->lb_nl_of[c.sub];
    const double omega = 2.0 * 3.14159265358979323846 * opt.contact_stabilization_freq;
    const double kp = omega * omega, kd = 2.0 * omega;
    // element e of the env's double-buffered a = Omega F, spread over the env's lanes of the shared-memory region
    // (row i of buffer p sits in field p * nrow + i / L of lane i % L: every lane updates rows in its own column)
    double* const a_base = jb_smem + KP->bd_off * 32 + (c.lane - c.sub);
    const int lsh = KP->bd_lsh, lmask = L - 1, nrow = (D + L - 1) >> lsh;
#define AVS(p, i) (a_base[((p) * nrow + ((i) >> lsh)) * 32 + ((i) & lmask)])
    unsigned body_bits = 0;   // contact body of each contact frame, 2 bits each
    for (int k = 0; k < n_cc; ++k) body_bits |= static_cast<unsigned>(KP->bd_of_contact[k]) << (2 * k);
    double pl_all[12 * BD_MAX_CONTACTS];   // per contact: lambda (4), y (4), y_prev (4)
    double Fv[6 * BD_MAX_BODIES];
    __syncwarp(c.gmask);      // the enabled flags below were written by the lanes owning the contact frames
    unsigned enabled = 0;
    for (int k = 0; k < n_cc; ++k) if (CST(cs_contact(k)) != 0.0) enabled |= 1u << k;
    lb_prepare(c, w, lw, status);
    // ---------------- B. the contact bodies this lane owns: Jacobian at the body origin, X, G, H, c
    for (int e = c.sub; e < D * D; e += L) SHW(ws.OM + e) = 0.0;
    for (int b = 0; b < nb; ++b) {
        if (KP->bd_owner[b] != c.sub) continue;
        double* const cb = lw + wl.CB + wl.cb_stride * KP->bd_slot[b];
        const int rb = KP->bd_rec[b];
        const V3 ob = ld3(lw + w.KI + 24 * rb + 9);
        for (int e = 0; e < 6 * (nl + 1); ++e) cb[wl.JBL + e] = 0.0;
        for (int e = 0; e < 6 * (nt + 1); ++e) cb[wl.JBT + e] = 0.0;
        for (int j = rb; j >= 0; j = rint[j * L].parent_rec) {
            const Xf oMj = lb_load_xf(lw + w.KI + 24 * j);
            const V3 axj = ld3(rdbl[j * L].axis);
            const int kj = rint[j * L].kind, ndj = lb_ndof(rint, j, L), j0 = dof0[j * L];
            for (int d = 0; d < ndj; ++d) {
                const Mot Jw = motion_act(oMj, subspace_col(kj, axj, d));   // world-frame Jacobian column
                const V3 lin = Jw.l - cross(ob, Jw.a);                      // ... at the body origin
                double* col = (j >= ntrunk) ? cb + wl.JBL + j0 + d : cb + wl.JBT + j0 + d;
                const int ld = (j >= ntrunk) ? nl : nt;
                col[0] = lin.x; col[ld] = lin.y; col[2 * ld] = lin.z; col[3 * ld] = Jw.a.x; col[4 * ld] = Jw.a.y; col[5 * ld] = Jw.a.z;
            }
        }
        for (int k = 0; k < 6; ++k) {
            const int xo = static_cast<int>(cb - lw) + wl.XB + k * nl;
            for (int i = 0; i < my_nl; ++i) LBW(xo + i) = cb[wl.JBL + k * nl + i];
            lb_solve(lw, w.ML, my_nl, nl, xo);
            double* const g = sh + ws.GS + (6 * b + k) * nt;
            double* const h = sh + ws.HS + (6 * b + k) * nt;
            for (int t = 0; t < nt; ++t) {
                double s = cb[wl.JBT + k * nt + t];
                for (int i = 0; i < my_nl; ++i) s -= cb[wl.JBL + k * nl + i] * LBW(w.WW + i * nt + t);
                g[t] = s;
                LBW(w.TT + t) = s;
            }
            lb_solve(lw, w.SS, nt, nt, w.TT);
            for (int t = 0; t < nt; ++t) h[t] = LBW(w.TT + t);
            // c = J ddq_free
            double jd = 0.0;
            for (int q = 0; q < nrec; ++q) {
                const int kq = rint[q * L].kind;
                if (kq == REC_PAD) continue;
                const double* rq = jb_smem + KP->rec_off[q] * 32 + c.lane;
                const int ndq = lb_ndof(rint, q, L), q0 = dof0[q * L];
                for (int d = 0; d < ndq; ++d) {
                    const double acc = (kq == REC_FREE) ? rq[(RF_A + d) * 32] : rq[R1_A * 32];
                    jd += (q < ntrunk ? cb[wl.JBT + k * nt + q0 + d] : cb[wl.JBL + k * nl + q0 + d]) * acc;
                }
            }
            cb[wl.CF + k] = jd;
        }
    }
    __syncwarp(c.gmask);
    // Omega: same-lane part by the owner, trunk-coupled part dealt round-robin
    for (int b = 0; b < nb; ++b) {
        if (KP->bd_owner[b] != c.sub) continue;
        const double* const cb = lw + wl.CB + wl.cb_stride * KP->bd_slot[b];
        for (int b2 = 0; b2 < nb; ++b2) {
            if (KP->bd_owner[b2] != c.sub) continue;
            const double* const cb2 = lw + wl.CB + wl.cb_stride * KP->bd_slot[b2];
            for (int k = 0; k < 6; ++k)
                for (int k2 = 0; k2 < 6; ++k2) {
                    double s = 0.0;
                    for (int i = 0; i < my_nl; ++i) s += cb[wl.JBL + k * nl + i] * cb2[wl.XB + k2 * nl + i];
                    SHW(ws.OM + (6 * b + k) * D + 6 * b2 + k2) = s;
                }
        }
    }
    __syncwarp(c.gmask);
    for (int e = c.sub; e < D * D; e += L) {
        const int i = e / D, j = e - i * D;
        double s = SHW(ws.OM + e);
        for (int t = 0; t < nt; ++t) s += SHW(ws.GS + i * nt + t) * SHW(ws.HS + j * nt + t);
        SHW(ws.OM + e) = s;
    }
    __syncwarp(c.gmask);
    // ---------------- C. per enabled contact (owner lane): lever, drift, b, diagonal of A
    for (int k = 0; k < n_cc; ++k) {
        const int o = cs_contact(k);
        if (CST(o) == 0.0) continue;
        const ContactMap cm = KP->cmap[k];
        if ((cm.trunk ? 0 : cm.sub) != c.sub) continue;
        const int b = KP->bd_of_contact[k];
        const double* const cb = lw + wl.CB + wl.cb_stride * KP->bd_slot[b];
        const double* kj = lw + w.KI + 24 * KP->bd_rec[b];
        const Xf oM = lb_load_xf(kj);
        Xf P;
#pragma unroll
        for (int e = 0; e < 9; ++e) P.R[e] = cm.placement[e];
        P.p = ld3(cm.placement + 9);
        double Rf[9];
        mat3mul(oM.R, P.R, Rf);
        const V3 r = rmul(oM.R, P.p), pf = oM.p + r;
        const Mot vLoc = motion_act_inv(P, lb_load_mot(kj + 12)), aLoc = motion_act_inv(P, lb_load_mot(kj + 18));
        const V3 vl = rmul(Rf, vLoc.l), va = rmul(Rf, vLoc.a);
        V3 dl = rmul(Rf, aLoc.l) + cross(va, vl), da = rmul(Rf, aLoc.a);
        double RrT[9], Rref[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) Rref[e] = CST(o + 5 + e);
#pragma unroll
        for (int a2 = 0; a2 < 3; ++a2)
#pragma unroll
            for (int b2 = 0; b2 < 3; ++b2) RrT[3 * a2 + b2] = Rf[3 * a2] * Rref[3 * b2] + Rf[3 * a2 + 1] * Rref[3 * b2 + 1] + Rf[3 * a2 + 2] * Rref[3 * b2 + 2];
        dl = dl + kp * (pf - mk(CST(o + 14), CST(o + 15), CST(o + 16))) + kd * vl;
        da = da + kp * cons_log3(RrT) + kd * va;
        // E c : acceleration of the contact point from the unconstrained joint accelerations
        const V3 cl = ld3(cb + wl.CF), ca = ld3(cb + wl.CF + 3);
        const V3 el = cl + cross(ca, r);
        double* const pb = sh + ws.PB + BD_PB * k;
        pb[0] = r.x; pb[1] = r.y; pb[2] = r.z;
        pb[3] = -dl.x - el.x; pb[4] = -dl.y - el.y; pb[5] = -dl.z - el.z; pb[6] = -da.z - ca.z;
        // diagonal: e_k^T Omega_bb e_k with the rows of E_i
        const double* const Ob = sh + ws.OM + (6 * b) * D + 6 * b;
        const double rows[4][6] = {{1.0, 0.0, 0.0, 0.0, r.z, -r.y}, {0.0, 1.0, 0.0, -r.z, 0.0, r.x},
                                   {0.0, 0.0, 1.0, r.y, -r.x, 0.0}, {0.0, 0.0, 0.0, 0.0, 0.0, 1.0}};
        for (int q = 0; q < 4; ++q) {
            double s = 0.0;
            for (int i = 0; i < 6; ++i) {
                double t = 0.0;
                for (int j = 0; j < 6; ++j) t += Ob[i * D + j] * rows[q][j];
                s += rows[q][i] * t;
            }
            const double reg = fmax(s * opt.constraint_regularization, CONS_MIN_REGULARIZER);
            pb[7 + q] = s + reg; pb[11 + q] = reg;
        }
        // the sweep divides by these at every update: keep the reciprocals (a double-precision division is ~200 cycles of
        // dependent instructions on the critical path of a Gauss-Seidel step): [7] 1 / max(A_xx, A_yy), [9] 1 / A_zz, [10] 1 / A_tt
        pb[7] = 1.0 / fmax(pb[7], pb[8]); pb[9] = 1.0 / pb[9]; pb[10] = 1.0 / pb[10];
    }
    // ---------------- D. warm start: private lambda, F, then a = Omega F (rows shared out)
    for (int e = 0; e < D; ++e) Fv[e] = 0.0;
    __syncwarp(c.gmask);
    for (int k = 0; k < n_cc; ++k) {
        const int o = cs_contact(k);
        if (!(enabled >> k & 1u)) continue;
        const double* const pb = sh + ws.PB + BD_PB * k;
        double* const pl = pl_all + 12 * k;
        for (int e = 0; e < 4; ++e) { pl[e] = CST(o + 1 + e); pl[4 + e] = 0.0; pl[8 + e] = 0.0; }
        const V3 f = mk(pl[0], pl[1], pl[2]), r = ld3(pb);
        const V3 tq = cross(r, f);
        double* const F = Fv + 6 * KP->bd_of_contact[k];
        F[0] += f.x; F[1] += f.y; F[2] += f.z; F[3] += tq.x; F[4] += tq.y; F[5] += tq.z + pl[3];
    }
    // torsion switched off and no torsional multiplier to clear: the second block of the sweep is a no-op
    bool skip_torsion = opt.contact_torsion < D_EPS;
    for (int k = 0; k < n_cc; ++k) if ((enabled >> k & 1u) && pl_all[12 * k + 3] != 0.0) skip_torsion = false;
    int cur = 0;
    for (int i = c.sub; i < D; i += L) {
        double s = 0.0;
        for (int j = 0; j < D; ++j) s += SHW(ws.OM + i * D + j) * Fv[j];
        AVS(0, i) = s;
    }
    __syncwarp(c.gmask);
    // ---------------- E. projected Gauss-Seidel sweep (constraint_solvers.cc:107-318), redundantly on every lane
    bool ok = false;
    for (int iter = 0; iter < CONS_PGS_MAX_ITER && !ok; ++iter) {
        const double ratio = (static_cast<double>(CONS_PGS_MAX_ITER - CONS_RELAX_MIN_ITER) - iter) /
                             (CONS_PGS_MAX_ITER - CONS_RELAX_MIN_ITER - CONS_RELAX_MAX_ITER);
        double wr = CONS_RELAX_MAX;
        if (ratio < 1.0) {
            wr = CONS_RELAX_MIN;
            if (ratio > 0.0) wr += (CONS_RELAX_MAX - CONS_RELAX_MIN) * (ratio * ratio);
        }
        for (int pass = 0; pass < 3; ++pass) {
            if (pass == 1 && skip_torsion) continue;
            for (int k = 0; k < n_cc; ++k) {
                if (!(enabled >> k & 1u)) continue;
                const double* const pb = sh + ws.PB + BD_PB * k;
                double* const pl = pl_all + 12 * k;
                const int b = (body_bits >> (2 * k)) & 3, ae = 6 * b;
                const V3 r = ld3(pb), al = mk(AVS(cur, ae), AVS(cur, ae + 1), AVS(cur, ae + 2)), aa = mk(AVS(cur, ae + 3), AVS(cur, ae + 4), AVS(cur, ae + 5));
                if (pass == 0) { pl[8] = pl[4]; pl[9] = pl[5]; pl[10] = pl[6]; pl[11] = pl[7]; }   // y_prev = y
                const V3 ea = al + cross(aa, r);   // E a, linear rows
                V3 df = mk(0.0, 0.0, 0.0);
                double d3 = 0.0;
                if (pass == 0) {                    // normal force: lambda_z >= 0
                    const double y = pb[5] - ea.z - pb[13] * pl[2];
                    pl[6] = y;
                    const double e = fmax(pl[2] + wr * y * pb[9], 0.0);
                    df.z = e - pl[2];
                    pl[2] = e;
                } else if (pass == 1) {             // torsional friction |lambda_3| <= torsion * lambda_z
                    double e;
                    if (opt.contact_torsion < D_EPS) e = pl[3] * 0.0;
                    else {
                        const double y = pb[6] - aa.z - pb[14] * pl[3];
                        pl[7] = y;
                        const double thr = opt.contact_torsion * pl[2];
                        e = fmin(fmax(pl[3] + wr * y * pb[10], -thr), thr);
                    }
                    d3 = e - pl[3];
                    pl[3] = e;
                } else {                            // Coulomb cone |(lambda_x, lambda_y)| <= friction * lambda_z
                    double e0, e1;
                    if (opt.contact_friction < D_EPS) { e0 = pl[0] * 0.0; e1 = pl[1] * 0.0; }
                    else {
                        const double y0 = pb[3] - ea.x - pb[11] * pl[0], y1 = pb[4] - ea.y - pb[12] * pl[1];
                        pl[4] = y0; pl[5] = y1;
                        const double iA_max = pb[7];
                        e0 = pl[0] + wr * y0 * iA_max;
                        e1 = pl[1] + wr * y1 * iA_max;
                        const double thr = opt.contact_friction * pl[2];
                        const double sq = e0 * e0 + e1 * e1;
                        { const double scale = sq > thr * thr ? thr * rsqrt(sq) : 1.0; e0 *= scale; e1 *= scale; }   // (thr / sqrt(sq), branch-free)
                    }
                    df.x = e0 - pl[0]; df.y = e1 - pl[1];
                    pl[0] = e0; pl[1] = e1;
                }
                if (df.x == 0.0 && df.y == 0.0 && df.z == 0.0 && d3 == 0.0) continue;   // same decision on every lane
                // dF = E^T dlambda, a += Omega[:, body b] dF : this lane's rows into the other buffer
                const V3 tq = cross(r, df);
                const double dF[6] = {df.x, df.y, df.z, tq.x, tq.y, tq.z + d3};
                double* const F = Fv + 6 * b;
#pragma unroll
                for (int e = 0; e < 6; ++e) F[e] += dF[e];
                {
                    const double* const a0 = a_base + cur * nrow * 32 + c.sub;
                    double* const a1 = a_base + (1 - cur) * nrow * 32 + c.sub;
                    const double* const Ob = sh + ws.OM + c.sub * D + 6 * b;
                    const int ostep = L * D;
#pragma unroll
                    for (int q = 0; q < 6; ++q) {      // independent rows: unrolled so that their chains overlap
                        if (c.sub + q * L < D) {
                            const double* const Oi = Ob + q * ostep;
                            a1[q * 32] = a0[q * 32] + (Oi[0] * dF[0] + Oi[1] * dF[1] + Oi[2] * dF[2] + Oi[3] * dF[3] + Oi[4] * dF[4] + Oi[5] * dF[5]);
                        }
                    }
                    for (int q = 6; c.sub + q * L < D; ++q) {
                        const double* const Oi = Ob + q * ostep;
                        a1[q * 32] = a0[q * 32] + (Oi[0] * dF[0] + Oi[1] * dF[1] + Oi[2] * dF[2] + Oi[3] * dF[3] + Oi[4] * dF[4] + Oi[5] * dF[5]);
                    }
                }
                cur = 1 - cur;
                __syncwarp(c.gmask);
            }
        }
        double ymax = 0.0;
        for (int k = 0; k < n_cc; ++k) {
            if (!(enabled >> k & 1u)) continue;
            const double* const pl = pl_all + 12 * k;
            for (int e = 4; e < 8; ++e) ymax = fmax(ymax, fabs(pl[e]));
        }
        const double tol = opt.tol_abs + opt.tol_rel * ymax + D_EPS;
        ok = true;
        for (int k = 0; k < n_cc && ok; ++k) {
            if (!(enabled >> k & 1u)) continue;
            const double* const pl = pl_all + 12 * k;
            for (int e = 0; e < 4; ++e) ok = ok && (fabs(pl[4 + e] - pl[8 + e]) < tol);
        }
    }
    // ---------------- F. accelerations: z = sum_b G_b^T F_b ; ddq_t += S^-1 z ; ddq_l += sum_b X_b F_b - W S^-1 z
    for (int t = 0; t < nt; ++t) {
        double s = 0.0;
        for (int i = 0; i < D; ++i) s += SHW(ws.GS + i * nt + t) * Fv[i];
        LBW(w.TT + t) = s;
    }
    lb_solve(lw, w.SS, nt, nt, w.TT);
    for (int q = 0; q < nrec; ++q) {
        const int kq = rint[q * L].kind;
        if (kq == REC_PAD) continue;
        double* rq = jb_smem + KP->rec_off[q] * 32 + c.lane;
        const int ndq = lb_ndof(rint, q, L), q0 = dof0[q * L];
        for (int d = 0; d < ndq; ++d) {
            const int id = q0 + d;
            double x;
            if (q < ntrunk) x = LBW(w.TT + id);
            else {
                x = 0.0;
                for (int b = 0; b < nb; ++b) {
                    if (KP->bd_owner[b] != c.sub) continue;
                    const double* const cb = lw + wl.CB + wl.cb_stride * KP->bd_slot[b];
                    for (int k = 0; k < 6; ++k) x += cb[wl.XB + k * nl + id] * Fv[6 * b + k];
                }
                for (int t = 0; t < nt; ++t) x -= LBW(w.WW + id * nt + t) * LBW(w.TT + t);
            }
            if (kq == REC_FREE) rq[(RF_A + d) * 32] += x; else rq[R1_A * 32] += x;
        }
    }
    // multipliers back into the constraints; contact wrenches in the parent joint frame (engine.cc:3790-3822)
    for (int k = 0; k < n_cc; ++k) {
        const int o = cs_contact(k);
        if (CST(o) == 0.0) continue;
        const ContactMap cm = KP->cmap[k];
        const bool mine = (cm.trunk ? 0 : cm.sub) == c.sub;
        const double* const pl = pl_all + 12 * k;
        if (mine) for (int e = 0; e < 4; ++e) CST(o + 1 + e) = pl[e];
        if (mine || cm.trunk) {
            const Xf oM = lb_load_xf(lw + w.KI + 24 * KP->jmap[cm.joint].rec);
            const V3 Fl = rtmul(oM.R, mk(pl[0], pl[1], pl[2]));
            const V3 Tl = rtmul(oM.R, mk(0.0, 0.0, pl[3]));
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
#undef SHW
#undef AVS
