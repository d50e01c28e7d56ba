// TEST INFRASTRUCTURE -- CPU oracle, not product code.  See oracle/README.md.
//
// The constraint path of the jiminy step: what Engine::computeAcceleration does once a kinematic
// constraint is enabled (a joint left its position bounds, or contacts.model == "constraint" and a
// contact frame touches the ground).  Restated function by function from the reference:
//   Engine::computeAcceleration ............ core/src/engine/engine.cc:3709-3866
//   Model::computeConstraints .............. core/src/robot/model.cc:1238-1287
//   pinocchio_overload::crba ............... core/include/jiminy/core/robot/pinocchio_overload_algorithms.h:99-124
//   computeJMinvJt / solveJMinvJtv ......... same file :491-551
//   JointConstraint ........................ core/src/constraints/joint_constraint.cc:55-164
//   FrameConstraint ........................ core/src/constraints/frame_constraint.cc:71-183
//   PGSSolver .............................. core/src/solver/constraint_solvers.cc:107-448
//   computePositionLimitsForcesAlgo ........ core/src/engine/engine.cc:3253-3338
//   computeContactDynamicsAtFrame .......... core/src/engine/engine.cc:3117-3195 (constraint branch)
// Pinocchio 2.7 primitives used by those (crba, computeJointJacobians, nonLinearEffects,
// cholesky::decompose/solve, getFrameVelocity/Acceleration, log3) are restated from their published
// algorithms in dense form: the inertia matrix is factored with a plain dense LL^T instead of
// Pinocchio's sparse U D U^T, which is the same mathematics in a different summation order.
#include <cmath>
#include <cstring>

#include "engine.hpp"

namespace orc {

namespace {
constexpr double MIN_REGULARIZER = 1.0e-11;       // constraint_solvers.cc:15
constexpr double RELAX_MIN = 0.01, RELAX_MAX = 1.0;
constexpr uint32_t RELAX_MIN_ITER_NUM = 20, RELAX_MAX_ITER_NUM = 30;
constexpr double RELAX_SLOPE_ORDER = 2.0;
constexpr uint32_t PGS_MAX_ITERATIONS = 100;      // engine.cc:62

// dense LL^T, in place in the lower triangle; returns false if not positive definite
bool llt(int n, std::vector<double>& A) {
    for (int j = 0; j < n; ++j) {
        double s = A[j * n + j];
        for (int k = 0; k < j; ++k) s -= A[j * n + k] * A[j * n + k];
        if (!(s > 0.0)) return false;
        const double d = std::sqrt(s);
        A[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double t = A[i * n + j];
            for (int k = 0; k < j; ++k) t -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = t / d;
        }
    }
    return true;
}
void llt_forward(int n, const std::vector<double>& L, double* x) {   // L y = x
    for (int i = 0; i < n; ++i) {
        double s = x[i];
        for (int k = 0; k < i; ++k) s -= L[i * n + k] * x[k];
        x[i] = s / L[i * n + i];
    }
}
void llt_backward(int n, const std::vector<double>& L, double* x) {  // L^T y = x
    for (int i = n - 1; i >= 0; --i) {
        double s = x[i];
        for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * x[k];
        x[i] = s / L[i * n + i];
    }
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// registry: one JointConstraint per mechanical joint (model.cc:335-360), then one FrameConstraint
// {x, y, z, rot z} per contact frame (model.cc:813-822) -- the iteration order of
// ConstraintTree::foreach (model.h:43-46), which the PGS sweep follows.
void Engine::buildConstraints() {
    constraints.clear();
    for (int i = 1; i < model.njoints; ++i) {
        if (model.jtype[i] == JB_JOINT_FREEFLYER) continue;   // 'root_joint' is not a mechanical joint
        if (model.jtype[i] == JB_JOINT_SPHERICAL) continue;   // flexibility joints carry no bound constraint (model.cc:335-360)
        Constraint c;
        c.kind = 0; c.joint = i; c.dim = 1;
        c.jac.assign(model.nv, 0.0);
        constraints.push_back(c);
    }
    for (int k = 0; k < model.ncontacts; ++k) {
        Constraint c;
        c.kind = 1; c.joint = model.contact_joint[k]; c.contact = k; c.dim = 4;
        c.jac.assign(4 * static_cast<size_t>(model.nv), 0.0);
        constraints.push_back(c);
    }
    rowsMax = 0;
    for (const Constraint& c : constraints) rowsMax += c.dim;
    solverJ.assign(static_cast<size_t>(rowsMax) * model.nv, 0.0);
    solverGamma.assign(rowsMax, 0.0); solverLambda.assign(rowsMax, 0.0);
    solverB.assign(rowsMax, 0.0); solverY.assign(rowsMax, 0.0); solverYPrev.assign(rowsMax, 0.0);
    Mmat.assign(static_cast<size_t>(model.nv) * model.nv, 0.0); Mchol = Mmat;
    Jworld.assign(6 * static_cast<size_t>(model.nv), 0.0);
    nle.assign(model.nv, 0.0); torqueResidual.assign(model.nv, 0.0);
    aDrift.assign(model.njoints, Motion{});
}

bool Engine::hasConstraints() const {   // Model::hasConstraints (model.cc:1009-1024)
    for (const Constraint& c : constraints) if (c.enabled) return true;
    return false;
}

static void set_normal(Engine::Constraint& c, const V3& n) {   // FrameConstraint::setNormal (frame_constraint.cc:61-67)
    c.normal = n;
    V3 c1 = cross(n, V3(1.0, 0.0, 0.0));
    c1 = (1.0 / norm(c1)) * c1;
    const V3 c0 = cross(c1, n);
    for (int r = 0; r < 3; ++r) { c.rotationLocal(r, 0) = c0[r]; c.rotationLocal(r, 1) = c1[r]; c.rotationLocal(r, 2) = n[r]; }
}

// Model::resetConstraints (model.cc:1026-1045) + the start-time configuration of engine.cc:1268-1309
void Engine::resetConstraints(const double* qv) {
    const double omega = 2.0 * M_PI * opt.contact_stabilization_freq;   // setBaumgarteFreq (abstract_constraint.cc:88-99)
    for (Constraint& c : constraints) {
        std::fill(c.jac.begin(), c.jac.end(), 0.0);
        for (double& x : c.drift) x = 0.0;
        for (double& x : c.lambda) x = 0.0;
        if (c.kind == 0) {
            c.jac[model.idx_v[c.joint]] = c.reversed ? -1.0 : 1.0;
            c.qRef = qv[model.idx_q[c.joint]];
        } else {
            c.transformRef = data.oMi[c.joint] * model.contact_placement[c.contact];
            c.rotationLocal = M3::identity();
        }
        c.enabled = false;
        c.kp = omega * omega; c.kd = 2.0 * omega;
        if (opt.contact_model == JB_CONTACT_CONSTRAINT) {
            if (c.kind == 0 && c.reversed) {   // setRotationDir(false)
                for (double& x : c.jac) x = -x;
                c.reversed = false;
            }
            c.enabled = true;
        }
    }
    successiveSolveFailed = 0;
    std::fill(solverLambda.begin(), solverLambda.end(), 0.0);
}

// computePositionLimitsForcesAlgo (engine.cc:3253-3338) for every bounded-joint constraint
void Engine::updateJointBoundConstraints(const double* qv) {
    for (Constraint& c : constraints) {
        if (c.kind != 0) continue;
        const int t = model.jtype[c.joint];
        if (Model::is_unbounded(t)) { c.enabled = false; for (double& x : c.lambda) x = 0.0; continue; }
        const int iq = model.idx_q[c.joint];
        const double qJoint = qv[iq], qMin = model.q_lower[iq], qMax = model.q_upper[iq];
        const double eps = opt.contact_transition_eps;
        if (qMax < qJoint || qJoint < qMin) {
            c.qRef = std::min(std::max(qJoint, qMin), qMax);
            const bool rev = qMax < qJoint;
            if (rev != c.reversed) { for (double& x : c.jac) x = -x; c.reversed = rev; }
            c.enabled = true;
            status |= JB_ENV_JOINT_LIMIT;
        } else if (qMin + eps < qJoint && qJoint < qMax - eps) {
            c.enabled = false;
            for (double& x : c.lambda) x = 0.0;
        }
    }
}

// computeContactDynamicsAtFrame, contacts.model == "constraint" (engine.cc:3133-3194)
void Engine::updateContactConstraint(int contact) {
    Constraint* cp = nullptr;
    for (Constraint& c : constraints) if (c.kind == 1 && c.contact == contact) cp = &c;
    Constraint& c = *cp;
    const SE3 oMf = data.oMi[c.joint] * model.contact_placement[contact];
    const double heightGround = 0.0;
    V3 normalGround(0.0, 0.0, 1.0);
    normalGround = (1.0 / norm(normalGround)) * normalGround;
    const double depth = (oMf.p.z - heightGround) * normalGround.z;
    if (depth < 0.0) c.enabled = true;
    else if (depth > opt.contact_transition_eps) { c.enabled = false; for (double& x : c.lambda) x = 0.0; }
    if (c.enabled) {
        c.transformRef.R = oMf.R;
        c.transformRef.p = oMf.p - depth * normalGround;
        set_normal(c, normalGround);
    }
}

// pinocchio_overload::crba (overload.h:99-124): joint-space inertia with rotor inertia on the diagonal
// (composite-rigid-body algorithm) + pinocchio::computeJointJacobians (world-frame columns).
void Engine::computeCrba() {
    const int n = model.njoints, nv = model.nv;
    std::fill(Mmat.begin(), Mmat.end(), 0.0);
    std::vector<M6> Ycrb(n);
    for (int i = 1; i < n; ++i) Ycrb[i] = inertia_matrix(model.inertia[i]);
    for (int i = n - 1; i > 0; --i) {
        const JointData& jd = data.joints[i];
        const int nvj = Model::nvj(model.jtype[i]), iv = model.idx_v[i];
        for (int c = 0; c < nvj; ++c) {
            double s6[6], f6[6];
            for (int r = 0; r < 6; ++r) s6[r] = jd.S[r][c];
            mul6(Ycrb[i], s6, f6);   // F = Ycrb S_c, a spatial force in the frame of joint i
            for (int r = 0; r < nvj; ++r) {
                double s = 0.0;
                for (int k = 0; k < 6; ++k) s += jd.S[k][r] * f6[k];
                Mmat[(iv + r) * nv + iv + c] = s;
            }
            Force F = force6(f6);
            int j = i;
            while (model.parent[j] > 0) {
                F = act(data.liMi[j], F);
                j = model.parent[j];
                const JointData& jp = data.joints[j];
                const int nvp = Model::nvj(model.jtype[j]), ivp = model.idx_v[j];
                double g6[6];
                to6(F, g6);
                for (int r = 0; r < nvp; ++r) {
                    double s = 0.0;
                    for (int k = 0; k < 6; ++k) s += jp.S[k][r] * g6[k];
                    Mmat[(ivp + r) * nv + iv + c] = s;
                    Mmat[(iv + c) * nv + ivp + r] = s;
                }
            }
        }
        const int p = model.parent[i];
        if (p > 0) {
            const M6 T = se3_act_on(data.liMi[i], Ycrb[i]);
            for (int k = 0; k < 36; ++k) Ycrb[p].m[k] += T.m[k];
        }
    }
    for (int k = 0; k < nv; ++k) Mmat[k * nv + k] += model.rotor[k];
    for (int i = 1; i < n; ++i) {
        const JointData& jd = data.joints[i];
        const int nvj = Model::nvj(model.jtype[i]), iv = model.idx_v[i];
        for (int c = 0; c < nvj; ++c) {
            double s6[6];
            for (int r = 0; r < 6; ++r) s6[r] = jd.S[r][c];
            double w6[6];
            to6(act(data.oMi[i], motion6(s6)), w6);
            for (int r = 0; r < 6; ++r) Jworld[r * nv + iv + c] = w6[r];
        }
    }
}

// pinocchio::nonLinearEffects = rnea(q, v, 0) with gravity
void Engine::computeNle() {
    const int n = model.njoints;
    std::vector<Motion> ag(n);
    std::vector<Force> f(n);
    ag[0] = Motion{V3(-opt.gravity[0], -opt.gravity[1], -opt.gravity[2]), V3(-opt.gravity[3], -opt.gravity[4], -opt.gravity[5])};
    for (int i = 1; i < n; ++i) {
        ag[i] = cross(data.v[i], data.joints[i].vJ) + act_inv(data.liMi[i], ag[model.parent[i]]);
        f[i] = model.inertia[i] * ag[i] + cross(data.v[i], model.inertia[i] * data.v[i]);
    }
    for (int i = n - 1; i > 0; --i) {
        const JointData& jd = data.joints[i];
        const int nvj = Model::nvj(model.jtype[i]), iv = model.idx_v[i];
        double f6[6];
        to6(f[i], f6);
        for (int c = 0; c < nvj; ++c) {
            double s = 0.0;
            for (int r = 0; r < 6; ++r) s += jd.S[r][c] * f6[r];
            nle[iv + c] = s;
        }
        if (model.parent[i] > 0) f[model.parent[i]] += act(data.liMi[i], f[i]);
    }
}

// Model::computeConstraints (model.cc:1238-1287)
void Engine::computeConstraints(const double* qv, const double* vv) {
    if (!hasConstraints()) return;
    computeCrba();
    // joint spatial accelerations with ddq = 0 and no gravity: the constraint drift
    aDrift[0] = Motion{};
    for (int i = 1; i < model.njoints; ++i) {
        aDrift[i] = cross(data.v[i], data.joints[i].vJ);
        if (model.parent[i] > 0) aDrift[i] += act_inv(data.liMi[i], aDrift[model.parent[i]]);
    }
    const int nv = model.nv;
    for (Constraint& c : constraints) {
        if (!c.enabled) continue;
        if (c.kind == 0) {
            // JointConstraint::computeJacobianAndDrift (joint_constraint.cc:141-163)
            const double deltaPosition = qv[model.idx_q[c.joint]] - c.qRef;
            c.drift[0] = c.kp * deltaPosition + c.kd * vv[model.idx_v[c.joint]];
            if (c.reversed) c.drift[0] *= -1.0;
            continue;
        }
        // FrameConstraint::computeJacobianAndDrift (frame_constraint.cc:103-183)
        const SE3& P = model.contact_placement[c.contact];
        const SE3 framePose = data.oMi[c.joint] * P;
        SE3 transformLocal;
        transformLocal.R = c.rotationLocal; transformLocal.p = framePose.p;
        double frameJac[6][64];
        std::vector<int> support;
        for (int j = c.joint; j > 0; j = model.parent[j])
            for (int k = 0; k < Model::nvj(model.jtype[j]); ++k) support.push_back(model.idx_v[j] + k);
        for (int col : support) {
            double w6[6];
            for (int r = 0; r < 6; ++r) w6[r] = Jworld[r * nv + col];
            double o6[6];
            to6(act_inv(transformLocal, motion6(w6)), o6);
            for (int r = 0; r < 6; ++r) frameJac[r][col] = o6[r];
        }
        const V3 deltaPosition = framePose.p - c.transformRef.p;
        double theta;
        const V3 deltaRotation = log3(framePose.R * transpose(c.transformRef.R), theta);
        // frame velocity / "drift" acceleration, LOCAL_WORLD_ALIGNED
        const Motion vLoc = act_inv(P, data.v[c.joint]);
        const Motion velocity{framePose.R * vLoc.lin, framePose.R * vLoc.ang};
        const Motion aLoc = act_inv(P, aDrift[c.joint]);
        Motion frameDrift{framePose.R * aLoc.lin, framePose.R * aLoc.ang};
        frameDrift.lin += cross(velocity.ang, velocity.lin);   // classical acceleration
        frameDrift.lin += c.kp * deltaPosition;
        frameDrift.ang += c.kp * deltaRotation;
        frameDrift.lin += c.kd * velocity.lin;
        frameDrift.ang += c.kd * velocity.ang;
        frameDrift.lin = tmul(c.rotationLocal, frameDrift.lin);
        frameDrift.ang = tmul(c.rotationLocal, frameDrift.ang);
        double d6[6];
        to6(frameDrift, d6);
        static const int dofsFixed[4] = {0, 1, 2, 5};
        for (int i = 0; i < 4; ++i) {
            for (int col : support) c.jac[static_cast<size_t>(i) * nv + col] = frameJac[dofsFixed[i]][col];
            c.drift[i] = d6[dofsFixed[i]];
        }
    }
}

// PGSSolver::ProjectedGaussSeidelIter (constraint_solvers.cc:107-221).  A is m x m (row-major, full).
void Engine::pgsIter(int m, const std::vector<double>& A, const double* b, double w, double* x) {
    struct Block { double lo, hi; bool isZero; int fIndex[3]; int fSize; };
    for (int i = 0; i < 3; ++i) {
        for (const Constraint& c : constraints) {
            if (!c.enabled) continue;
            const int nBlocks = c.kind == 0 ? 1 : 3;
            if (nBlocks <= i) continue;
            Block blk{};
            if (c.kind == 0) blk = Block{0.0, INF, false, {0, 0, 0}, 1};
            else if (i == 0) blk = Block{0.0, INF, false, {2, 0, 0}, 1};
            else if (i == 1) blk = Block{0.0, opt.contact_torsion, opt.contact_torsion < EPS, {3, 2, 0}, 2};
            else blk = Block{0.0, opt.contact_friction, opt.contact_friction < EPS, {0, 1, 2}, 3};
            const int o = c.startIndex;
            const int i0 = o + blk.fIndex[0];
            double& e = x[i0];
            if (blk.isZero) {
                e *= 0;
                for (int j = 1; j < blk.fSize - 1; ++j) x[o + blk.fIndex[j]] *= 0;
                continue;
            }
            auto residual = [&](int k) {
                double s = 0.0;
                for (int r = 0; r < m; ++r) s += A[r * m + k] * x[r];   // A.col(k).dot(x)
                return b[k] - s;
            };
            double A_max = A[i0 * m + i0];
            solverY[i0] = residual(i0);
            for (int j = 1; j < blk.fSize - 1; ++j) {
                const int k = o + blk.fIndex[j];
                solverY[k] = residual(k);
                A_max = std::max(A_max, A[k * m + k]);
            }
            e += w * solverY[i0] / A_max;
            for (int j = 1; j < blk.fSize - 1; ++j) {
                const int k = o + blk.fIndex[j];
                x[k] += w * solverY[k] / A_max;
            }
            if (blk.fSize == 1) e = std::min(std::max(e, blk.lo), blk.hi);
            else {
                const double thr = blk.hi * x[o + blk.fIndex[blk.fSize - 1]];
                if (blk.fSize == 2) e = std::min(std::max(e, -thr), thr);
                else {
                    double squaredNorm = e * e;
                    for (int j = 1; j < blk.fSize - 1; ++j) { const double f = x[o + blk.fIndex[j]]; squaredNorm += f * f; }
                    if (squaredNorm > thr * thr) {
                        const double scale = thr / std::sqrt(squaredNorm);
                        e *= scale;
                        for (int j = 1; j < blk.fSize - 1; ++j) x[o + blk.fIndex[j]] *= scale;
                    }
                }
            }
        }
    }
}

// PGSSolver::ProjectedGaussSeidelSolver (constraint_solvers.cc:223-318)
bool Engine::pgsSolve(int m, const std::vector<double>& A, const double* b, double* x) {
    std::fill(solverY.begin(), solverY.end(), 0.0);
    for (uint32_t iter = 0; iter < PGS_MAX_ITERATIONS; ++iter) {
        solverYPrev = solverY;
        const double ratio = (static_cast<double>(PGS_MAX_ITERATIONS - RELAX_MIN_ITER_NUM) - iter) /
                             (PGS_MAX_ITERATIONS - RELAX_MIN_ITER_NUM - RELAX_MAX_ITER_NUM);
        double w = RELAX_MAX;
        if (ratio < 1.0) {
            w = RELAX_MIN;
            if (ratio > 0.0) w += (RELAX_MAX - RELAX_MIN) * std::pow(ratio, RELAX_SLOPE_ORDER);
        }
        pgsIter(m, A, b, w, x);
        double ymax = 0.0;
        for (double y : solverY) ymax = std::max(ymax, std::fabs(y));
        const double tol = opt.tol_abs + opt.tol_rel * ymax + EPS;
        bool converged = true;
        for (int k = 0; k < rowsMax; ++k) if (!(std::fabs(solverY[k] - solverYPrev[k]) < tol)) { converged = false; break; }
        ++pgsIterations;
        if (converged) { if (keepPgsHistory) pgsHistory.push_back(static_cast<int32_t>(iter + 1)); return true; }
    }
    if (keepPgsHistory) pgsHistory.push_back(static_cast<int32_t>(PGS_MAX_ITERATIONS));
    return false;
}

// PGSSolver::SolveBoxedForwardDynamics (constraint_solvers.cc:320-447)
bool Engine::solveBoxedForwardDynamics(double dampingInv, bool isStateUpToDate, bool ignoreBounds) {
    const int nv = model.nv;
    int m = 0;
    for (Constraint& c : constraints) {
        if (!c.enabled) continue;
        if (!isStateUpToDate) {
            std::memcpy(&solverJ[static_cast<size_t>(m) * nv], c.jac.data(), sizeof(double) * c.dim * nv);
            for (int k = 0; k < c.dim; ++k) { solverGamma[m + k] = c.drift[k]; solverLambda[m + k] = c.lambda[k]; }
        }
        c.startIndex = m;
        m += c.dim;
    }
    if (!isStateUpToDate) {
        // computeJMinvJt (overload.h:491-537): M = L L^T, Y = L^-1 J^T, A = Y^T Y
        Mchol = Mmat;
        if (!llt(nv, Mchol)) { status |= JB_ENV_NAN; }
        std::vector<double> Y(static_cast<size_t>(nv) * m);
        std::vector<double> col(nv);
        for (int r = 0; r < m; ++r) {
            for (int k = 0; k < nv; ++k) col[k] = solverJ[static_cast<size_t>(r) * nv + k];
            llt_forward(nv, Mchol, col.data());
            for (int k = 0; k < nv; ++k) Y[static_cast<size_t>(k) * m + r] = col[k];
        }
        solverA.assign(static_cast<size_t>(m) * m, 0.0);
        for (int r = 0; r < m; ++r)
            for (int c2 = 0; c2 <= r; ++c2) {
                double s = 0.0;
                for (int k = 0; k < nv; ++k) s += Y[static_cast<size_t>(k) * m + r] * Y[static_cast<size_t>(k) * m + c2];
                solverA[r * m + c2] = s; solverA[c2 * m + r] = s;
            }
        for (int r = 0; r < m; ++r) solverA[r * m + r] += std::max(solverA[r * m + r] * dampingInv, MIN_REGULARIZER);
    }
    // dynamic drift: torque_residual = M^-1 (u - nle)
    for (int k = 0; k < nv; ++k) torqueResidual[k] = data.u[k] - nle[k];
    llt_forward(nv, Mchol, torqueResidual.data());
    llt_backward(nv, Mchol, torqueResidual.data());
    for (int r = 0; r < m; ++r) {
        double s = 0.0;
        for (int k = 0; k < nv; ++k) s += solverJ[static_cast<size_t>(r) * nv + k] * torqueResidual[k];
        solverB[r] = -solverGamma[r] - s;
    }
    bool isSuccess = false;
    if (ignoreBounds) {
        // solveJMinvJtv (overload.h:539-551): lambda = (J M^-1 J^T)^-1 b, dense LL^T
        std::vector<double> L = solverA;
        if (llt(m, L)) {
            for (int r = 0; r < m; ++r) solverLambda[r] = solverB[r];
            llt_forward(m, L, solverLambda.data());
            llt_backward(m, L, solverLambda.data());
        }
        isSuccess = true;
    } else {
        isSuccess = pgsSolve(m, solverA, solverB.data(), solverLambda.data());
    }
    int row = 0;
    for (Constraint& c : constraints) {
        if (!c.enabled) continue;
        for (int k = 0; k < c.dim; ++k) c.lambda[k] = solverLambda[row + k];
        row += c.dim;
    }
    // ddq = M^-1 J^T lambda + torque_residual
    std::vector<double> rhs(nv, 0.0);
    for (int r = 0; r < m; ++r)
        for (int k = 0; k < nv; ++k) rhs[k] += solverJ[static_cast<size_t>(r) * nv + k] * solverLambda[r];
    llt_forward(nv, Mchol, rhs.data());
    llt_backward(nv, Mchol, rhs.data());
    for (int k = 0; k < nv; ++k) data.ddq[k] = rhs[k] + torqueResidual[k];
    return isSuccess;
}

// Engine::computeAcceleration (engine.cc:3709-3866)
const std::vector<double>& Engine::computeAcceleration(const double* qv, const double* vv, std::vector<double>& u,
                                                      std::vector<Force>& fext, bool isStateUpToDate, bool ignoreBounds) {
    if (!hasConstraints()) return aba(qv, vv, u, fext);
    const int nv = model.nv;
    if (!isStateUpToDate) {
        computeConstraints(qv, vv);
        computeNle();
    }
    // project the external forces to joint space: data.u = u + sum_i J_i^T fext_i (LOCAL joint jacobians)
    data.u = u;
    for (int i = 1; i < model.njoints; ++i) {
        double f6[6];
        to6(fext[i], f6);
        bool any = false;
        for (double x : f6) any = any || std::fabs(x) > EPS;
        if (!any) continue;
        for (int j = i; j > 0; j = model.parent[j])
            for (int k = 0; k < Model::nvj(model.jtype[j]); ++k) {
                const int col = model.idx_v[j] + k;
                double w6[6];
                for (int r = 0; r < 6; ++r) w6[r] = Jworld[r * nv + col];
                double l6[6];
                to6(act_inv(data.oMi[i], motion6(w6)), l6);
                double s = 0.0;
                for (int r = 0; r < 6; ++r) s += l6[r] * f6[r];
                data.u[col] += s;
            }
    }
    const bool isSuccess = solveBoxedForwardDynamics(opt.constraint_regularization, isStateUpToDate, ignoreBounds);
    if (isSuccess) successiveSolveFailed = 0; else ++successiveSolveFailed;
    // restore the bounds efforts (engine.cc:3770-3788; adds lambda itself, whatever the constraint direction)
    for (const Constraint& c : constraints) {
        if (c.kind != 0 || !c.enabled) continue;
        const int iv = model.idx_v[c.joint];
        state.uInternal[iv] += c.lambda[0];
        u[iv] += c.lambda[0];
    }
    // contact forces from the multipliers (engine.cc:3790-3822)
    for (const Constraint& c : constraints) {
        if (c.kind != 1 || !c.enabled) continue;
        const V3 linLocal(c.lambda[0], c.lambda[1], c.lambda[2]), angLocal(0.0, 0.0, c.lambda[3]);
        const Force fWorld{c.rotationLocal * linLocal, c.rotationLocal * angLocal};
        const SE3 oMf = data.oMi[c.joint] * model.contact_placement[c.contact];
        contactForces[c.contact].lin = tmul(oMf.R, fWorld.lin);
        contactForces[c.contact].ang = tmul(oMf.R, fWorld.ang);
        fext[c.joint] += convertForceGlobalFrameToJoint(c.joint, model.contact_placement[c.contact].p, fWorld);
    }
    return data.ddq;
}

}  // namespace orc
