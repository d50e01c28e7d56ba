// TEST INFRASTRUCTURE -- CPU oracle, not product code.  See oracle/README.md.
//
// Spatial algebra restated from Pinocchio 2.7.0 (third-party, NOT vendored in the reference; the
// reference pins it in build_tools/build_install_deps_unix.sh:227).  Conventions as Pinocchio:
// spatial vectors are [linear(3); angular(3)], SE3 M=(R,p) maps child coordinates to parent
// coordinates (`act`), `actInv` is the inverse map.  Plain scalar fp64, no SIMD intrinsics, no
// fused-multiply-add contraction assumptions (compiled with -ffp-contract=off).
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

struct V3 {
    double x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(const V3& a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(double s, const V3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(const V3& a, double s) { return {s * a.x, s * a.y, s * a.z}; }
inline V3& operator+=(V3& a, const V3& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
inline V3& operator-=(V3& a, const V3& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3& a, const V3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(const V3& a) { return std::sqrt(dot(a, a)); }

struct M3 {
    double m[9];  // row-major
    double operator()(int i, int j) const { return m[3 * i + j]; }
    double& operator()(int i, int j) { return m[3 * i + j]; }
    static M3 identity() { M3 r; std::memset(r.m, 0, sizeof r.m); r.m[0] = r.m[4] = r.m[8] = 1.0; return r; }
    static M3 zero() { M3 r; std::memset(r.m, 0, sizeof r.m); return r; }
};
inline V3 operator*(const M3& A, const V3& v) {
    return {A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
            A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
inline V3 tmul(const M3& A, const V3& v) {  // A^T v
    return {A.m[0] * v.x + A.m[3] * v.y + A.m[6] * v.z, A.m[1] * v.x + A.m[4] * v.y + A.m[7] * v.z,
            A.m[2] * v.x + A.m[5] * v.y + A.m[8] * v.z};
}
inline M3 operator*(const M3& A, const M3& B) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[3 * i + j] = A(i, 0) * B(0, j) + A(i, 1) * B(1, j) + A(i, 2) * B(2, j);
    return r;
}
inline M3 transpose(const M3& A) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[3 * i + j] = A(j, i);
    return r;
}

struct SE3 {
    M3 R;
    V3 p;
    static SE3 identity() { SE3 s; s.R = M3::identity(); s.p = V3(); return s; }
};
inline SE3 operator*(const SE3& a, const SE3& b) { SE3 r; r.R = a.R * b.R; r.p = a.p + a.R * b.p; return r; }
inline V3 act_point(const SE3& M, const V3& x) { return M.p + M.R * x; }

struct Motion {
    V3 lin, ang;
};
struct Force {
    V3 lin, ang;
};
inline Motion operator+(const Motion& a, const Motion& b) { return {a.lin + b.lin, a.ang + b.ang}; }
inline Motion& operator+=(Motion& a, const Motion& b) { a.lin += b.lin; a.ang += b.ang; return a; }
inline Force operator+(const Force& a, const Force& b) { return {a.lin + b.lin, a.ang + b.ang}; }
inline Force& operator+=(Force& a, const Force& b) { a.lin += b.lin; a.ang += b.ang; return a; }
inline Force& operator-=(Force& a, const Force& b) { a.lin -= b.lin; a.ang -= b.ang; return a; }

// SE3::act / actInv on motions and forces (pinocchio/spatial/{motion,force}-dense.hpp)
inline Motion act(const SE3& M, const Motion& m) {
    V3 w = M.R * m.ang;
    return {M.R * m.lin + cross(M.p, w), w};
}
inline Motion act_inv(const SE3& M, const Motion& m) {
    return {tmul(M.R, m.lin - cross(M.p, m.ang)), tmul(M.R, m.ang)};
}
inline Force act(const SE3& M, const Force& f) {
    V3 fl = M.R * f.lin;
    return {fl, M.R * f.ang + cross(M.p, fl)};
}
inline Force act_inv(const SE3& M, const Force& f) {
    return {tmul(M.R, f.lin), tmul(M.R, f.ang - cross(M.p, f.lin))};
}
// Motion x Motion and Motion x* Force
inline Motion cross(const Motion& a, const Motion& b) {
    return {cross(a.lin, b.ang) + cross(a.ang, b.lin), cross(a.ang, b.ang)};
}
inline Force cross(const Motion& v, const Force& f) {
    return {cross(v.ang, f.lin), cross(v.ang, f.ang) + cross(v.lin, f.lin)};
}

struct Inertia {
    double mass;
    V3 c;         // lever
    double I[6];  // Symmetric3: xx, xy, yy, xz, yz, zz
};
inline V3 sym_mul(const double I[6], const V3& w) {
    return {I[0] * w.x + I[1] * w.y + I[3] * w.z, I[1] * w.x + I[2] * w.y + I[4] * w.z,
            I[3] * w.x + I[4] * w.y + I[5] * w.z};
}
// InertiaTpl::__mult__ (pinocchio/spatial/inertia.hpp)
inline Force operator*(const Inertia& Y, const Motion& v) {
    Force f;
    f.lin = Y.mass * (v.lin - cross(Y.c, v.ang));
    f.ang = sym_mul(Y.I, v.ang) + cross(Y.c, f.lin);
    return f;
}
inline double vtiv(const Inertia& Y, const Motion& v) {  // v^T I v
    V3 cxw = cross(Y.c, v.ang);
    V3 d = v.lin - cxw;
    return Y.mass * dot(d, d) + dot(v.ang, sym_mul(Y.I, v.ang));
}

// InertiaTpl::se3Action_impl: Y expressed in the parent frame (pinocchio/spatial/inertia.hpp)
inline Inertia act(const SE3& M, const Inertia& Y) {
    Inertia r;
    r.mass = Y.mass;
    r.c = M.p + M.R * Y.c;
    // Symmetric3::rotate: R S R^T
    M3 S;
    S.m[0] = Y.I[0]; S.m[1] = Y.I[1]; S.m[2] = Y.I[3];
    S.m[3] = Y.I[1]; S.m[4] = Y.I[2]; S.m[5] = Y.I[4];
    S.m[6] = Y.I[3]; S.m[7] = Y.I[4]; S.m[8] = Y.I[5];
    const M3 T = M.R * S * transpose(M.R);
    r.I[0] = T.m[0]; r.I[1] = T.m[1]; r.I[2] = T.m[4]; r.I[3] = T.m[2]; r.I[4] = T.m[5]; r.I[5] = T.m[8];
    return r;
}
// InertiaTpl::__pequ__: composite of two rigid bodies expressed in the same frame
inline void add_inertia(Inertia& Ya, const Inertia& Yb) {
    const double mab = Ya.mass + Yb.mass;
    const double mab_inv = 1.0 / std::max(mab, 2.220446049250313e-16);
    const V3 AB = Ya.c - Yb.c;
    Ya.c = (Ya.mass * mab_inv) * Ya.c + (Yb.mass * mab_inv) * Yb.c;
    // I += Ib - (ma mb / mab) [AB]x^2,  [v]x^2 = v v^T - |v|^2 1
    const double k = Ya.mass * Yb.mass * mab_inv, n2 = dot(AB, AB);
    Ya.I[0] += Yb.I[0] - k * (AB.x * AB.x - n2);
    Ya.I[1] += Yb.I[1] - k * (AB.x * AB.y);
    Ya.I[2] += Yb.I[2] - k * (AB.y * AB.y - n2);
    Ya.I[3] += Yb.I[3] - k * (AB.x * AB.z);
    Ya.I[4] += Yb.I[4] - k * (AB.y * AB.z);
    Ya.I[5] += Yb.I[5] - k * (AB.z * AB.z - n2);
    Ya.mass = mab;
}

struct M6 {
    double m[36];  // row-major, rows/cols: [lin(3), ang(3)]
    double operator()(int i, int j) const { return m[6 * i + j]; }
    double& operator()(int i, int j) { return m[6 * i + j]; }
    static M6 zero() { M6 r; std::memset(r.m, 0, sizeof r.m); return r; }
};
inline M3 skew(const V3& v) {
    M3 r = M3::zero();
    r(0, 1) = -v.z; r(0, 2) = v.y;
    r(1, 0) = v.z;  r(1, 2) = -v.x;
    r(2, 0) = -v.y; r(2, 1) = v.x;
    return r;
}
// InertiaTpl::matrix(): [[m 1, -m[c]x], [m[c]x, I - m[c]x[c]x]]
inline M6 inertia_matrix(const Inertia& Y) {
    M6 M = M6::zero();
    for (int i = 0; i < 3; ++i) M(i, i) = Y.mass;
    M3 mc = skew(Y.mass * Y.c);
    M3 cx = skew(Y.c);
    M3 mcc = mc * cx;
    const double I[9] = {Y.I[0], Y.I[1], Y.I[3], Y.I[1], Y.I[2], Y.I[4], Y.I[3], Y.I[4], Y.I[5]};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            M(3 + i, j) = mc(i, j);
            M(i, 3 + j) = -mc(i, j);
            M(3 + i, 3 + j) = I[3 * i + j] - mcc(i, j);
        }
    return M;
}
inline void mul6(const M6& A, const double x[6], double y[6]) {
    for (int i = 0; i < 6; ++i) {
        double s = 0.0;
        for (int j = 0; j < 6; ++j) s += A(i, j) * x[j];
        y[i] = s;
    }
}
// Dual action matrix of M: force transform child -> parent, [[R, 0], [[p]x R, R]]
inline M6 dual_action_matrix(const SE3& M) {
    M6 X = M6::zero();
    M3 pR = skew(M.p) * M.R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            X(i, j) = M.R(i, j);
            X(3 + i, 3 + j) = M.R(i, j);
            X(3 + i, j) = pR(i, j);
        }
    return X;
}
// pinocchio::internal::SE3actOn: articulated inertia child -> parent, X* Ia X^{-1} = X* Ia (X*)^T
inline M6 se3_act_on(const SE3& M, const M6& Ia) {
    M6 X = dual_action_matrix(M), T = M6::zero(), R = M6::zero();
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += X(i, k) * Ia(k, j);
            T(i, j) = s;
        }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += T(i, k) * X(j, k);
            R(i, j) = s;
        }
    return R;
}
inline void to6(const Motion& m, double x[6]) { x[0] = m.lin.x; x[1] = m.lin.y; x[2] = m.lin.z; x[3] = m.ang.x; x[4] = m.ang.y; x[5] = m.ang.z; }
inline void to6(const Force& f, double x[6]) { x[0] = f.lin.x; x[1] = f.lin.y; x[2] = f.lin.z; x[3] = f.ang.x; x[4] = f.ang.y; x[5] = f.ang.z; }
inline Motion motion6(const double x[6]) { return {V3(x[0], x[1], x[2]), V3(x[3], x[4], x[5])}; }
inline Force force6(const double x[6]) { return {V3(x[0], x[1], x[2]), V3(x[3], x[4], x[5])}; }

// ---------------------------------------------------------------- quaternions / exp / log
// Eigen::Quaternion::toRotationMatrix, coefficient order (x, y, z, w) as stored in q
inline M3 quat_to_matrix(const double q[4]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    M3 R;
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz;       R(0, 2) = txz + twy;
    R(1, 0) = txy + twz;       R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy;       R(2, 1) = tyz + twx;       R(2, 2) = 1 - (txx + tyy);
    return R;
}
// Eigen rotation matrix -> quaternion (Quaternion::operator=(Matrix3), Shoemake), as used by
// pinocchio::quaternion::assignQuaternion
inline void matrix_to_quat(const M3& R, double q[4]) {
    double t = R(0, 0) + R(1, 1) + R(2, 2);
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R(2, 1) - R(1, 2)) * t;
        q[1] = (R(0, 2) - R(2, 0)) * t;
        q[2] = (R(1, 0) - R(0, 1)) * t;
    } else {
        int i = 0;
        if (R(1, 1) > R(0, 0)) i = 1;
        if (R(2, 2) > R(i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R(k, j) - R(j, k)) * t;
        q[j] = (R(j, i) + R(i, j)) * t;
        q[k] = (R(k, i) + R(i, k)) * t;
    }
}
constexpr double TAYLOR_PREC3 = 1.220703125e-4;  // eps^(1/4): TaylorSeriesExpansion<double>::precision<3>()

// pinocchio::exp6 (explog.hpp, v2.7.0)
inline SE3 exp6(const Motion& nu) {
    const V3& v = nu.lin;
    const V3& w = nu.ang;
    const double t2 = dot(w, w);
    const double t = std::sqrt(t2);
    const double st = std::sin(t), ct = std::cos(t);
    const double inv_t2 = 1.0 / t2;
    const bool small = t < TAYLOR_PREC3;
    const double alpha_wxv = small ? 0.5 - t2 / 24.0 : (1.0 - ct) * inv_t2;
    const double alpha_v = small ? 1.0 - t2 / 6.0 : st / t;
    const double alpha_w = small ? 1.0 / 6.0 - t2 / 120.0 : (1.0 - alpha_v) * inv_t2;
    const double diag = small ? 1.0 - t2 / 2.0 : ct;
    SE3 M;
    M.p = alpha_v * v + (alpha_w * dot(w, v)) * w + alpha_wxv * cross(w, v);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M.R(i, j) = alpha_wxv * w[i] * w[j];
    M.R(0, 1) -= alpha_v * w.z; M.R(1, 0) += alpha_v * w.z;
    M.R(0, 2) += alpha_v * w.y; M.R(2, 0) -= alpha_v * w.y;
    M.R(1, 2) -= alpha_v * w.x; M.R(2, 1) += alpha_v * w.x;
    M.R(0, 0) += diag; M.R(1, 1) += diag; M.R(2, 2) += diag;
    return M;
}
// pinocchio::log3 (explog.hpp), returns w and theta
inline V3 log3(const M3& R, double& theta) {
    const double PI = 3.14159265358979323846;
    const double tr = R(0, 0) + R(1, 1) + R(2, 2);
    if (tr >= 3.0) theta = 0.0;
    else if (tr <= -1.0) theta = PI;
    else theta = std::acos((tr - 1.0) / 2.0);
    V3 res;
    if (theta >= PI - 1e-2) {
        // 1e-2: A low value is not required since the computation is using explicit formulas
        const double cphi = -(tr - 1.0) / 2.0;
        const double beta = theta * theta / (1.0 + cphi);
        V3 tmp((R(0, 0) + cphi) * beta, (R(1, 1) + cphi) * beta, (R(2, 2) + cphi) * beta);
        res.x = (R(2, 1) > R(1, 2) ? 1.0 : -1.0) * (tmp.x > 0.0 ? std::sqrt(tmp.x) : 0.0);
        res.y = (R(0, 2) > R(2, 0) ? 1.0 : -1.0) * (tmp.y > 0.0 ? std::sqrt(tmp.y) : 0.0);
        res.z = (R(1, 0) > R(0, 1) ? 1.0 : -1.0) * (tmp.z > 0.0 ? std::sqrt(tmp.z) : 0.0);
    } else {
        const double t = ((theta > TAYLOR_PREC3) ? theta / std::sin(theta) : 1.0) / 2.0;
        res = V3(t * (R(2, 1) - R(1, 2)), t * (R(0, 2) - R(2, 0)), t * (R(1, 0) - R(0, 1)));
    }
    return res;
}
// pinocchio::log6 (explog.hpp)
inline Motion log6(const SE3& M) {
    double t;
    const V3 w = log3(M.R, t);
    const double t2 = t * t;
    double alpha, beta;
    if (t < TAYLOR_PREC3) {
        alpha = 1.0 - t2 / 12.0 - t2 * t2 / 720.0;
        beta = 1.0 / 12.0 + t2 / 720.0;
    } else {
        const double st = std::sin(t), ct = std::cos(t);
        alpha = t * st / (2.0 * (1.0 - ct));
        beta = 1.0 / t2 - st / (2.0 * t * (1.0 - ct));
    }
    Motion m;
    m.lin = alpha * M.p - 0.5 * cross(w, M.p) + (beta * dot(w, M.p)) * w;
    m.ang = w;
    return m;
}

// ---------------------------------------------------------------- unit-quaternion Lie group (JointModelSpherical)
// Pinocchio 2.7.0 `explog-quaternion.hpp` and `SpecialOrthogonalOperationTpl<3>` (liegroup/special-orthogonal.hpp),
// restated from the published algorithm (the library is not vendored in the reference); quaternions as (x, y, z, w).
constexpr double TAYLOR_PREC2 = 6.0554544523933395e-6;  // eps^(1/3): TaylorSeriesExpansion<double>::precision<2>()
constexpr double DBL_EPS_ = 2.220446049250313e-16;

inline void quat_mul(const double a[4], const double b[4], double out[4]) {   // Eigen: a * b
    const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    out[0] = x; out[1] = y; out[2] = z; out[3] = w;
}
// quaternion::exp3
inline void quat_exp3(const V3& v, double out[4]) {
    const double t2 = dot(v, v);
    const double t = std::sqrt(t2 + DBL_EPS_ * DBL_EPS_);
    if (t2 > TAYLOR_PREC3 * TAYLOR_PREC3) {
        // Eigen::Quaternion(AngleAxis(t, v / t))
        const double s = std::sin(0.5 * t), c = std::cos(0.5 * t);
        out[0] = s * (v.x / t); out[1] = s * (v.y / t); out[2] = s * (v.z / t); out[3] = c;
    } else {
        const double t2_2 = t2 / 4.0;
        const double k = 0.5 * (1.0 - t2_2 / 6.0 + t2_2 * t2_2 / 120.0);
        out[0] = k * v.x; out[1] = k * v.y; out[2] = k * v.z;
        out[3] = 1.0 - t2_2 / 2.0 + t2_2 * t2_2 / 24.0;
    }
}
// quaternion::log3: angle-axis vector of a unit quaternion, theta in [0, pi]... [0, 2 pi) before the sign flip
inline V3 quat_log3(const double q[4], double& theta) {
    const double norm_squared = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    const double norm = std::sqrt(norm_squared + DBL_EPS_ * DBL_EPS_);
    const double pos_neg = q[3] >= 0.0 ? 1.0 : -1.0;
    const double w = pos_neg * q[3];
    const V3 vec(pos_neg * q[0], pos_neg * q[1], pos_neg * q[2]);
    const double theta_2 = std::atan2(norm, w);
    const double y_x = norm / w;
    const double y_x_sq = norm_squared / (w * w);
    const bool small = norm_squared < TAYLOR_PREC2;
    theta = small ? 2.0 * (1.0 - y_x_sq / 3.0) * y_x : 2.0 * theta_2;
    const double th2_2 = theta * theta / 4.0;
    const double inv_sinc = small ? 2.0 * (1.0 + th2_2 / 6.0 + 7.0 / 360.0 * th2_2 * th2_2) : theta / std::sin(theta_2);
    return inv_sinc * vec;
}
// pinocchio::Jlog3(theta, log, Jlog)
inline M3 Jlog3(double theta, const V3& lg) {
    const double st = std::sin(theta), ct = std::cos(theta);
    const double st_1mct = st / (1.0 - ct);
    const bool small = theta < TAYLOR_PREC3;
    const double alpha = small ? 1.0 / 12.0 + theta * theta / 720.0 : 1.0 / (theta * theta) - st_1mct / (2.0 * theta);
    const double diag = small ? 0.5 * (2.0 - theta * theta / 6.0) : 0.5 * (theta * st_1mct);
    M3 J;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) J(i, j) = alpha * lg[i] * lg[j];
    J(0, 0) += diag; J(1, 1) += diag; J(2, 2) += diag;
    const V3 h = 0.5 * lg;   // addSkew(0.5 * log, Jlog)
    J(0, 1) -= h.z; J(0, 2) += h.y; J(1, 0) += h.z; J(1, 2) -= h.x; J(2, 0) -= h.y; J(2, 1) += h.x;
    return J;
}

}  // namespace orc
