// TEST INFRASTRUCTURE -- CPU oracle, not product code.  See oracle/README.md and engine.hpp.
#include "engine.hpp"

#include <cassert>
#include <cmath>
#include <stdexcept>

namespace orc {

// ============================================================================ model
Model make_model(const JbModelDesc& d) {
    Model m;
    m.njoints = d.njoints; m.nq = d.nq; m.nv = d.nv;
    m.jtype.assign(d.joint_type, d.joint_type + d.njoints);
    m.parent.assign(d.parent, d.parent + d.njoints);
    m.idx_q.assign(d.idx_q, d.idx_q + d.njoints);
    m.idx_v.assign(d.idx_v, d.idx_v + d.njoints);
    for (int i = 0; i < d.njoints; ++i) {
        m.placement.push_back(se3_from12(d.placement + 12 * i));
        m.axis.push_back(V3(d.axis[3 * i], d.axis[3 * i + 1], d.axis[3 * i + 2]));
        Inertia Y;
        const double* y = d.inertia + 10 * i;
        Y.mass = y[0]; Y.c = V3(y[1], y[2], y[3]);
        for (int k = 0; k < 6; ++k) Y.I[k] = y[4 + k];
        m.inertia.push_back(Y);
    }
    m.rotor.assign(d.rotor_inertia, d.rotor_inertia + d.nv);
    m.q_lower.assign(d.q_lower, d.q_lower + d.nq);
    m.q_upper.assign(d.q_upper, d.q_upper + d.nq);
    if (d.flexibility) m.flexibility.assign(d.flexibility, d.flexibility + 6 * d.njoints);
    m.nmotors = d.nmotors;
    m.motor_joint.assign(d.motor_joint, d.motor_joint + d.nmotors);
    m.motor_flags.assign(d.motor_flags, d.motor_flags + d.nmotors);
    m.motor_params.assign(d.motor_params, d.motor_params + 10 * d.nmotors);
    m.ncontacts = d.ncontacts;
    m.contact_joint.assign(d.contact_joint, d.contact_joint + d.ncontacts);
    for (int i = 0; i < d.ncontacts; ++i) m.contact_placement.push_back(se3_from12(d.contact_placement + 12 * i));
    m.nimu = d.nimu; m.nforce = d.nforce; m.nenc = d.nencoder; m.neff = d.neffort; m.ncs = d.ncontact_sensor;
    m.imu_joint.assign(d.imu_joint, d.imu_joint + d.nimu);
    for (int i = 0; i < d.nimu; ++i) m.imu_placement.push_back(se3_from12(d.imu_placement + 12 * i));
    m.force_joint.assign(d.force_joint, d.force_joint + d.nforce);
    for (int i = 0; i < d.nforce; ++i) m.force_placement.push_back(se3_from12(d.force_placement + 12 * i));
    m.enc_joint.assign(d.encoder_joint, d.encoder_joint + d.nencoder);
    m.enc_reduction.assign(d.encoder_reduction, d.encoder_reduction + d.nencoder);
    m.eff_motor.assign(d.effort_motor, d.effort_motor + d.neffort);
    m.cs_index.assign(d.contact_sensor_index, d.contact_sensor_index + d.ncontact_sensor);
    // ForceSensor::refreshProxies (basic_sensors.cc:324-351)
    m.force_contacts.resize(m.nforce);
    for (int s = 0; s < m.nforce; ++s)
        for (int c = 0; c < m.ncontacts; ++c)
            if (m.contact_joint[c] == m.force_joint[s])
                m.force_contacts[s].emplace_back(c, se3_inverse(m.force_placement[s]) * m.contact_placement[c]);
    JbSensorLayout& L = m.layout;
    L.imu_offset = 0;
    L.force_offset = L.imu_offset + 6 * m.nimu;
    L.encoder_offset = L.force_offset + 6 * m.nforce;
    L.effort_offset = L.encoder_offset + 2 * m.nenc;
    L.contact_offset = L.effort_offset + m.neff;
    L.width = L.contact_offset + 3 * m.ncs;
    return m;
}

Engine::Engine(const JbModelDesc& d, const JbOptions& o) : model(make_model(d)) {
    const int n = model.njoints, nv = model.nv, nq = model.nq;
    data.liMi.assign(n, SE3::identity()); data.oMi.assign(n, SE3::identity());
    data.v.assign(n, Motion{}); data.a.assign(n, Motion{}); data.a_gf.assign(n, Motion{});
    data.f.assign(n, Force{}); data.h.assign(n, Force{});
    data.Ycrb.assign(n, Inertia{}); data.com.assign(n, V3()); data.vcom.assign(n, V3());
    // subtree masses: data.mass as left by pinocchio::centerOfMass(model, data, qNeutral) (model.cc:269); mass[0] = total
    data.mass.assign(n, 0.0);
    for (int i = 1; i < n; ++i) data.mass[i] = model.inertia[i].mass;
    for (int i = n - 1; i > 0; --i) data.mass[model.parent[i]] += data.mass[i];
    data.Yaba.assign(n, M6::zero()); data.joints.resize(n);
    data.u.assign(nv, 0.0); data.ddq.assign(nv, 0.0);
    auto init_state = [&](RobotState& s) {
        s.q.assign(nq, 0.0); s.v.assign(nv, 0.0); s.a.assign(nv, 0.0);
        s.command.assign(model.nmotors, 0.0); s.u.assign(nv, 0.0);
        s.uMotor.assign(model.nmotors, 0.0); s.uTransmission.assign(model.nmotors, 0.0);
        s.uInternal.assign(nv, 0.0); s.uCustom.assign(nv, 0.0);
        s.fExternal.assign(n, Force{});
    };
    init_state(state); init_state(statePrev);
    q.assign(nq, 0.0); v.assign(nv, 0.0); a.assign(nv, 0.0);
    contactFrameForces.assign(model.ncontacts, Force{}); contactForces.assign(model.ncontacts, Force{});
    contactForcesPrev.assign(model.ncontacts, Force{});
    fPrev.assign(n, Force{}); aPrev.assign(n, Motion{}); fExtBuffer.assign(n, Force{});
    sensors.assign(model.layout.width, 0.0);
    buildConstraints();
    spring_k.assign(nv, 0.0); spring_d.assign(nv, 0.0);
    ki.resize(7);
    for (auto& k : ki) { k.v.assign(nv, 0.0); k.a.assign(nv, 0.0); }
    for (Deriv* dd : {&inc, &scale, &err}) { dd->v.assign(nv, 0.0); dd->a.assign(nv, 0.0); }
    qBuf.assign(nq, 0.0); vBuf.assign(nv, 0.0); qCand.assign(nq, 0.0); vCand.assign(nv, 0.0);
    qOther.assign(nq, 0.0); vOther.assign(nv, 0.0); aOut.assign(nv, 0.0);
    set_options(o);
}

void Engine::set_options(const JbOptions& o) {
    static_cast<JbOptions&>(opt) = o;
    // stepperUpdatePeriod_ = min strictly positive of the two periods (engine.cc:2699-2715, :2794)
    const double sp = o.sensors_update_period, cp = o.controller_update_period;
    (void)sp; (void)cp;
    refreshStepperUpdatePeriod();
}

// isGcdIncluded over the controller / sensor / profile-force periods (engine.cc:2492-2516, :2551-2562):
// the caller guarantees they are multiples of each other, the breakpoint period is the smallest one.
void Engine::refreshStepperUpdatePeriod() {
    stepperUpdatePeriod = INF;
    if (opt.sensors_update_period > EPS) stepperUpdatePeriod = std::min(stepperUpdatePeriod, opt.sensors_update_period);
    if (opt.controller_update_period > EPS) stepperUpdatePeriod = std::min(stepperUpdatePeriod, opt.controller_update_period);
    for (const ProfileForce& pf : profileForces)
        if (pf.updatePeriod > EPS) stepperUpdatePeriod = std::min(stepperUpdatePeriod, pf.updatePeriod);
}

// Engine::registerImpulseForce (engine.cc:2450-2491)
int Engine::registerImpulseForce(int joint, const double* p, double tf, double dtf, const double* F) {
    if (running) return JB_ERR_BAD_CONTROL_FLOW;
    if (dtf < STEPPER_MIN_TIMESTEP || tf < 0.0 || joint <= 0 || joint >= model.njoints) return JB_ERR_INVALID_ARGUMENT;
    ImpulseForce f{joint, V3(p[0], p[1], p[2]), tf, dtf, Force{V3(F[0], F[1], F[2]), V3(F[3], F[4], F[5])}, false};
    impulseForces.push_back(f);
    for (double b : {tf, tf + dtf}) {
        auto it = std::lower_bound(impulseForceBreakpoints.begin(), impulseForceBreakpoints.end(), b);
        if (it == impulseForceBreakpoints.end() || *it != b) impulseForceBreakpoints.insert(it, b);
    }
    return JB_OK;
}

// Engine::registerProfileForce (engine.cc:2518-2567); the force function is "return the caller's buffer"
int Engine::registerProfileForce(int joint, const double* p, double updatePeriod) {
    if (running) return JB_ERR_BAD_CONTROL_FLOW;
    if (joint <= 0 || joint >= model.njoints) return JB_ERR_INVALID_ARGUMENT;
    if (EPS < updatePeriod && updatePeriod < SIMULATION_MIN_TIMESTEP) return JB_ERR_INVALID_ARGUMENT;
    profileForces.push_back(ProfileForce{joint, V3(p[0], p[1], p[2]), updatePeriod, Force{}, Force{}});
    refreshStepperUpdatePeriod();
    return static_cast<int>(profileForces.size()) - 1;
}

void Engine::removeAllForces() {
    impulseForces.clear(); impulseForceBreakpoints.clear(); profileForces.clear();
    refreshStepperUpdatePeriod();
}

// convertForceGlobalFrameToJoint (core/src/utilities/pinocchio.cc:794-809)
Force Engine::convertForceGlobalFrameToJoint(int joint, const V3& p, const Force& F) const {
    Force out;
    out.lin = tmul(data.oMi[joint].R, F.lin);
    out.ang = tmul(data.oMi[joint].R, F.ang) + cross(p, out.lin);
    return out;
}

// Engine::computeExternalForces (engine.cc:3455-3495)
void Engine::computeExternalForces(std::vector<Force>& fext) {
    for (const ImpulseForce& f : impulseForces)
        if (f.active) fext[f.joint] += convertForceGlobalFrameToJoint(f.joint, f.p, f.F);
    for (ProfileForce& pf : profileForces) {
        if (pf.updatePeriod < EPS) pf.force = pf.pending;   // profileForce.func(t, q, v)
        fext[pf.joint] += convertForceGlobalFrameToJoint(pf.joint, pf.p, pf.force);
    }
}

// ============================================================================ joint calc
// JointModel*::calc(jdata, q, v) of Pinocchio 2.7 for the joint types the URDF parser emits.
void Engine::jointCalc(int i, const double* qv, const double* vv) {
    const int t = model.jtype[i];
    JointData& jd = data.joints[i];
    const double* qj = qv + model.idx_q[i];
    jd.M = SE3::identity();
    jd.vJ = Motion{};
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) jd.S[r][c] = 0.0;
    if (t == JB_JOINT_FREEFLYER) {
        jd.M.R = quat_to_matrix(qj + 3);
        jd.M.p = V3(qj[0], qj[1], qj[2]);
        for (int k = 0; k < 6; ++k) jd.S[k][k] = 1.0;
        if (vv) { const double* vj = vv + model.idx_v[i]; jd.vJ = motion6(vj); }
        return;
    }
    if (t == JB_JOINT_SPHERICAL) {
        // JointModelSphericalTpl::calc: M = (quat.matrix(), 0), S = [0; 1_3], v = (0, omega), c = 0
        jd.M.R = quat_to_matrix(qj);
        for (int k = 0; k < 3; ++k) jd.S[3 + k][k] = 1.0;
        if (vv) { const double* vj = vv + model.idx_v[i]; jd.vJ.ang = V3(vj[0], vj[1], vj[2]); }
        return;
    }
    const V3 ax = model.axis[i];
    const double qd = vv ? vv[model.idx_v[i]] : 0.0;
    if (Model::is_revolute(t)) {
        double ca, sa;
        if (Model::is_unbounded(t)) { ca = qj[0]; sa = qj[1]; }
        else { ca = std::cos(qj[0]); sa = std::sin(qj[0]); }
        const int k = (t == JB_JOINT_RX || t == JB_JOINT_RUBX) ? 0
                    : (t == JB_JOINT_RY || t == JB_JOINT_RUBY) ? 1
                    : (t == JB_JOINT_RZ || t == JB_JOINT_RUBZ) ? 2 : -1;
        M3& R = jd.M.R;
        if (k == 0) { R(1, 1) = ca; R(1, 2) = -sa; R(2, 1) = sa; R(2, 2) = ca; }
        else if (k == 1) { R(0, 0) = ca; R(0, 2) = sa; R(2, 0) = -sa; R(2, 2) = ca; }
        else if (k == 2) { R(0, 0) = ca; R(0, 1) = -sa; R(1, 0) = sa; R(1, 1) = ca; }
        else {
            // Eigen::AngleAxis::toRotationMatrix as used by JointModelRevoluteUnaligned::calc
            const V3 sin_axis = sa * ax;
            const V3 cos1_axis = (1.0 - ca) * ax;
            double tmp;
            tmp = cos1_axis.x * ax.y; R(0, 1) = tmp - sin_axis.z; R(1, 0) = tmp + sin_axis.z;
            tmp = cos1_axis.x * ax.z; R(0, 2) = tmp + sin_axis.y; R(2, 0) = tmp - sin_axis.y;
            tmp = cos1_axis.y * ax.z; R(1, 2) = tmp - sin_axis.x; R(2, 1) = tmp + sin_axis.x;
            R(0, 0) = cos1_axis.x * ax.x + ca; R(1, 1) = cos1_axis.y * ax.y + ca; R(2, 2) = cos1_axis.z * ax.z + ca;
        }
        jd.S[3][0] = ax.x; jd.S[4][0] = ax.y; jd.S[5][0] = ax.z;
        jd.vJ.ang = qd * ax;
    } else {  // prismatic
        jd.M.p = qj[0] * ax;
        jd.S[0][0] = ax.x; jd.S[1][0] = ax.y; jd.S[2][0] = ax.z;
        jd.vJ.lin = qd * ax;
    }
}

// pinocchio::forwardKinematics(model, data, q, v, a) (second-order), called at engine.cc:2969
void Engine::forwardKinematics(const double* qv, const double* vv, const double* av) {
    data.v[0] = Motion{}; data.a[0] = Motion{};
    for (int i = 1; i < model.njoints; ++i) {
        const int p = model.parent[i];
        jointCalc(i, qv, vv);
        const JointData& jd = data.joints[i];
        data.v[i] = jd.vJ;
        data.liMi[i] = model.placement[i] * jd.M;
        if (p > 0) {
            data.oMi[i] = data.oMi[p] * data.liMi[i];
            data.v[i] += act_inv(data.liMi[i], data.v[p]);
        } else {
            data.oMi[i] = data.liMi[i];
        }
        // a = S * a_j + c + (v x v_J) ; a += liMi.actInv(a[parent])
        double sa[6] = {0, 0, 0, 0, 0, 0};
        const int nvj = Model::nvj(model.jtype[i]);
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < nvj; ++c) sa[r] += jd.S[r][c] * av[model.idx_v[i] + c];
        data.a[i] = motion6(sa) + cross(data.v[i], jd.vJ);
        data.a[i] += act_inv(data.liMi[i], data.a[p]);
    }
}

// ============================================================================ contacts
// Engine::computeContactDynamics (engine.cc:3197-3238)
V3 Engine::computeContactDynamics(const V3& nGround, double depth, const V3& vContactInWorld) const {
    V3 fextInWorld;
    if (depth < 0.0) {
        const double vDepth = dot(vContactInWorld, nGround);
        const double fextNormal = -std::min(opt.contact_stiffness * depth + opt.contact_damping * vDepth, 0.0);
        fextInWorld = fextNormal * nGround;
        const V3 vTangential = vContactInWorld - vDepth * nGround;
        const double vRatio = std::min(norm(vTangential) / opt.contact_transition_velocity, 1.0);
        const double fextTangential = opt.contact_friction * vRatio * fextNormal;
        fextInWorld -= fextTangential * vTangential;
        if (opt.contact_transition_eps > EPS) {
            const double blendingFactor = -depth / opt.contact_transition_eps;
            const double blendingLaw = std::tanh(2.0 * blendingFactor);
            fextInWorld = blendingLaw * fextInWorld;
        }
    }
    return fextInWorld;
}

// Engine::computeContactDynamicsAtFrame (engine.cc:3117-3195), spring-damper model, flat ground
// (engine.h:292-302), + convertForceGlobalFrameToJoint (utilities/pinocchio.cc:794-809)
void Engine::computeContactDynamicsAtFrame(int c, Force& fextLocal) const {
    const int j = model.contact_joint[c];
    const SE3& P = model.contact_placement[c];
    const SE3 oMf = data.oMi[j] * P;
    const double heightGround = 0.0;
    V3 normalGround(0.0, 0.0, 1.0);
    normalGround = (1.0 / norm(normalGround)) * normalGround;
    const V3& posFrame = oMf.p;
    const double depth = (posFrame.z - heightGround) * normalGround.z;
    if (depth < 0.0) {
        const V3 motionFrameLocal = act_inv(P, data.v[j]).lin;  // getFrameVelocity(LOCAL).linear()
        const V3 vContactInWorld = oMf.R * motionFrameLocal;
        const V3 fw = computeContactDynamics(normalGround, depth, vContactInWorld);
        // convertForceGlobalFrameToJoint, zero torque at the contact point
        fextLocal.lin = tmul(data.oMi[j].R, fw);
        fextLocal.ang = tmul(data.oMi[j].R, V3()) + cross(P.p, fextLocal.lin);
    } else {
        fextLocal = Force{};
    }
}

// Engine::computeCollisionForces (engine.cc:3394-3453), contact frames only
void Engine::computeCollisionForces(std::vector<Force>& fext, bool isStateUpToDate) {
    for (int c = 0; c < model.ncontacts; ++c) {
        Force& fextLocal = contactFrameForces[c];
        if (!isStateUpToDate) {
            if (opt.contact_model == JB_CONTACT_CONSTRAINT) updateContactConstraint(c);
            else computeContactDynamicsAtFrame(c, fextLocal);
        }
        fext[model.contact_joint[c]] += fextLocal;
        contactForces[c] = act_inv(model.contact_placement[c], fextLocal);
    }
}

// Engine::computeInternalDynamics (engine.cc:3340-3392): joint position bounds -> JointConstraint
// enable / disable, then the spring-damper of the flexibility joints (:3367-3391)
void Engine::computeInternalDynamics(const double* qv, const double* vv, std::vector<double>& uInternal) {
    updateJointBoundConstraints(qv);
    if (model.flexibility.empty()) return;
    const double PI = 3.14159265358979323846;
    for (int i = 1; i < model.njoints; ++i) {
        if (model.jtype[i] != JB_JOINT_SPHERICAL) continue;
        const int iq = model.idx_q[i], iv = model.idx_v[i];
        const double* kd = model.flexibility.data() + 6 * i;
        double angle;
        const V3 angleAxis = quat_log3(qv + iq, angle);
        if (angle > 0.95 * PI) { flexAngleError = true; }   // "Flexible joint angle must be smaller than 0.95 * pi."
        const M3 rotJlog3 = Jlog3(angle, angleAxis);
        const V3 t = rotJlog3 * V3(kd[0] * angleAxis.x, kd[1] * angleAxis.y, kd[2] * angleAxis.z);
        uInternal[iv] -= t.x; uInternal[iv + 1] -= t.y; uInternal[iv + 2] -= t.z;
        for (int k = 0; k < 3; ++k) uInternal[iv + k] -= kd[3 + k] * vv[iv + k];
    }
}

// Engine::computeAllTerms (engine.cc:3538-3583)
void Engine::computeAllTerms(double /*t*/, const double* qv, const double* vv, bool isStateUpToDate) {
    for (Force& f : state.fExternal) f = Force{};
    std::fill(state.uInternal.begin(), state.uInternal.end(), 0.0);
    computeInternalDynamics(qv, vv, state.uInternal);
    computeCollisionForces(state.fExternal, isStateUpToDate);
    computeExternalForces(state.fExternal);
}

// Engine::computeCommand (engine.cc:3240-3251).  Without a functor the command buffer is a
// zero-order hold of what the caller wrote (the batched boundary).
void Engine::computeCommand(double tt, const double* qv, const double* vv, std::vector<double>& command) {
    if (pdf_enabled) {
        // PDController.compute_command (proportional_derivative_controller.py:492-535): encoder data = motor-side
        // position / velocity; the command state restarts from the (clipped) measurement while no simulation is
        // running, and is integrated over one controller period otherwise
        const int nm = model.nmotors;
        std::vector<double> enc(2 * nm), elim(nm), vlim(nm);
        for (int m = 0; m < nm; ++m) {
            const int j = model.motor_joint[m];
            const double red = model.motor_params[10 * m];
            const double pos = Model::is_unbounded(model.jtype[j]) ? std::atan2(qv[model.idx_q[j] + 1], qv[model.idx_q[j]]) : qv[model.idx_q[j]];
            enc[m] = pos * red; enc[nm + m] = vv[model.idx_v[j]] * red;
            elim[m] = model.motor_params[10 * m + 1]; vlim[m] = model.motor_params[10 * m + 2];
        }
        if (!simStarted)
            for (int k = 0; k < 2 * nm; ++k) pdf_state[k] = std::min(std::max(enc[k], pdf_lower[k]), pdf_upper[k]);
        for (int m = 0; m < nm; ++m) pdf_state[2 * nm + m] = pdf_action[m];
        pd_controller(enc.data(), pdf_state.data(), pdf_lower.data(), pdf_upper.data(), pdf_kp.data(), pdf_kd.data(), elim.data(),
                      nm, simStarted ? opt.controller_update_period : 0.0, command.data());
        if (pdf_safety)
            apply_safety_limits(command.data(), enc.data(), enc.data() + nm, pdf_skp.data(), pdf_skd.data(), pdf_slo.data(),
                                pdf_shi.data(), pdf_svlim.data(), elim.data(), nm, command.data());
        return;
    }
    if (pd_enabled) {
        // gym_jiminy.common.blocks.pd_controller (python/gym_jiminy/common/gym_jiminy/common/blocks/
        // proportional_derivative_controller.py:101-165) with a zero-order-held position target and zero
        // target velocity: tau = clip(kp * ((q_des - q_enc) + kd * (0 - v_enc)), +-effort_limit).
        // Encoder data = motor-side position / velocity at the last sensor refresh (same instant).
        for (int m = 0; m < model.nmotors; ++m) {
            const int j = model.motor_joint[m];
            const double red = model.motor_params[10 * m], lim = model.motor_params[10 * m + 1];
            double pos = Model::is_unbounded(model.jtype[j]) ? std::atan2(qv[model.idx_q[j] + 1], qv[model.idx_q[j]]) : qv[model.idx_q[j]];
            const double q_enc = pos * red, v_enc = vv[model.idx_v[j]] * red;
            const double tau = pd_kp[m] * ((pd_target[m] - q_enc) + pd_kd[m] * (0.0 - v_enc));
            command[m] = std::min(std::max(tau, -lim), lim);
        }
        return;
    }
    if (!controller) return;
    std::fill(command.begin(), command.end(), 0.0);
    controller(ctx, tt, qv, vv, sensors.data(), command.data());
}

void Engine::computeCustom(double tt, const double* qv, const double* vv) {
    std::fill(state.uCustom.begin(), state.uCustom.end(), 0.0);
    for (int i = 1; i < model.njoints; ++i) {
        const int t = model.jtype[i];
        if (Model::nvj(t) != 1 || Model::is_unbounded(t)) continue;
        const int iv = model.idx_v[i], iq = model.idx_q[i];
        if (spring_k[iv] != 0.0 || spring_d[iv] != 0.0) state.uCustom[iv] = -spring_k[iv] * qv[iq] - spring_d[iv] * vv[iv];
    }
    if (internalDyn) internalDyn(ctx, tt, qv, vv, sensors.data(), state.uCustom.data());
}

// SimpleMotor::computeEffort over all motors (basic_motors.cc:83-143, abstract_motor.cc:459-493)
void Engine::computeMotorEfforts(const double* vv, const std::vector<double>& command) {
    for (int m = 0; m < model.nmotors; ++m) {
        const double* P = &model.motor_params[10 * m];
        const int flags = model.motor_flags[m];
        const double reduction = P[0], effortLimit = P[1], velocityLimit = P[2], invSlope = P[3];
        const double vj = vv[model.idx_v[model.motor_joint[m]]];
        const double vMotor = reduction * vj;
        double effortMin = -INF, effortMax = INF;
        if (flags & 1) {
            effortMin = -effortLimit; effortMax = effortLimit;
            if (flags & 2) {
                const double velocityDelta = effortLimit * invSlope;
                if (velocityDelta > 0.0) {
                    const double velocityThr = std::max(velocityLimit - velocityDelta, 0.0);
                    effortMin *= std::clamp((velocityLimit + vMotor) / (velocityLimit - velocityThr), 0.0, 1.0);
                    effortMax *= std::clamp((velocityLimit - vMotor) / (velocityLimit - velocityThr), 0.0, 1.0);
                }
            }
        }
        double uMotor = std::clamp(command[m], effortMin, effortMax);
        double uTransmission = reduction * uMotor;
        if (flags & 4) {
            if (vj > 0.0) uTransmission += P[4] * vj + P[6] * std::tanh(P[8] * vj);
            else uTransmission += P[5] * vj + P[7] * std::tanh(P[8] * vj);
        }
        state.uMotor[m] = uMotor;
        state.uTransmission[m] = uTransmission;
    }
}

// ============================================================================ ABA
static void invert_spd(int n, const double A[6][6], double Ainv[6][6]) {
    // PerformStYSInversion: StYS.llt().solveInPlace(Identity)
    double L[6][6] = {};
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = A[i][j];
            for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
            L[i][j] = (i == j) ? std::sqrt(s) : s / L[j][j];
        }
    for (int c = 0; c < n; ++c) {
        double y[6], x[6];
        for (int i = 0; i < n; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
            y[i] = s / L[i][i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int k = i + 1; k < n; ++k) s -= L[k][i] * x[k];
            x[i] = s / L[i][i];
        }
        for (int i = 0; i < n; ++i) Ainv[i][c] = x[i];
    }
}

// pinocchio_overload::aba (pinocchio_overload_algorithms.h:446-489) with AbaForwardStep1,
// AbaBackwardStep (:126-167, calc_aba :169-413) and AbaForwardStep2 of Pinocchio 2.7.
const std::vector<double>& Engine::aba(const double* qv, const double* vv, const std::vector<double>& tau,
                                       const std::vector<Force>& fext) {
    const int n = model.njoints;
    data.v[0] = Motion{};
    data.a_gf[0] = Motion{V3(-opt.gravity[0], -opt.gravity[1], -opt.gravity[2]),
                          V3(-opt.gravity[3], -opt.gravity[4], -opt.gravity[5])};
    data.u = tau;
    // Pass 1
    for (int i = 1; i < n; ++i) {
        const int p = model.parent[i];
        jointCalc(i, qv, vv);
        const JointData& jd = data.joints[i];
        data.liMi[i] = model.placement[i] * jd.M;
        data.v[i] = jd.vJ;
        if (p > 0) data.v[i] += act_inv(data.liMi[i], data.v[p]);
        data.a_gf[i] = cross(data.v[i], jd.vJ);  // + c (== 0)
        data.Yaba[i] = inertia_matrix(model.inertia[i]);
        data.f[i] = cross(data.v[i], model.inertia[i] * data.v[i]);  // vxiv
        data.f[i] -= fext[i];
    }
    // Pass 2
    for (int i = n - 1; i > 0; --i) {
        const int p = model.parent[i];
        JointData& jd = data.joints[i];
        const int nvj = Model::nvj(model.jtype[i]);
        const int iv = model.idx_v[i];
        M6& Ia = data.Yaba[i];
        double f6[6];
        to6(data.f[i], f6);
        for (int c = 0; c < nvj; ++c) {
            double s = 0.0;
            for (int r = 0; r < 6; ++r) s += jd.S[r][c] * f6[r];
            data.u[iv + c] -= s;
        }
        // calc_aba: U = Ia S ; StU = S^T U + Im ; Dinv = StU^-1 ; UDinv = U Dinv
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < nvj; ++c) {
                double s = 0.0;
                for (int k = 0; k < 6; ++k) s += Ia(r, k) * jd.S[k][c];
                jd.U[r][c] = s;
            }
        double StU[6][6];
        for (int r = 0; r < nvj; ++r)
            for (int c = 0; c < nvj; ++c) {
                double s = 0.0;
                for (int k = 0; k < 6; ++k) s += jd.S[k][r] * jd.U[k][c];
                StU[r][c] = s;
            }
        for (int c = 0; c < nvj; ++c) StU[c][c] += model.rotor[iv + c];
        if (nvj == 1) jd.Dinv[0][0] = 1.0 / StU[0][0];
        else invert_spd(nvj, StU, jd.Dinv);
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < nvj; ++c) {
                double s = 0.0;
                for (int k = 0; k < nvj; ++k) s += jd.U[r][k] * jd.Dinv[k][c];
                jd.UDinv[r][c] = s;
            }
        if (p > 0) {
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 6; ++c) {
                    double s = 0.0;
                    for (int k = 0; k < nvj; ++k) s += jd.UDinv[r][k] * jd.U[c][k];
                    Ia(r, c) -= s;
                }
            double ag[6], pa[6], tmp[6];
            to6(data.a_gf[i], ag);
            mul6(Ia, ag, tmp);
            for (int r = 0; r < 6; ++r) pa[r] = f6[r] + tmp[r];
            for (int r = 0; r < 6; ++r) {
                double s = 0.0;
                for (int k = 0; k < nvj; ++k) s += jd.UDinv[r][k] * data.u[iv + k];
                pa[r] += s;
            }
            data.f[i] = force6(pa);
            const M6 T = se3_act_on(data.liMi[i], Ia);
            for (int k = 0; k < 36; ++k) data.Yaba[p].m[k] += T.m[k];
            data.f[p] += act(data.liMi[i], data.f[i]);
        }
    }
    // Pass 3
    for (int i = 1; i < n; ++i) {
        const int p = model.parent[i];
        const JointData& jd = data.joints[i];
        const int nvj = Model::nvj(model.jtype[i]);
        const int iv = model.idx_v[i];
        data.a_gf[i] += act_inv(data.liMi[i], data.a_gf[p]);
        double ag[6];
        to6(data.a_gf[i], ag);
        for (int c = 0; c < nvj; ++c) {
            double s = 0.0;
            for (int k = 0; k < nvj; ++k) s += jd.Dinv[c][k] * data.u[iv + k];
            double s2 = 0.0;
            for (int r = 0; r < 6; ++r) s2 += jd.UDinv[r][c] * ag[r];
            data.ddq[iv + c] = s - s2;
        }
        for (int r = 0; r < 6; ++r) {
            double s = 0.0;
            for (int c = 0; c < nvj; ++c) s += jd.S[r][c] * data.ddq[iv + c];
            ag[r] += s;
        }
        data.a_gf[i] = motion6(ag);
    }
    return data.ddq;
}

// ============================================================================ sensors
// <Sensor>::set() of IMU / Force / Encoder / Effort / Contact (basic_sensors.cc:142-164, :267,
// :368-386, :509-537, :604), noise-free, zero delay, zero bias.
void Engine::computeSensorMeasurements(const double* qv, const double* vv, const std::vector<double>& uMotor) {
    const JbSensorLayout& L = model.layout;
    double* s = sensors.data();
    for (int k = 0; k < model.nimu; ++k) {
        const int j = model.imu_joint[k];
        const SE3& P = model.imu_placement[k];
        const Motion velocity = act_inv(P, data.v[j]);
        Motion acceleration = act_inv(P, data.a[j]);
        acceleration.lin += cross(velocity.ang, velocity.lin);  // classical acceleration
        const M3 rot = (data.oMi[j] * P).R;
        const V3 g(opt.gravity[0], opt.gravity[1], opt.gravity[2]);
        const V3 acc = acceleration.lin - tmul(rot, g);
        const double val[6] = {velocity.ang.x, velocity.ang.y, velocity.ang.z, acc.x, acc.y, acc.z};
        for (int f = 0; f < 6; ++f) s[L.imu_offset + f * model.nimu + k] = val[f];
    }
    for (int k = 0; k < model.nforce; ++k) {
        Force sum{};
        for (const auto& cp : model.force_contacts[k]) sum += act(cp.second, contactForces[cp.first]);
        double val[6];
        to6(sum, val);
        for (int f = 0; f < 6; ++f) s[L.force_offset + f * model.nforce + k] = val[f];
    }
    for (int k = 0; k < model.nenc; ++k) {
        const int j = model.enc_joint[k];
        const int t = model.jtype[j];
        double pos;
        if (Model::is_unbounded(t)) pos = std::atan2(qv[model.idx_q[j] + 1], qv[model.idx_q[j]]);
        else pos = qv[model.idx_q[j]];
        const double vel = vv[model.idx_v[j]];
        s[L.encoder_offset + 0 * model.nenc + k] = pos * model.enc_reduction[k];
        s[L.encoder_offset + 1 * model.nenc + k] = vel * model.enc_reduction[k];
    }
    for (int k = 0; k < model.neff; ++k) s[L.effort_offset + k] = uMotor[model.eff_motor[k]];
    for (int k = 0; k < model.ncs; ++k) {
        const V3& f = contactForces[model.cs_index[k]].lin;
        s[L.contact_offset + 0 * model.ncs + k] = f.x;
        s[L.contact_offset + 1 * model.ncs + k] = f.y;
        s[L.contact_offset + 2 * model.ncs + k] = f.z;
    }
    if (sensorPipeline) measureSensors();
}

// ---- measurement pipeline: Robot::reset -> resetAll seeds, setAll ring, interpolateData + measureData (sensor_noise.hpp)
void Engine::setSensorOptions(int type, int index, const double* noiseStd, const double* bias, double delay, double jitter, uint32_t order) {
    const int counts[N_SENSOR_TYPES] = {model.nimu, model.nforce, model.nenc, model.neff, model.ncs};
    const int offs[N_SENSOR_TYPES] = {model.layout.imu_offset, model.layout.force_offset, model.layout.encoder_offset,
                                      model.layout.effort_offset, model.layout.contact_offset};
    if (!sensorPipeline) {
        for (int t = 0; t < N_SENSOR_TYPES; ++t) {
            SensorGroup& g = sensorGroups[t];
            g.nf = SENSOR_FIELDS[t]; g.ns = counts[t]; g.offset = offs[t];
            g.opt.assign(g.ns, SensorOptions{});
        }
        measurements.assign(sensors.size(), 0.0);
        sensorPipeline = true;
    }
    SensorOptions& o = sensorGroups[type].opt.at(index);
    const int nf = SENSOR_FIELDS[type];
    o.noiseStd.clear(); o.bias.clear();
    if (noiseStd) o.noiseStd.assign(noiseStd, noiseStd + nf);
    if (bias) o.bias.assign(bias, bias + nf);
    o.delay = delay; o.jitter = jitter; o.delayInterpolationOrder = order;
}
void Engine::resetSensorPipeline() {
    // Engine::reset: generator_.seed(seed_seq(randomSeedSeq)) (engine.cc:756-757), robot->reset(generator_) (:763): one
    // draw per sensor type that has sensors (robot.cc:137-144), in the fixed type order documented in sensor_noise.hpp
    std::seed_seq seq{engineSeed};
    PCG32 g = pcg32_from_seed_seq(seq);
    for (SensorGroup& grp : sensorGroups)
        if (grp.ns > 0) grp.reset(g());
}
void Engine::measureSensors() {
    for (SensorGroup& grp : sensorGroups) {
        if (!grp.ns) continue;
        grp.push(sensorClock, sensors.data() + grp.offset);
        for (int k = 0; k < grp.ns; ++k) grp.measure(k, measurements.data() + grp.offset);
    }
}

// ============================================================================ RHS
// Engine::computeRobotsDynamics (engine.cc:3585-3708)
void Engine::computeRobotsDynamics(double tt, const double* qv, const double* vv, std::vector<double>& aOutV,
                                   bool isStateUpToDate) {
    ++rhs_count;
    if (!isStateUpToDate) forwardKinematics(qv, vv, statePrev.a.data());
    computeAllTerms(tt, qv, vv, isStateUpToDate);
    if (!isStateUpToDate && opt.sensors_update_period < EPS) {
        // Roll back to forces and accelerations computed at previous iteration (engine.cc:3658-3672)
        contactForcesPrev.swap(contactForces); fPrev.swap(data.f); aPrev.swap(data.a);
        computeSensorMeasurements(qv, vv, statePrev.uMotor);
        contactForcesPrev.swap(contactForces); fPrev.swap(data.f); aPrev.swap(data.a);
    }
    if (opt.controller_update_period < EPS) computeCommand(tt, qv, vv, state.command);
    computeMotorEfforts(vv, state.command);
    computeCustom(tt, qv, vv);
    for (int k = 0; k < model.nv; ++k) state.u[k] = state.uInternal[k] + state.uCustom[k];
    for (int m = 0; m < model.nmotors; ++m) state.u[model.idx_v[model.motor_joint[m]]] += state.uTransmission[m];
    aOutV = computeAcceleration(qv, vv, state.u, state.fExternal, isStateUpToDate, false);
}

// computeExtraTerms (engine.cc:800-905): energies, true joint accelerations `data.a`, joint
// internal wrenches `data.f` (subtree inertia / CoM / centroidal terms are analysis-only outputs
// not consumed by the step path and are not restated).
void Engine::computeExtraTerms() {
    const int n = model.njoints;
    double kin = 0.0;
    for (int i = 1; i < n; ++i) kin += vtiv(model.inertia[i], data.v[i]);
    kin *= 0.5;
    double rot = 0.0;
    for (int k = 0; k < model.nv; ++k) rot += model.rotor[k] * (state.v[k] * state.v[k]);
    data.kinetic_energy = kin + 0.5 * rot;
    double pot = 0.0;
    const V3 g(opt.gravity[0], opt.gravity[1], opt.gravity[2]);
    for (int i = 1; i < n; ++i) {
        const V3 com = act_point(data.oMi[i], model.inertia[i].c);
        pot -= model.inertia[i].mass * dot(g, com);
    }
    data.potential_energy = pot;

    std::vector<Force>& fExt = fExtBuffer;
    data.h[0] = Force{}; fExt[0] = Force{}; data.f[0] = Force{}; data.a[0] = Motion{};
    data.a_gf[0] = Motion{V3(-opt.gravity[0], -opt.gravity[1], -opt.gravity[2]),
                          V3(-opt.gravity[3], -opt.gravity[4], -opt.gravity[5])};
    for (int i = 1; i < n; ++i) {
        const JointData& jd = data.joints[i];
        const int nvj = Model::nvj(model.jtype[i]);
        // ForwardKinematicsAccelerationStep (engine.cc:776-791)
        data.a[i] = cross(data.v[i], jd.vJ);
        double sa[6] = {0, 0, 0, 0, 0, 0};
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < nvj; ++c) sa[r] += jd.S[r][c] * state.a[model.idx_v[i] + c];
        data.a[i] += motion6(sa);
        const int p = model.parent[i];
        data.a_gf[i] = data.a[i];
        data.a[i] += act_inv(data.liMi[i], data.a[p]);
        data.a_gf[i] += act_inv(data.liMi[i], data.a_gf[p]);
        data.h[i] = model.inertia[i] * data.v[i];
        fExt[i] = model.inertia[i] * data.a[i];
        data.f[i] = cross(data.v[i], data.h[i]);
        fExt[i] += data.f[i];
        data.f[i] += model.inertia[i] * data.a_gf[i];
        data.f[i] -= state.fExternal[i];
    }
    for (int i = n - 1; i > 0; --i) {
        const int p = model.parent[i];
        fExt[p] += act(data.liMi[i], fExt[i]);
        data.h[p] += act(data.liMi[i], data.h[i]);
        if (p > 0) data.f[p] += act(data.liMi[i], data.f[i]);
    }
    // subtree (composite) inertias (engine.cc:817-832; with constraints the reference's CRBA leaves the same quantity)
    for (int i = 1; i < n; ++i) data.Ycrb[i] = model.inertia[i];
    for (int i = n - 1; i > 0; --i) {
        const int p = model.parent[i];
        if (p > 0) add_inertia(data.Ycrb[p], act(data.liMi[i], data.Ycrb[i]));
    }
    // position and velocity of the centre of mass of each subtree (engine.cc:890-898)
    for (int i = 0; i < n; ++i) {
        if (i > 0) data.com[i] = data.Ycrb[i].c;
        data.vcom[i] = V3(data.h[i].lin.x / data.mass[i], data.h[i].lin.y / data.mass[i], data.h[i].lin.z / data.mass[i]);
    }
    data.com[0] = n > 1 ? act_point(data.liMi[1], data.com[1]) : V3();
    // centroidal momentum and its derivative (engine.cc:900-904)
    data.hg = data.h[0];
    data.hg.ang += cross(data.hg.lin, data.com[0]);
    data.dhg = fExt[0];
    data.dhg.ang += cross(data.dhg.lin, data.com[0]);
}

// syncAccelerationsAndForces (engine.cc:920-950)
void Engine::syncAccelerationsAndForces() {
    for (int c = 0; c < model.ncontacts; ++c) contactForcesPrev[c] = contactForces[c];
    for (int i = 0; i < model.njoints; ++i) { fPrev[i] = data.f[i]; aPrev[i] = data.a[i]; }
}

// ============================================================================ Lie group
// pinocchio::integrate (liegroup/{vector-space,special-orthogonal,special-euclidean}.hpp)
void Engine::integrate(const double* q0, const double* vel, double* out) const {
    for (int i = 1; i < model.njoints; ++i) {
        const int t = model.jtype[i], iq = model.idx_q[i], iv = model.idx_v[i];
        if (t == JB_JOINT_FREEFLYER) {
            SE3 M0; M0.R = quat_to_matrix(q0 + iq + 3); M0.p = V3(q0[iq], q0[iq + 1], q0[iq + 2]);
            const SE3 M1 = M0 * exp6(motion6(vel + iv));
            out[iq] = M1.p.x; out[iq + 1] = M1.p.y; out[iq + 2] = M1.p.z;
            double quat[4];
            matrix_to_quat(M1.R, quat);
            double dp = 0.0;
            for (int k = 0; k < 4; ++k) dp += quat[k] * q0[iq + 3 + k];
            if (dp < 0.0) for (int k = 0; k < 4; ++k) quat[k] = -quat[k];
            double N2 = 0.0;
            for (int k = 0; k < 4; ++k) N2 += quat[k] * quat[k];
            const double alpha = (3.0 - N2) / 2.0;  // quaternion::firstOrderNormalize
            for (int k = 0; k < 4; ++k) out[iq + 3 + k] = quat[k] * alpha;
        } else if (t == JB_JOINT_SPHERICAL) {
            // SpecialOrthogonalOperationTpl<3>::integrate_impl: quat * exp3(omega), firstOrderNormalize
            double pOmega[4], quat[4];
            quat_exp3(V3(vel[iv], vel[iv + 1], vel[iv + 2]), pOmega);
            quat_mul(q0 + iq, pOmega, quat);
            double N2 = 0.0;
            for (int k = 0; k < 4; ++k) N2 += quat[k] * quat[k];
            const double alpha = (3.0 - N2) / 2.0;
            for (int k = 0; k < 4; ++k) out[iq + k] = quat[k] * alpha;
        } else if (Model::is_unbounded(t)) {
            const double ca = q0[iq], sa = q0[iq + 1], omega = vel[iv];
            const double cosOmega = std::cos(omega), sinOmega = std::sin(omega);
            double o0 = cosOmega * ca - sinOmega * sa, o1 = sinOmega * ca + cosOmega * sa;
            const double norm2 = o0 * o0 + o1 * o1;
            const double k = (3.0 - norm2) / 2.0;
            out[iq] = o0 * k; out[iq + 1] = o1 * k;
        } else {
            out[iq] = q0[iq] + vel[iv];
        }
    }
}

// pinocchio::difference(q0, q1)
void Engine::difference(const double* q0, const double* q1, double* out) const {
    const double PI = 3.14159265358979323846;
    for (int i = 1; i < model.njoints; ++i) {
        const int t = model.jtype[i], iq = model.idx_q[i], iv = model.idx_v[i];
        if (t == JB_JOINT_FREEFLYER) {
            SE3 M0; M0.R = quat_to_matrix(q0 + iq + 3); M0.p = V3(q0[iq], q0[iq + 1], q0[iq + 2]);
            SE3 M1; M1.R = quat_to_matrix(q1 + iq + 3); M1.p = V3(q1[iq], q1[iq + 1], q1[iq + 2]);
            const Motion d = log6(se3_inverse(M0) * M1);
            to6(d, out + iv);
        } else if (t == JB_JOINT_SPHERICAL) {
            // SpecialOrthogonalOperationTpl<3>::difference_impl: quaternion::log3(q0.conjugate() * q1)
            const double q0c[4] = {-q0[iq], -q0[iq + 1], -q0[iq + 2], q0[iq + 3]};
            double dq[4], theta;
            quat_mul(q0c, q1 + iq, dq);
            const V3 w = quat_log3(dq, theta);
            out[iv] = w.x; out[iv + 1] = w.y; out[iv + 2] = w.z;
        } else if (Model::is_unbounded(t)) {
            // SpecialOrthogonalOperationTpl<2>::difference_impl + log
            const double R00 = q0[iq] * q1[iq] + q0[iq + 1] * q1[iq + 1];
            const double R10 = q0[iq] * q1[iq + 1] - q0[iq + 1] * q1[iq];
            const double tr = 2.0 * R00;
            const bool pos = R10 > 0.0;
            double theta;
            if (tr > 2.0) theta = 0.0;
            else if (tr < -2.0) theta = pos ? PI : -PI;
            else if (tr > 2.0 - 1e-2) theta = std::asin((R10 - (-R10)) / 2.0);
            else theta = pos ? std::acos(tr / 2.0) : -std::acos(tr / 2.0);
            out[iv] = theta;
        } else {
            out[iv] = q1[iq] - q0[iq];
        }
    }
}

void Engine::neutral(double* qn) const {
    for (int i = 1; i < model.njoints; ++i) {
        const int t = model.jtype[i], iq = model.idx_q[i];
        for (int k = 0; k < Model::nqj(t); ++k) qn[iq + k] = 0.0;
        if (t == JB_JOINT_FREEFLYER) qn[iq + 6] = 1.0;
        else if (t == JB_JOINT_SPHERICAL) qn[iq + 3] = 1.0;
        else if (Model::is_unbounded(t)) qn[iq] = 1.0;
    }
}

void Engine::normalize(double* qn) const {
    for (int i = 1; i < model.njoints; ++i) {
        const int t = model.jtype[i], iq = model.idx_q[i];
        int off = -1, len = 0;
        if (t == JB_JOINT_FREEFLYER) { off = iq + 3; len = 4; }
        else if (t == JB_JOINT_SPHERICAL) { off = iq; len = 4; }
        else if (Model::is_unbounded(t)) { off = iq; len = 2; }
        if (off < 0) continue;
        double n2 = 0.0;
        for (int k = 0; k < len; ++k) n2 += qn[off + k] * qn[off + k];
        const double nn = std::sqrt(n2);
        for (int k = 0; k < len; ++k) qn[off + k] /= nn;
    }
}

// ============================================================================ steppers
// AbstractStepper::f (abstract_stepper.cc:64-69)
void Engine::f(double tt, const std::vector<double>& qq, const std::vector<double>& vv, Deriv& out) {
    computeRobotsDynamics(tt, qq.data(), vv.data(), aOut, false);
    out.a = aOut;
    out.v = vv;
}

// EulerExplicitStepper::tryStepImpl (euler_explicit_stepper.cc:6-22); state = (qBuf,vBuf), ki[0] = derivative
bool Engine::tryStepEuler(double tt, double& dtt) {
    const int nv = model.nv;
    for (int k = 0; k < nv; ++k) { inc.v[k] = dtt * ki[0].v[k]; inc.a[k] = dtt * ki[0].a[k]; }
    integrate(qBuf.data(), inc.v.data(), qCand.data());
    for (int k = 0; k < nv; ++k) vCand[k] = vBuf[k] + inc.a[k];
    qBuf = qCand; vBuf = vCand;
    f(tt + dtt, qBuf, vBuf, ki[0]);
    dtt = INF;
    return true;
}

namespace rk4 {
const double A[4][4] = {{0, 0, 0, 0}, {0.5, 0, 0, 0}, {0, 0.5, 0, 0}, {0, 0, 1.0, 0}};
const double c[4] = {0.0, 0.5, 0.5, 1.0};
const double b[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
}
namespace dopri {
const double A[7][7] = {
    {0, 0, 0, 0, 0, 0, 0},
    {1.0 / 5.0, 0, 0, 0, 0, 0, 0},
    {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0, 0, 0},
    {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0, 0, 0},
    {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0, 0, 0},
    {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, 0, 0},
    {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0, 0}};
const double c[7] = {0.0, 2.0 / 10.0, 3.0 / 10.0, 4.0 / 5.0, 8.0 / 9.0, 1.0, 1.0};
const double b[7] = {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0, 0.0};
const double e[7] = {5179.0 / 57600.0, 0.0, 7571.0 / 16695.0, 393.0 / 640.0, -92097.0 / 339200.0,
                     187.0 / 2100.0, 1.0 / 40.0};
const double STEPPER_ORDER = 5.0, SAFETY = 0.8, ERROR_THRESHOLD = 0.5, MIN_FACTOR = 0.2, MAX_FACTOR = 5.0;
}

// RungeKuttaDOPRIStepper::computeError (runge_kutta_dopri_stepper.cc:58-82)
double Engine::computeErrorDopri(double dtt) {
    const int nv = model.nv;
    // scale = tolAbs + tolRel * |initialState (-) neutral|
    neutral(qOther.data());
    std::fill(vOther.begin(), vOther.end(), 0.0);
    difference(qBuf.data(), qOther.data(), scale.v.data());  // initialState.difference(otherSolution_, scale_)
    for (int k = 0; k < nv; ++k) scale.a[k] = vBuf[k] - vOther[k];
    for (int k = 0; k < nv; ++k) {
        // NB: Engine::start builds `RungeKuttaDOPRIStepper(robotOde, robots, stepper.tolAbs, stepper.tolRel)`
        // (engine.cc:1161-1163) while the constructor takes (tolRel, tolAbs): the two options reach the
        // error scale swapped.  Restated as the reference behaves, not as the option names suggest.
        const double tolRel_ = opt.tol_abs, tolAbs_ = opt.tol_rel;
        scale.v[k] = std::fabs(scale.v[k]) * tolRel_ + tolAbs_;
        scale.a[k] = std::fabs(scale.a[k]) * tolRel_ + tolAbs_;
    }
    std::fill(inc.v.begin(), inc.v.end(), 0.0);
    std::fill(inc.a.begin(), inc.a.end(), 0.0);
    for (int i = 0; i < 7; ++i) {
        const double w = dtt * dopri::e[i];
        for (int k = 0; k < nv; ++k) { inc.v[k] += w * ki[i].v[k]; inc.a[k] += w * ki[i].a[k]; }
    }
    integrate(qBuf.data(), inc.v.data(), qOther.data());
    for (int k = 0; k < nv; ++k) vOther[k] = vBuf[k] + inc.a[k];
    // solution.difference(otherSolution_, error_)
    difference(qCand.data(), qOther.data(), err.v.data());
    for (int k = 0; k < nv; ++k) err.a[k] = vCand[k] - vOther[k];
    double e = 0.0;
    bool isnan = false;
    for (int k = 0; k < nv; ++k) {
        const double ev = std::fabs(err.v[k] / scale.v[k]), ea = std::fabs(err.a[k] / scale.a[k]);
        if (ev != ev || ea != ea) isnan = true;
        e = std::max(e, std::max(ev, ea));
    }
    return isnan ? std::nan("") : e;
}

// RungeKuttaDOPRIStepper::adjustStep (runge_kutta_dopri_stepper.cc:18-56)
bool Engine::adjustStepDopri(double& dtt) {
    const double error = computeErrorDopri(dtt);
    if (std::isnan(error)) throw std::runtime_error("The estimated integration error contains 'nan'.");
    using namespace dopri;
    if (error < 1.0) {
        if (error < std::min(ERROR_THRESHOLD, std::pow(SAFETY, STEPPER_ORDER))) {
            const double clippedError = std::max(error, std::pow(MAX_FACTOR / SAFETY, -STEPPER_ORDER));
            dtt *= SAFETY * std::pow(clippedError, -1.0 / STEPPER_ORDER);
        }
        return true;
    }
    dtt *= std::max(SAFETY * std::pow(error, -1.0 / (STEPPER_ORDER - 2)), MIN_FACTOR);
    return false;
}

// AbstractRungeKuttaStepper::tryStepImpl (abstract_runge_kutta_stepper.cc:25-77)
bool Engine::tryStepRK(double tt, double& dtt) {
    const bool isDopri = opt.ode_solver == JB_SOLVER_RUNGE_KUTTA_DOPRI;
    const int ns = isDopri ? 7 : 4;
    const int nv = model.nv;
    auto Acoef = [&](int i, int j) { return isDopri ? dopri::A[i][j] : rk4::A[i][j]; };
    const double* cc = isDopri ? dopri::c : rk4::c;
    const double* bb = isDopri ? dopri::b : rk4::b;
    // ki[0] already holds the provided stateDerivative
    std::vector<double>& qS = qCand;  // reuse as stage buffer before the candidate is formed
    std::vector<double>& vS = vCand;
    for (int i = 1; i < ns; ++i) {
        std::fill(inc.v.begin(), inc.v.end(), 0.0);
        std::fill(inc.a.begin(), inc.a.end(), 0.0);
        for (int j = 0; j < i; ++j) {
            const double w = dtt * Acoef(i, j);
            for (int k = 0; k < nv; ++k) { inc.v[k] += w * ki[j].v[k]; inc.a[k] += w * ki[j].a[k]; }
        }
        integrate(qBuf.data(), inc.v.data(), qS.data());
        for (int k = 0; k < nv; ++k) vS[k] = vBuf[k] + inc.a[k];
        f(tt + cc[i] * dtt, qS, vS, ki[i]);
    }
    std::fill(inc.v.begin(), inc.v.end(), 0.0);
    std::fill(inc.a.begin(), inc.a.end(), 0.0);
    for (int i = 0; i < ns; ++i) {
        const double w = dtt * bb[i];
        for (int k = 0; k < nv; ++k) { inc.v[k] += w * ki[i].v[k]; inc.a[k] += w * ki[i].a[k]; }
    }
    integrate(qBuf.data(), inc.v.data(), qCand.data());
    for (int k = 0; k < nv; ++k) vCand[k] = vBuf[k] + inc.a[k];
    const double t_next = tt + dtt;
    bool hasSucceeded = true;
    if (isDopri) hasSucceeded = adjustStepDopri(dtt);
    else dtt = INF;
    if (hasSucceeded) {
        qBuf = qCand; vBuf = vCand;
        if (isDopri) ki[0] = ki[ns - 1];
        else f(t_next, qBuf, vBuf, ki[0]);
    }
    return hasSucceeded;
}

// AbstractStepper::tryStep (abstract_stepper.cc:16-62)
Engine::RC Engine::tryStep(double& tt, double& dtt) {
    const double t_next = tt + dtt;
    qBuf = q; vBuf = v;
    ki[0].v = v; ki[0].a = a;
    try {
        bool ok = (opt.ode_solver == JB_SOLVER_EULER_EXPLICIT) ? tryStepEuler(tt, dtt) : tryStepRK(tt, dtt);
        if (!ok) return IS_FAILURE;
        for (double x : ki[0].a)
            if (x != x) throw std::runtime_error("The integrated acceleration contains 'nan'.");
    } catch (...) {
        return IS_ERROR;
    }
    tt = t_next;
    q = qBuf; v = vBuf; a = ki[0].a;
    return IS_SUCCESS;
}

// MahonyFilter.refresh_observation (blocks/mahony_filter.py:337-393) with exact_init = True, ignore_twist = False:
// at start the estimate is the true orientation of the IMU frame, afterwards one filter iteration per sensor refresh
void Engine::mahonyInit() {
    const int M = model.nimu;
    mahony_q.assign(4 * M, 0.0); mahony_bias.assign(3 * M, 0.0); mahony_omega.assign(3 * M, 0.0);
    for (int k = 0; k < M; ++k) {
        const M3 R = (data.oMi[model.imu_joint[k]] * model.imu_placement[k]).R;
        double quat[4];
        matrix_to_quat_ref(R.m, quat);
        for (int e = 0; e < 4; ++e) mahony_q[e * M + k] = quat[e];
    }
}
void Engine::mahonyUpdate() {
    const int M = model.nimu;
    if (!M) return;
    const double* s = sensorOutput().data() + model.layout.imu_offset;   // [6][nimu]: gyro (3), accel (3), as measured
    mahony_filter(mahony_q.data(), mahony_omega.data(), s, s + 3 * M, mahony_bias.data(), M, mahony_kp, mahony_ki, opt.sensors_update_period);
}

// ============================================================================ start
// Engine::start (engine.cc:952-1533), single robot, spring-damper contact model
int Engine::start(const double* q0, const double* v0) {
    const int nv = model.nv, nq = model.nq;
    status = JB_ENV_OK;
    std::vector<double> qn(q0, q0 + nq);
    for (int k = 0; k < nq; ++k) {
        if (EPS < qn[k] - model.q_upper[k] || EPS < model.q_lower[k] - qn[k]) return JB_ERR_INVALID_ARGUMENT;
    }
    normalize(qn.data());
    simStarted = false;   // is_simulation_running becomes true at the very end of Engine::start (engine.cc:1532)
    if (sensorPipeline) resetSensorPipeline();
    sensorClock = 0.0;
    q = qn; v.assign(v0, v0 + nv); a.assign(nv, 0.0);
    iter = 0; iterFailed = 0; t = 0.0; tPrev = 0.0; tError = 0.0;
    dt = SIMULATION_MIN_TIMESTEP; dtLargest = dt; dtLargestPrev = dt;
    for (Force& f : contactForcesPrev) f = Force{};
    for (Force& f : fPrev) f = Force{};
    for (Motion& m : aPrev) m = Motion{};
    for (Force& f : contactForces) f = Force{};
    // syncRobotsStateWithStepper
    state.q = q; state.v = v; state.a = a;
    std::fill(state.u.begin(), state.u.end(), 0.0);
    std::fill(state.uMotor.begin(), state.uMotor.end(), 0.0);
    std::fill(state.uTransmission.begin(), state.uTransmission.end(), 0.0);
    std::fill(state.uCustom.begin(), state.uCustom.end(), 0.0);
    // impulse forces: breakpoint iterator, active set (engine.cc:1214-1238); a profile force with a
    // finite update period is first evaluated by the first `step` (its value is zero until then)
    impulseForceBreakpointNext = 0;
    for (ImpulseForce& f : impulseForces) f.active = f.t < STEPPER_MIN_TIMESTEP;
    for (ProfileForce& pf : profileForces) pf.force = Force{};
    forwardKinematics(state.q.data(), state.v.data(), state.a.data());
    resetConstraints(state.q.data());
    double forceMax = 0.0;
    for (int c = 0; c < model.ncontacts; ++c) {
        contactFrameForces[c] = Force{};
        if (opt.contact_model == JB_CONTACT_SPRING_DAMPER) {
            computeContactDynamicsAtFrame(c, contactFrameForces[c]);
            forceMax = std::max(forceMax, norm(contactFrameForces[c].lin));
        }
    }
    if (forceMax > 1e5) { status |= JB_ENV_CONTACT_FORCE; return JB_ERR_INVALID_ARGUMENT; }
    running = true;
    computeAllTerms(t, q.data(), v.data(), false);
    const std::vector<Force> fextNoConst = state.fExternal;
    const std::vector<double> uInternalConst = state.uInternal;
    for (int i = 0; i < INIT_ITERATIONS; ++i) {
        state.fExternal = fextNoConst;
        state.uInternal = uInternalConst;
        state.a = computeAcceleration(state.q.data(), state.v.data(), state.u, state.fExternal, i > 0, i == 0);
        for (double x : state.a) if (x != x) { status |= JB_ENV_NAN; return JB_ERR_RUNTIME; }
        computeExtraTerms();
        computeSensorMeasurements(state.q.data(), state.v.data(), state.uMotor);
        computeCommand(t, state.q.data(), state.v.data(), state.command);
        computeMotorEfforts(state.v.data(), state.command);
        computeCustom(t, state.q.data(), state.v.data());
        for (int k = 0; k < nv; ++k) state.u[k] = state.uInternal[k] + state.uCustom[k];
        for (int m = 0; m < model.nmotors; ++m) state.u[model.idx_v[model.motor_joint[m]]] += state.uTransmission[m];
    }
    computeSensorMeasurements(state.q.data(), state.v.data(), state.uMotor);
    if (mahony_enabled) mahonyInit();
    syncAccelerationsAndForces();
    q = state.q; v = state.v; a = state.a;  // syncStepperStateWithRobots
    statePrev = state;
    simStarted = true;
    return JB_OK;
}

// ============================================================================ step
// Engine::step (engine.cc:1724-2417).  No telemetry or timeout.
int Engine::step(double stepSize) {
    if (!running) return JB_ERR_BAD_CONTROL_FLOW;
    for (double x : q) if (x != x) { status |= JB_ENV_NAN; return JB_ERR_RUNTIME; }
    for (double x : v) if (x != x) { status |= JB_ENV_NAN; return JB_ERR_RUNTIME; }
    for (double x : a) if (x != x) { status |= JB_ENV_NAN; return JB_ERR_RUNTIME; }
    if (stepSize > EPS && stepSize < SIMULATION_MIN_TIMESTEP) return JB_ERR_INVALID_ARGUMENT;
    if (stepSize < EPS) {
        if (opt.controller_update_period > EPS) stepSize = opt.controller_update_period;
        else if (opt.sensors_update_period > EPS) stepSize = opt.sensors_update_period;
        else stepSize = opt.dt_max;
    }
    // Kahan-compensated end time (engine.cc:1793-1795)
    const double stepSizeCorrected = stepSize - tError;
    const double tEnd = t + stepSizeCorrected;
    tError = (tEnd - t) - stepSizeCorrected;

    uint32_t successiveIterTooLarge = 0, successiveIterFailed = 0;
    RC rc = IS_SUCCESS;
    bool isBreakpointReached = false;
    bool hasDynamicsChanged = false;
    const uint32_t failedMax = static_cast<uint32_t>(opt.successive_iter_failed_max);
    const bool finitePeriod = std::isfinite(stepperUpdatePeriod);

    auto onSuccess = [&]() {
        successiveIterTooLarge = 0; successiveIterFailed = 0;
        state.q = q; state.v = v; state.a = a;  // syncRobotsStateWithStepper
        computeExtraTerms();
        syncAccelerationsAndForces();
        ++iter;
        if (isBreakpointReached) {
            const double dtRestoreThresholdAbs = dtLargestPrev * opt.dt_restore_threshold_rel;
            if (dt < dtLargest && dtLargest < dtRestoreThresholdAbs) dtLargest = dtLargestPrev;
        }
        tPrev = t;
        dtLargestPrev = dtLargest;
        statePrev = state;
    };
    uint32_t successiveSolveFailedBackup = successiveSolveFailed;   // engine.cc:2104-2112, :2245-2252
    auto onFailure = [&]() {
        if (rc == IS_ERROR) dtLargest *= 0.1;
        if (rc == IS_FAILURE) ++successiveIterTooLarge;
        ++successiveIterFailed;
        ++iterFailed;
        successiveSolveFailed = successiveSolveFailedBackup;       // engine.cc:2211-2217
    };

    while (tEnd - t >= STEPPER_MIN_TIMESTEP) {
        double tNext = t;
        // active set and next breakpoint of the impulse forces (engine.cc:1843-1890)
        double tImpulseForceNext = INF;
        for (ImpulseForce& f : impulseForces) {
            if (t > f.t - STEPPER_MIN_TIMESTEP) { f.active = true; hasDynamicsChanged = true; }
            if (t >= f.t + f.dt - STEPPER_MIN_TIMESTEP) { f.active = false; hasDynamicsChanged = true; }
        }
        while (impulseForceBreakpointNext < impulseForceBreakpoints.size() &&
               impulseForceBreakpoints[impulseForceBreakpointNext] - t < STEPPER_MIN_TIMESTEP)
            ++impulseForceBreakpointNext;
        if (impulseForceBreakpointNext < impulseForceBreakpoints.size())
            tImpulseForceNext = std::min(tImpulseForceNext, impulseForceBreakpoints[impulseForceBreakpointNext]);
        // profile forces with a finite update period (engine.cc:1892-1917)
        if (finitePeriod) {
            for (ProfileForce& pf : profileForces) {
                if (pf.updatePeriod > EPS) {
                    const double dtNextForceUpdatePeriod = pf.updatePeriod - std::fmod(t, pf.updatePeriod);
                    if (dtNextForceUpdatePeriod < SIMULATION_MIN_TIMESTEP ||
                        pf.updatePeriod - dtNextForceUpdatePeriod < STEPPER_MIN_TIMESTEP) {
                        pf.force = pf.pending;
                        hasDynamicsChanged = true;
                    }
                }
            }
        }
        // Controller update (engine.cc:1920-1940)
        if (finitePeriod && opt.controller_update_period > EPS) {
            const double cp = opt.controller_update_period;
            const double dtNextControllerUpdatePeriod = cp - std::fmod(t, cp);
            if (dtNextControllerUpdatePeriod < SIMULATION_MIN_TIMESTEP ||
                cp - dtNextControllerUpdatePeriod < STEPPER_MIN_TIMESTEP) {
                computeCommand(t, state.q.data(), state.v.data(), state.command);
                hasDynamicsChanged = true;
            }
        }
        // Fix the FSAL issue if the dynamics has changed (continuous case, engine.cc:1973-1983)
        if (!finitePeriod && hasDynamicsChanged) {
            computeRobotsDynamics(t, q.data(), v.data(), a, true);
            syncAccelerationsAndForces();
            state.a = a;  // syncRobotsStateWithStepper(true)
            hasDynamicsChanged = false;
        }
        if (finitePeriod) {
            double dtNextGlobal;
            const double dtNextUpdatePeriod = stepperUpdatePeriod - std::fmod(t, stepperUpdatePeriod);
            if (dtNextUpdatePeriod < SIMULATION_MIN_TIMESTEP)
                dtNextGlobal = std::min(dtNextUpdatePeriod + stepperUpdatePeriod, tImpulseForceNext - t);
            else
                dtNextGlobal = std::min(dtNextUpdatePeriod, tImpulseForceNext - t);
            if (tEnd - t - STEPPER_MIN_TIMESTEP < dtNextGlobal) dtNextGlobal = tEnd - t;
            tNext += dtNextGlobal;
            while (tNext - t > STEPPER_MIN_TIMESTEP) {
                if (hasDynamicsChanged) {
                    computeRobotsDynamics(t, q.data(), v.data(), a, true);
                    syncAccelerationsAndForces();
                    state.a = a;
                    hasDynamicsChanged = false;
                }
                if (dt < STEPPER_MIN_TIMESTEP) break;
                double dtResidualThr = STEPPER_MIN_TIMESTEP;
                if (successiveIterTooLarge == 0)
                    dtResidualThr = std::clamp(0.1 * dt, STEPPER_MIN_TIMESTEP, SIMULATION_MIN_TIMESTEP);
                if (tNext - t < dt || (successiveIterTooLarge <= 1 && tNext - t < dt + dtResidualThr)) dt = tNext - t;
                if (dt > SIMULATION_MIN_TIMESTEP) {
                    const double dtResidual = std::fmod(dt, SIMULATION_MIN_TIMESTEP);
                    if (dtResidual > STEPPER_MIN_TIMESTEP &&
                        dtResidual < SIMULATION_MIN_TIMESTEP - STEPPER_MIN_TIMESTEP &&
                        dt - dtResidual > STEPPER_MIN_TIMESTEP)
                        dt -= dtResidual;
                }
                if (successiveIterFailed > failedMax) break;
                successiveSolveFailedBackup = successiveSolveFailed;
                isBreakpointReached = (dtLargest > dt);
                dtLargest = dt;
                rc = tryStep(t, dtLargest);
                if (rc == IS_SUCCESS) onSuccess(); else onFailure();
                dt = std::min(dtLargest, opt.dt_max);
            }
        } else {
            dt = std::min({dt, tEnd - t, tImpulseForceNext - t});
            isBreakpointReached = (dtLargest > dt);
            bool isStepSuccessful = false;
            while (!isStepSuccessful) {
                if (successiveIterFailed > failedMax) break;
                successiveSolveFailedBackup = successiveSolveFailed;
                dtLargest = dt;
                rc = tryStep(t, dtLargest);
                isStepSuccessful = (rc == IS_SUCCESS);
                if (isStepSuccessful) onSuccess(); else onFailure();
                dt = std::min(dtLargest, opt.dt_max);
            }
        }
        if (successiveIterFailed > failedMax) { status |= JB_ENV_ITER_FAILED; return JB_ERR_RUNTIME; }
        if (successiveSolveFailed > failedMax) { status |= JB_ENV_SOLVER_FAILED; return JB_ERR_RUNTIME; }
        if (dt < STEPPER_MIN_TIMESTEP) { status |= JB_ENV_DT_UNDERFLOW; return JB_ERR_RUNTIME; }
        // Sensors update (engine.cc:2386-2410)
        const double sp = opt.sensors_update_period;
        const double dtNextSensorsUpdatePeriod = sp - std::fmod(t, sp);
        bool mustUpdateSensors = sp < EPS;
        if (!mustUpdateSensors)
            mustUpdateSensors = dtNextSensorsUpdatePeriod < SIMULATION_MIN_TIMESTEP ||
                                sp - dtNextSensorsUpdatePeriod < STEPPER_MIN_TIMESTEP;
        if (mustUpdateSensors) {
            sensorClock = t;
            computeSensorMeasurements(state.q.data(), state.v.data(), state.uMotor);
            if (mahony_enabled) mahonyUpdate();
        }
    }
    t = tEnd;
    return JB_OK;
}

}  // namespace orc
