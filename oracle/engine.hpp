// TEST INFRASTRUCTURE -- CPU oracle, not product code.  See oracle/README.md.
//
// Scalar, single-robot, single-thread C++17 restatement of the jiminy step path, following the
// reference function by function (citations relative to /root/reference):
//   Engine::step scheduler ............ core/src/engine/engine.cc:1724-2417
//   Engine::start ..................... core/src/engine/engine.cc:952-1533
//   computeRobotsDynamics (the RHS) ... core/src/engine/engine.cc:3585-3708
//   computeForwardKinematics .......... core/src/engine/engine.cc:2957-3014
//   contact model ..................... core/src/engine/engine.cc:3117-3238, utilities/pinocchio.cc:794-809
//   computeExtraTerms ................. core/src/engine/engine.cc:800-905
//   ABA with rotor inertia ............ core/include/jiminy/core/robot/pinocchio_overload_algorithms.h:126-489
//   steppers .......................... core/src/stepper/*.cc, lie_group.h:446-471
//   SimpleMotor ....................... core/src/hardware/basic_motors.cc:83-143
//   sensors ........................... core/src/hardware/basic_sensors.cc (set() of each type)
// The Pinocchio 2.7.0 primitives it relies on (not vendored in the reference) are restated in
// spatial.hpp and below (joint calc, forwardKinematics, AbaForwardStep1/2, integrate, difference).
//
// PARITY STATUS: pinned by the reference's analytical tests re-implemented in tests/ (pendulum /
// spring-mass vs expm, contact equilibrium, friction steady state, IMU analytics, energy drift).
// Parity against trajectories *produced by the reference binary* is UNPINNED: jiminy cannot be
// built or imported in this image (no Eigen/Boost/Pinocchio, no network) and ships no golden
// trajectories.
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <vector>

#include "../include/jiminy_b200.h"
#include "spatial.hpp"
#include "sensor_noise.hpp"

namespace orc {

constexpr double INF = std::numeric_limits<double>::infinity();
constexpr double EPS = std::numeric_limits<double>::epsilon();
constexpr double STEPPER_MIN_TIMESTEP = 1e-10;     // core/include/jiminy/core/constants.h:18-20
constexpr double SIMULATION_MIN_TIMESTEP = 1e-6;
constexpr double SIMULATION_MAX_TIMESTEP = 0.02;
constexpr int INIT_ITERATIONS = 4;                 // engine.cc:61

typedef void (*ControllerFn)(void* ctx, double t, const double* q, const double* v, const double* sensors,
                             double* command);
typedef void (*InternalDynFn)(void* ctx, double t, const double* q, const double* v, const double* sensors,
                              double* u_custom);

struct Model {
    int njoints = 0, nq = 0, nv = 0;
    std::vector<int> jtype, parent, idx_q, idx_v;
    std::vector<SE3> placement;
    std::vector<V3> axis;
    std::vector<Inertia> inertia;
    std::vector<double> rotor, q_lower, q_upper;
    std::vector<double> flexibility;   // [njoints][6] stiffness | damping of the spherical flexibility joints (empty: none)
    int nmotors = 0;
    std::vector<int> motor_joint, motor_flags;
    std::vector<double> motor_params;
    int ncontacts = 0;
    std::vector<int> contact_joint;
    std::vector<SE3> contact_placement;
    int nimu = 0, nforce = 0, nenc = 0, neff = 0, ncs = 0;
    std::vector<int> imu_joint, force_joint, enc_joint, eff_motor, cs_index;
    std::vector<SE3> imu_placement, force_placement;
    std::vector<double> enc_reduction;
    // ForceSensor::refreshProxies (basic_sensors.cc:324-351)
    std::vector<std::vector<std::pair<int, SE3>>> force_contacts;
    JbSensorLayout layout{};

    static int nvj(int t) { return t == JB_JOINT_UNIVERSE ? 0 : (t == JB_JOINT_FREEFLYER ? 6 : (t == JB_JOINT_SPHERICAL ? 3 : 1)); }
    static int nqj(int t) {
        if (t == JB_JOINT_UNIVERSE) return 0;
        if (t == JB_JOINT_FREEFLYER) return 7;
        if (t == JB_JOINT_SPHERICAL) return 4;
        return (t >= JB_JOINT_RUBX && t <= JB_JOINT_RUBU) ? 2 : 1;
    }
    static bool is_unbounded(int t) { return t >= JB_JOINT_RUBX && t <= JB_JOINT_RUBU; }
    static bool is_revolute(int t) { return t >= JB_JOINT_RX && t <= JB_JOINT_RUBU; }
    static bool is_prismatic(int t) { return t >= JB_JOINT_PX && t <= JB_JOINT_PU; }
};

inline SE3 se3_from12(const double* x) {
    SE3 M;
    for (int i = 0; i < 9; ++i) M.R.m[i] = x[i];
    M.p = V3(x[9], x[10], x[11]);
    return M;
}
inline SE3 se3_inverse(const SE3& M) { SE3 r; r.R = transpose(M.R); r.p = -(tmul(M.R, M.p)); return r; }

Model make_model(const JbModelDesc& d);

// Per-joint scratch, the subset of pinocchio::Data + JointData the path touches.
struct JointData {
    SE3 M;       // jdata.M()
    Motion vJ;   // jdata.v()
    // jdata.c() == 0 for every supported joint type
    double S[6][6];     // motion subspace columns (nv_j columns used)
    double U[6][6], Dinv[6][6], UDinv[6][6];
};

struct Data {
    std::vector<SE3> liMi, oMi;
    std::vector<Motion> v, a, a_gf;
    std::vector<Force> f, h;
    std::vector<M6> Yaba;
    std::vector<JointData> joints;
    std::vector<double> u, ddq;
    double kinetic_energy = 0, potential_energy = 0;
    // computeExtraTerms (engine.cc:817-832, :890-904): subtree inertias, subtree centres of mass and their velocities,
    // centroidal momentum and its derivative; `mass` = subtree masses (pinocchio::centerOfMass at model set-up, model.cc:269)
    std::vector<Inertia> Ycrb;
    std::vector<V3> com, vcom;
    std::vector<double> mass;
    Force hg, dhg;
};

struct RobotState {
    std::vector<double> q, v, a, command, u, uMotor, uTransmission, uInternal, uCustom;
    std::vector<Force> fExternal;
};

struct Options : JbOptions {};

struct Engine {
    Model model;
    Options opt;
    Data data;
    RobotState state, statePrev;
    // stepper state (engine.h:216-250)
    int64_t iter = 0, iterFailed = 0;
    double t = 0, tPrev = 0, tError = 0, dt = INF, dtLargest = INF, dtLargestPrev = INF;
    std::vector<double> q, v, a;  // StepperState.qSplit[0], vSplit[0], aSplit[0]
    double stepperUpdatePeriod = INF;
    bool running = false;
    bool flexAngleError = false;   // a flexibility joint went beyond 0.95 pi (the reference throws, engine.cc:3381-3385)
    int status = JB_ENV_NOT_STARTED;

    std::vector<Force> contactFrameForces;  // RobotData::contactFrameForces (in parent joint frame)
    std::vector<Force> contactForces;       // robot->contactForces_ (in contact frame)
    std::vector<Force> contactForcesPrev, fPrev;
    std::vector<Motion> aPrev;
    std::vector<Force> fExtBuffer;          // `fPrev_` argument of computeExtraTerms used as fExt buffer
    std::vector<double> sensors;            // true values written by the sensors' set() (AbstractSensorTpl::data())
    // measurement pipeline (oracle/sensor_noise.hpp): off until a sensor option is set; `measurements` = what
    // robot.sensor_measurements exposes (delayed, noisy, biased), `sensors` keeps the true values
    bool sensorPipeline = false;
    uint32_t engineSeed = 0;                // stepper.randomSeedSeq = [engineSeed]
    std::array<SensorGroup, N_SENSOR_TYPES> sensorGroups;
    std::vector<double> measurements;
    double sensorClock = 0.0;               // time of the refresh in progress
    void setSensorOptions(int type, int index, const double* noiseStd, const double* bias, double delay, double jitter, uint32_t order);
    void resetSensorPipeline();
    void measureSensors();
    const std::vector<double>& sensorOutput() const { return sensorPipeline ? measurements : sensors; }
    // ---- kinematic constraints (oracle/constraints.cpp)
    struct Constraint {
        int kind = 0;          // 0: JointConstraint (bounds), 1: FrameConstraint {x, y, z, rot z} (contact frame)
        int joint = 0, contact = -1, dim = 1;
        bool enabled = false;
        double kp = 0.0, kd = 0.0;                  // Baumgarte gains
        double lambda[4] = {0, 0, 0, 0};
        double qRef = 0.0; bool reversed = false;   // JointConstraint
        SE3 transformRef = SE3::identity(); V3 normal; M3 rotationLocal = M3::identity();   // FrameConstraint
        std::vector<double> jac;                    // dim x nv
        double drift[4] = {0, 0, 0, 0};
        int startIndex = 0;
    };
    std::vector<Constraint> constraints;
    int rowsMax = 0;
    uint32_t successiveSolveFailed = 0;
    int64_t pgsIterations = 0;
    std::vector<int32_t> pgsHistory;            // iterations of every solve (diagnostics: tools/pgs_iteration_stats.py)
    bool keepPgsHistory = false;
    std::vector<double> Mmat, Mchol, Jworld, nle, torqueResidual;
    std::vector<Motion> aDrift;
    std::vector<double> solverJ, solverGamma, solverLambda, solverB, solverY, solverYPrev, solverA;
    void buildConstraints();
    bool hasConstraints() const;
    void resetConstraints(const double* q);
    void updateJointBoundConstraints(const double* q);
    void updateContactConstraint(int contact);
    void computeCrba();
    void computeNle();
    void computeConstraints(const double* q, const double* v);
    void pgsIter(int m, const std::vector<double>& A, const double* b, double w, double* x);
    bool pgsSolve(int m, const std::vector<double>& A, const double* b, double* x);
    bool solveBoxedForwardDynamics(double dampingInv, bool isStateUpToDate, bool ignoreBounds);
    const std::vector<double>& computeAcceleration(const double* q, const double* v, std::vector<double>& u,
                                                   std::vector<Force>& fext, bool isStateUpToDate, bool ignoreBounds);

    ControllerFn controller = nullptr;      // computeCommand functor (may be null: ZOH of `state.command`)
    InternalDynFn internalDyn = nullptr;    // internalDynamics functor
    void* ctx = nullptr;
    // Engine::registerImpulseForce / registerProfileForce (engine.cc:2450-2567).  A frame is given by its
    // parent joint and its translation in the joint frame (the wrench is expressed in world-aligned axes at
    // the frame origin, so the frame rotation never matters).  A profile force holds the value its
    // "function" (the caller's buffer, `pending`) returned at the last update.
    struct ImpulseForce { int joint; V3 p; double t, dt; Force F; bool active = false; };
    struct ProfileForce { int joint; V3 p; double updatePeriod; Force pending, force; };
    std::vector<ImpulseForce> impulseForces;
    std::vector<ProfileForce> profileForces;
    std::vector<double> impulseForceBreakpoints;   // sorted, unique (std::set in the reference)
    size_t impulseForceBreakpointNext = 0;
    int registerImpulseForce(int joint, const double* p, double t, double dt, const double* F);
    int registerProfileForce(int joint, const double* p, double updatePeriod);
    void removeAllForces();
    void refreshStepperUpdatePeriod();
    Force convertForceGlobalFrameToJoint(int joint, const V3& p, const Force& F) const;
    void computeExternalForces(std::vector<Force>& fext);

    std::vector<double> spring_k, spring_d; // built-in linear internal dynamics u = -k q - d v (1-dof joints)
    // built-in discrete PD controller (gym_jiminy PDController, order-0 target): command buffer = targets
    bool pd_enabled = false;
    // gym_jiminy PDController block (+ optional MotorSafetyLimit), oracle/controllers.cpp: the command buffer holds
    // target motor accelerations; (position, velocity, acceleration) targets are integrated at controller updates
    bool pdf_enabled = false, pdf_safety = false, simStarted = false;
    // gym_jiminy MahonyFilter observer (oracle/controllers.cpp): attitude estimate of every IMU, refreshed with the sensors
    bool mahony_enabled = false;
    double mahony_kp = 1.0, mahony_ki = 0.1;
    std::vector<double> mahony_q, mahony_bias, mahony_omega;   // [4][nimu], [3][nimu], [3][nimu]
    void mahonyInit();
    void mahonyUpdate();
    std::vector<double> pdf_kp, pdf_kd, pdf_lower, pdf_upper, pdf_state, pdf_action, pdf_skp, pdf_skd, pdf_slo, pdf_shi, pdf_svlim;
    std::vector<double> pd_kp, pd_kd, pd_target;
    int64_t rhs_count = 0;

    // stepper buffers
    struct Deriv { std::vector<double> v, a; };
    std::vector<Deriv> ki;
    Deriv inc, scale, err;
    std::vector<double> qBuf, vBuf, qCand, vCand, qOther, vOther, aOut;

    Engine(const JbModelDesc& d, const JbOptions& o);
    void set_options(const JbOptions& o);
    int start(const double* q0, const double* v0);
    int step(double stepSize);
    void stop() { running = false; }

    // ---- physics
    void jointCalc(int i, const double* q, const double* v);
    void forwardKinematics(const double* q, const double* v, const double* a);
    void computeContactDynamicsAtFrame(int c, Force& fextLocal) const;
    V3 computeContactDynamics(const V3& nGround, double depth, const V3& vContactInWorld) const;
    void computeInternalDynamics(const double* q, const double* v, std::vector<double>& uInternal);
    void computeCollisionForces(std::vector<Force>& fext, bool isStateUpToDate);
    void computeAllTerms(double t, const double* q, const double* v, bool isStateUpToDate);
    void computeCommand(double t, const double* q, const double* v, std::vector<double>& command);
    void computeMotorEfforts(const double* v, const std::vector<double>& command);
    void computeCustom(double t, const double* q, const double* v);
    const std::vector<double>& aba(const double* q, const double* v, const std::vector<double>& tau,
                                   const std::vector<Force>& fext);
    void computeRobotsDynamics(double t, const double* q, const double* v, std::vector<double>& aOut,
                               bool isStateUpToDate);
    void computeExtraTerms();
    void syncAccelerationsAndForces();
    void computeSensorMeasurements(const double* q, const double* v, const std::vector<double>& uMotor);

    // ---- Lie group (pinocchio::integrate / difference / normalize)
    void integrate(const double* q, const double* vel, double* out) const;
    void difference(const double* q0, const double* q1, double* out) const;
    void neutral(double* q) const;
    void normalize(double* q) const;

    // ---- steppers
    enum RC { IS_SUCCESS, IS_FAILURE, IS_ERROR };
    RC tryStep(double& t, double& dt);
    bool tryStepEuler(double t, double& dt);
    bool tryStepRK(double t, double& dt);
    bool adjustStepDopri(double& dt);
    double computeErrorDopri(double dt);
    void f(double t, const std::vector<double>& q, const std::vector<double>& v, Deriv& out);
};

// gym_jiminy controller blocks (oracle/controllers.cpp)
void integrate_zoh(double* state, const double* state_min, const double* state_max, int n, double dt);
void pd_controller(const double* encoder_data, double* command_state, const double* lower, const double* upper,
                   const double* kp, const double* kd, const double* effort_limit, int n, double control_dt, double* out);
void mahony_filter(double* q, double* omega, const double* gyro, const double* acc, double* bias_hat, int M, double kp, double ki, double dt);
void matrix_to_quat_ref(const double* R, double* out);
void apply_safety_limits(const double* command, const double* q, const double* v, const double* kp, const double* kd,
                         const double* soft_lower, const double* soft_upper, const double* velocity_limit,
                         const double* effort_limit, int n, double* out);

}  // namespace orc
