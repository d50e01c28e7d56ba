// Host-side lane planner, see jb_plan.h.
#include "jb_plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <numeric>
#include <sstream>
#include <stdexcept>

namespace jb {

static int joint_nv(int t) { return t == JB_JOINT_UNIVERSE ? 0 : (t == JB_JOINT_FREEFLYER ? 6 : (t == JB_JOINT_SPHERICAL ? 3 : 1)); }

static int rec_kind(int t) {
    switch (t) {
        case JB_JOINT_RX: case JB_JOINT_RY: case JB_JOINT_RZ: case JB_JOINT_RU: return REC_REV;
        case JB_JOINT_RUBX: case JB_JOINT_RUBY: case JB_JOINT_RUBZ: case JB_JOINT_RUBU: return REC_REVU;
        case JB_JOINT_PX: case JB_JOINT_PY: case JB_JOINT_PZ: case JB_JOINT_PU: return REC_PRISM;
        case JB_JOINT_FREEFLYER: return REC_FREE;
        case JB_JOINT_SPHERICAL: return REC_SPH;
        default: throw std::invalid_argument("unknown joint type");
    }
}

static void se3_mul(const double* a, const double* b, double* out) {  // out = a * b (R row-major, p)
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            out[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
    for (int i = 0; i < 3; ++i) out[9 + i] = a[9 + i] + a[3 * i] * b[9] + a[3 * i + 1] * b[10] + a[3 * i + 2] * b[11];
}
static void se3_inv(const double* a, double* out) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out[3 * i + j] = a[3 * j + i];
    for (int i = 0; i < 3; ++i) out[9 + i] = -(a[i] * a[9] + a[3 + i] * a[10] + a[6 + i] * a[11]);
}

Plan build_plan(const JbModelDesc& m, int lanes, int n_hist) {
    const int n = m.njoints;
    if (n < 2) throw std::invalid_argument("model has no joint");
    std::vector<std::vector<int>> children(n);
    for (int j = 1; j < n; ++j) {
        if (m.parent[j] < 0 || m.parent[j] >= j) throw std::invalid_argument("parents must precede children");
        if (m.joint_type[j] == JB_JOINT_FREEFLYER && m.parent[j] != 0)
            throw std::invalid_argument("a free-flyer joint must be attached to the universe");
        children[m.parent[j]].push_back(j);
    }
    std::vector<int> size(n, 1), leaves(n, 0);
    for (int j = n - 1; j >= 1; --j) {
        if (children[j].empty()) leaves[j] = 1;
        size[m.parent[j]] += size[j];
        leaves[m.parent[j]] += leaves[j];
    }

    // ---- choose trunk / subtrees
    std::vector<char> is_trunk(n, 0);
    std::vector<int> roots = children[0];
    const int want = lanes > 0 ? lanes : 8;
    while (static_cast<int>(roots.size()) < want) {
        // split the largest subtree that still branches somewhere
        int best = -1;
        for (size_t k = 0; k < roots.size(); ++k)
            if (leaves[roots[k]] >= 2 && (best < 0 || size[roots[k]] > size[roots[best]])) best = static_cast<int>(k);
        if (best < 0) break;
        const int r = roots[best];
        is_trunk[r] = 1;
        roots.erase(roots.begin() + best);
        for (int c : children[r]) roots.push_back(c);
    }
    int L = lanes;
    if (L <= 0) {
        L = 1;
        while (L * 2 <= static_cast<int>(roots.size()) && L < 8) L *= 2;
        // 5..7 subtrees: 4 lanes with two subtrees sharing a lane beats 8 lanes with idle ones
    }
    if (L != 1 && L != 2 && L != 4 && L != 8) throw std::invalid_argument("lanes must be 1, 2, 4 or 8");
    if (L == 1) {  // no trunk needed: everything private to the single lane
        std::fill(is_trunk.begin(), is_trunk.end(), 0);
        roots = children[0];
    }
    // LPT assignment of subtrees to lanes
    std::vector<int> order(roots.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return size[roots[a]] > size[roots[b]]; });
    std::vector<int> load(L, 0);
    std::vector<int> joint_lane(n, -1);
    std::function<void(int, int)> mark = [&](int j, int lane) {
        joint_lane[j] = lane;
        for (int c : children[j]) mark(c, lane);
    };
    for (int k : order) {
        int lane = static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin());
        load[lane] += size[roots[k]];
        mark(roots[k], lane);
    }

    Plan P;
    P.L = L;
    P.joint_lane = joint_lane;
    std::vector<int> trunk;
    for (int j = 1; j < n; ++j)
        if (is_trunk[j]) trunk.push_back(j);
    P.ntrunk = static_cast<int>(trunk.size());
    std::vector<std::vector<int>> priv(L);
    for (int j = 1; j < n; ++j)
        if (!is_trunk[j]) priv[joint_lane[j]].push_back(j);
    size_t nslot = 0;
    for (auto& p : priv) nslot = std::max(nslot, p.size());
    P.nrec = P.ntrunk + static_cast<int>(nslot);

    // per-lane joint list (record -> joint, -1 pad) and joint -> record
    std::vector<std::vector<int>> rec_joint(L, std::vector<int>(P.nrec, -1));
    std::vector<std::vector<int>> joint_rec(L, std::vector<int>(n, -1));
    for (int s = 0; s < L; ++s) {
        for (int k = 0; k < P.ntrunk; ++k) rec_joint[s][k] = trunk[k];
        for (size_t k = 0; k < priv[s].size(); ++k) rec_joint[s][P.ntrunk + k] = priv[s][k];
        for (int r = 0; r < P.nrec; ++r)
            if (rec_joint[s][r] >= 0) joint_rec[s][rec_joint[s][r]] = r;
    }

    // ---- pool entries.  A joint needs a pool entry when some child cannot use the register
    // carry (child record != parent record + 1) or when it is a trunk joint.  Pool indices are
    // assigned per lane; trunk joints get the same index on every lane.
    std::vector<std::vector<int>> pool_of(L, std::vector<int>(n, -1));
    int npool_trunk = 0;
    for (int j : trunk) {
        for (int s = 0; s < L; ++s) pool_of[s][j] = npool_trunk;
        ++npool_trunk;
    }
    int npool = npool_trunk;
    for (int s = 0; s < L; ++s) {
        int next = npool_trunk;
        for (int j : priv[s]) {
            bool need = false;
            for (int c : children[j])
                if (joint_rec[s][c] != joint_rec[s][j] + 1) need = true;
            if (need) pool_of[s][j] = next++;
        }
        npool = std::max(npool, next);
    }
    P.npool = npool;

    // ---- contact slots per lane
    std::vector<std::vector<int>> lane_contacts(L);
    for (int c = 0; c < m.ncontacts; ++c) {
        const int j = m.contact_joint[c];
        if (j <= 0) continue;  // contact frame attached to the universe: never moves, ignored
        for (int s = 0; s < L; ++s)
            if (is_trunk[j] || joint_lane[j] == s) lane_contacts[s].push_back(c);
    }
    // group by joint so that each record's contacts are contiguous
    for (int s = 0; s < L; ++s)
        std::stable_sort(lane_contacts[s].begin(), lane_contacts[s].end(),
                         [&](int a, int b) { return m.contact_joint[a] < m.contact_joint[b]; });
    size_t ncs = 0;
    for (auto& lc : lane_contacts) ncs = std::max(ncs, lc.size());
    P.ncslot = static_cast<int>(ncs);

    // ---- record layout (lane-uniform offsets)
    P.rec_off.resize(P.nrec);
    P.rec_free.assign(P.nrec, 0);
    P.trunk_reduce.assign(P.nrec, 0);
    int off = 0;
    for (int r = 0; r < P.nrec; ++r) {
        bool any_free = false, any_sph = false;
        for (int s = 0; s < L; ++s) {
            if (rec_joint[s][r] >= 0 && m.joint_type[rec_joint[s][r]] == JB_JOINT_FREEFLYER) any_free = true;
            if (rec_joint[s][r] >= 0 && m.joint_type[rec_joint[s][r]] == JB_JOINT_SPHERICAL) any_sph = true;
        }
        P.rec_free[r] = any_free || any_sph;
        P.rec_off[r] = off;
        off += (any_free || any_sph) ? (RF_KA + 6 * n_hist + (any_sph ? RS_EXTRA : 0)) : (R1_KA + n_hist);
    }
    P.sph_off = RF_KA + 6 * n_hist;
    P.pool_off = off;
    off += POOL_SIZE * P.npool;
    P.cslot_off = off;
    off += CSLOT_SIZE * P.ncslot;
    // IMU capture slots: one per IMU-carrying joint of the lane
    {
        int nmax = 0;
        for (int s = 0; s < L; ++s) {
            int cnt = 0;
            std::vector<char> seen(n, 0);
            for (int i = 0; i < m.nimu; ++i) {
                const int j = m.imu_joint[i];
                if (j <= 0) throw std::invalid_argument("IMU attached to the universe is not supported");
                if (seen[j]) throw std::invalid_argument("several IMUs on one joint are not supported");
                seen[j] = 1;
                if (is_trunk[j] || joint_lane[j] == s) ++cnt;
            }
            nmax = std::max(nmax, cnt);
        }
        P.nimuslot = nmax;
    }
    P.imu_off = off;
    off += IMUSLOT_SIZE * P.nimuslot;
    P.nfields = off;

    // ---- tables
    std::vector<double> subtree_mass(m.njoints, 0.0);
    for (int j = 1; j < m.njoints; ++j) subtree_mass[j] = m.inertia[10 * j];
    for (int j = m.njoints - 1; j > 0; --j) subtree_mass[m.parent[j]] += subtree_mass[j];
    P.total_mass = subtree_mass[0];
    P.rint.assign(static_cast<size_t>(P.nrec) * L, RecInt{});
    P.rdbl.assign(static_cast<size_t>(P.nrec) * L, RecDbl{});
    P.cslots.assign(static_cast<size_t>(P.ncslot) * L, ContactSlot{});
    std::vector<int> lane_imu_count(L, 0);
    for (int s = 0; s < L; ++s) {
        for (size_t k = 0; k < static_cast<size_t>(P.ncslot); ++k) {
            ContactSlot& cs = P.cslots[k * L + s];
            cs.contact = -1; cs.sensor = -1; cs.force = -1;
            if (k >= lane_contacts[s].size()) continue;
            const int c = lane_contacts[s][k];
            cs.contact = c;
            std::memcpy(cs.placement, m.contact_placement + 12 * c, sizeof cs.placement);
            for (int i = 0; i < m.ncontact_sensor; ++i)
                if (m.contact_sensor_index[i] == c) cs.sensor = i;
            for (int f = 0; f < m.nforce; ++f)
                if (m.force_joint[f] == m.contact_joint[c]) {
                    // ForceSensor::refreshProxies: contactPlacementRel = frame.placement.actInv(contact.placement)
                    // NB: a contact feeds every force sensor on the same joint; the table keeps the first,
                    //     additional sensors on the same joint are rejected at batch creation.
                    if (cs.force < 0) {
                        double inv[12], rel[12];
                        se3_inv(m.force_placement + 12 * f, inv);
                        se3_mul(inv, m.contact_placement + 12 * c, rel);
                        cs.force = f;
                        std::memcpy(cs.force_R, rel, sizeof cs.force_R);
                        std::memcpy(cs.force_p, rel + 9, sizeof cs.force_p);
                    }
                }
        }
        for (int r = 0; r < P.nrec; ++r) {
            RecInt& ri = P.rint[static_cast<size_t>(r) * L + s];
            RecDbl& rd = P.rdbl[static_cast<size_t>(r) * L + s];
            const int j = rec_joint[s][r];
            ri.kind = REC_PAD; ri.joint = -1; ri.parent_rec = -1; ri.pool = -1; ri.parent_pool = -1;
            ri.motor = -1; ri.imu = -1; ri.encoder = -1; ri.effort = -1; ri.contact0 = 0; ri.ncontact = 0;
            ri.imu_slot = -1;
            if (j < 0) continue;
            const int p = m.parent[j];
            ri.kind = rec_kind(m.joint_type[j]);
            if (ri.kind == REC_REV && std::fabs(m.axis[3 * j]) == 1.0 && m.axis[3 * j + 1] == 0.0 && m.axis[3 * j + 2] == 0.0)
                ri.kind = REC_REVX;
            ri.joint = j;
            ri.parent_rec = p > 0 ? joint_rec[s][p] : -1;
            const bool trunk_j = is_trunk[j];
            // forward carry: parent is the previous record (never across the trunk/private border for
            // simplicity: trunk children always read the pool)
            ri.carry_in = (p > 0 && ri.parent_rec == r - 1 && !is_trunk[p]) ? 1 : 0;
            if (p > 0 && trunk_j && ri.parent_rec == r - 1) ri.carry_in = 0;  // trunk -> trunk goes through pool
            ri.pool = pool_of[s][j];
            ri.parent_pool = p > 0 ? pool_of[s][p] : -1;
            ri.carry_out = ri.carry_in;
            if (p > 0 && !ri.carry_in && ri.parent_pool < 0) throw std::logic_error("planner: missing parent pool");
            ri.idx_q = m.idx_q[j];
            ri.idx_v = m.idx_v[j];
            ri.owner = (!trunk_j || s == 0) ? 1 : 0;
            ri.has_limit = (ri.kind == REC_REV || ri.kind == REC_REVX || ri.kind == REC_PRISM) ? 1 : 0;
            std::memcpy(rd.placement, m.placement + 12 * j, sizeof rd.placement);
            std::memcpy(rd.axis, m.axis + 3 * j, sizeof rd.axis);
            std::memcpy(rd.inertia, m.inertia + 10 * j, sizeof rd.inertia);
            if (joint_nv(m.joint_type[j]) == 1) {
                rd.armature = m.rotor_inertia[m.idx_v[j]];
                rd.q_lo = m.q_lower[m.idx_q[j]];
                rd.q_hi = m.q_upper[m.idx_q[j]];
            } else if (m.joint_type[j] == JB_JOINT_SPHERICAL) {
                for (int k = 0; k < 3; ++k) rd.axis[k] = m.rotor_inertia[m.idx_v[j] + k];
                for (int k = 0; k < 6; ++k) rd.motor[k] = m.flexibility ? m.flexibility[6 * j + k] : 0.0;
            } else {
                for (int k = 0; k < 6; ++k)
                    if (m.rotor_inertia[m.idx_v[j] + k] != 0.0)
                        throw std::invalid_argument("rotor inertia on free-flyer dofs is not supported");
            }
            rd.enc_reduction = 1.0;
            rd.subtree_mass = subtree_mass[j];
            for (int mm = 0; mm < m.nmotors; ++mm)
                if (m.motor_joint[mm] == j) {
                    if (ri.motor >= 0) throw std::invalid_argument("several motors on one joint are not supported");
                    ri.motor = mm;
                    ri.motor_flags = m.motor_flags[mm];
                    std::memcpy(rd.motor, m.motor_params + 10 * mm, sizeof rd.motor);
                    {   // reciprocal of the velocity taper span of SimpleMotor::computeEffort (basic_motors.cc:110-118)
                        const double effLim = rd.motor[1], velLim = rd.motor[2], velocityDelta = effLim * rd.motor[3];
                        const double thr = std::max(velLim - velocityDelta, 0.0);
                        rd.motor[9] = 1.0 / (velLim - thr);
                        rd.pad = thr;   // |vMotor| <= thr: no taper
                    }
                }
            for (int e = 0; e < m.nencoder; ++e)
                if (m.encoder_joint[e] == j && ri.encoder < 0) { ri.encoder = e; rd.enc_reduction = m.encoder_reduction[e]; }
            for (int e = 0; e < m.neffort; ++e)
                if (ri.motor >= 0 && m.effort_motor[e] == ri.motor && ri.effort < 0) ri.effort = e;
            for (int i = 0; i < m.nimu; ++i)
                if (m.imu_joint[i] == j && ri.imu < 0) { ri.imu = i; ri.imu_slot = lane_imu_count[s]++; }
            // contacts of this joint in the lane's slot list
            int c0 = -1, nc = 0;
            for (size_t k = 0; k < lane_contacts[s].size(); ++k)
                if (m.contact_joint[lane_contacts[s][k]] == j) { if (c0 < 0) c0 = static_cast<int>(k); ++nc; }
            ri.contact0 = c0 < 0 ? 0 : c0;
            ri.ncontact = nc;
        }
        // take_carry: record r consumes the carry produced by record r+1 when r+1 carries out to r
        for (int r = 0; r + 1 < P.nrec; ++r) {
            RecInt& ri = P.rint[static_cast<size_t>(r) * L + s];
            const RecInt& rn = P.rint[static_cast<size_t>(r + 1) * L + s];
            ri.take_carry = (rn.joint >= 0 && rn.carry_out && rn.parent_rec == r) ? 1 : 0;
        }
    }
    // Trunk records all-reduce their pool accumulator when they have any child
    for (int k = 0; k < P.ntrunk; ++k) P.trunk_reduce[k] = children[trunk[k]].empty() ? 0 : 1;
    return P;
}

std::string Plan::describe() const {
    std::ostringstream os;
    os << "lanes=" << L << " nrec=" << nrec << " ntrunk=" << ntrunk << " npool=" << npool << " ncslot=" << ncslot
       << " fields/lane=" << nfields << " (" << nfields * 8 * 32 << " B shared per warp)";
    return os.str();
}

}  // namespace jb
