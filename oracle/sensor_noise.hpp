// TEST INFRASTRUCTURE -- CPU oracle, not product code.
//
// Restatement of the measurement pipeline every jiminy sensor goes through after `set()` has written its true value
// (paths relative to /root/reference/core):
//   AbstractSensorTpl<T>::setAll ............ include/jiminy/core/hardware/abstract_sensor.hxx:445-522  (ring of past true values)
//   AbstractSensorTpl<T>::interpolateData ... include/jiminy/core/hardware/abstract_sensor.hxx:305-430  (delay + jitter, order 0 / 1)
//   AbstractSensorBase::measureData ......... src/hardware/abstract_sensor.cc:71-85                      (white noise, bias)
//   AbstractSensorTpl<T>::resetAll .......... include/jiminy/core/hardware/abstract_sensor.hxx:149-232   (per-sensor seeds)
//   PCG32, uniform, normal (ziggurat) ....... src/utilities/random.cc:10-170, include/jiminy/core/utilities/random.hxx:16-57
// The random streams follow the reference's algorithms (PCG32 `pcg32_fast`, float ziggurat, std::seed_seq seeding chain);
// the ONE thing its source does not define is the order in which the sensor TYPES draw their seeds from the engine's
// generator (it iterates a std::unordered_map, robot.cc:137-144): here the fixed order Imu, Force, Encoder, Effort, Contact.
// PARITY STATUS: unpinned against the reference binary (its tests hold no golden noise sequence); pinned by distribution
// moments and by the closed-form behaviour of the delay line (tests/test_sensor_pipeline.py).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <random>
#include <stdexcept>
#include <vector>

namespace orc {

struct PCG32 {   // random.cc:10-37: pcg32_fast (64-bit multiplicative congruential state, xorshift-high + random shift output)
    using result_type = uint32_t;
    uint64_t state;
    explicit PCG32(uint64_t s = 0xcafef00dd15ea5e5ULL) : state(s | 3ULL) {}
    static constexpr result_type min() { return 0; }
    static constexpr result_type max() { return 0xffffffffu; }
    result_type operator()() {
        state *= 6364136223846793005ULL;
        uint64_t s = state;
        const unsigned rshift = static_cast<unsigned>(s >> 61) & 7u;   // opBits = 3
        s ^= s >> 22;                                                  // xShift
        return static_cast<uint32_t>(s >> (22 + rshift));              // bottomSpare - randShiftMax + rshift
    }
};
// PCG32(SeedSeq&&): two 32-bit words of the sequence, low word first (random.hxx:16-46)
inline PCG32 pcg32_from_seed_seq(std::seed_seq& seq) {
    std::array<uint32_t, 2> buf;
    seq.generate(buf.begin(), buf.end());
    return PCG32(static_cast<uint64_t>(buf[0]) | (static_cast<uint64_t>(buf[1]) << 32));
}
inline float uniform01(PCG32& g) { return std::generate_canonical<float, std::numeric_limits<float>::digits>(g); }   // random.cc:41-44
inline float uniform(PCG32& g, float lo, float hi) { return std::uniform_real_distribution<float>(lo, hi)(g); }       // random.cc:46-49

struct ZigguratNormalData {   // random.cc:62-98
    std::array<uint32_t, 128> kn{};
    std::array<float, 128> fn{}, wn{};
    ZigguratNormalData() {
        constexpr double m1 = 2147483648.0, vn = 9.91256303526217e-03;
        double dn = 3.442619855899, tn = dn;
        const double q = vn / std::exp(-0.5 * dn * dn);
        kn[0] = static_cast<uint32_t>((dn / q) * m1);
        kn[1] = 0;
        wn[0] = static_cast<float>(q / m1);
        wn[127] = static_cast<float>(dn / m1);
        fn[0] = 1.0F;
        fn[127] = static_cast<float>(std::exp(-0.5 * dn * dn));
        for (uint8_t i = 126; 1 <= i; i--) {
            dn = std::sqrt(-2.0 * std::log(vn / dn + std::exp(-0.5 * dn * dn)));
            kn[i + 1] = static_cast<uint32_t>((dn / tn) * m1);
            tn = dn;
            fn[i] = static_cast<float>(std::exp(-0.5 * dn * dn));
            wn[i] = static_cast<float>(dn / m1);
        }
    }
};
inline const ZigguratNormalData& ziggurat() { static const ZigguratNormalData z; return z; }

inline float normal(PCG32& g) {   // internal::normal, random.cc:108-160
    const auto& Z = ziggurat();
    constexpr float r = 3.442620F;
    int32_t hz = static_cast<int32_t>(g());
    uint32_t iz = static_cast<uint32_t>(hz) & 127UL;
    float x, y;
    if (std::fabs(hz) < Z.kn[iz]) return static_cast<float>(hz) * Z.wn[iz];
    while (true) {
        if (iz == 0) {
            while (true) {
                x = -0.2904764F * std::log(uniform01(g));
                y = -std::log(uniform01(g));
                if (x * x <= y + y) break;
            }
            return hz <= 0 ? -r - x : r + x;
        }
        x = static_cast<float>(hz) * Z.wn[iz];
        if (Z.fn[iz] + uniform01(g) * (Z.fn[iz - 1] - Z.fn[iz]) < std::exp(-0.5F * x * x)) return x;
        hz = static_cast<int32_t>(g());
        iz = hz & 127;
        if (std::fabs(hz) < Z.kn[iz]) return static_cast<float>(hz) * Z.wn[iz];
    }
}
inline float normal(PCG32& g, float mean, float stddev) { return normal(g) * stddev + mean; }   // random.cc:163-166

constexpr int N_SENSOR_TYPES = 5;                       // Imu, Force, Encoder, Effort, Contact (the order of the observation row)
constexpr int SENSOR_FIELDS[N_SENSOR_TYPES] = {6, 6, 2, 1, 3};

struct SensorOptions {                                   // AbstractSensorOptions (abstract_sensor.h:66-100)
    std::vector<double> noiseStd, bias;                  // empty = off
    double delay = 0.0, jitter = 0.0;
    uint32_t delayInterpolationOrder = 1U;
};

// Shared storage of one sensor type (SensorSharedStorage, abstract_sensor.h:30-58): ring of past TRUE values and their times
struct SensorGroup {
    int nf = 0, ns = 0, offset = 0;                      // fields, sensors, first column in the observation row
    std::vector<SensorOptions> opt;
    std::vector<PCG32> gen;
    std::vector<double> times;                           // oldest -> newest
    std::vector<std::vector<double>> data;               // each [nf][ns], the row segment
    double delayMax = 0.0;

    void reset(uint32_t seed) {                          // resetAll (abstract_sensor.hxx:199-232)
        times.assign(1, 0.0);
        data.assign(1, std::vector<double>(static_cast<size_t>(nf) * ns, 0.0));
        delayMax = 0.0;
        for (const auto& o : opt) delayMax = std::max(delayMax, o.delay + o.jitter);
        std::seed_seq seq{seed};
        std::vector<uint32_t> seeds(ns);
        seq.generate(seeds.begin(), seeds.end());
        gen.clear();
        for (int i = 0; i < ns; ++i) gen.emplace_back(static_cast<uint64_t>(seeds[i]));
    }
    // setAll (abstract_sensor.hxx:445-510): make room for the sample at time t, `truth` = row segment of the true values
    void push(double t, const double* truth) {
        constexpr double EPS = std::numeric_limits<double>::epsilon(), SIMULATION_MAX_TIMESTEP = 0.02;
        const double timeMin = t - delayMax - SIMULATION_MAX_TIMESTEP;
        if (t + EPS > times.back()) {
            if (timeMin > times.front()) {
                std::rotate(times.begin(), times.begin() + 1, times.end());
                std::rotate(data.begin(), data.begin() + 1, data.end());
            } else {
                times.push_back(std::numeric_limits<double>::infinity());
                data.push_back(data.back());
            }
        } else {
            while (t + EPS < times.back() && times.size() > 1) { times.pop_back(); data.pop_back(); }
        }
        times.back() = t;
        std::copy(truth, truth + static_cast<size_t>(nf) * ns, data.back().begin());
    }
    // interpolateData + measureData of sensor k -> `out` (row segment of the measurements)
    void measure(int k, double* out) {
        constexpr double EPS = std::numeric_limits<double>::epsilon(), STEPPER_MIN_TIMESTEP = 1e-10;
        const SensorOptions& o = opt[k];
        PCG32& g = gen[k];
        const double delay = o.delay + uniform(g, 0.0F, static_cast<float>(o.jitter));
        double timeDesired = times.back() - delay;
        if (o.delayInterpolationOrder == 0) timeDesired += STEPPER_MIN_TIMESTEP;
        const std::ptrdiff_t n = static_cast<std::ptrdiff_t>(times.size());
        auto bisectLeft = [&]() -> std::ptrdiff_t {
            std::ptrdiff_t left = 0, right = n - 1, mid = 0;
            if (timeDesired >= times.back()) return right;
            if (timeDesired < times.front()) return -1;
            while (left < right) {
                mid = (left + right) / 2;
                if (timeDesired < times[mid]) right = mid;
                else if (timeDesired > times[mid]) left = mid + 1;
                else return mid;
            }
            return timeDesired < times[mid] ? mid - 1 : mid;
        };
        const std::ptrdiff_t idxLeft = bisectLeft();
        auto col = [&](std::ptrdiff_t i, int f) { return data[i][static_cast<size_t>(f) * ns + k]; };
        for (int f = 0; f < nf; ++f) {
            double val;
            if (timeDesired >= 0.0 && idxLeft + 1 < n) {
                if (idxLeft < 0) throw std::runtime_error("No data old enough is available.");
                if (o.delayInterpolationOrder == 0) val = col(idxLeft, f);
                else {
                    const double ratio = (timeDesired - times[idxLeft]) / (times[idxLeft + 1] - times[idxLeft]);
                    val = col(idxLeft, f) + ratio * (col(idxLeft + 1, f) - col(idxLeft, f));
                }
            } else if (o.delay > EPS || o.jitter > EPS) {
                // the buffer is not old enough yet: oldest non-initial value
                std::ptrdiff_t index = n - 1;
                for (std::ptrdiff_t i = 0; i < n; ++i) if (times[i] > 0) { index = std::max<std::ptrdiff_t>(0, i - 1); break; }
                val = col(index, f);
            } else val = col(n - 1, f);
            out[static_cast<size_t>(f) * ns + k] = val;
        }
        if (!o.noiseStd.empty())
            for (int f = 0; f < nf; ++f) out[static_cast<size_t>(f) * ns + k] += static_cast<double>(normal(g, 0.0F, static_cast<float>(o.noiseStd[f])));
        if (!o.bias.empty())
            for (int f = 0; f < nf; ++f) out[static_cast<size_t>(f) * ns + k] += o.bias[f];
    }
};

}  // namespace orc
