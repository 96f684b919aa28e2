// host_rng.h — host-side random state: the xoroshiro128+ seeder, the lookup tables and the parts of
// GapsRng that only ever run on the host (range draws for the proposal generator, Poisson update
// counts for the run loop, the same-bin exchange's truncated gamma).
//
// Reference restated: math/Random.cpp:79-175 (ranges, Poisson, exponential), :194-200 (truncGammaUpper),
// :216-260 (Xoroshiro128plus), :264-305 (GapsRandomState).
#ifndef CGB_HOST_RNG_H
#define CGB_HOST_RNG_H

#include "gaps_math.h"
#include "../../include/cogaps_b200.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace cgb {

// math/Random.cpp:216-248 — keeps exactly one previous state so a failed proposal can hand its seed back
class Xoroshiro128plus
{
public:
    explicit Xoroshiro128plus(uint64_t seed)
    {
        mState[0] = seed | 1;
        mState[1] = seed | 1;
        mPrev[0] = mPrev[1] = 0;
        for (unsigned i = 0; i < 5000; ++i) { next(); }
    }
    uint64_t next()
    {
        mPrev[0] = mState[0];
        mPrev[1] = mState[1];
        const uint64_t s0 = mState[0];
        uint64_t s1 = mState[1];
        const uint64_t result = s0 + s1;
        s1 ^= s0;
        mState[0] = rotl(s0, 24) ^ s1 ^ (s1 << 16);
        mState[1] = rotl(s1, 37);
        return result;
    }
    void rollBackOnce()
    {
        mState[0] = mPrev[0];
        mState[1] = mPrev[1];
    }
    // what operator<< archives (math/Random.cpp:250-260): the current state, not the roll-back copy
    void getState(uint64_t out[2]) const { out[0] = mState[0]; out[1] = mState[1]; }
    void setState(const uint64_t in[2]) { mState[0] = in[0]; mState[1] = in[1]; }
private:
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t mState[2];
    uint64_t mPrev[2];
};

// Lookup tables (math/Random.cpp:269-295).  The reference fills them through Boost.Math; a host that
// has Boost hands its own tables in through cgb_randstate_set_tables.  The built-in generator restates
// the definitions in f64: Phi via erfc, quantiles by bisection to a fixed point, rounded to f32 with
// the reference's float intermediates.
struct LookupTables
{
    float erf[CGB_ERF_TABLE_SIZE];
    float erfinv[CGB_ERFINV_TABLE_SIZE];
    float qgamma[CGB_QGAMMA_TABLE_SIZE];

    static double normCdf(double x) { return 0.5 * std::erfc(-x / 1.4142135623730950488016887242096980785696718753769); }
    template <class F>
    static double invert(F cdf, double p, double lo, double hi)
    {
        for (int it = 0; it < 400; ++it)
        {
            double mid = 0.5 * (lo + hi);
            if (mid == lo || mid == hi) { break; }
            if (cdf(mid) < p) { lo = mid; } else { hi = mid; }
        }
        return 0.5 * (lo + hi);
    }
    static double gamma2Cdf(double x) { return x <= 0.0 ? 0.0 : 1.0 - (1.0 + x) * std::exp(-x); }
    static float pNorm(float x) { return static_cast<float>(normCdf(static_cast<double>(x))); }
    static float qNorm(float q) { return static_cast<float>(invert(normCdf, static_cast<double>(q), -40.0, 40.0)); }
    static float qGamma2(float q)
    {
        if (q < 0.000001f) { return 0.f; } // Q_GAMMA_THRESHOLD, math/Math.cpp:10,57-60
        double hi = 1.0;
        while (gamma2Cdf(hi) < static_cast<double>(q) && hi < 1e300) { hi *= 2.0; }
        return static_cast<float>(invert(gamma2Cdf, static_cast<double>(q), 0.0, hi));
    }
    void generate()
    {
        for (unsigned i = 0; i < CGB_ERF_TABLE_SIZE; ++i)
        {
            float x = static_cast<float>(i) / 1000.f;
            erf[i] = 2.f * pNorm(x * kSqrt2) - 1.f;
        }
        for (unsigned i = 0; i < CGB_ERFINV_TABLE_SIZE - 1; ++i)
        {
            float x = static_cast<float>(i) / static_cast<float>(CGB_ERFINV_TABLE_SIZE - 1);
            erfinv[i] = qNorm((1.f + x) / 2.f) / kSqrt2;
        }
        erfinv[CGB_ERFINV_TABLE_SIZE - 1] = qNorm(1.9998f / 2.f) / kSqrt2;
        qgamma[0] = 0.f;
        for (unsigned i = 1; i < CGB_QGAMMA_TABLE_SIZE - 1; ++i)
        {
            float x = static_cast<float>(i) / static_cast<float>(CGB_QGAMMA_TABLE_SIZE - 1);
            qgamma[i] = qGamma2(x);
        }
        qgamma[CGB_QGAMMA_TABLE_SIZE - 1] = qGamma2(0.9998f);
    }
    // the built-in tables, generated once per process (13 000 bisections: 15 ms, which a cgb_run used to pay on every call)
    static const LookupTables &builtin()
    {
        static const LookupTables once = [] { LookupTables t; t.generate(); return t; }();
        return once;
    }
};

// GapsRng on the host: the PCG stream plus the draws the device never makes
struct HostRng : public Pcg
{
    HostRng() { state = 0; }
    explicit HostRng(Xoroshiro128plus &seeder)
    {
        state = seeder.next();
        advance(); // math/Random.cpp:32-38
    }
    double uniformd() { return static_cast<double>(next()) / 4294967295.0; }
    // inclusive ranges with rejection, math/Random.cpp:79-123
    uint32_t uniform32(uint32_t a, uint32_t b)
    {
        if (b == a) { return a; }
        uint32_t range = b + 1 - a;
        uint32_t x = next();
        uint32_t iPart = 0xFFFFFFFFu / range;
        while (x >= range * iPart) { x = next(); }
        return x / iPart + a;
    }
    uint64_t uniform64()
    {
        uint64_t high = (static_cast<uint64_t>(next()) << 32) & 0xFFFFFFFF00000000ull;
        uint64_t low = next();
        return high | low;
    }
    uint64_t uniform64(uint64_t a, uint64_t b)
    {
        if (b == a) { return a; }
        uint64_t range = b + 1 - a;
        uint64_t x = uniform64();
        uint64_t iPart = 0xFFFFFFFFFFFFFFFFull / range;
        while (x >= range * iPart) { x = uniform64(); }
        return x / iPart + a;
    }
    // uniform64(a, b) with iPart = UINT64_MAX / (b + 1 - a) supplied by a caller whose range never changes
    uint64_t uniform64Fixed(uint64_t a, uint64_t b, uint64_t iPart)
    {
        if (b == a) { return a; }
        const uint64_t range = b + 1 - a;
        uint64_t x = uniform64();
        while (x >= range * iPart) { x = uniform64(); }
        return (iPart == 1 ? x : x / iPart) + a;
    }
    // math/Random.cpp:125-170
    int poisson(double lambda)
    {
        if (lambda <= 5.0)
        {
            int x = 0;
            double p = uniformd();
            double cutoff = std::exp(-lambda);
            while (p >= cutoff)
            {
                p *= uniformd();
                ++x;
            }
            return x;
        }
        double c = 0.767 - 3.36 / lambda;
        double beta = 3.1415926535897932384626433832795 / std::sqrt(3.0 * lambda);
        double alpha = beta * lambda;
        double k = std::log(c) - lambda - std::log(beta);
        for (;;)
        {
            double u = uniformd();
            double x = (alpha - std::log((1.0 - u) / u)) / beta;
            double n = std::floor(x + 0.5);
            if (n < 0.0) { continue; }
            double v = uniformd();
            double y = alpha - beta * x;
            double w = 1.0 + std::exp(y);
            double lhs = y + std::log(v / (w * w));
            double rhs = k + n * std::log(lambda) - ::lgamma(n + 1);
            if (lhs <= rhs) { return static_cast<int>(n); }
        }
    }
    // math/Random.cpp:172-175 with the portable log (see gaps_math.h)
    float exponential(float lambda) { return -1.f * portable_logf(uniform()) / lambda; }
    // math/Random.cpp:194-200 — shape fixed at 2
    float truncGammaUpper(const float *qgammaTable, float b, float scale)
    {
        float upper = 1.f - std::exp(-b / scale) * (1.f + b / scale);
        const unsigned ndx = static_cast<unsigned>(uniform(0.f, upper * 5000.f));
        return qgammaTable[ndx] * scale;
    }
};

} // namespace cgb

// GapsRandomState (math/Random.h:79-98)
struct cgb_randstate
{
    explicit cgb_randstate(uint32_t seed) : seeder(seed), tables(cgb::LookupTables::builtin()), dErf(nullptr), dErfinv(nullptr), dQgamma(nullptr), device(-1) { }
    cgb::Xoroshiro128plus seeder;
    cgb::LookupTables tables;
    float *dErf;     // device copies, uploaded lazily by the first sampler
    float *dErfinv;
    float *dQgamma;
    int device;
};

struct cgb_rng
{
    cgb::HostRng rng;
    cgb_randstate *rs;
};

#endif // CGB_HOST_RNG_H
