// gaps_math.h — scalar arithmetic shared by the host generator and the device epilogue.
//
// Everything here is written so that the host (g++/nvcc host pass) and the device (sm_100a) produce
// the SAME BITS: only IEEE-754 round-to-nearest add/sub/mul/div/sqrt/fma, spelled with the CUDA
// `__f*_rn` intrinsics on the device so ptxas can never contract a mul+add pair into an FMA, and with
// plain operators on the host (the library is built with -Xcompiler -ffp-contract=off).
//
// Reference semantics restated (file:line under the reference's src/):
//   PCG32 XSH-RR stream ............ math/Random.cpp:32-66
//   uniform / ranges ............... math/Random.cpp:58-123
//   p_norm_fast / q_norm_fast ...... math/Random.cpp:307-345
//   truncNormal .................... math/Random.cpp:178-191
//   gibbsMass ...................... gibbs_sampler/AlphaParameters.cpp:27-48
#ifndef CGB_GAPS_MATH_H
#define CGB_GAPS_MATH_H

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define CGB_HD __host__ __device__ __forceinline__
#else
#define CGB_HD inline
#endif

namespace cgb {

static const float kEpsilon = 1.0e-5f;                          // math/Math.h:11
static const float kSqrt2 = 1.4142135623730950488016887242097f; // math/Math.h:14

// ---- rounding-exact primitives --------------------------------------------------------------
CGB_HD float fadd(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
CGB_HD float fsub(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
CGB_HD float fmul(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
CGB_HD float fdiv(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
CGB_HD float fsqrt(float a)
{
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}
CGB_HD double dadd(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
CGB_HD double dmul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
CGB_HD double ddiv(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
CGB_HD double dfma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// gaps::min / gaps::max (math/Math.cpp:13-31) — ternaries, not fminf/fmaxf (NaN behaviour differs)
CGB_HD float gmin(float a, float b) { return a < b ? a : b; }
CGB_HD float gmax(float a, float b) { return a < b ? b : a; }

// ---- portable log ------------------------------------------------------------------------------
// std::log(float) in the reference resolves to the platform libm, which no GPU reproduces bit for
// bit.  This log is computed in f64 from IEEE basic operations only (log x = e ln2 + 2 atanh((m-1)/(m+1)),
// m in [sqrt(1/2), sqrt 2), 14-term odd series) and rounded once to f32: the correctly rounded fp32
// logarithm on every input we have tried, within 1 ulp of glibc's 0.818-ulp logf, and identical on
// host and device.  The oracle restates the same series (oracle/cogaps_oracle.c portable_log_f64).
CGB_HD double portable_log_f64(double x)
{
#if defined(__CUDA_ARCH__)
    uint64_t u = static_cast<uint64_t>(__double_as_longlong(x));
#else
    union { double d; uint64_t u; } cv;
    cv.d = x;
    uint64_t u = cv.u;
#endif
    int e = static_cast<int>((u >> 52) & 0x7ff) - 1023;
    u = (u & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
#if defined(__CUDA_ARCH__)
    double m = __longlong_as_double(static_cast<long long>(u));
#else
    cv.u = u;
    double m = cv.d;
#endif
    if (m > 1.4142135623730951)
    {
        m = dmul(m, 0.5);
        e += 1;
    }
    double f = ddiv(dadd(m, -1.0), dadd(m, 1.0));
    double f2 = dmul(f, f);
    double p = 1.0 / 27.0;
    p = dfma(p, f2, 1.0 / 25.0);
    p = dfma(p, f2, 1.0 / 23.0);
    p = dfma(p, f2, 1.0 / 21.0);
    p = dfma(p, f2, 1.0 / 19.0);
    p = dfma(p, f2, 1.0 / 17.0);
    p = dfma(p, f2, 1.0 / 15.0);
    p = dfma(p, f2, 1.0 / 13.0);
    p = dfma(p, f2, 1.0 / 11.0);
    p = dfma(p, f2, 1.0 / 9.0);
    p = dfma(p, f2, 1.0 / 7.0);
    p = dfma(p, f2, 1.0 / 5.0);
    p = dfma(p, f2, 1.0 / 3.0);
    p = dfma(p, f2, 1.0);
    double lm = dmul(2.0, dmul(f, p));
    return dfma(static_cast<double>(e), 0.6931471805599453094, lm);
}

CGB_HD float portable_logf(float x)
{
    if (x != x) { return x; }
    if (x < 0.f)
    {
#if defined(__CUDA_ARCH__)
        return __int_as_float(0x7fc00000);
#else
        return NAN;
#endif
    }
    if (x == 0.f)
    {
#if defined(__CUDA_ARCH__)
        return __int_as_float(0xff800000);
#else
        return -INFINITY;
#endif
    }
    if (x > 3.0e38f && x + x == x) { return x; } // +inf
    return static_cast<float>(portable_log_f64(static_cast<double>(x)));
}


// ---- portable exp ------------------------------------------------------------------------------
// exp() from IEEE +,*,fma and floor only (same bits on host, device and in the oracle): x = n ln2 + r with |r| <= ln2/2,
// degree-13 Taylor polynomial in f64, scaled by 2^n, rounded once to f32 by the caller.  Used by the sweep's same-bin
// exchange (GapsRng::truncGammaUpper, math/Random.cpp:194-200), whose std::exp no GPU reproduces bit for bit.
CGB_HD double portable_exp_f64(double x)
{
    if (x != x) { return x; }
    if (x > 700.0)
    {
#if defined(__CUDA_ARCH__)
        return __longlong_as_double(0x7ff0000000000000ll);
#else
        return INFINITY;
#endif
    }
    if (x < -700.0) { return 0.0; }
    const double n = floor(dfma(x, 1.4426950408889634074, 0.5));
    double r = dfma(-n, 0.69314718036912381649, x);
    r = dfma(-n, 1.9082149292705877e-10, r);
    double p = 1.0 / 6227020800.0;
    p = dfma(p, r, 1.0 / 479001600.0);
    p = dfma(p, r, 1.0 / 39916800.0);
    p = dfma(p, r, 1.0 / 3628800.0);
    p = dfma(p, r, 1.0 / 362880.0);
    p = dfma(p, r, 1.0 / 40320.0);
    p = dfma(p, r, 1.0 / 5040.0);
    p = dfma(p, r, 1.0 / 720.0);
    p = dfma(p, r, 1.0 / 120.0);
    p = dfma(p, r, 1.0 / 24.0);
    p = dfma(p, r, 1.0 / 6.0);
    p = dfma(p, r, 0.5);
    p = dfma(p, r, 1.0);
    p = dfma(p, r, 1.0);
    const uint64_t bits = static_cast<uint64_t>(static_cast<int64_t>(n) + 1023) << 52; // 2^n, |n| <= 1010 here
#if defined(__CUDA_ARCH__)
    return dmul(p, __longlong_as_double(static_cast<long long>(bits)));
#else
    union { double d; uint64_t u; } sc;
    sc.u = bits;
    return p * sc.d;
#endif
}

CGB_HD float portable_expf(float x) { return static_cast<float>(portable_exp_f64(static_cast<double>(x))); }

// ---- PCG32 XSH-RR, increment 55 (math/Random.cpp:40-56) -----------------------------------------
struct Pcg
{
    uint64_t state;
    CGB_HD void advance() { state = state * 6364136223846793005ull + (54u | 1); }
    CGB_HD uint32_t get() const
    {
        uint32_t xorshifted = static_cast<uint32_t>(((state >> 18u) ^ state) >> 27u);
        uint32_t rot = static_cast<uint32_t>(state >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31));
    }
    CGB_HD uint32_t next()
    {
        advance();
        return get();
    }
    // uniform(): u32 / float(UINT32_MAX); float(UINT32_MAX) rounds to 2^32 (Random.cpp:10,63-66)
    CGB_HD float uniform()
    {
#if defined(__CUDA_ARCH__)
        return fmul(__uint2float_rn(next()), 2.3283064365386963e-10f); // x / 2^32, exact scaling
#else
        return static_cast<float>(next()) * 2.3283064365386963e-10f;
#endif
    }
    CGB_HD float uniform(float a, float b) { return fadd(fmul(uniform(), fsub(b, a)), a); } // :68-71
};

// ---- lookup-table distribution functions -------------------------------------------------------
CGB_HD unsigned f2u(float x)
{
#if defined(__CUDA_ARCH__)
    return __float2uint_rz(x);
#else
    return static_cast<unsigned>(x);
#endif
}

// math/Random.cpp:307-326
CGB_HD float p_norm_fast(const float *erfTable, float p, float mean, float sd)
{
    float term = fdiv(fsub(p, mean), fmul(sd, kSqrt2));
    float erf = 0.f;
    if (term < 0.f)
    {
        term = gmax(term, -3.f);
        const unsigned ndx = f2u(fmul(-term, 1000.f));
        erf = -erfTable[ndx];
    }
    else
    {
        term = gmin(term, 3.f);
        const unsigned ndx = f2u(fmul(term, 1000.f));
        erf = erfTable[ndx];
    }
    return fmul(0.5f, fadd(1.f, erf));
}

// math/Random.cpp:328-345
CGB_HD float q_norm_fast(const float *erfinvTable, float q, float mean, float sd)
{
    float term = fsub(fmul(2.f, q), 1.f);
    float erfinv = 0.f;
    if (term < 0.f)
    {
        const unsigned ndx = f2u(fmul(-term, 5000.f));
        erfinv = -erfinvTable[ndx];
    }
    else
    {
        const unsigned ndx = f2u(fmul(term, 5000.f));
        erfinv = erfinvTable[ndx];
    }
    return fadd(mean, fmul(fmul(sd, kSqrt2), erfinv));
}

// math/Random.cpp:178-191; returns false when too far in the tail (OptionalFloat without a value)
CGB_HD bool trunc_normal(Pcg &rng, const float *erfTable, const float *erfinvTable, float a, float b,
                         float mean, float sd, float *out)
{
    float pLower = p_norm_fast(erfTable, a, mean, sd);
    float pUpper = p_norm_fast(erfTable, b, mean, sd);
    if (!(pLower > 0.95f || pUpper < 0.05f))
    {
        float z = q_norm_fast(erfinvTable, rng.uniform(pLower, pUpper), mean, sd);
        z = gmax(a, gmin(z, b));
        *out = z;
        return true;
    }
    *out = 0.f;
    return false;
}

// gibbs_sampler/AlphaParameters.cpp:27-48 (useLambda selects the 5-argument overload)
CGB_HD bool gibbs_mass(Pcg &rng, const float *erfTable, const float *erfinvTable, float s, float s_mu,
                       float a, float b, bool useLambda, float lambda, float *out)
{
    if (s > kEpsilon)
    {
        float mean = useLambda ? fdiv(fsub(s_mu, lambda), s) : fdiv(s_mu, s);
        float sd = fdiv(1.f, fsqrt(s));
        return trunc_normal(rng, erfTable, erfinvTable, a, b, mean, sd, out);
    }
    *out = 0.f;
    return false;
}

} // namespace cgb

#endif // CGB_GAPS_MATH_H
