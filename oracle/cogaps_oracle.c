/*
 * cogaps_oracle.c — CPU restatement of the CoGAPS atomic-domain Gibbs sampler (dense normal model).
 *
 * TEST INFRASTRUCTURE ONLY — see cogaps_oracle.h.  Parity status: PINNED against the reference
 * compiled in place (oracle/_ref) by tests/test_oracle_vs_reference.py.
 *
 * Every function cites the reference code it restates, as file:line under /root/reference/src.
 * The structure is deliberately plain (sorted array + swap-erase vector for the atomic domain,
 * dense row-per-sampler-row matrices) — it has to be obviously the same algorithm, not fast.
 *
 * Compile with -ffp-contract=off: the reference's default build has no FMA and every mul/add pair
 * below must round twice.
 */
#include "cogaps_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define EPSILON 1.0e-5f                                   /* math/Math.h:11 */
#define SQRT2F 1.4142135623730950488016887242097f         /* math/Math.h:14 */
#define PI_DOUBLE 3.1415926535897932384626433832795       /* math/Math.h:13 */
#define ERF_N CGB_ERF_TABLE_SIZE
#define ERFINV_N CGB_ERFINV_TABLE_SIZE
#define QGAMMA_N CGB_QGAMMA_TABLE_SIZE

/* gaps::min / gaps::max for floats (math/Math.cpp:13-31): note the NaN behaviour of the ternaries */
static float fminr(float a, float b) { return a < b ? a : b; }
static float fmaxr(float a, float b) { return a < b ? b : a; }

/* ============================== portable log ======================================= */
/* f64 log via log(x) = e*ln2 + 2*atanh((m-1)/(m+1)), m in [sqrt(1/2), sqrt(2)); only IEEE
 * +,*,/ and fma, so the CUDA device function with the same constants returns the same bits. */
static double portable_log_f64(double x)
{
    union { double d; uint64_t u; } v;
    v.d = x;
    int e = (int)((v.u >> 52) & 0x7ff) - 1023;
    v.u = (v.u & 0x000fffffffffffffull) | 0x3ff0000000000000ull; /* m in [1,2) */
    double m = v.d;
    if (m > 1.4142135623730951)
    {
        m = m * 0.5;
        e += 1;
    }
    double f = (m - 1.0) / (m + 1.0);
    double f2 = f * f;
    double p = 1.0 / 27.0;
    p = fma(p, f2, 1.0 / 25.0);
    p = fma(p, f2, 1.0 / 23.0);
    p = fma(p, f2, 1.0 / 21.0);
    p = fma(p, f2, 1.0 / 19.0);
    p = fma(p, f2, 1.0 / 17.0);
    p = fma(p, f2, 1.0 / 15.0);
    p = fma(p, f2, 1.0 / 13.0);
    p = fma(p, f2, 1.0 / 11.0);
    p = fma(p, f2, 1.0 / 9.0);
    p = fma(p, f2, 1.0 / 7.0);
    p = fma(p, f2, 1.0 / 5.0);
    p = fma(p, f2, 1.0 / 3.0);
    p = fma(p, f2, 1.0);
    double lm = 2.0 * (f * p);
    return fma((double)e, 0.6931471805599453094, lm);
}

float cogaps_oracle_portable_logf(float x)
{
    if (x != x) { return x; }
    if (x < 0.f) { return NAN; }
    if (x == 0.f) { return -INFINITY; }
    if (isinf(x)) { return x; }
    return (float)portable_log_f64((double)x); /* float subnormals are normal doubles */
}

/* ============================== RNG (math/Random.cpp) ============================== */

typedef struct
{
    uint64_t s[2];
    uint64_t prev[2];
} xoro_t;

static uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

/* Xoroshiro128plus::next, Random.cpp:231-242 */
static uint64_t xoro_next(xoro_t *g)
{
    g->prev[0] = g->s[0];
    g->prev[1] = g->s[1];
    const uint64_t s0 = g->s[0];
    uint64_t s1 = g->s[1];
    uint64_t result = s0 + s1;
    s1 ^= s0;
    g->s[0] = rotl64(s0, 24) ^ s1 ^ (s1 << 16);
    g->s[1] = rotl64(s1, 37);
    return result;
}

/* Xoroshiro128plus ctor, Random.cpp:221-229 */
static void xoro_init(xoro_t *g, uint64_t seed)
{
    g->s[0] = seed | 1;
    g->s[1] = seed | 1;
    g->prev[0] = g->prev[1] = 0;
    for (unsigned i = 0; i < 5000; ++i) { xoro_next(g); }
}

/* Xoroshiro128plus::rollBackOnce, Random.cpp:244-248 */
static void xoro_rollback(xoro_t *g)
{
    g->s[0] = g->prev[0];
    g->s[1] = g->prev[1];
}

/* GapsRandomState, Random.h:79-98 */
typedef struct
{
    xoro_t seeder;
    float erf[ERF_N];
    float erfinv[ERFINV_N];
    float qgamma[QGAMMA_N];
    int mathMode;
} randstate_t;

/* Third-party arithmetic: the reference fills its tables through Boost.Math (R package BH,
 * version unpinned; call sites math/Math.cpp:43-86).  Boost is absent here, so the published
 * definitions are restated: Phi(x) = erfc(-x/sqrt2)/2 in f64, quantiles by f64 bisection to a
 * fixed point — the same arithmetic as oracle/ref_shim, so oracle and oracle/_ref share tables
 * bit-for-bit (tests/test_oracle_vs_reference.py::test_tables). */
static double norm_cdf(double x) { return 0.5 * erfc(-x / 1.4142135623730950488016887242096980785696718753769); }

static double norm_quantile(double p)
{
    double lo = -40.0, hi = 40.0;
    for (int it = 0; it < 400; ++it)
    {
        double mid = 0.5 * (lo + hi);
        if (mid == lo || mid == hi) { break; }
        if (norm_cdf(mid) < p) { lo = mid; } else { hi = mid; }
    }
    return 0.5 * (lo + hi);
}

static double gamma2_cdf(double x) { return x <= 0.0 ? 0.0 : 1.0 - (1.0 + x) * exp(-x); }

static double gamma2_quantile(double p)
{
    double lo = 0.0, hi = 1.0;
    while (gamma2_cdf(hi) < p && hi < 1e300) { hi *= 2.0; }
    for (int it = 0; it < 400; ++it)
    {
        double mid = 0.5 * (lo + hi);
        if (mid == lo || mid == hi) { break; }
        if (gamma2_cdf(mid) < p) { lo = mid; } else { hi = mid; }
    }
    return 0.5 * (lo + hi);
}

/* gaps::p_norm / q_norm / q_gamma, math/Math.cpp:55-81: float in, f64 inside, float out */
static float p_norm(float p, float mean, float sd) { return (float)norm_cdf(((double)p - (double)mean) / (double)sd); }
static float q_norm(float q, float mean, float sd) { return (float)((double)mean + (double)sd * norm_quantile((double)q)); }
static float q_gamma2(float q)
{
    if (q < 0.000001f) { return 0.f; }
    return (float)(1.0 * gamma2_quantile((double)q));
}

/* GapsRandomState::initLookupTables, Random.cpp:269-295 */
static void init_tables(randstate_t *rs)
{
    for (unsigned i = 0; i < ERF_N; ++i)
    {
        float x = (float)i / 1000.f;
        rs->erf[i] = 2.f * p_norm(x * SQRT2F, 0.f, 1.f) - 1.f;
    }
    for (unsigned i = 0; i < ERFINV_N - 1; ++i)
    {
        float x = (float)i / (float)(ERFINV_N - 1);
        rs->erfinv[i] = q_norm((1.f + x) / 2.f, 0.f, 1.f) / SQRT2F;
    }
    rs->erfinv[ERFINV_N - 1] = q_norm(1.9998f / 2.f, 0.f, 1.f) / SQRT2F;
    rs->qgamma[0] = 0.f;
    for (unsigned i = 1; i < QGAMMA_N - 1; ++i)
    {
        float x = (float)i / (float)(QGAMMA_N - 1);
        rs->qgamma[i] = q_gamma2(x);
    }
    rs->qgamma[QGAMMA_N - 1] = q_gamma2(0.9998f);
}

static void randstate_init(randstate_t *rs, uint32_t seed, const oracle_options *opt)
{
    xoro_init(&rs->seeder, (uint64_t)seed);
    init_tables(rs);
    rs->mathMode = opt ? opt->mathMode : ORACLE_MATH_LIBM;
    if (opt && opt->erf) { memcpy(rs->erf, opt->erf, sizeof(rs->erf)); }
    if (opt && opt->erfinv) { memcpy(rs->erfinv, opt->erfinv, sizeof(rs->erfinv)); }
    if (opt && opt->qgamma) { memcpy(rs->qgamma, opt->qgamma, sizeof(rs->qgamma)); }
}

static float rs_logf(const randstate_t *rs, float x)
{
    return rs->mathMode == ORACLE_MATH_PORTABLE ? cogaps_oracle_portable_logf(x) : logf(x);
}

/* GapsRandomState::p_norm_fast, Random.cpp:307-326 */
static float p_norm_fast(const randstate_t *rs, float p, float mean, float sd)
{
    float term = (p - mean) / (sd * SQRT2F);
    float erf_ = 0.f;
    if (term < 0.f)
    {
        term = fmaxr(term, -3.f);
        const unsigned ndx = (unsigned)(-term * 1000.f);
        erf_ = -rs->erf[ndx];
    }
    else
    {
        term = fminr(term, 3.f);
        const unsigned ndx = (unsigned)(term * 1000.f);
        erf_ = rs->erf[ndx];
    }
    return 0.5f * (1.f + erf_);
}

/* GapsRandomState::q_norm_fast, Random.cpp:328-345 */
static float q_norm_fast(const randstate_t *rs, float q, float mean, float sd)
{
    float term = 2.f * q - 1.f;
    float erfinv_ = 0.f;
    if (term < 0.f)
    {
        const unsigned ndx = (unsigned)(-term * (float)(ERFINV_N - 1));
        erfinv_ = -rs->erfinv[ndx];
    }
    else
    {
        const unsigned ndx = (unsigned)(term * (float)(ERFINV_N - 1));
        erfinv_ = rs->erfinv[ndx];
    }
    return mean + sd * SQRT2F * erfinv_;
}

/* GapsRng, Random.cpp:32-66: PCG32 XSH-RR, increment 55 */
typedef struct
{
    const randstate_t *rs;
    uint64_t state;
} rng_t;

static void rng_advance(rng_t *r) { r->state = r->state * 6364136223846793005ull + (54u | 1); }

static uint32_t rng_get(const rng_t *r)
{
    uint32_t xorshifted = (uint32_t)(((r->state >> 18u) ^ r->state) >> 27u);
    uint32_t rot = (uint32_t)(r->state >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
}

static void rng_init(rng_t *r, randstate_t *rs)
{
    r->rs = rs;
    r->state = xoro_next(&rs->seeder);
    rng_advance(r);
}

static uint32_t rng_u32(rng_t *r)
{
    rng_advance(r);
    return rng_get(r);
}

static float rng_uniform(rng_t *r) { return (float)rng_u32(r) / 4294967296.0f; } /* (float)UINT32_MAX == 2^32 */
static double rng_uniformd(rng_t *r) { return (double)rng_u32(r) / 4294967295.0; }
static float rng_uniform_ab(rng_t *r, float a, float b) { return rng_uniform(r) * (b - a) + a; }

/* Random.cpp:79-96 */
static uint32_t rng_u32_range(rng_t *r, uint32_t a, uint32_t b)
{
    if (b == a) { return a; }
    uint32_t range = b + 1 - a;
    uint32_t x = rng_u32(r);
    uint32_t iPart = 0xFFFFFFFFu / range;
    while (x >= range * iPart) { x = rng_u32(r); }
    return x / iPart + a;
}

/* Random.cpp:98-103 */
static uint64_t rng_u64(rng_t *r)
{
    uint64_t high = ((uint64_t)rng_u32(r) << 32) & 0xFFFFFFFF00000000ull;
    uint64_t low = rng_u32(r);
    return high | low;
}

/* Random.cpp:106-123 */
static uint64_t rng_u64_range(rng_t *r, uint64_t a, uint64_t b)
{
    if (b == a) { return a; }
    uint64_t range = b + 1 - a;
    uint64_t x = rng_u64(r);
    uint64_t iPart = 0xFFFFFFFFFFFFFFFFull / range;
    while (x >= range * iPart) { x = rng_u64(r); }
    return x / iPart + a;
}

/* Random.cpp:125-170 */
static int rng_poisson(rng_t *r, double lambda)
{
    if (lambda <= 5.0)
    {
        int x = 0;
        double p = rng_uniformd(r);
        double cutoff = exp(-lambda);
        while (p >= cutoff)
        {
            p *= rng_uniformd(r);
            ++x;
        }
        return x;
    }
    double c = 0.767 - 3.36 / lambda;
    double beta = PI_DOUBLE / sqrt(3.0 * lambda);
    double alpha = beta * lambda;
    double k = log(c) - lambda - log(beta);
    for (;;)
    {
        double u = rng_uniformd(r);
        double x = (alpha - log((1.0 - u) / u)) / beta;
        double n = floor(x + 0.5);
        if (n < 0.0) { continue; }
        double v = rng_uniformd(r);
        double y = alpha - beta * x;
        double w = 1.0 + exp(y);
        double lhs = y + log(v / (w * w));
        double rhs = k + n * log(lambda) - lgamma(n + 1); /* boost::math::lgamma, Math.cpp:83-86 */
        if (lhs <= rhs) { return (int)n; }
    }
}

/* Random.cpp:172-175 */
static float rng_exponential(rng_t *r, float lambda) { return -1.f * rs_logf(r->rs, rng_uniform(r)) / lambda; }

/* Random.cpp:178-191; returns 1 and *out on success */
static int rng_trunc_normal(rng_t *r, float a, float b, float mean, float sd, float *out)
{
    float pLower = p_norm_fast(r->rs, a, mean, sd);
    float pUpper = p_norm_fast(r->rs, b, mean, sd);
    if (!(pLower > 0.95f || pUpper < 0.05f))
    {
        float z = q_norm_fast(r->rs, rng_uniform_ab(r, pLower, pUpper), mean, sd);
        z = fmaxr(a, fminr(z, b));
        *out = z;
        return 1;
    }
    *out = 0.f;
    return 0;
}

/* Random.cpp:194-200 */
static float rng_trunc_gamma_upper(rng_t *r, float b, float scale)
{
    float upper = 1.f - expf(-b / scale) * (1.f + b / scale);
    const unsigned ndx = (unsigned)rng_uniform_ab(r, 0.f, upper * 5000.f);
    return r->rs->qgamma[ndx] * scale;
}

/* ============================ atomic domain ======================================== */
/* ConcurrentAtomicDomain (atomic/ConcurrentAtomicDomain.cpp:14-132) and AtomicDomain
 * (atomic/AtomicDomain.cpp:12-119) keep the same two views: atoms sorted by position (std::map)
 * and an insertion-ordered vector with swap-with-last erase used for uniform picks.  Restated
 * with a sorted array of atom ids and a plain vector of atom ids. */
typedef struct
{
    uint64_t pos;
    float mass;
    uint32_t vecIndex; /* ConcurrentAtom::mIndex */
    uint32_t alive;
} atom_t;

typedef struct
{
    atom_t *pool;
    uint32_t poolSize, poolCap;
    uint32_t *freeList;
    uint32_t nFree, freeCap;
    uint32_t *sorted; /* atom ids by increasing pos  (mAtomMap) */
    uint32_t *vec;    /* atom ids in vector order    (mAtoms)   */
    uint32_t n, cap;
    uint32_t *eraseCache;
    uint32_t nErase, eraseCap;
    uint64_t domainLength;
} domain_t;

#define NO_ATOM 0xFFFFFFFFu

static void domain_init(domain_t *d, uint64_t nBins)
{
    memset(d, 0, sizeof(*d));
    uint64_t binLength = 0xFFFFFFFFFFFFFFFFull / nBins;
    d->domainLength = binLength * nBins; /* ConcurrentAtomicDomain.cpp:14-18 */
}

static void domain_free(domain_t *d)
{
    free(d->pool); free(d->freeList); free(d->sorted); free(d->vec); free(d->eraseCache);
    memset(d, 0, sizeof(*d));
}

/* index in sorted[] of the first atom with pos >= p */
static uint32_t domain_lower_bound(const domain_t *d, uint64_t p)
{
    uint32_t lo = 0, hi = d->n;
    while (lo < hi)
    {
        uint32_t mid = lo + (hi - lo) / 2;
        if (d->pool[d->sorted[mid]].pos < p) { lo = mid + 1; } else { hi = mid; }
    }
    return lo;
}

static int domain_occupied(const domain_t *d, uint64_t p)
{
    uint32_t i = domain_lower_bound(d, p);
    return i < d->n && d->pool[d->sorted[i]].pos == p;
}

static uint32_t domain_sorted_index(const domain_t *d, uint32_t id)
{
    return domain_lower_bound(d, d->pool[id].pos);
}

static uint32_t domain_left(const domain_t *d, uint32_t id)
{
    uint32_t i = domain_sorted_index(d, id);
    return i > 0 ? d->sorted[i - 1] : NO_ATOM;
}

static uint32_t domain_right(const domain_t *d, uint32_t id)
{
    uint32_t i = domain_sorted_index(d, id);
    return i + 1 < d->n ? d->sorted[i + 1] : NO_ATOM;
}

static uint32_t domain_front(const domain_t *d) { return d->sorted[0]; }

/* ConcurrentAtomicDomain::insert, ConcurrentAtomicDomain.cpp:82-105 */
static uint32_t domain_insert(domain_t *d, uint64_t pos, float mass)
{
    uint32_t id;
    if (d->nFree > 0)
    {
        id = d->freeList[--d->nFree];
    }
    else
    {
        if (d->poolSize == d->poolCap)
        {
            d->poolCap = d->poolCap ? d->poolCap * 2 : 1024;
            d->pool = (atom_t*)realloc(d->pool, sizeof(atom_t) * d->poolCap);
        }
        id = d->poolSize++;
    }
    if (d->n == d->cap)
    {
        d->cap = d->cap ? d->cap * 2 : 1024;
        d->sorted = (uint32_t*)realloc(d->sorted, sizeof(uint32_t) * d->cap);
        d->vec = (uint32_t*)realloc(d->vec, sizeof(uint32_t) * d->cap);
    }
    uint32_t i = domain_lower_bound(d, pos);
    memmove(d->sorted + i + 1, d->sorted + i, sizeof(uint32_t) * (d->n - i));
    d->sorted[i] = id;
    d->pool[id].pos = pos;
    d->pool[id].mass = mass;
    d->pool[id].vecIndex = d->n;
    d->pool[id].alive = 1;
    d->vec[d->n] = id;
    d->n++;
    return id;
}

/* ConcurrentAtomicDomain::erase, ConcurrentAtomicDomain.cpp:108-123 */
static void domain_erase(domain_t *d, uint32_t id)
{
    uint32_t i = domain_sorted_index(d, id);
    memmove(d->sorted + i, d->sorted + i + 1, sizeof(uint32_t) * (d->n - i - 1));
    uint32_t vi = d->pool[id].vecIndex;
    d->vec[vi] = d->vec[d->n - 1];
    d->pool[d->vec[vi]].vecIndex = vi;
    d->n--;
    d->pool[id].alive = 0;
    if (d->nFree == d->freeCap)
    {
        d->freeCap = d->freeCap ? d->freeCap * 2 : 1024;
        d->freeList = (uint32_t*)realloc(d->freeList, sizeof(uint32_t) * d->freeCap);
    }
    d->freeList[d->nFree++] = id;
}

/* ConcurrentAtomicDomain::cacheErase / flushEraseCache, ConcurrentAtomicDomain.cpp:62-79 */
static void domain_cache_erase(domain_t *d, uint32_t id)
{
    if (d->nErase == d->eraseCap)
    {
        d->eraseCap = d->eraseCap ? d->eraseCap * 2 : 256;
        d->eraseCache = (uint32_t*)realloc(d->eraseCache, sizeof(uint32_t) * d->eraseCap);
    }
    d->eraseCache[d->nErase++] = id;
}

static const domain_t *g_sortDomain;
static int cmp_atoms_by_pos(const void *a, const void *b)
{
    uint64_t pa = g_sortDomain->pool[*(const uint32_t*)a].pos;
    uint64_t pb = g_sortDomain->pool[*(const uint32_t*)b].pos;
    return pa < pb ? -1 : (pa > pb ? 1 : 0);
}

static void domain_flush_erase_cache(domain_t *d)
{
    g_sortDomain = d;
    qsort(d->eraseCache, d->nErase, sizeof(uint32_t), cmp_atoms_by_pos); /* positions are distinct */
    for (uint32_t i = 0; i < d->nErase; ++i) { domain_erase(d, d->eraseCache[i]); }
    d->nErase = 0;
}

/* ConcurrentAtomicDomain::move, ConcurrentAtomicDomain.cpp:126-132 (never crosses a neighbour) */
static void domain_move(domain_t *d, uint32_t id, uint64_t newPos) { d->pool[id].pos = newPos; }

/* ConcurrentAtomicDomain::randomFreePosition, :53-60 */
static uint64_t domain_random_free_position(const domain_t *d, rng_t *rng)
{
    uint64_t pos = rng_u64_range(rng, 1, d->domainLength);
    while (domain_occupied(d, pos)) { pos = rng_u64_range(rng, 1, d->domainLength); }
    return pos;
}

/* ConcurrentAtomicDomain::randomAtom / randomAtomWithNeighbors, :34-51 */
static uint32_t domain_random_atom(const domain_t *d, rng_t *rng)
{
    uint32_t index = rng_u32_range(rng, 0, d->n - 1);
    return d->vec[index];
}

/* static_cast<uint64_t>(double) as the reference's default (SSE2) build executes it: the domain
 * length as a double is often exactly 2^64 (ProposalQueue.cpp:31,214; SingleThreadedGibbsSampler.h:190),
 * out of range for the cast; cvttsd2si-based code yields 0 there.  (An AVX-512 build yields 2^64-1.) */
static uint64_t ref_double_to_u64(double x)
{
    if (x >= 18446744073709551616.0) { return 0; }
    return (uint64_t)x;
}

/* ============================ dense normal model =================================== */
/* DenseNormalModel (gibbs_sampler/DenseNormalModel.h:15-97).  The reference stores D, S, AP
 * column-major with one column per sampler row; here row r of the sampler is the contiguous
 * slice [r*L, (r+1)*L).  "other" is the other sampler's factor matrix, column c contiguous. */
typedef struct model_s
{
    uint32_t nRows;     /* rows of the factor matrix this sampler owns (genes for A) */
    uint32_t L;         /* scan length (samples for A) */
    uint32_t k;
    float *D, *S, *AP;  /* nRows x L */
    float *M;           /* factor matrix, column-major: M[c*nRows + r]  (mMatrix) */
    const struct model_s *other;
    float maxGibbsMass, annealingTemp, lambda;
    int reduceMode;
    cgb_reduction_order order;
    /* SparseNormalModel (gibbs_sampler/SparseNormalModel.h:16-66) */
    int sparse;
    float *Mrows;       /* HybridMatrix::mRows: nRows x k row-major, never zeroed below epsilon          */
                        /* (M above is HybridMatrix::mCols: values below epsilon stored as 0, flag off) */
    float *Z1, *Z2;     /* k and k x k lookup tables of the other factor (generateLookupTables)        */
    float beta;
} model_t;

/* Matrix(const Matrix&, genesInCols, subsetGenes, indices), data_structures/Matrix.cpp:30-69.
 * Output: out[j*nG + i] = result(i, j)  (column j contiguous), dims returned. */
static float *load_matrix(const float *data, uint32_t nrow, uint32_t ncol, int genesInCols, int subsetGenes,
                          const uint32_t *indices, uint32_t nIdx, uint32_t *outRows, uint32_t *outCols)
{
    int subsetData = nIdx > 0;
    uint32_t nGenes = (subsetData && subsetGenes) ? nIdx : (genesInCols ? ncol : nrow);
    uint32_t nSamples = (subsetData && !subsetGenes) ? nIdx : (genesInCols ? nrow : ncol);
    float *out = (float*)malloc(sizeof(float) * (size_t)nGenes * nSamples);
    for (uint32_t j = 0; j < nSamples; ++j)
    {
        for (uint32_t i = 0; i < nGenes; ++i)
        {
            uint32_t dataRow = (subsetData && (subsetGenes != genesInCols))
                ? indices[genesInCols ? j : i] - 1 : (genesInCols ? j : i);
            uint32_t dataCol = (subsetData && (subsetGenes == genesInCols))
                ? indices[genesInCols ? i : j] - 1 : (genesInCols ? i : j);
            out[(size_t)j * nGenes + i] = data[(size_t)dataRow * ncol + dataCol];
        }
    }
    *outRows = nGenes;
    *outCols = nSamples;
    return out;
}

/* DenseNormalModel ctor, DenseNormalModel.h:66-88 */
static void model_init(model_t *m, const float *data, uint32_t nrow, uint32_t ncol, int transpose,
                       int subsetRows, const cgb_params *p, float alpha, float maxGibbsMass,
                       const oracle_options *opt, int isA)
{
    memset(m, 0, sizeof(*m));
    uint32_t R, Cc;
    m->D = load_matrix(data, nrow, ncol, transpose, subsetRows, p->subsetIndices, p->nSubsetIndices, &R, &Cc);
    m->L = R;      /* mDMatrix.nRow() */
    m->nRows = Cc; /* mDMatrix.nCol() */
    m->k = p->nPatterns;
    size_t n = (size_t)m->nRows * m->L;
    m->S = (float*)malloc(sizeof(float) * n);
    m->AP = (float*)calloc(n, sizeof(float));
    m->M = (float*)calloc((size_t)m->nRows * m->k, sizeof(float));
    for (size_t i = 0; i < n; ++i) { m->S[i] = fmaxr(m->D[i] * 0.1f, 0.1f); } /* gaps::pmax, MatrixMath.cpp:74-84 */
    m->maxGibbsMass = maxGibbsMass;
    m->annealingTemp = 1.f;
    /* gaps::nonZeroMean, MatrixMath.cpp:39-55: sums everything, divides by the count of positives */
    float sum = 0.f;
    unsigned nnz = 0;
    for (size_t i = 0; i < n; ++i)
    {
        sum += m->D[i];
        if (m->D[i] > 0.f) { ++nnz; }
    }
    float meanD = sum / (float)nnz;
    m->lambda = alpha * sqrtf((float)(uint64_t)m->k / meanD);
    m->maxGibbsMass = m->maxGibbsMass / m->lambda;
    m->reduceMode = opt ? opt->reduceMode : ORACLE_REDUCE_SCALAR;
    if (opt) { m->order = isA ? opt->orderA : opt->orderP; }
    m->sparse = p->useSparseOptimization != 0;
    m->beta = 100.f;                                         /* SparseNormalModel.h:77 */
    m->Mrows = (float*)calloc((size_t)m->nRows * m->k, sizeof(float));
    m->Z1 = (float*)calloc(m->k, sizeof(float));
    m->Z2 = (float*)calloc((size_t)m->k * m->k, sizeof(float));
}

static void model_free(model_t *m)
{
    free(m->D); free(m->S); free(m->AP); free(m->M); free(m->Mrows); free(m->Z1); free(m->Z2);
    memset(m, 0, sizeof(*m));
}

/* setUncertainty, DenseNormalModel.h:90-95 */
static void model_set_uncertainty(model_t *m, const float *unc, uint32_t nrow, uint32_t ncol, int transpose,
                                  int subsetRows, const cgb_params *p)
{
    uint32_t R, Cc;
    if (m->sparse) { return; } /* SparseNormalModel::setUncertainty is a nop (SparseNormalModel.h:92-98) */
    free(m->S);
    m->S = load_matrix(unc, nrow, ncol, transpose, subsetRows, p->subsetIndices, p->nSubsetIndices, &R, &Cc);
}

/* setMatrix, DenseNormalModel.cpp:9-12; mat is rows x k row-major */
static void model_set_matrix(model_t *m, const float *mat)
{
    for (uint32_t r = 0; r < m->nRows; ++r)
    {
        for (uint32_t c = 0; c < m->k; ++c)
        {
            float v = mat[(size_t)r * m->k + c];
            m->Mrows[(size_t)r * m->k + c] = v;
            /* HybridMatrix::operator=(Matrix): add(-old) then add(new) on the column copy */
            m->M[(size_t)c * m->nRows + r] = (m->sparse && v < EPSILON) ? 0.f : v;
        }
    }
}

/* sync, DenseNormalModel.cpp:20-36: AP <- transpose(other.AP) */
static void sparse_generate_tables(model_t *m);

static void model_sync(model_t *m, const model_t *o)
{
    if (m->sparse)
    {
        m->other = o; /* SparseNormalModel::sync, SparseNormalModel.cpp:27-31 */
        sparse_generate_tables(m);
        return;
    }
    for (uint32_t r = 0; r < m->nRows; ++r)
    {
        for (uint32_t l = 0; l < m->L; ++l) { m->AP[(size_t)r * m->L + l] = o->AP[(size_t)l * o->L + r]; }
    }
    m->other = o;
}

/* extraInitialization, DenseNormalModel.cpp:38-54 */
static void model_extra_initialization(model_t *m)
{
    if (m->sparse) { return; } /* SparseNormalModel.cpp:33-37 */
    const model_t *o = m->other;
    for (uint32_t r = 0; r < m->nRows; ++r)
    {
        for (uint32_t l = 0; l < m->L; ++l)
        {
            float acc = 0.f;
            for (uint32_t c = 0; c < m->k; ++c)
            {
                acc += o->M[(size_t)c * o->nRows + l] * m->M[(size_t)c * m->nRows + r];
            }
            m->AP[(size_t)r * m->L + l] = acc;
        }
    }
}

/* chiSq, DenseNormalModel.cpp:56-68: outer loop over mDMatrix rows (= scan index), inner over columns */
static float sparse_chisq(const model_t *m);

static float model_chisq(const model_t *m)
{
    if (m->sparse) { return sparse_chisq(m); }
    float chisq = 0.f;
    for (uint32_t l = 0; l < m->L; ++l)
    {
        for (uint32_t r = 0; r < m->nRows; ++r)
        {
            size_t i = (size_t)r * m->L + l;
            float t = (m->D[i] - m->AP[i]) / m->S[i];
            chisq += t * t;
        }
    }
    return chisq;
}

/* dataSparsity, DenseNormalModel.cpp:70-73 + MatrixMath.cpp:6-21 */
static float model_data_sparsity(const model_t *m)
{
    unsigned nnz = 0;
    size_t n = (size_t)m->nRows * m->L;
    for (size_t i = 0; i < n; ++i) { if (m->D[i] > 0.f) { ++nnz; } }
    float size = (float)(m->nRows * m->L);
    return 1.f - (float)nnz / size;
}

/* canUseGibbs, DenseNormalModel.cpp:100-108 + gaps::isVectorZero, VectorMath.cpp */
static int model_can_use_gibbs(const model_t *m, uint32_t col)
{
    const model_t *o = m->other;
    const float *v = o->M + (size_t)col * o->nRows;
    for (uint32_t i = 0; i < o->nRows; ++i) { if (v[i] > 0.f) { return 1; } }
    return 0;
}

/* ---- the association order of the two scan sums ---- */
static float reduce_terms_mode(int reduceMode, const cgb_reduction_order *o, const float *t, uint32_t L, float *scratch)
{
    if (reduceMode == ORACLE_REDUCE_SCALAR)
    {
        /* math/SIMD.h scalar path: PackedFloat is one float, one running sum from 0 */
        float acc = 0.f;
        for (uint32_t i = 0; i < L; ++i) { acc += t[i]; }
        return acc;
    }
    if (reduceMode == ORACLE_REDUCE_AVX8)
    {
        /* math/SIMD.h:8-19: 8 lanes, chunk c feeds lane j with element 8c+j; the loop runs over
         * ceil(L/8) chunks and reads the pads (data 0, S 1 -> term +0); scalar() = two hadds then
         * ra[0] + ra[4] (SIMD.h:101-114) */
        float lane[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        uint32_t nChunks = (L + 7) / 8;
        for (uint32_t c = 0; c < nChunks; ++c)
        {
            for (uint32_t j = 0; j < 8; ++j)
            {
                uint32_t i = c * 8 + j;
                lane[j] += (i < L) ? t[i] : 0.f;
            }
        }
        float lo = (lane[0] + lane[1]) + (lane[2] + lane[3]);
        float hi = (lane[4] + lane[5]) + (lane[6] + lane[7]);
        return lo + hi;
    }
    /* ORACLE_REDUCE_DEVICE: see cgb_reduction_order in include/cogaps_b200.h */
    uint32_t T = o->threadsPerSegment, V = o->vectorWidth;
    float total = 0.f;
    for (uint32_t q = 0; q < o->nSegments; ++q)
    {
        uint32_t base = q * o->segmentLength;
        uint32_t len = base >= L ? 0 : (L - base < o->segmentLength ? L - base : o->segmentLength);
        float *lanes = scratch;
        for (uint32_t i = 0; i < T; ++i) { lanes[i] = 0.f; }
        for (uint32_t e = 0; e < len; ++e) { lanes[(e / V) % T] += t[base + e]; }
        float wt[32];
        for (uint32_t i = 0; i < 32; ++i) { wt[i] = 0.f; }
        for (uint32_t w = 0; w < T / 32; ++w)
        {
            float v[32], nv[32];
            memcpy(v, lanes + w * 32, sizeof(v));
            for (uint32_t off = 16; off >= 1; off >>= 1)
            {
                for (uint32_t i = 0; i < 32; ++i) { nv[i] = v[i] + v[i ^ off]; }
                memcpy(v, nv, sizeof(v));
            }
            wt[w] = v[0];
        }
        /* warp totals in the low lanes of one warp (zeros above), same butterfly again */
        {
            float nv[32];
            for (uint32_t off = 16; off >= 1; off >>= 1)
            {
                for (uint32_t i = 0; i < 32; ++i) { nv[i] = wt[i] + wt[i ^ off]; }
                memcpy(wt, nv, sizeof(wt));
            }
        }
        float segTotal = wt[0];
        total = (q == 0) ? segTotal : total + segTotal;
    }
    return total;
}

static float reduce_terms(const model_t *m, const float *t, uint32_t L, float *scratch)
{
    return reduce_terms_mode(m->reduceMode, &m->order, t, L, scratch);
}

typedef struct { float s, s_mu; } alpha_t;

/* scan core shared by the three PERFORMANCE CRITICAL variants, DenseNormalModel.cpp:162-240.
 * v2 == NULL: v = other[:,c1]; else v = other[:,c1] - other[:,c2].  ch != 0 path is WithChange. */
static alpha_t model_scan(const model_t *m, uint32_t row, const float *v1, const float *v2, int withChange, float ch)
{
    uint32_t L = m->L;
    const float *D = m->D + (size_t)row * L;
    const float *S = m->S + (size_t)row * L;
    const float *AP = m->AP + (size_t)row * L;
    uint32_t T = m->reduceMode == ORACLE_REDUCE_DEVICE ? m->order.threadsPerSegment : 0;
    float *ts = (float*)malloc(sizeof(float) * ((size_t)L * 2 + T + 1));
    float *tmu = ts + L;
    float *scratch = tmu + L;
    for (uint32_t i = 0; i < L; ++i)
    {
        float mat = v2 ? (v1[i] - v2[i]) : v1[i];
        float ratio = mat / (S[i] * S[i]);
        ts[i] = mat * ratio;
        if (withChange)
        {
            tmu[i] = ratio * (D[i] - (AP[i] + ch * v1[i]));
        }
        else
        {
            tmu[i] = ratio * (D[i] - AP[i]);
        }
    }
    alpha_t a;
    a.s = reduce_terms(m, ts, L, scratch);
    a.s_mu = reduce_terms(m, tmu, L, scratch);
    free(ts);
    return a;
}

static const float *other_col(const model_t *m, uint32_t c) { return m->other->M + (size_t)c * m->other->nRows; }

/* alphaParameters(row, col), DenseNormalModel.cpp:162-183 */
static alpha_t sparse_alpha(const model_t *m, uint32_t row, uint32_t col, int withChange, float ch);
static alpha_t sparse_alpha2(const model_t *m, uint32_t r1, uint32_t c1, uint32_t r2, uint32_t c2);
static void sparse_change_matrix(model_t *m, uint32_t row, uint32_t col, float delta);
static void sparse_safely_change_matrix(model_t *m, uint32_t row, uint32_t col, float delta);

static alpha_t model_alpha(const model_t *m, uint32_t row, uint32_t col)
{
    if (m->sparse) { return sparse_alpha(m, row, col, 0, 0.f); }
    return model_scan(m, row, other_col(m, col), NULL, 0, 0.f);
}

/* alphaParameters(r1,c1,r2,c2), DenseNormalModel.cpp:186-214; operator+ AlphaParameters.cpp:11-14 */
static alpha_t model_alpha2(const model_t *m, uint32_t r1, uint32_t c1, uint32_t r2, uint32_t c2)
{
    if (m->sparse) { return sparse_alpha2(m, r1, c1, r2, c2); }
    if (r1 == r2) { return model_scan(m, r1, other_col(m, c1), other_col(m, c2), 0, 0.f); }
    alpha_t a = model_alpha(m, r1, c1);
    alpha_t b = model_alpha(m, r2, c2);
    alpha_t r;
    r.s = a.s + b.s;
    r.s_mu = a.s_mu - b.s_mu; /* "minus sign not a typo" */
    return r;
}

/* alphaParametersWithChange, DenseNormalModel.cpp:217-240 */
static alpha_t model_alpha_with_change(const model_t *m, uint32_t row, uint32_t col, float ch)
{
    if (m->sparse) { return sparse_alpha(m, row, col, 1, ch); }
    return model_scan(m, row, other_col(m, col), NULL, 1, ch);
}

/* updateAPMatrix, DenseNormalModel.cpp:243-258 */
static void model_update_ap(model_t *m, uint32_t row, uint32_t col, float delta)
{
    const float *o = other_col(m, col);
    float *ap = m->AP + (size_t)row * m->L;
    for (uint32_t i = 0; i < m->L; ++i) { ap[i] = ap[i] + delta * o[i]; }
}

/* changeMatrix / safelyChangeMatrix, DenseNormalModel.cpp:110-123 */
static void model_change_matrix(model_t *m, uint32_t row, uint32_t col, float delta)
{
    if (m->sparse) { sparse_change_matrix(m, row, col, delta); return; }
    m->M[(size_t)col * m->nRows + row] += delta;
    model_update_ap(m, row, col, delta);
}

static void model_safely_change_matrix(model_t *m, uint32_t row, uint32_t col, float delta)
{
    if (m->sparse) { sparse_safely_change_matrix(m, row, col, delta); return; }
    float *el = &m->M[(size_t)col * m->nRows + row];
    float newVal = fmaxr(*el + delta, 0.f);
    model_update_ap(m, row, col, newVal - *el);
    *el = newVal;
}


/* ============================ sparse normal model ================================== */
/* SparseNormalModel (gibbs_sampler/SparseNormalModel.cpp).  The reference walks 64-bit index flags
 * of the data column and of the factor column (common = d_flags & v_flags) and a packed value
 * array; here the same elements are visited in the same ascending order by testing D > 0 and the
 * factor's column copy != 0 (its flag is set exactly when the stored value is non-zero,
 * data_structures/HybridVector.cpp:55-86). */

/* gaps::dot (math/VectorMath.h:40-98): the 25-case fall-through switch accumulates the chunks in
 * DESCENDING order when there are at most 25 of them, ascending otherwise */
static float ref_dot(const float *a, const float *b, uint32_t n, int mode)
{
    if (mode == ORACLE_REDUCE_AVX8)
    {
        float lane[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        uint32_t nChunks = 1 + (n - 1) / 8;
        for (uint32_t q = 0; q < nChunks; ++q)
        {
            uint32_t c = (nChunks <= 25) ? nChunks - 1 - q : q;
            for (uint32_t j = 0; j < 8; ++j)
            {
                uint32_t i = c * 8 + j;
                lane[j] = lane[j] + ((i < n) ? a[i] * b[i] : 0.f);
            }
        }
        float lo = (lane[0] + lane[1]) + (lane[2] + lane[3]);
        float hi = (lane[4] + lane[5]) + (lane[6] + lane[7]);
        return lo + hi;
    }
    float acc = 0.f;
    if (mode == ORACLE_REDUCE_SCALAR && n <= 25)
    {
        for (uint32_t q = 0; q < n; ++q) { uint32_t i = n - 1 - q; acc = acc + a[i] * b[i]; }
        return acc;
    }
    for (uint32_t i = 0; i < n; ++i) { acc = acc + a[i] * b[i]; }
    return acc;
}

/* gaps::dot_diff (math/VectorMath.h:136-154): always ascending */
static float ref_dot_diff(const float *a, const float *b, const float *c, uint32_t n, int mode)
{
    if (mode == ORACLE_REDUCE_AVX8)
    {
        float lane[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (uint32_t i0 = 0; i0 < n; i0 += 8)
        {
            for (uint32_t j = 0; j < 8; ++j)
            {
                uint32_t i = i0 + j;
                lane[j] += (i < n) ? a[i] * (b[i] - c[i]) : 0.f;
            }
        }
        float lo = (lane[0] + lane[1]) + (lane[2] + lane[3]);
        float hi = (lane[4] + lane[5]) + (lane[6] + lane[7]);
        return lo + hi;
    }
    float acc = 0.f;
    for (uint32_t i = 0; i < n; ++i) { acc += a[i] * (b[i] - c[i]); }
    return acc;
}

/* the device's association order for the sparse sums: element e of the visited list goes to lane
 * e % 256, lanes and warps combine by the usual butterflies (one segment) */
static const cgb_reduction_order SPARSE_ORDER = {256u, 1u, 1u, 0x7fffffffu};

static float sparse_reduce(const model_t *m, const float *t, uint32_t n)
{
    float scratch[256];
    return reduce_terms_mode(ORACLE_REDUCE_DEVICE, &SPARSE_ORDER, t, n, scratch);
}

static const float *mrow(const model_t *m, uint32_t r) { return m->Mrows + (size_t)r * m->k; }

/* generateLookupTables, SparseNormalModel.cpp:294-311 */
static void sparse_generate_tables(model_t *m)
{
    const model_t *o = m->other;
    uint32_t k = m->k, n = o->nRows;
    float *t = (float*)malloc(sizeof(float) * (n + 1));
    for (uint32_t i = 0; i < k; ++i)
    {
        if (m->reduceMode == ORACLE_REDUCE_DEVICE)
        {
            for (uint32_t r = 0; r < n; ++r) { float v = mrow(o, r)[i]; t[r] = v * v; }
            m->Z1[i] = sparse_reduce(m, t, n);
        }
        else
        {
            m->Z1[i] = 0.f;
            for (uint32_t r = 0; r < n; ++r) { float v = mrow(o, r)[i]; m->Z1[i] += v * v; }
        }
        for (uint32_t j = i; j < k; ++j)
        {
            const float *a = o->M + (size_t)i * n, *b = o->M + (size_t)j * n;
            float d;
            if (m->reduceMode == ORACLE_REDUCE_DEVICE)
            {
                for (uint32_t r = 0; r < n; ++r) { t[r] = a[r] * b[r]; }
                d = sparse_reduce(m, t, n);
            }
            else
            {
                d = ref_dot(a, b, n, m->reduceMode);
            }
            m->Z2[(size_t)j * k + i] = d;
            m->Z2[(size_t)i * k + j] = d;
        }
    }
    free(t);
}

/* alphaParameters(row,col) / alphaParametersWithChange, SparseNormalModel.cpp:153-236 */
static alpha_t sparse_alpha(const model_t *m, uint32_t row, uint32_t col, int withChange, float ch)
{
    const model_t *o = m->other;
    uint32_t L = m->L, k = m->k;
    const float *D = m->D + (size_t)row * L;
    const float *V = o->M + (size_t)col * o->nRows;
    const float *Z2col = m->Z2 + (size_t)col * k;
    int dev = m->reduceMode == ORACLE_REDUCE_DEVICE;
    float s = m->Z1[col];
    float s_mu = -1.f * ref_dot(mrow(m, row), Z2col, k, m->reduceMode);
    if (withChange) { s_mu -= ch * m->Z2[(size_t)col * k + col]; }
    float *ts = NULL, *tm = NULL, *tm2 = NULL;
    uint32_t n = 0;
    if (dev)
    {
        ts = (float*)malloc(sizeof(float) * 3 * (L + 1));
        tm = ts + L + 1;
        tm2 = tm + L + 1;
    }
    for (uint32_t l = 0; l < L; ++l)
    {
        if (D[l] > 0.f && V[l] != 0.f)
        {
            float v_val = V[l], d_val = D[l];
            float term1 = v_val / d_val;
            float term2 = v_val - term1 / d_val;
            float dotv = ref_dot(mrow(m, row), mrow(o, l), k, m->reduceMode);
            float es = term1 * term1 - v_val * v_val;
            float em = term1 + term2 * dotv;
            float em2 = withChange ? term2 * mrow(o, l)[col] * ch : 0.f;
            if (dev)
            {
                ts[n] = es; tm[n] = em; tm2[n] = em2; ++n;
            }
            else
            {
                s += es;
                s_mu += em;
                if (withChange) { s_mu += em2; }
            }
        }
    }
    if (dev)
    {
        /* lane e % 256 adds its elements in order (both s_mu contributions of an element back to back) */
        float lanesS[256], lanesM[256];
        for (int i = 0; i < 256; ++i) { lanesS[i] = 0.f; lanesM[i] = 0.f; }
        for (uint32_t e = 0; e < n; ++e)
        {
            lanesS[e % 256] += ts[e];
            lanesM[e % 256] += tm[e];
            if (withChange) { lanesM[e % 256] += tm2[e]; }
        }
        /* feed the lane totals to the generic butterfly (one element per lane) */
        s = s + sparse_reduce(m, lanesS, 256);
        s_mu = s_mu + sparse_reduce(m, lanesM, 256);
        free(ts);
    }
    alpha_t a;
    a.s = s * m->beta;
    a.s_mu = s_mu * m->beta;
    return a;
}

/* alphaParameters(r1,c1,r2,c2), SparseNormalModel.cpp:239-292 */
static alpha_t sparse_alpha2(const model_t *m, uint32_t r1, uint32_t c1, uint32_t r2, uint32_t c2)
{
    if (r1 != r2)
    {
        alpha_t a = sparse_alpha(m, r1, c1, 0, 0.f);
        alpha_t b = sparse_alpha(m, r2, c2, 0, 0.f);
        alpha_t r;
        r.s = a.s + b.s;
        r.s_mu = a.s_mu - b.s_mu;
        return r;
    }
    const model_t *o = m->other;
    uint32_t L = m->L, k = m->k;
    const float *D = m->D + (size_t)r1 * L;
    const float *V1 = o->M + (size_t)c1 * o->nRows;
    const float *V2 = o->M + (size_t)c2 * o->nRows;
    int dev = m->reduceMode == ORACLE_REDUCE_DEVICE;
    float s = m->Z1[c1] - 2.f * m->Z2[(size_t)c2 * k + c1] + m->Z1[c2];
    float s_mu = -1.f * ref_dot_diff(mrow(m, r1), m->Z2 + (size_t)c1 * k, m->Z2 + (size_t)c2 * k, k, m->reduceMode);
    float lanesS[256], lanesM[256];
    for (int i = 0; i < 256; ++i) { lanesS[i] = 0.f; lanesM[i] = 0.f; }
    uint32_t n = 0;
    for (uint32_t l = 0; l < L; ++l)
    {
        if (D[l] > 0.f && (V1[l] != 0.f || V2[l] != 0.f))
        {
            float d_recip = 1.f / D[l];
            float term1 = 1.f - d_recip * d_recip;
            float v_diff = V1[l] - V2[l];
            float ap = ref_dot(mrow(m, r1), mrow(o, l), k, m->reduceMode);
            float es = v_diff * v_diff * term1;
            float em = v_diff * (ap * term1 + d_recip);
            if (dev)
            {
                lanesS[n % 256] += es;
                lanesM[n % 256] += em;
                ++n;
            }
            else
            {
                s -= es;
                s_mu += em;
            }
        }
    }
    if (dev)
    {
        s = s - sparse_reduce(m, lanesS, 256);
        s_mu = s_mu + sparse_reduce(m, lanesM, 256);
    }
    alpha_t a;
    a.s = s * m->beta;
    a.s_mu = s_mu * m->beta;
    return a;
}

/* HybridMatrix::add / set (data_structures/HybridMatrix.cpp:25-39, HybridVector.cpp:55-86) */
static void sparse_change_matrix(model_t *m, uint32_t row, uint32_t col, float delta)
{
    m->Mrows[(size_t)row * m->k + col] += delta;
    float *c = &m->M[(size_t)col * m->nRows + row];
    if (*c + delta < EPSILON) { *c = 0.f; } else { *c += delta; }
}

static void sparse_safely_change_matrix(model_t *m, uint32_t row, uint32_t col, float delta)
{
    float newVal = fmaxr(m->Mrows[(size_t)row * m->k + col] + delta, 0.f);
    m->Mrows[(size_t)row * m->k + col] = newVal;
    m->M[(size_t)col * m->nRows + row] = (newVal < EPSILON) ? 0.f : newVal;
}

/* chiSq, SparseNormalModel.cpp:39-60 */
static float sparse_chisq(const model_t *m)
{
    const model_t *o = m->other;
    float chisq = 0.f;
    for (uint32_t j = 0; j < m->nRows; ++j)
    {
        for (uint32_t i = 0; i < m->L; ++i)
        {
            float dotv = ref_dot(mrow(m, j), mrow(o, i), m->k, m->reduceMode);
            chisq += dotv * dotv;
        }
        for (uint32_t i = 0; i < m->L; ++i)
        {
            float d = m->D[(size_t)j * m->L + i];
            if (d > 0.f)
            {
                float dotv = ref_dot(mrow(m, j), mrow(o, i), m->k, m->reduceMode);
                float dsq = d * d;
                chisq += 1 + dotv * (dotv - 2 * d - dsq * dotv) / dsq;
            }
        }
    }
    return chisq * m->beta;
}

/* gibbsMass, gibbs_sampler/AlphaParameters.cpp:27-48 */
static int gibbs_mass(alpha_t a, float lo, float hi, rng_t *rng, int useLambda, float lambda, float *out)
{
    if (a.s > EPSILON)
    {
        float mean = useLambda ? (a.s_mu - lambda) / a.s : a.s_mu / a.s;
        float sd = 1.f / sqrtf(a.s);
        return rng_trunc_normal(rng, lo, hi, mean, sd, out);
    }
    *out = 0.f;
    return 0;
}

static alpha_t alpha_scale(alpha_t a, float t)
{
    a.s *= t;
    a.s_mu *= t;
    return a;
}

/* ============================ proposal queue ======================================= */
/* AtomicProposal, atomic/ProposalQueue.h:15-28 */
typedef struct
{
    rng_t rng;
    uint64_t pos;
    uint32_t atom1, atom2;
    uint32_t r1, c1, r2, c2;
    char type;
} proposal_t;

typedef struct
{
    proposal_t *q;
    uint32_t n, cap;
    uint32_t *usedRows;      /* FixedHashSetU32, data_structures/HashSets.cpp:5-37 */
    uint32_t usedKey;
    uint64_t *usedAtoms;     /* SmallHashSetU64, :39-69 */
    uint32_t nUsedAtoms, usedAtomsCap;
    uint64_t *movesA, *movesB; /* SmallPairedHashSetU64, :71-113 */
    uint32_t nMoves, movesCap;
    randstate_t *rs;
    rng_t rng;
    uint64_t minAtoms, maxAtoms, binLength, numCols;
    double alpha, domainLength, numBins;
    float lambda, u1, u2;
    unsigned numProcessed;
    int useCachedRng;
} queue_t;

/* ProposalQueue ctor, ProposalQueue.cpp:19-37 */
static void queue_init(queue_t *q, uint64_t nElements, uint64_t nPatterns, randstate_t *rs)
{
    memset(q, 0, sizeof(*q));
    q->usedRows = (uint32_t*)calloc((size_t)(nElements / nPatterns), sizeof(uint32_t));
    q->usedKey = 1;
    q->rs = rs;
    rng_init(&q->rng, rs);
    q->binLength = 0xFFFFFFFFFFFFFFFFull / nElements;
    q->numCols = nPatterns;
    q->domainLength = (double)(q->binLength * nElements);
    q->numBins = (double)nElements;
}

static void queue_free(queue_t *q)
{
    free(q->q); free(q->usedRows); free(q->usedAtoms); free(q->movesA); free(q->movesB);
    memset(q, 0, sizeof(*q));
}

static void queue_push(queue_t *q, const proposal_t *p)
{
    if (q->n == q->cap)
    {
        q->cap = q->cap ? q->cap * 2 : 256;
        q->q = (proposal_t*)realloc(q->q, sizeof(proposal_t) * q->cap);
    }
    q->q[q->n++] = *p;
}

static void queue_use_atom(queue_t *q, uint64_t pos)
{
    if (q->nUsedAtoms == q->usedAtomsCap)
    {
        q->usedAtomsCap = q->usedAtomsCap ? q->usedAtomsCap * 2 : 256;
        q->usedAtoms = (uint64_t*)realloc(q->usedAtoms, sizeof(uint64_t) * q->usedAtomsCap);
    }
    q->usedAtoms[q->nUsedAtoms++] = pos;
}

static int queue_atom_used(const queue_t *q, uint64_t pos)
{
    for (uint32_t i = 0; i < q->nUsedAtoms; ++i) { if (q->usedAtoms[i] == pos) { return 1; } }
    return 0;
}

static void queue_add_move(queue_t *q, uint64_t a, uint64_t b)
{
    if (q->nMoves == q->movesCap)
    {
        q->movesCap = q->movesCap ? q->movesCap * 2 : 256;
        q->movesA = (uint64_t*)realloc(q->movesA, sizeof(uint64_t) * q->movesCap);
        q->movesB = (uint64_t*)realloc(q->movesB, sizeof(uint64_t) * q->movesCap);
    }
    q->movesA[q->nMoves] = a < b ? a : b;
    q->movesB[q->nMoves] = a < b ? b : a;
    q->nMoves++;
}

static int queue_move_overlap(const queue_t *q, uint64_t pos)
{
    for (uint32_t i = 0; i < q->nMoves; ++i) { if (q->movesA[i] < pos && pos < q->movesB[i]) { return 1; } }
    return 0;
}

/* ProposalQueue::clear, ProposalQueue.cpp:78-85 */
static void queue_clear(queue_t *q)
{
    q->n = 0;
    q->usedKey++;
    q->nUsedAtoms = 0;
    q->nMoves = 0;
}

/* ProposalQueue::deathProb, ProposalQueue.cpp:123-127 */
static float queue_death_prob(const queue_t *q, double nAtoms)
{
    double numer = nAtoms * q->domainLength;
    return (float)(numer / (numer + q->alpha * q->numBins * (q->domainLength - nAtoms)));
}

static void proposal_init(proposal_t *p, char t, randstate_t *rs)
{
    memset(p, 0, sizeof(*p));
    rng_init(&p->rng, rs); /* AtomicProposal ctor pulls one seed, ProposalQueue.cpp:12-15 */
    p->atom1 = p->atom2 = NO_ATOM;
    p->type = t;
}

/* ProposalQueue::birth, ProposalQueue.cpp:162-187 */
static int queue_birth(queue_t *q, domain_t *d)
{
    proposal_t prop;
    proposal_init(&prop, 'B', q->rs);
    uint64_t pos = domain_random_free_position(d, &prop.rng);
    if (queue_move_overlap(q, pos))
    {
        xoro_rollback(&q->rs->seeder);
        return 0;
    }
    prop.r1 = (uint32_t)((pos / q->binLength) / q->numCols);
    prop.c1 = (uint32_t)((pos / q->binLength) % q->numCols);
    if (q->usedRows[prop.r1] == q->usedKey)
    {
        xoro_rollback(&q->rs->seeder);
        return 0;
    }
    prop.atom1 = domain_insert(d, pos, 0.f);
    q->usedRows[prop.r1] = q->usedKey;
    queue_use_atom(q, pos);
    queue_push(q, &prop);
    ++q->maxAtoms;
    return 1;
}

/* ProposalQueue::death, ProposalQueue.cpp:189-207 */
static int queue_death(queue_t *q, domain_t *d)
{
    proposal_t prop;
    proposal_init(&prop, 'D', q->rs);
    prop.atom1 = domain_random_atom(d, &prop.rng);
    uint64_t p1 = d->pool[prop.atom1].pos;
    prop.r1 = (uint32_t)((p1 / q->binLength) / q->numCols);
    prop.c1 = (uint32_t)((p1 / q->binLength) % q->numCols);
    if (q->usedRows[prop.r1] == q->usedKey)
    {
        xoro_rollback(&q->rs->seeder);
        return 0;
    }
    q->usedRows[prop.r1] = q->usedKey;
    queue_use_atom(q, p1);
    queue_push(q, &prop);
    --q->minAtoms;
    return 1;
}

/* ProposalQueue::move, ProposalQueue.cpp:209-248 */
static int queue_move(queue_t *q, domain_t *d)
{
    proposal_t prop;
    proposal_init(&prop, 'M', q->rs);
    prop.atom1 = domain_random_atom(d, &prop.rng); /* randomAtomWithNeighbors: same single draw */
    uint32_t left = domain_left(d, prop.atom1);
    uint32_t right = domain_right(d, prop.atom1);
    uint64_t lbound = left != NO_ATOM ? d->pool[left].pos : 0;
    uint64_t rbound = right != NO_ATOM ? d->pool[right].pos : ref_double_to_u64(q->domainLength);
    if (queue_atom_used(q, lbound) || queue_atom_used(q, rbound))
    {
        xoro_rollback(&q->rs->seeder);
        return 0;
    }
    prop.pos = rng_u64_range(&prop.rng, lbound + 1, rbound - 1);
    uint64_t p1 = d->pool[prop.atom1].pos;
    prop.r1 = (uint32_t)((p1 / q->binLength) / q->numCols);
    prop.c1 = (uint32_t)((p1 / q->binLength) % q->numCols);
    prop.r2 = (uint32_t)((prop.pos / q->binLength) / q->numCols);
    prop.c2 = (uint32_t)((prop.pos / q->binLength) % q->numCols);
    if (q->usedRows[prop.r1] == q->usedKey || q->usedRows[prop.r2] == q->usedKey)
    {
        xoro_rollback(&q->rs->seeder);
        return 0;
    }
    if (prop.r1 == prop.r2 && prop.c1 == prop.c2)
    {
        domain_move(d, prop.atom1, prop.pos);
        return 1;
    }
    queue_push(q, &prop);
    q->usedRows[prop.r1] = q->usedKey;
    q->usedRows[prop.r2] = q->usedKey;
    queue_use_atom(q, p1);
    queue_add_move(q, p1, prop.pos);
    return 1;
}

/* ProposalQueue::exchange, ProposalQueue.cpp:250-283 */
static int queue_exchange(queue_t *q, domain_t *d)
{
    proposal_t prop;
    proposal_init(&prop, 'E', q->rs);
    prop.atom1 = domain_random_atom(d, &prop.rng);
    uint32_t right = domain_right(d, prop.atom1);
    prop.atom2 = right != NO_ATOM ? right : domain_front(d);
    uint64_t p1 = d->pool[prop.atom1].pos, p2 = d->pool[prop.atom2].pos;
    prop.r1 = (uint32_t)((p1 / q->binLength) / q->numCols);
    prop.c1 = (uint32_t)((p1 / q->binLength) % q->numCols);
    prop.r2 = (uint32_t)((p2 / q->binLength) / q->numCols);
    prop.c2 = (uint32_t)((p2 / q->binLength) % q->numCols);
    if (q->usedRows[prop.r1] == q->usedKey || q->usedRows[prop.r2] == q->usedKey)
    {
        xoro_rollback(&q->rs->seeder);
        return 0;
    }
    if (prop.r1 == prop.r2 && prop.c1 == prop.c2)
    {
        atom_t *a1 = &d->pool[prop.atom1], *a2 = &d->pool[prop.atom2];
        float newMass = rng_trunc_gamma_upper(&prop.rng, a1->mass + a2->mass, 1.f / q->lambda);
        float delta = (a1->mass > a2->mass) ? newMass - a1->mass : a2->mass - newMass;
        if (a1->mass + delta > EPSILON && a2->mass - delta > EPSILON)
        {
            float m1 = a1->mass + delta, m2 = a2->mass - delta;
            a1->mass = m1;
            a2->mass = m2;
        }
        return 1;
    }
    queue_push(q, &prop);
    q->usedRows[prop.r1] = q->usedKey;
    q->usedRows[prop.r2] = q->usedKey;
    return 1;
}

/* ProposalQueue::makeProposal, ProposalQueue.cpp:129-160 */
static int queue_make_proposal(queue_t *q, domain_t *d)
{
    q->u1 = q->useCachedRng ? q->u1 : rng_uniform(&q->rng);
    q->u2 = q->useCachedRng ? q->u2 : rng_uniform(&q->rng);
    q->useCachedRng = 0;
    if (q->minAtoms < 2 && q->maxAtoms >= 2) { return 0; }
    if (q->maxAtoms < 2) { return queue_birth(q, d); }
    float lowerBound = queue_death_prob(q, (double)q->minAtoms);
    float upperBound = queue_death_prob(q, (double)q->maxAtoms);
    if (q->u1 < 0.5f)
    {
        if (q->u2 < lowerBound) { return queue_death(q, d); }
        if (q->u2 >= upperBound) { return queue_birth(q, d); }
        return 0;
    }
    return (q->u1 < 0.75f) ? queue_move(q, d) : queue_exchange(q, d);
}

/* ProposalQueue::populate, ProposalQueue.cpp:53-76 */
static void queue_populate(queue_t *q, domain_t *d, unsigned limit)
{
    int success = 1;
    q->numProcessed = 0;
    while (q->numProcessed < limit && success)
    {
        if (!queue_make_proposal(q, d))
        {
            success = 0;
            q->useCachedRng = 1;
        }
        else
        {
            ++q->numProcessed;
        }
    }
}

/* ============================ samplers ============================================= */
typedef struct
{
    oracle_trace_record *rec;
    uint64_t capacity, count;
    uint32_t phase, iter;
} trace_t;

typedef struct
{
    model_t model;
    domain_t domain;
    queue_t queue;       /* asynchronous sampler */
    rng_t rng;           /* sequential sampler: SingleThreadedGibbsSampler::mRng */
    randstate_t *rs;
    float avgQueueLength, numQueueSamples;
    /* SingleThreadedGibbsSampler.h:52-61 */
    uint64_t binLength, numPatterns;
    double numBins, domainLength, alpha;
    int asynchronous;
    uint32_t side;
    trace_t *trace;
    /* row-parallel sweep mode (cgb_params.updateMode == CGB_UPDATE_SWEEP; no reference counterpart, see "sweep" below) */
    int sweep;
    uint64_t *swPos;      /* [nRows][swCap] atom positions relative to the row's domain segment, ascending */
    float *swMass;        /* [nRows][swCap] */
    uint32_t *swCount;    /* [nRows] */
    uint32_t swCap;
    uint64_t swTotal;     /* atoms in all rows */
    uint64_t swSteps, swScans, swOverflow, swUpdates;
} sampler_t;

static void sampler_init(sampler_t *s, const float *data, uint32_t nrow, uint32_t ncol, int transpose, int subsetRows,
                         float alpha, float maxGibbsMass, const cgb_params *p, randstate_t *rs,
                         const oracle_options *opt, int isA, trace_t *trace)
{
    memset(s, 0, sizeof(*s));
    model_init(&s->model, data, nrow, ncol, transpose, subsetRows, p, alpha, maxGibbsMass, opt, isA);
    uint64_t nElements = (uint64_t)s->model.nRows * s->model.k;
    domain_init(&s->domain, nElements);
    s->rs = rs;
    s->asynchronous = p->asynchronousUpdates != 0;
    s->side = isA ? 'A' : 'P';
    s->trace = trace;
    s->sweep = (p->updateMode == 1); /* CGB_UPDATE_SWEEP */
    s->alpha = (double)alpha;
    if (s->asynchronous)
    {
        /* AsynchronousGibbsSampler ctor, AsynchronousGibbsSampler.h:63-76 */
        queue_init(&s->queue, nElements, s->model.k, rs);
        s->queue.alpha = (double)alpha;
        s->queue.lambda = s->model.lambda;
    }
    else
    {
        /* SingleThreadedGibbsSampler ctor, SingleThreadedGibbsSampler.h:66-81 */
        rng_init(&s->rng, rs);
        s->numBins = (double)nElements;
        s->binLength = 0xFFFFFFFFFFFFFFFFull / nElements;
        s->numPatterns = s->model.k;
        s->domainLength = (double)(s->binLength * nElements);
        s->alpha = (double)alpha;
    }
}

static void sampler_free(sampler_t *s)
{
    model_free(&s->model);
    domain_free(&s->domain);
    if (s->asynchronous) { queue_free(&s->queue); }
    free(s->swPos);
    free(s->swMass);
    free(s->swCount);
}

static oracle_trace_record *trace_begin(sampler_t *s, const proposal_t *p, uint32_t batch)
{
    trace_t *t = s->trace;
    if (t == NULL || t->count >= t->capacity) { if (t) { t->count++; } return NULL; }
    oracle_trace_record *r = &t->rec[t->count++];
    memset(r, 0, sizeof(*r));
    r->phase = t->phase;
    r->iter = t->iter;
    r->side = s->side;
    r->batch = batch;
    r->type = (uint32_t)p->type;
    r->r1 = p->r1; r->c1 = p->c1; r->r2 = p->r2; r->c2 = p->c2;
    r->pos = p->pos;
    r->rngState = p->rng.state;
    if (p->atom1 != NO_ATOM) { r->atom1Pos = s->domain.pool[p->atom1].pos; r->mass1 = s->domain.pool[p->atom1].mass; }
    if (p->atom2 != NO_ATOM) { r->atom2Pos = s->domain.pool[p->atom2].pos; r->mass2 = s->domain.pool[p->atom2].mass; }
    return r;
}

/* AsynchronousGibbsSampler::birth, AsynchronousGibbsSampler.h:126-144 */
static void async_birth(sampler_t *s, proposal_t *p, oracle_trace_record *tr)
{
    model_t *m = &s->model;
    float mass = 0.f;
    int has;
    if (model_can_use_gibbs(m, p->c1))
    {
        alpha_t a = alpha_scale(model_alpha(m, p->r1, p->c1), m->annealingTemp); /* sampleBirth, DenseNormalModel.cpp:132-136 */
        if (tr) { tr->s = a.s; tr->s_mu = a.s_mu; }
        has = gibbs_mass(a, 0.f, m->maxGibbsMass, &p->rng, 1, m->lambda, &mass);
    }
    else
    {
        mass = rng_exponential(&p->rng, m->lambda);
        has = 1;
    }
    if (has && mass >= EPSILON)
    {
        ++s->queue.minAtoms; /* acceptBirth */
        s->domain.pool[p->atom1].mass = mass;
        model_change_matrix(m, p->r1, p->c1, mass);
        if (tr) { tr->accepted = 1; tr->newMass1 = mass; }
        return;
    }
    --s->queue.maxAtoms; /* rejectBirth */
    domain_cache_erase(&s->domain, p->atom1);
}

/* AsynchronousGibbsSampler::death, :147-180 */
static void async_death(sampler_t *s, proposal_t *p, oracle_trace_record *tr)
{
    model_t *m = &s->model;
    atom_t *a1 = &s->domain.pool[p->atom1];
    float rebirthMass = a1->mass;
    alpha_t a = alpha_scale(model_alpha_with_change(m, p->r1, p->c1, -1.f * a1->mass), m->annealingTemp);
    if (tr) { tr->s = a.s; tr->s_mu = a.s_mu; }
    if (model_can_use_gibbs(m, p->c1))
    {
        float g;
        if (gibbs_mass(a, 0.f, m->maxGibbsMass, &p->rng, 1, m->lambda, &g)) { rebirthMass = g; }
    }
    float deltaLL = rebirthMass * (a.s_mu - a.s * rebirthMass / 2.f);
    if (rs_logf(s->rs, rng_uniform(&p->rng)) < deltaLL)
    {
        ++s->queue.minAtoms; /* rejectDeath */
        if (rebirthMass != a1->mass)
        {
            model_safely_change_matrix(m, p->r1, p->c1, rebirthMass - a1->mass);
            a1->mass = rebirthMass;
        }
        if (tr) { tr->accepted = 1; tr->newMass1 = a1->mass; }
    }
    else
    {
        --s->queue.maxAtoms; /* acceptDeath */
        model_safely_change_matrix(m, p->r1, p->c1, -1.f * a1->mass);
        domain_cache_erase(&s->domain, p->atom1);
    }
}

/* AsynchronousGibbsSampler::move, :183-196; deltaLogLikelihood DenseNormalModel.cpp:125-130 */
static void async_move(sampler_t *s, proposal_t *p, oracle_trace_record *tr)
{
    model_t *m = &s->model;
    atom_t *a1 = &s->domain.pool[p->atom1];
    alpha_t a = alpha_scale(model_alpha2(m, p->r1, p->c1, p->r2, p->c2), m->annealingTemp);
    if (tr) { tr->s = a.s; tr->s_mu = a.s_mu; }
    float deltaLL = -1.f * a1->mass * (a.s_mu + a.s * a1->mass / 2.f);
    if (rs_logf(s->rs, rng_uniform(&p->rng)) < deltaLL)
    {
        domain_move(&s->domain, p->atom1, p->pos);
        model_safely_change_matrix(m, p->r1, p->c1, -a1->mass);
        model_change_matrix(m, p->r2, p->c2, a1->mass);
        if (tr) { tr->accepted = 1; tr->newMass1 = a1->mass; }
    }
}

/* AsynchronousGibbsSampler::exchange, :200-219; sampleExchange DenseNormalModel.cpp:154-159 */
static void async_exchange(sampler_t *s, proposal_t *p, oracle_trace_record *tr)
{
    model_t *m = &s->model;
    atom_t *a1 = &s->domain.pool[p->atom1];
    atom_t *a2 = &s->domain.pool[p->atom2];
    if (model_can_use_gibbs(m, p->c1) || model_can_use_gibbs(m, p->c2))
    {
        alpha_t a = alpha_scale(model_alpha2(m, p->r1, p->c1, p->r2, p->c2), m->annealingTemp);
        if (tr) { tr->s = a.s; tr->s_mu = a.s_mu; }
        float mass;
        int has = gibbs_mass(a, -a1->mass, a2->mass, &p->rng, 0, 0.f, &mass);
        float newMass1 = a1->mass + mass;
        float newMass2 = a2->mass - mass;
        if (has && newMass1 > EPSILON && newMass2 > EPSILON)
        {
            model_safely_change_matrix(m, p->r1, p->c1, newMass1 - a1->mass);
            model_safely_change_matrix(m, p->r2, p->c2, newMass2 - a2->mass);
            a1->mass = newMass1;
            a2->mass = newMass2;
            if (tr) { tr->accepted = 1; tr->newMass1 = newMass1; tr->newMass2 = newMass2; }
        }
    }
}

/* AsynchronousGibbsSampler::update, AsynchronousGibbsSampler.h:88-122.  The OpenMP loop over the
 * queue is replaced by a plain loop: proposals of one batch touch disjoint rows and atoms, the
 * min/max counters commute, and the erase cache is sorted before use, so the order is immaterial. */
static void async_update(sampler_t *s, unsigned nSteps)
{
    unsigned n = 0;
    uint32_t batch = 0;
    while (n < nSteps)
    {
        queue_populate(&s->queue, &s->domain, nSteps - n);
        n += s->queue.numProcessed;
        if (n < nSteps)
        {
            s->numQueueSamples += 1.f;
            s->avgQueueLength *= (s->numQueueSamples - 1.f) / s->numQueueSamples;
            s->avgQueueLength += (float)s->queue.n / s->numQueueSamples;
        }
        for (uint32_t i = 0; i < s->queue.n; ++i)
        {
            proposal_t *p = &s->queue.q[i];
            oracle_trace_record *tr = trace_begin(s, p, batch);
            switch (p->type)
            {
                case 'B': async_birth(s, p, tr); break;
                case 'D': async_death(s, p, tr); break;
                case 'M': async_move(s, p, tr); break;
                case 'E': async_exchange(s, p, tr); break;
            }
        }
        queue_clear(&s->queue);
        domain_flush_erase_cache(&s->domain);
        ++batch;
    }
}

/* SingleThreadedGibbsSampler::getUpdateType, SingleThreadedGibbsSampler.h:94-111 */
static char seq_update_type(sampler_t *s)
{
    if (s->domain.n < 2) { return 'B'; }
    float u1 = rng_uniform(&s->rng);
    if (u1 < 0.5f)
    {
        double nAtoms = (double)s->domain.n;
        double numer = nAtoms * s->domainLength;
        float deathProb = (float)(numer / (numer + s->alpha * s->numBins * (s->domainLength - nAtoms)));
        return rng_uniform(&s->rng) < deathProb ? 'D' : 'B';
    }
    return u1 < 0.75f ? 'M' : 'E';
}

/* SingleThreadedGibbsSampler::birth/death/move/exchange, SingleThreadedGibbsSampler.h:130-257 */
static void seq_birth(sampler_t *s)
{
    model_t *m = &s->model;
    uint64_t pos = domain_random_free_position(&s->domain, &s->rng);
    uint32_t row = (uint32_t)((pos / s->binLength) / s->numPatterns);
    uint32_t col = (uint32_t)((pos / s->binLength) % s->numPatterns);
    float mass = 0.f;
    int has;
    if (model_can_use_gibbs(m, col))
    {
        alpha_t a = alpha_scale(model_alpha(m, row, col), m->annealingTemp);
        has = gibbs_mass(a, 0.f, m->maxGibbsMass, &s->rng, 1, m->lambda, &mass);
    }
    else
    {
        mass = rng_exponential(&s->rng, m->lambda);
        has = 1;
    }
    if (has && mass > EPSILON)
    {
        domain_insert(&s->domain, pos, mass);
        model_change_matrix(m, row, col, mass);
    }
}

static void seq_death(sampler_t *s)
{
    model_t *m = &s->model;
    uint32_t id = domain_random_atom(&s->domain, &s->rng);
    atom_t *a1 = &s->domain.pool[id];
    uint32_t row = (uint32_t)((a1->pos / s->binLength) / s->numPatterns);
    uint32_t col = (uint32_t)((a1->pos / s->binLength) % s->numPatterns);
    float rebirthMass = a1->mass;
    alpha_t a = alpha_scale(model_alpha_with_change(m, row, col, -1.f * a1->mass), m->annealingTemp);
    if (model_can_use_gibbs(m, col))
    {
        float g;
        if (gibbs_mass(a, 0.f, m->maxGibbsMass, &s->rng, 1, m->lambda, &g)) { rebirthMass = g; }
    }
    float deltaLL = rebirthMass * (a.s_mu - a.s * rebirthMass / 2.f);
    if (rs_logf(s->rs, rng_uniform(&s->rng)) < deltaLL)
    {
        if (rebirthMass != a1->mass)
        {
            model_safely_change_matrix(m, row, col, rebirthMass - a1->mass);
            a1->mass = rebirthMass;
        }
    }
    else
    {
        model_safely_change_matrix(m, row, col, -1.f * a1->mass);
        domain_erase(&s->domain, id);
    }
}

static void seq_move(sampler_t *s)
{
    model_t *m = &s->model;
    uint32_t id = domain_random_atom(&s->domain, &s->rng);
    atom_t *a1 = &s->domain.pool[id];
    uint32_t left = domain_left(&s->domain, id), right = domain_right(&s->domain, id);
    uint64_t lbound = left != NO_ATOM ? s->domain.pool[left].pos : 0;
    uint64_t rbound = right != NO_ATOM ? s->domain.pool[right].pos : ref_double_to_u64(s->domainLength);
    uint64_t pos = rng_u64_range(&s->rng, lbound + 1, rbound - 1);
    uint32_t r1 = (uint32_t)((a1->pos / s->binLength) / s->numPatterns);
    uint32_t c1 = (uint32_t)((a1->pos / s->binLength) % s->numPatterns);
    uint32_t r2 = (uint32_t)((pos / s->binLength) / s->numPatterns);
    uint32_t c2 = (uint32_t)((pos / s->binLength) % s->numPatterns);
    if (r1 == r2 && c1 == c2)
    {
        domain_move(&s->domain, id, pos);
        return;
    }
    alpha_t a = alpha_scale(model_alpha2(m, r1, c1, r2, c2), m->annealingTemp);
    float deltaLL = -1.f * a1->mass * (a.s_mu + a.s * a1->mass / 2.f);
    if (rs_logf(s->rs, rng_uniform(&s->rng)) < deltaLL)
    {
        domain_move(&s->domain, id, pos);
        model_safely_change_matrix(m, r1, c1, -a1->mass);
        model_change_matrix(m, r2, c2, a1->mass);
    }
}

static void seq_exchange(sampler_t *s)
{
    model_t *m = &s->model;
    uint32_t id1 = domain_random_atom(&s->domain, &s->rng);
    uint32_t right = domain_right(&s->domain, id1);
    uint32_t id2 = right != NO_ATOM ? right : domain_front(&s->domain);
    atom_t *a1 = &s->domain.pool[id1], *a2 = &s->domain.pool[id2];
    uint32_t r1 = (uint32_t)((a1->pos / s->binLength) / s->numPatterns);
    uint32_t c1 = (uint32_t)((a1->pos / s->binLength) % s->numPatterns);
    uint32_t r2 = (uint32_t)((a2->pos / s->binLength) / s->numPatterns);
    uint32_t c2 = (uint32_t)((a2->pos / s->binLength) % s->numPatterns);
    if ((r1 != r2 || c1 != c2) && (model_can_use_gibbs(m, c1) || model_can_use_gibbs(m, c2)))
    {
        alpha_t a = alpha_scale(model_alpha2(m, r1, c1, r2, c2), m->annealingTemp);
        float mass;
        int has = gibbs_mass(a, -a1->mass, a2->mass, &s->rng, 0, 0.f, &mass);
        float newMass1 = a1->mass + mass;
        float newMass2 = a2->mass - mass;
        if (has && newMass1 > EPSILON && newMass2 > EPSILON)
        {
            model_safely_change_matrix(m, r1, c1, newMass1 - a1->mass);
            model_safely_change_matrix(m, r2, c2, newMass2 - a2->mass);
            a1->mass = newMass1;
            a2->mass = newMass2;
        }
    }
}

/* SingleThreadedGibbsSampler::update, :113-126 */
static void seq_update(sampler_t *s, unsigned nSteps)
{
    for (unsigned i = 0; i < nSteps; ++i)
    {
        switch (seq_update_type(s))
        {
            case 'B': seq_birth(s); break;
            case 'D': seq_death(s); break;
            case 'M': seq_move(s); break;
            case 'E': seq_exchange(s); break;
        }
    }
}


/* ============================ row-parallel sweep =================================== */
/* NO REFERENCE COUNTERPART AS A WHOLE.  The north star's throughput mode: "a conflict-partitioned commit step (atoms
 * binned by row so non-conflicting updates apply in parallel within one sweep)" with counter-based device-side draws.
 * It is a DIFFERENT CHAIN from the reference's for the same seed and is validated against the reference
 * statistically; this restatement is what the CUDA sweep kernel (cogaps_b200/csrc/sweep.cuh) is held to bit for bit.
 *
 * Why it is a valid sampler: row r of the factor matrix owns the contiguous segment [r*k*binLength, (r+1)*k*binLength)
 * of the atomic domain (ProposalQueue.cpp:172-173), the other factor is constant during update()
 * (GapsRunner.cpp:201-222), and the likelihood terms of different rows share no element of D / AP
 * (DenseNormalModel.cpp:162-240 reads row `row` only).  Given the other factor, the rows are therefore independent, and
 * every row runs the reference's own four proposal types — same evaluation code as the asynchronous sampler
 * (AsynchronousGibbsSampler.h:126-219) — on its own segment, sequentially, all rows at once.  What changes:
 *   - the birth/death balance uses the total atom count frozen at the start of update() (the reference tracks it
 *     proposal by proposal, ProposalQueue.cpp:123-160);
 *   - moves are bounded by the row's segment and the exchange partner of a row's last atom is the row's first atom
 *     (the reference bounds by the whole domain and wraps at its end, ProposalQueue.cpp:213-214,254);
 *   - draws come from Philox4x32-10 keyed by one seeder value per update() and countered by (row, proposal), so the
 *     result does not depend on how rows are scheduled and the draws can be formed ahead of the proposals.
 * Expected proposals of a row: the reference gives each of its nSteps proposals to a birth with probability
 * (1 - pDeath)/2, spread evenly over the rows, and to a death / move / exchange of a uniformly chosen atom with
 * probability pDeath/2, 1/4, 1/4 (ProposalQueue.cpp:129-160).  A row holding m of the n atoms therefore expects
 * nSteps * ((1-pDeath)/2/nRows + m*(pDeath/2 + 1/2)/n) proposals; it takes that many, rounded stochastically, and picks
 * each proposal's type with the same weights evaluated at its current atom count. */
#define CGB_UPDATE_SWEEP 1

static const uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u, kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

static void philox_block(const uint32_t ctrIn[4], const uint32_t keyIn[2], uint32_t out[4])
{
    uint32_t c0 = ctrIn[0], c1 = ctrIn[1], c2 = ctrIn[2], c3 = ctrIn[3];
    uint32_t k0 = keyIn[0], k1 = keyIn[1];
    for (int round = 0; round < 10; ++round)
    {
        uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
        uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* known-answer probe for tests/test_sweep.py */
void cogaps_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox_block(ctr, key, out); }

/* The draws of one proposal: two Philox blocks countered by (row, step, block, stream), so every proposal's random words
 * are a function of (key, row, step) alone and can be formed ahead of time by any thread.
 *   w[0] -> uniform that picks the proposal type      w[1] -> atom pick
 *   w[2], w[3] -> position draw (low, high word)        w[4], w[5] -> state of the proposal's own PCG stream (low, high)
 * step 0xFFFFFFFF holds the row-level draw (w[0]: stochastic rounding of the row's proposal count).
 * stream 0: the rows; stream 1: the transport between adjacent rows. */
static void sweep_draws(uint64_t key, uint32_t row, uint32_t step, uint32_t stream, uint32_t w[8])
{
    uint32_t k[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
    uint32_t c[4] = {row, step, 0, stream};
    philox_block(c, k, w);
    c[2] = 1;
    philox_block(c, k, w + 4);
}

static float u32_uniform(uint32_t v) { return (float)v / 4294967296.0f; }
static uint64_t u32_pair(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

static uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
static uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* exp() from IEEE +,*,fma and floor only, so host, oracle and device agree bit for bit: x = n ln2 + r, |r| <= ln2/2,
 * degree-13 Taylor polynomial in f64, scaled by 2^n; rounded once to f32 by the caller */
static double portable_exp_f64(double x)
{
    if (x != x) { return x; }
    if (x > 700.0) { return INFINITY; }
    if (x < -700.0) { return 0.0; }
    double n = floor(fma(x, 1.4426950408889634074, 0.5));
    double r = fma(-n, 0.69314718036912381649, x);      /* ln2 high part: 32 significant bits */
    r = fma(-n, 1.9082149292705877e-10, r);             /* ln2 low part */
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    union { double d; uint64_t u; } sc;
    sc.u = (uint64_t)((int64_t)n + 1023) << 52;          /* 2^n, n in [-1010, 1010] here */
    return p * sc.d;
}

float cogaps_oracle_portable_expf(float x) { return (float)portable_exp_f64((double)x); }

/* GapsRng::truncGammaUpper (Random.cpp:194-200) with the portable exp */
static float sweep_trunc_gamma_upper(rng_t *r, float b, float scale)
{
    float q = b / scale;
    float upper = 1.f - cogaps_oracle_portable_expf(-q) * (1.f + q);
    const unsigned ndx = (unsigned)rng_uniform_ab(r, 0.f, upper * 5000.f);
    return r->rs->qgamma[ndx] * scale;
}

static void sweep_alloc(sampler_t *s, uint32_t cap)
{
    uint32_t R = s->model.nRows;
    uint64_t *pos = (uint64_t*)calloc((size_t)R * cap, sizeof(uint64_t));
    float *mass = (float*)calloc((size_t)R * cap, sizeof(float));
    if (s->swPos)
    {
        for (uint32_t r = 0; r < R; ++r)
        {
            memcpy(pos + (size_t)r * cap, s->swPos + (size_t)r * s->swCap, sizeof(uint64_t) * s->swCount[r]);
            memcpy(mass + (size_t)r * cap, s->swMass + (size_t)r * s->swCap, sizeof(float) * s->swCount[r]);
        }
        free(s->swPos);
        free(s->swMass);
    }
    else
    {
        s->swCount = (uint32_t*)calloc(R, sizeof(uint32_t));
    }
    s->swPos = pos;
    s->swMass = mass;
    s->swCap = cap;
}

/* the constants of one update(), formed in f64 by the host side of the product in exactly these operations */
typedef struct
{
    double birthRow, deathAtom, moveAtom, exchAtom;
} sweep_rates_t;

static sweep_rates_t sweep_rates(uint64_t nAtoms, uint32_t nRows, uint32_t k, uint64_t binLength, double alpha)
{
    sweep_rates_t w;
    if (nAtoms < 2)
    {
        /* "always birth when 0 or 1 atoms exist", ProposalQueue.cpp:139-142 */
        w.birthRow = 1.0 / (double)nRows;
        w.deathAtom = w.moveAtom = w.exchAtom = 0.0;
        return w;
    }
    /* deathProb, ProposalQueue.cpp:123-127 (float result) */
    uint64_t nElements = (uint64_t)nRows * k;
    double domainLength = (double)(binLength * nElements);
    double numer = (double)nAtoms * domainLength;
    float pDeath = (float)(numer / (numer + alpha * (double)nElements * (domainLength - (double)nAtoms)));
    w.birthRow = 0.5 * (1.0 - (double)pDeath) / (double)nRows;
    w.deathAtom = 0.5 * (double)pDeath / (double)nAtoms;
    w.moveAtom = 0.25 / (double)nAtoms;
    w.exchAtom = 0.25 / (double)nAtoms;
    return w;
}

static void sweep_insert(uint64_t *pos, float *mass, uint32_t m, uint32_t idx, uint64_t p, float v)
{
    for (uint32_t i = m; i > idx; --i) { pos[i] = pos[i - 1]; mass[i] = mass[i - 1]; }
    pos[idx] = p;
    mass[idx] = v;
}

static void sweep_erase(uint64_t *pos, float *mass, uint32_t m, uint32_t idx)
{
    for (uint32_t i = idx; i + 1 < m; ++i) { pos[i] = pos[i + 1]; mass[i] = mass[i + 1]; }
}

static void sweep_row(sampler_t *s, uint32_t row, uint64_t key, unsigned nSteps, const sweep_rates_t *w, uint64_t binLength)
{
    model_t *m = &s->model;
    const uint32_t k = m->k;
    const uint64_t Lseg = binLength * k;
    uint64_t *pos = s->swPos + (size_t)row * s->swCap;
    float *mass = s->swMass + (size_t)row * s->swCap;
    uint32_t cnt = s->swCount[row];
    uint32_t dr[8];
    const double perAtom = w->deathAtom + w->moveAtom + w->exchAtom;
    double lam = (double)nSteps * (w->birthRow + (double)cnt * perAtom);
    if (lam > 1.0e9) { lam = 1.0e9; }
    uint32_t steps = (uint32_t)lam;
    float frac = (float)(lam - (double)steps);
    sweep_draws(key, row, 0xFFFFFFFFu, 0, dr);
    if (u32_uniform(dr[0]) < frac) { steps += 1; }
    for (uint32_t step = 0; step < steps; ++step)
    {
        s->swSteps += 1;
        sweep_draws(key, row, step, 0, dr);
        const double b = w->birthRow, d = (double)cnt * w->deathAtom, mv = (double)cnt * w->moveAtom;
        const double tot = b + (double)cnt * perAtom;
        const double x = (double)u32_uniform(dr[0]) * tot;
        rng_t rng;
        rng.rs = s->rs;
        rng.state = u32_pair(dr[4], dr[5]);
        if (cnt == 0 || x < b)
        {
            /* ProposalQueue::birth (:162-187) within the row's segment + AsynchronousGibbsSampler::birth (:126-144) */
            const uint64_t p = 1 + mulhi64(u32_pair(dr[2], dr[3]), Lseg - 1);
            uint32_t idx = 0;
            while (idx < cnt && pos[idx] < p) { ++idx; }
            if (idx < cnt && pos[idx] == p) { continue; }              /* occupied (randomFreePosition would redraw) */
            if (cnt == s->swCap) { s->swOverflow += 1; continue; }     /* the row's atom store is full */
            const uint32_t col = (uint32_t)(p / binLength);
            float v = 0.f;
            int has;
            if (model_can_use_gibbs(m, col))
            {
                s->swScans += 1;
                alpha_t a = alpha_scale(model_alpha(m, row, col), m->annealingTemp);
                has = gibbs_mass(a, 0.f, m->maxGibbsMass, &rng, 1, m->lambda, &v);
            }
            else
            {
                v = rng_exponential(&rng, m->lambda);
                has = 1;
            }
            if (has && v >= EPSILON)
            {
                sweep_insert(pos, mass, cnt, idx, p, v);
                cnt += 1;
                model_change_matrix(m, row, col, v);
            }
        }
        else if (x < b + d)
        {
            /* ProposalQueue::death (:189-207) + AsynchronousGibbsSampler::death (:147-180) */
            const uint32_t idx = mulhi32(dr[1], cnt);
            const uint32_t col = (uint32_t)(pos[idx] / binLength);
            const float old = mass[idx];
            float rebirth = old;
            s->swScans += 1;
            alpha_t a = alpha_scale(model_alpha_with_change(m, row, col, -1.f * old), m->annealingTemp);
            if (model_can_use_gibbs(m, col))
            {
                float gm;
                if (gibbs_mass(a, 0.f, m->maxGibbsMass, &rng, 1, m->lambda, &gm)) { rebirth = gm; }
            }
            float deltaLL = rebirth * (a.s_mu - a.s * rebirth / 2.f);
            if (rs_logf(s->rs, rng_uniform(&rng)) < deltaLL)
            {
                if (rebirth != old)
                {
                    model_safely_change_matrix(m, row, col, rebirth - old);
                    mass[idx] = rebirth;
                }
            }
            else
            {
                model_safely_change_matrix(m, row, col, -1.f * old);
                sweep_erase(pos, mass, cnt, idx);
                cnt -= 1;
            }
        }
        else if (x < b + d + mv)
        {
            /* ProposalQueue::move (:209-248) bounded by the segment + AsynchronousGibbsSampler::move (:183-196) */
            const uint32_t idx = mulhi32(dr[1], cnt);
            const uint64_t draw = u32_pair(dr[2], dr[3]);
            const uint64_t lb = idx > 0 ? pos[idx - 1] : 0;
            const uint64_t rb = idx + 1 < cnt ? pos[idx + 1] : Lseg;
            if (rb - lb < 2) { continue; }
            const uint64_t p = lb + 1 + mulhi64(draw, rb - lb - 1);
            const uint32_t c1 = (uint32_t)(pos[idx] / binLength), c2 = (uint32_t)(p / binLength);
            if (c1 == c2)
            {
                pos[idx] = p; /* "automatically accept moves in same bin" */
                continue;
            }
            s->swScans += 1;
            alpha_t a = alpha_scale(model_alpha2(m, row, c1, row, c2), m->annealingTemp);
            const float am = mass[idx];
            float deltaLL = -1.f * am * (a.s_mu + a.s * am / 2.f);
            if (rs_logf(s->rs, rng_uniform(&rng)) < deltaLL)
            {
                pos[idx] = p;
                model_safely_change_matrix(m, row, c1, -am);
                model_change_matrix(m, row, c2, am);
            }
        }
        else
        {
            /* ProposalQueue::exchange (:250-283) within the row + AsynchronousGibbsSampler::exchange (:200-219) */
            const uint32_t idx = mulhi32(dr[1], cnt);
            if (cnt < 2) { continue; }
            const uint32_t j = idx + 1 < cnt ? idx + 1 : 0;
            const uint32_t c1 = (uint32_t)(pos[idx] / binLength), c2 = (uint32_t)(pos[j] / binLength);
            const float m1 = mass[idx], m2 = mass[j];
            if (c1 == c2)
            {
                float newMass = sweep_trunc_gamma_upper(&rng, m1 + m2, 1.f / m->lambda);
                float delta = (m1 > m2) ? newMass - m1 : m2 - newMass;
                if (m1 + delta > EPSILON && m2 - delta > EPSILON)
                {
                    mass[idx] = m1 + delta;
                    mass[j] = m2 - delta;
                }
                continue;
            }
            if (model_can_use_gibbs(m, c1) || model_can_use_gibbs(m, c2))
            {
                s->swScans += 1;
                alpha_t a = alpha_scale(model_alpha2(m, row, c1, row, c2), m->annealingTemp);
                float gm;
                int has = gibbs_mass(a, -m1, m2, &rng, 0, 0.f, &gm);
                float n1 = m1 + gm, n2 = m2 - gm;
                if (has && n1 > EPSILON && n2 > EPSILON)
                {
                    model_safely_change_matrix(m, row, c1, n1 - m1);
                    model_safely_change_matrix(m, row, c2, n2 - m2);
                    mass[idx] = n1;
                    mass[j] = n2;
                }
            }
        }
    }
    s->swTotal = s->swTotal + cnt - s->swCount[row];
    s->swCount[row] = cnt;
}


/* ---- transport between adjacent rows ----
 * What the row-local sweep cannot do is carry an atom from one row to another, which the reference's move does all the
 * time (its bounds are the atom's neighbours in the WHOLE domain, ProposalQueue.cpp:213-214) and which its early,
 * annealed iterations rely on.  After the rows have run, the pairs of adjacent rows (r, r+1) — those with even r on
 * even-numbered updates, those with odd r on odd-numbered ones, so that concurrently handled pairs never share a row and
 * every boundary has its turn every other update — give the two atoms at their common
 * boundary, a = the last atom of row r and b = the first atom of row r+1, the reference's own proposals on the
 * two-row segment: move a or b to a uniform position between its neighbours there (ProposalQueue.cpp:209-248; across the
 * boundary that is a two-row move, AsynchronousGibbsSampler.h:183-196 with alphaParameters(r1,c1,r2,c2),
 * DenseNormalModel.cpp:186-214), or exchange mass between a and b (ProposalQueue.cpp:250-283).  The proposal
 * distributions are symmetric exactly as the reference's are: a move never changes the gap it is drawn from.  Expected
 * numbers follow the per-atom weights of the update. */
static void sweep_pair(sampler_t *s, uint32_t r, uint64_t key, unsigned nSteps, const sweep_rates_t *w, uint64_t binLength)
{
    model_t *m = &s->model;
    const uint64_t Lseg = binLength * m->k;
    uint64_t *posL = s->swPos + (size_t)r * s->swCap, *posR = posL + s->swCap;
    float *massL = s->swMass + (size_t)r * s->swCap, *massR = massL + s->swCap;
    uint32_t cl = s->swCount[r], cr = s->swCount[r + 1];
    uint32_t dr[8];
    double lam = (double)nSteps * ((cl > 0 ? w->moveAtom : 0.0) + (cr > 0 ? w->moveAtom : 0.0) + (cl > 0 && cr > 0 ? w->exchAtom : 0.0));
    if (lam > 1.0e9) { lam = 1.0e9; }
    uint32_t steps = (uint32_t)lam;
    float frac = (float)(lam - (double)steps);
    sweep_draws(key, r, 0xFFFFFFFFu, 1, dr);
    if (u32_uniform(dr[0]) < frac) { steps += 1; }
    for (uint32_t step = 0; step < steps; ++step)
    {
        s->swSteps += 1;
        sweep_draws(key, r, step, 1, dr);
        const double wa = cl > 0 ? w->moveAtom : 0.0, wb = cr > 0 ? w->moveAtom : 0.0, we = (cl > 0 && cr > 0) ? w->exchAtom : 0.0;
        const double x = (double)u32_uniform(dr[0]) * (wa + wb + we);
        rng_t rng;
        rng.rs = s->rs;
        rng.state = u32_pair(dr[4], dr[5]);
        if (cl == 0 && cr == 0) { continue; }
        if (x < wa + wb)
        {
            /* move a (leftA) or b within the two-row segment; positions in pair coordinates: row r+1 starts at Lseg */
            const int moveA = (cr == 0) || (cl > 0 && x < wa);
            const uint64_t draw = u32_pair(dr[2], dr[3]);
            uint64_t from, lb, rb;
            if (moveA)
            {
                from = posL[cl - 1];
                lb = cl > 1 ? posL[cl - 2] : 0;
                rb = cr > 0 ? Lseg + posR[0] : 2 * Lseg;
            }
            else
            {
                from = Lseg + posR[0];
                lb = cl > 0 ? posL[cl - 1] : 0;
                rb = cr > 1 ? Lseg + posR[1] : 2 * Lseg;
            }
            if (rb - lb < 2) { continue; }
            const uint64_t to = lb + 1 + mulhi64(draw, rb - lb - 1);
            const uint32_t r1 = from < Lseg ? r : r + 1, r2 = to < Lseg ? r : r + 1;
            const uint32_t c1 = (uint32_t)((from < Lseg ? from : from - Lseg) / binLength);
            const uint32_t c2 = (uint32_t)((to < Lseg ? to : to - Lseg) / binLength);
            const float am = moveA ? massL[cl - 1] : massR[0];
            if (r1 == r2 && c1 == c2)
            {
                if (moveA) { posL[cl - 1] = to; } else { posR[0] = to - Lseg; }
                continue;
            }
            if (r1 != r2 && (moveA ? cr : cl) == s->swCap) { s->swOverflow += 1; continue; }
            s->swScans += 1;
            alpha_t a = alpha_scale(model_alpha2(m, r1, c1, r2, c2), m->annealingTemp);
            float deltaLL = -1.f * am * (a.s_mu + a.s * am / 2.f);
            if (rs_logf(s->rs, rng_uniform(&rng)) < deltaLL)
            {
                model_safely_change_matrix(m, r1, c1, -am);
                model_change_matrix(m, r2, c2, am);
                if (r1 == r2)
                {
                    if (moveA) { posL[cl - 1] = to; } else { posR[0] = to - Lseg; }
                }
                else if (moveA)
                {
                    sweep_insert(posR, massR, cr, 0, to - Lseg, am); /* the new first atom of row r+1 */
                    cr += 1;
                    cl -= 1;
                }
                else
                {
                    posL[cl] = to;                                   /* the new last atom of row r */
                    massL[cl] = am;
                    cl += 1;
                    sweep_erase(posR, massR, cr, 0);
                    cr -= 1;
                }
            }
        }
        else
        {
            const uint32_t c1 = (uint32_t)(posL[cl - 1] / binLength), c2 = (uint32_t)(posR[0] / binLength);
            const float m1 = massL[cl - 1], m2 = massR[0];
            if (model_can_use_gibbs(m, c1) || model_can_use_gibbs(m, c2))
            {
                s->swScans += 1;
                alpha_t a = alpha_scale(model_alpha2(m, r, c1, r + 1, c2), m->annealingTemp);
                float gm;
                int has = gibbs_mass(a, -m1, m2, &rng, 0, 0.f, &gm);
                float n1 = m1 + gm, n2 = m2 - gm;
                if (has && n1 > EPSILON && n2 > EPSILON)
                {
                    model_safely_change_matrix(m, r, c1, n1 - m1);
                    model_safely_change_matrix(m, r + 1, c2, n2 - m2);
                    massL[cl - 1] = n1;
                    massR[0] = n2;
                }
            }
        }
    }
    s->swCount[r] = cl;
    s->swCount[r + 1] = cr;
}

static void sweep_update(sampler_t *s, unsigned nSteps)
{
    model_t *m = &s->model;
    if (s->swPos == NULL) { sweep_alloc(s, 64); }
    const uint64_t nElements = (uint64_t)m->nRows * m->k;
    const uint64_t binLength = 0xFFFFFFFFFFFFFFFFull / nElements;
    const uint64_t key = xoro_next(&s->rs->seeder);       /* one seeder value per update() */
    const sweep_rates_t w = sweep_rates(s->swTotal, m->nRows, m->k, binLength, s->alpha);
    for (uint32_t row = 0; row < m->nRows; ++row) { sweep_row(s, row, key, nSteps, &w, binLength); }
    {
        /* pairs (r, r+1) with even r on even-numbered updates of this sampler, odd r on odd-numbered ones */
        const uint32_t colour = (uint32_t)(s->swUpdates++ & 1u);
        for (uint32_t row = colour; row + 1 < m->nRows; row += 2) { sweep_pair(s, row, key, nSteps, &w, binLength); }
    }
    /* the store grows between updates so that a row practically never fills up within one */
    uint32_t maxCount = 0;
    for (uint32_t row = 0; row < m->nRows; ++row) { if (s->swCount[row] > maxCount) { maxCount = s->swCount[row]; } }
    if (2 * maxCount > s->swCap) { sweep_alloc(s, (4 * maxCount + 31u) / 32u * 32u); }
}

static void sampler_update(sampler_t *s, unsigned nSteps)
{
    if (s->sweep) { sweep_update(s, nSteps); return; }
    if (s->asynchronous) { async_update(s, nSteps); } else { seq_update(s, nSteps); }
}

static unsigned sampler_n_atoms(const sampler_t *s) { return s->sweep ? (unsigned)s->swTotal : s->domain.n; }

/* ============================ statistics =========================================== */
/* GapsStatistics, GapsStatistics.h:17-64; sums stored column-major like the reference */
typedef struct
{
    uint32_t nGenes, nSamples, k;
    float *Amean, *Astd, *Pmean, *Pstd, *pump;
    unsigned statUpdates, pumpUpdates;
} stats_t;

static void stats_init(stats_t *st, uint32_t g, uint32_t s, uint32_t k)
{
    st->nGenes = g; st->nSamples = s; st->k = k;
    st->Amean = (float*)calloc((size_t)g * k, sizeof(float));
    st->Astd = (float*)calloc((size_t)g * k, sizeof(float));
    st->Pmean = (float*)calloc((size_t)s * k, sizeof(float));
    st->Pstd = (float*)calloc((size_t)s * k, sizeof(float));
    st->pump = (float*)calloc((size_t)g * k, sizeof(float));
    st->statUpdates = 0;
    st->pumpUpdates = 0;
}

static void stats_free(stats_t *st)
{
    free(st->Amean); free(st->Astd); free(st->Pmean); free(st->Pstd); free(st->pump);
}

/* GapsStatistics::update / updateA / updateP, GapsStatistics.h:129-185.  mode 0 both, 1 A only, 2 P only */
static void stats_update(stats_t *st, const model_t *A, const model_t *P, int mode)
{
    ++st->statUpdates;
    for (uint32_t j = 0; j < st->k; ++j)
    {
        const float *pc = P->M + (size_t)j * P->nRows;
        const float *ac = A->M + (size_t)j * A->nRows;
        float norm = 0.f; /* gaps::max(Vector), VectorMath.cpp: starts from 0 */
        for (uint32_t i = 0; i < P->nRows; ++i) { norm = (pc[i] > norm) ? pc[i] : norm; }
        norm = (norm == 0.f) ? 1.f : norm;
        if (mode != 0) { norm = 1.f; }
        if (mode == 0 || mode == 2)
        {
            for (uint32_t i = 0; i < P->nRows; ++i)
            {
                float quot = pc[i] / norm;
                st->Pmean[(size_t)j * P->nRows + i] += quot;
                st->Pstd[(size_t)j * P->nRows + i] += quot * quot;
            }
        }
        if (mode == 0 || mode == 1)
        {
            for (uint32_t i = 0; i < A->nRows; ++i)
            {
                float prod = ac[i] * norm;
                st->Amean[(size_t)j * A->nRows + i] += prod;
                st->Astd[(size_t)j * A->nRows + i] += prod * prod;
            }
        }
    }
}

/* pumpMatrixCutThreshold / UniqueThreshold (identical bodies), GapsStatistics.h:66-117 */
static void pump_threshold(const float *Mcolmajor, uint32_t nRows, uint32_t k, float *stat)
{
    float *maxValues = (float*)calloc(nRows, sizeof(float));
    uint32_t *maxIdx = (uint32_t*)calloc(nRows, sizeof(uint32_t));
    for (uint32_t j = 0; j < k; ++j)
    {
        for (uint32_t i = 0; i < nRows; ++i)
        {
            float v = Mcolmajor[(size_t)j * nRows + i];
            if (maxValues[i] < v) { maxValues[i] = v; maxIdx[i] = j; }
        }
    }
    for (uint32_t i = 0; i < nRows; ++i) { stat[(size_t)maxIdx[i] * nRows + i] += 1.f; }
    free(maxValues);
    free(maxIdx);
}

static void out_rowmajor(const float *colmajor, uint32_t rows, uint32_t k, float *out, float div)
{
    if (!out) { return; }
    for (uint32_t i = 0; i < rows; ++i)
    {
        for (uint32_t j = 0; j < k; ++j) { out[(size_t)i * k + j] = colmajor[(size_t)j * rows + i] / div; }
    }
}

/* Asd / Psd, GapsStatistics.cpp:23-61 */
static void out_sd(const float *meanSum, const float *sqSum, uint32_t rows, uint32_t k, unsigned n, float *out)
{
    if (!out) { return; }
    for (uint32_t i = 0; i < rows; ++i)
    {
        for (uint32_t j = 0; j < k; ++j)
        {
            float ms = meanSum[(size_t)j * rows + i];
            float meanTerm = (ms * ms) / (float)n;
            float numer = fmaxr(0.f, sqSum[(size_t)j * rows + i] - meanTerm);
            out[(size_t)i * k + j] = sqrtf(numer / ((float)n - 1.f));
        }
    }
}

/* meanChiSq(DenseNormalModel), GapsStatistics.cpp:63-87; model = P sampler (D is genes x samples) */
static float stats_mean_chisq(const stats_t *st, const model_t *P)
{
    float chisq = 0.f;
    uint32_t nG = P->L, nS = P->nRows;
    float n = (float)st->statUpdates;
    for (uint32_t i = 0; i < nG; ++i)
    {
        for (uint32_t j = 0; j < nS; ++j)
        {
            float mm = 0.f;
            for (uint32_t c = 0; c < st->k; ++c)
            {
                mm += st->Amean[(size_t)c * nG + i] * st->Pmean[(size_t)c * nS + j];
            }
            mm /= (n * n);
            float d = P->D[(size_t)j * P->L + i];
            float s = P->S[(size_t)j * P->L + i];
            chisq += ((d - mm) * (d - mm)) / (s * s);
        }
    }
    return chisq;
}

/* ============================ checkpoints ========================================== */
/* The Archive wire format (utils/Archive.h:16-87): raw little-endian scalars behind a u32 magic number, written in
 * the order createCheckpoint streams them (GapsRunner.cpp:237-240).  Asynchronous sampler only: the reference's
 * SingleThreadedGibbsSampler does not save its rng and its operator>> writes instead of reading
 * (SingleThreadedGibbsSampler.h:260-273), so a sequential run cannot be resumed there either. */
#define ARCHIVE_MAGIC 0xB123AA4Du

typedef struct { FILE *f; int ok; } archive_t;

static void ar_put(archive_t *ar, const void *v, size_t n) { if (ar->ok && fwrite(v, 1, n, ar->f) != n) { ar->ok = 0; } }
static void ar_get(archive_t *ar, void *v, size_t n) { if (ar->ok && fread(v, 1, n, ar->f) != n) { ar->ok = 0; } }
#define AR_PUT(ar, type, val) do { type tmp_ = (type)(val); ar_put((ar), &tmp_, sizeof(tmp_)); } while (0)
#define AR_GET(ar, lvalue) ar_get((ar), &(lvalue), sizeof(lvalue))

/* Matrix << (Matrix.cpp:182-190): nRows nCols, then every column as a Vector (Vector.cpp:90-98: size, floats) */
static void ar_put_matrix(archive_t *ar, const float *colMajor, uint32_t rows, uint32_t cols)
{
    AR_PUT(ar, uint32_t, rows);
    AR_PUT(ar, uint32_t, cols);
    for (uint32_t c = 0; c < cols; ++c)
    {
        AR_PUT(ar, uint32_t, rows);
        ar_put(ar, colMajor + (size_t)c * rows, sizeof(float) * rows);
    }
}

static void ar_get_matrix(archive_t *ar, float *colMajor, uint32_t rows, uint32_t cols)
{
    uint32_t nr = 0, nc = 0;
    AR_GET(ar, nr);
    AR_GET(ar, nc);
    if (nr != rows || nc != cols) { ar->ok = 0; return; } /* GAPS_ASSERT in Matrix.cpp:196-197 */
    for (uint32_t c = 0; c < cols && ar->ok; ++c)
    {
        uint32_t sz = 0;
        AR_GET(ar, sz);
        if (sz != rows) { ar->ok = 0; return; }
        ar_get(ar, colMajor + (size_t)c * rows, sizeof(float) * rows);
    }
}

/* DenseNormalModel << (DenseNormalModel.cpp:260-264) / SparseNormalModel << (SparseNormalModel.cpp:313-317 ->
 * HybridMatrix.cpp:85-97: rows as Vectors, columns as HybridVectors with their flag words, then beta) */
static void ar_put_model(archive_t *ar, const model_t *m)
{
    if (!m->sparse) { ar_put_matrix(ar, m->M, m->nRows, m->k); return; }
    AR_PUT(ar, uint32_t, m->nRows);
    AR_PUT(ar, uint32_t, m->k);
    for (uint32_t r = 0; r < m->nRows; ++r)
    {
        AR_PUT(ar, uint32_t, m->k);
        ar_put(ar, m->Mrows + (size_t)r * m->k, sizeof(float) * m->k);
    }
    uint32_t nWords = m->nRows / 64 + 1; /* HybridVector.cpp:12 */
    uint64_t *flags = (uint64_t*)malloc(sizeof(uint64_t) * nWords);
    for (uint32_t c = 0; c < m->k; ++c)
    {
        const float *col = m->M + (size_t)c * m->nRows;
        memset(flags, 0, sizeof(uint64_t) * nWords);
        for (uint32_t r = 0; r < m->nRows; ++r) { if (col[r] != 0.f) { flags[r / 64] |= 1ull << (r % 64); } }
        AR_PUT(ar, uint32_t, m->nRows);
        ar_put(ar, flags, sizeof(uint64_t) * nWords);
        ar_put(ar, col, sizeof(float) * m->nRows);
    }
    free(flags);
    AR_PUT(ar, float, m->beta);
}

static void ar_get_model(archive_t *ar, model_t *m)
{
    if (!m->sparse) { ar_get_matrix(ar, m->M, m->nRows, m->k); return; }
    uint32_t nr = 0, nc = 0;
    AR_GET(ar, nr);
    AR_GET(ar, nc);
    if (nr != m->nRows || nc != m->k) { ar->ok = 0; return; }
    for (uint32_t r = 0; r < m->nRows && ar->ok; ++r)
    {
        uint32_t sz = 0;
        AR_GET(ar, sz);
        if (sz != m->k) { ar->ok = 0; return; }
        ar_get(ar, m->Mrows + (size_t)r * m->k, sizeof(float) * m->k);
    }
    uint32_t nWords = m->nRows / 64 + 1;
    uint64_t *flags = (uint64_t*)malloc(sizeof(uint64_t) * nWords);
    for (uint32_t c = 0; c < m->k && ar->ok; ++c)
    {
        uint32_t sz = 0;
        AR_GET(ar, sz);
        if (sz != m->nRows) { ar->ok = 0; break; }
        ar_get(ar, flags, sizeof(uint64_t) * nWords); /* implied by the values: set iff the value is not 0 */
        ar_get(ar, m->M + (size_t)c * m->nRows, sizeof(float) * m->nRows);
    }
    free(flags);
    AR_GET(ar, m->beta);
}

/* AsynchronousGibbsSampler << (AsynchronousGibbsSampler.h:221-226): model, ConcurrentAtomicDomain
 * (ConcurrentAtomicDomain.cpp:134-142: atoms in pick-vector order), ProposalQueue (ProposalQueue.cpp:285-291) */
static void ar_put_sampler(archive_t *ar, const sampler_t *s)
{
    ar_put_model(ar, &s->model);
    AR_PUT(ar, uint64_t, s->domain.domainLength);
    AR_PUT(ar, uint64_t, s->domain.n);
    for (uint32_t i = 0; i < s->domain.n; ++i)
    {
        const atom_t *a = &s->domain.pool[s->domain.vec[i]];
        AR_PUT(ar, uint64_t, a->pos);
        AR_PUT(ar, float, a->mass);
    }
    const queue_t *q = &s->queue;
    AR_PUT(ar, uint64_t, q->rng.state);
    AR_PUT(ar, uint64_t, q->minAtoms);
    AR_PUT(ar, uint64_t, q->maxAtoms);
    AR_PUT(ar, uint64_t, q->binLength);
    AR_PUT(ar, uint64_t, q->numCols);
    AR_PUT(ar, double, q->alpha);
    AR_PUT(ar, double, q->domainLength);
    AR_PUT(ar, double, q->numBins);
    AR_PUT(ar, float, q->lambda);
    AR_PUT(ar, uint8_t, q->useCachedRng ? 1 : 0);
    AR_PUT(ar, float, q->u1);
    AR_PUT(ar, float, q->u2);
}

static void ar_get_sampler(archive_t *ar, sampler_t *s)
{
    ar_get_model(ar, &s->model);
    uint64_t n = 0;
    AR_GET(ar, s->domain.domainLength);
    AR_GET(ar, n);
    for (uint64_t i = 0; i < n && ar->ok; ++i) /* ConcurrentAtomicDomain.cpp:144-155: inserted in archive order */
    {
        uint64_t pos = 0;
        float mass = 0.f;
        AR_GET(ar, pos);
        AR_GET(ar, mass);
        if (ar->ok) { domain_insert(&s->domain, pos, mass); }
    }
    queue_t *q = &s->queue;
    uint8_t cached = 0;
    AR_GET(ar, q->rng.state);
    AR_GET(ar, q->minAtoms);
    AR_GET(ar, q->maxAtoms);
    AR_GET(ar, q->binLength);
    AR_GET(ar, q->numCols);
    AR_GET(ar, q->alpha);
    AR_GET(ar, q->domainLength);
    AR_GET(ar, q->numBins);
    AR_GET(ar, q->lambda);
    AR_GET(ar, cached);
    AR_GET(ar, q->u1);
    AR_GET(ar, q->u2);
    q->useCachedRng = cached != 0;
}

/* GapsParameters << (GapsParameters.cpp:82-88) */
static void ar_put_params(archive_t *ar, const cgb_params *p, uint32_t nGenes, uint32_t nSamples, uint32_t interval)
{
    AR_PUT(ar, uint32_t, p->seed);
    AR_PUT(ar, uint32_t, nGenes);
    AR_PUT(ar, uint32_t, nSamples);
    AR_PUT(ar, uint32_t, p->nPatterns);
    AR_PUT(ar, uint32_t, p->nIterations);
    AR_PUT(ar, float, p->alphaA);
    AR_PUT(ar, float, p->alphaP);
    AR_PUT(ar, float, p->maxGibbsMassA);
    AR_PUT(ar, float, p->maxGibbsMassP);
    AR_PUT(ar, uint8_t, p->useSparseOptimization ? 1 : 0);
    AR_PUT(ar, uint32_t, interval);
}

static void ar_get_params(archive_t *ar, cgb_params *p, uint32_t *nGenes, uint32_t *nSamples, uint32_t *interval)
{
    uint8_t sparse = 0;
    AR_GET(ar, p->seed);
    AR_GET(ar, *nGenes);
    AR_GET(ar, *nSamples);
    AR_GET(ar, p->nPatterns);
    AR_GET(ar, p->nIterations);
    AR_GET(ar, p->alphaA);
    AR_GET(ar, p->alphaP);
    AR_GET(ar, p->maxGibbsMassA);
    AR_GET(ar, p->maxGibbsMassP);
    AR_GET(ar, sparse);
    AR_GET(ar, *interval);
    p->useSparseOptimization = sparse != 0;
}

static int archive_open(archive_t *ar, const char *path, int write)
{
    ar->f = fopen(path, write ? "wb" : "rb");
    ar->ok = ar->f != NULL;
    if (!ar->ok) { return 0; }
    if (write) { AR_PUT(ar, uint32_t, ARCHIVE_MAGIC); }
    else
    {
        uint32_t magic = 0;
        AR_GET(ar, magic);
        if (magic != ARCHIVE_MAGIC) { ar->ok = 0; } /* "incompatible checkpoint file", Archive.h:33-36 */
    }
    return ar->ok;
}

static int archive_close(archive_t *ar)
{
    if (ar->f && fclose(ar->f) != 0) { ar->ok = 0; }
    ar->f = NULL;
    return ar->ok;
}

/* ============================ run loop ============================================= */
static void snapshot(const model_t *m, float *dst)
{
    if (!dst) { return; }
    for (uint32_t r = 0; r < m->nRows; ++r)
    {
        for (uint32_t c = 0; c < m->k; ++c) { dst[(size_t)r * m->k + c] = m->sparse ? m->Mrows[(size_t)r * m->k + c] : m->M[(size_t)c * m->nRows + r]; }
    }
}

/* runCoGAPSAlgorithm + runOnePhase + updateSampler + displayStatus, GapsRunner.cpp:161-222,272-327,381-503 */
int cogaps_oracle_run_trace(const float *data, uint32_t nrow, uint32_t ncol, const float *uncertainty,
                            const cgb_params *p0, cgb_result *r, const oracle_options *opt,
                            oracle_trace_record *traceBuf, uint64_t capacity, uint64_t *count)
{
    cgb_params pv = *p0;
    const cgb_params *p = &pv;
    uint32_t ckInterval = opt ? opt->checkpointInterval : 0;
    const char *ckIn = (opt && opt->checkpointInFile && opt->checkpointInFile[0]) ? opt->checkpointInFile : NULL;
    const char *ckOut = (opt && opt->checkpointOutFile && opt->checkpointOutFile[0]) ? opt->checkpointOutFile : "gaps_checkpoint.out";
    if ((ckInterval > 0 || ckIn) && !p->asynchronousUpdates) { return -5; } /* see "checkpoints" above */
    randstate_t *rs = (randstate_t*)malloc(sizeof(randstate_t));
    randstate_init(rs, p->seed, opt);
    if (ckIn)
    {
        /* run_helper, GapsRunner.cpp:99-105: parameters and the seeder come from the file before anything is built */
        archive_t ar;
        uint32_t fileGenes = 0, fileSamples = 0;
        if (!archive_open(&ar, ckIn, 0)) { archive_close(&ar); free(rs); return -1; }
        ar_get_params(&ar, &pv, &fileGenes, &fileSamples, &ckInterval);
        AR_GET(&ar, rs->seeder.s[0]);
        AR_GET(&ar, rs->seeder.s[1]);
        if (!archive_close(&ar)) { free(rs); return -1; }
    }
    trace_t trace;
    memset(&trace, 0, sizeof(trace));
    trace.rec = traceBuf;
    trace.capacity = traceBuf ? capacity : 0;

    uint32_t nGenes = p->transposeData ? ncol : nrow;
    uint32_t nSamples = p->transposeData ? nrow : ncol;
    if (p->nSubsetIndices && p->subsetGenes) { nGenes = p->nSubsetIndices; }
    if (p->nSubsetIndices && !p->subsetGenes) { nSamples = p->nSubsetIndices; }

    sampler_t A, P;
    sampler_init(&A, data, nrow, ncol, !p->transposeData, !p->subsetGenes, p->alphaA, p->maxGibbsMassA, p, rs, opt, 1, traceBuf ? &trace : NULL);
    sampler_init(&P, data, nrow, ncol, p->transposeData, p->subsetGenes, p->alphaP, p->maxGibbsMassP, p, rs, opt, 0, traceBuf ? &trace : NULL);
    if (uncertainty)
    {
        model_set_uncertainty(&A.model, uncertainty, nrow, ncol, !p->transposeData, !p->subsetGenes, p);
        model_set_uncertainty(&P.model, uncertainty, nrow, ncol, p->transposeData, p->subsetGenes, p);
    }
    int fixed = p->whichMatrixFixed;
    int useFixed = p->fixedPatterns != NULL && fixed != 'N';
    if (useFixed)
    {
        if (fixed == 'A') { model_set_matrix(&A.model, p->fixedPatterns); }
        if (fixed == 'P') { model_set_matrix(&P.model, p->fixedPatterns); }
    }
    stats_t st;
    stats_init(&st, nGenes, nSamples, p->nPatterns);
    rng_t rng;
    rng_init(&rng, rs);
    int rc = 0;
    uint32_t nCheckpoints = 0;
    int startPhase = CGB_PHASE_EQUILIBRATION;
    uint32_t startIter = 0;
    if (ckIn)
    {
        /* processCheckpoint, GapsRunner.cpp:258-270 */
        archive_t ar;
        cgb_params again = pv;
        uint32_t g2 = 0, s2 = 0, interval2 = 0;
        int32_t iPhase = 0;
        archive_open(&ar, ckIn, 0);
        ar_get_params(&ar, &again, &g2, &s2, &interval2);
        if (ar.ok && (g2 != nGenes || s2 != nSamples)) { ar.ok = 0; } /* the data is not what the checkpoint was made from */
        AR_GET(&ar, rs->seeder.s[0]);
        AR_GET(&ar, rs->seeder.s[1]);
        if (ar.ok) { ar_get_sampler(&ar, &A); }
        if (ar.ok) { ar_get_sampler(&ar, &P); }
        if (ar.ok) { ar_get_matrix(&ar, st.Amean, nGenes, p->nPatterns); }
        if (ar.ok) { ar_get_matrix(&ar, st.Astd, nGenes, p->nPatterns); }
        if (ar.ok) { ar_get_matrix(&ar, st.Pmean, nSamples, p->nPatterns); }
        if (ar.ok) { ar_get_matrix(&ar, st.Pstd, nSamples, p->nPatterns); }
        AR_GET(&ar, st.statUpdates);
        AR_GET(&ar, st.k);
        AR_GET(&ar, iPhase);
        AR_GET(&ar, startIter);
        AR_GET(&ar, rng.state);
        if (!archive_close(&ar) || (iPhase != CGB_PHASE_EQUILIBRATION && iPhase != CGB_PHASE_SAMPLING)) { rc = -1; goto done; }
        startPhase = iPhase;
    }

    model_sync(&A.model, &P.model);
    model_sync(&P.model, &A.model);
    model_extra_initialization(&A.model);
    model_extra_initialization(&P.model);

    uint64_t totalUpdates = 0;
    uint32_t nHist = 0, nSnapEq = 0, nSnapSamp = 0;
    for (int phase = startPhase; phase <= CGB_PHASE_SAMPLING; ++phase)
    {
        for (uint32_t iter = (phase == startPhase) ? startIter : 0; iter < p->nIterations; ++iter)
        {
            trace.phase = (uint32_t)phase;
            trace.iter = iter;
            /* createCheckpoint, GapsRunner.cpp:226-256 */
            if (ckInterval > 0 && ((iter + 1) % ckInterval) == 0 && p->nSubsetIndices == 0)
            {
                size_t len = strlen(ckOut);
                char *backup = (char*)malloc(len + 8);
                memcpy(backup, ckOut, len);
                memcpy(backup + len, ".backup", 8);
                rename(ckOut, backup);
                archive_t ar;
                archive_open(&ar, ckOut, 1);
                ar_put_params(&ar, p, nGenes, nSamples, ckInterval);
                AR_PUT(&ar, uint64_t, rs->seeder.s[0]);
                AR_PUT(&ar, uint64_t, rs->seeder.s[1]);
                ar_put_sampler(&ar, &A);
                ar_put_sampler(&ar, &P);
                ar_put_matrix(&ar, st.Amean, nGenes, p->nPatterns);
                ar_put_matrix(&ar, st.Astd, nGenes, p->nPatterns);
                ar_put_matrix(&ar, st.Pmean, nSamples, p->nPatterns);
                ar_put_matrix(&ar, st.Pstd, nSamples, p->nPatterns);
                AR_PUT(&ar, uint32_t, st.statUpdates);
                AR_PUT(&ar, uint32_t, st.k);
                AR_PUT(&ar, int32_t, phase);
                AR_PUT(&ar, uint32_t, iter);
                AR_PUT(&ar, uint64_t, rng.state);
                int written = archive_close(&ar);
                remove(backup);
                free(backup);
                if (!written) { rc = -1; goto done; }
                if (opt && opt->stopAfterCheckpoints > 0 && ++nCheckpoints == opt->stopAfterCheckpoints) { rc = -7; goto done; }
                /* "running the extra initialization here allows for consistency with runs started from a checkpoint" */
                model_extra_initialization(&A.model);
                model_extra_initialization(&P.model);
            }
            if (phase == CGB_PHASE_EQUILIBRATION)
            {
                float temp = (float)(2 * iter) / (float)p->nIterations;
                A.model.annealingTemp = fminr(1.f, temp);
                P.model.annealingTemp = fminr(1.f, temp);
            }
            unsigned atomsA = sampler_n_atoms(&A), atomsP = sampler_n_atoms(&P);
            unsigned nA = (unsigned)rng_poisson(&rng, (double)(atomsA < 10 ? 10u : atomsA));
            unsigned nP = (unsigned)rng_poisson(&rng, (double)(atomsP < 10 ? 10u : atomsP));
            /* updateSampler, GapsRunner.cpp:201-222 */
            if (fixed != 'A')
            {
                sampler_update(&A, nA);
                if (fixed != 'P') { model_sync(&P.model, &A.model); }
            }
            if (fixed != 'P')
            {
                sampler_update(&P, nP);
                if (fixed != 'A') { model_sync(&A.model, &P.model); }
            }
            totalUpdates += (A.sweep || P.sweep) ? 0 : nA + nP;
            if (phase == CGB_PHASE_SAMPLING)
            {
                if (useFixed)
                {
                    if (fixed == 'A') { stats_update(&st, &A.model, &P.model, 2); }
                    else { stats_update(&st, &A.model, &P.model, 1); }
                }
                else
                {
                    stats_update(&st, &A.model, &P.model, 0);
                    if (p->takePumpSamples)
                    {
                        ++st.pumpUpdates;
                        pump_threshold(A.model.M, A.model.nRows, A.model.k, st.pump);
                    }
                }
            }
            if ((int)p->snapshotPhase == phase || p->snapshotPhase == CGB_PHASE_ALL)
            {
                if (p->snapshotFrequency > 0 && ((iter + 1) % p->snapshotFrequency) == 0)
                {
                    uint32_t slot = nSnapEq + nSnapSamp;
                    if (slot < r->snapshotCapacity)
                    {
                        snapshot(&A.model, r->snapshotsA ? r->snapshotsA + (size_t)slot * nGenes * p->nPatterns : NULL);
                        snapshot(&P.model, r->snapshotsP ? r->snapshotsP + (size_t)slot * nSamples * p->nPatterns : NULL);
                    }
                    if (phase == CGB_PHASE_EQUILIBRATION) { ++nSnapEq; } else { ++nSnapSamp; }
                }
            }
            if (p->outputFrequency > 0 && ((iter + 1) % p->outputFrequency) == 0)
            {
                float cs = (fixed == 'P') ? model_chisq(&A.model) : model_chisq(&P.model);
                if (nHist < r->historyCapacity)
                {
                    if (r->chisqHistory) { r->chisqHistory[nHist] = cs; }
                    if (r->atomHistoryA) { r->atomHistoryA[nHist] = sampler_n_atoms(&A); }
                    if (r->atomHistoryP) { r->atomHistoryP[nHist] = sampler_n_atoms(&P); }
                    ++nHist;
                }
            }
        }
    }

    out_rowmajor(st.Amean, nGenes, p->nPatterns, r->Amean, (float)st.statUpdates);
    out_rowmajor(st.Pmean, nSamples, p->nPatterns, r->Pmean, (float)st.statUpdates);
    out_sd(st.Amean, st.Astd, nGenes, p->nPatterns, st.statUpdates, r->Asd);
    out_sd(st.Pmean, st.Pstd, nSamples, p->nPatterns, st.statUpdates, r->Psd);
    r->nHistory = nHist;
    r->nSnapshotsEquilibration = nSnapEq;
    r->nSnapshotsSampling = nSnapSamp;
    r->seed = p->seed;
    r->totalUpdates = (A.sweep || P.sweep) ? A.swSteps + P.swSteps : totalUpdates; /* sweep: proposals actually made */
    r->totalRunningTime = 0.0;
    r->averageQueueLengthA = A.asynchronous ? A.avgQueueLength : 0.f;
    r->averageQueueLengthP = P.asynchronous ? P.avgQueueLength : 0.f;
    r->meanChiSq = (fixed != 'N') ? 0.f : stats_mean_chisq(&st, &P.model);
    if (p->takePumpSamples)
    {
        float denom = st.pumpUpdates != 0 ? (float)st.pumpUpdates : 1.f;
        out_rowmajor(st.pump, nGenes, p->nPatterns, r->pumpMatrix, denom);
        if (r->meanPatternAssignment)
        {
            /* meanPattern, GapsStatistics.cpp:119-131 on Amean() */
            size_t n = (size_t)nGenes * p->nPatterns;
            float *am = (float*)calloc(n, sizeof(float));
            float *mp = (float*)calloc(n, sizeof(float));
            for (size_t i = 0; i < n; ++i) { am[i] = st.Amean[i] / (float)st.statUpdates; }
            pump_threshold(am, nGenes, p->nPatterns, mp);
            out_rowmajor(mp, nGenes, p->nPatterns, r->meanPatternAssignment, 1.f);
            free(am);
            free(mp);
        }
    }
    if (count) { *count = trace.count; }
done:
    stats_free(&st);
    sampler_free(&A);
    sampler_free(&P);
    free(rs);
    return rc;
}

int cogaps_oracle_run(const float *data, uint32_t nrow, uint32_t ncol, const float *uncertainty,
                      const cgb_params *params, cgb_result *result, const oracle_options *opt)
{
    return cogaps_oracle_run_trace(data, nrow, ncol, uncertainty, params, result, opt, NULL, 0, NULL);
}

int cogaps_oracle_tables(float *erf_, float *erfinv_, float *qgamma_)
{
    randstate_t *rs = (randstate_t*)malloc(sizeof(randstate_t));
    init_tables(rs);
    memcpy(erf_, rs->erf, sizeof(rs->erf));
    memcpy(erfinv_, rs->erfinv, sizeof(rs->erfinv));
    memcpy(qgamma_, rs->qgamma, sizeof(rs->qgamma));
    free(rs);
    return 0;
}

int cogaps_oracle_rng_stream(uint32_t seed, int kind, uint32_t n, uint64_t a, uint64_t b, double lambda,
                             float f0, float f1, float f2, float f3, uint64_t *out)
{
    randstate_t *rs = (randstate_t*)malloc(sizeof(randstate_t));
    randstate_init(rs, seed, NULL);
    int rc = 0;
    if (kind == 0)
    {
        for (uint32_t i = 0; i < n; ++i) { out[i] = xoro_next(&rs->seeder); }
        free(rs);
        return 0;
    }
    rng_t rng;
    rng_init(&rng, rs);
    for (uint32_t i = 0; i < n && rc == 0; ++i)
    {
        float f = 0.f;
        uint32_t bits = 0;
        switch (kind)
        {
            case 1: out[i] = rng_u32(&rng); break;
            case 2: out[i] = rng_u32_range(&rng, (uint32_t)a, (uint32_t)b); break;
            case 3: out[i] = rng_u64_range(&rng, a, b); break;
            case 4: f = rng_uniform(&rng); memcpy(&bits, &f, 4); out[i] = bits; break;
            case 5: out[i] = (uint64_t)(int64_t)rng_poisson(&rng, lambda); break;
            case 6: f = rng_exponential(&rng, f0); memcpy(&bits, &f, 4); out[i] = bits; break;
            case 7:
            {
                int has = rng_trunc_normal(&rng, f0, f1, f2, f3, &f);
                memcpy(&bits, &f, 4);
                out[i] = has ? bits : 0xFFFFFFFFFFFFFFFFull;
                break;
            }
            case 8: f = rng_trunc_gamma_upper(&rng, f0, f1); memcpy(&bits, &f, 4); out[i] = bits; break;
            default: rc = -1;
        }
    }
    free(rs);
    return rc;
}

static int alpha_parameters_probe(int sparse, const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                                  const float *Amat, const float *Pmat, const float *uncertainty,
                                  uint32_t n, const int32_t *variant, const uint32_t *r1,
                                  const uint32_t *c1, const uint32_t *r2, const uint32_t *c2,
                                  const float *ch, float *s_out, float *smu_out, float *ap_out,
                                  const oracle_options *opt)
{
    cgb_params p;
    memset(&p, 0, sizeof(p));
    p.useSparseOptimization = sparse;
    p.nPatterns = k;
    p.alphaA = p.alphaP = 0.01f;
    p.maxGibbsMassA = p.maxGibbsMassP = 100.f;
    model_t A, P;
    model_init(&A, data, nGenes, nSamples, 1, 1, &p, p.alphaA, p.maxGibbsMassA, opt, 1);
    model_init(&P, data, nGenes, nSamples, 0, 0, &p, p.alphaP, p.maxGibbsMassP, opt, 0);
    if (uncertainty)
    {
        model_set_uncertainty(&A, uncertainty, nGenes, nSamples, 1, 1, &p);
        model_set_uncertainty(&P, uncertainty, nGenes, nSamples, 0, 0, &p);
    }
    model_set_matrix(&A, Amat);
    model_set_matrix(&P, Pmat);
    model_sync(&A, &P);
    model_sync(&P, &A);
    model_extra_initialization(&A);
    model_extra_initialization(&P);
    int rc = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        alpha_t a = {0.f, 0.f};
        switch (variant[i])
        {
            case 0: a = model_alpha(&A, r1[i], c1[i]); break;
            case 1: a = model_alpha2(&A, r1[i], c1[i], r2[i], c2[i]); break;
            case 2: a = model_alpha_with_change(&A, r1[i], c1[i], ch[i]); break;
            default: rc = -1;
        }
        s_out[i] = a.s;
        smu_out[i] = a.s_mu;
    }
    if (ap_out) { memcpy(ap_out, A.AP, sizeof(float) * (size_t)nGenes * nSamples); }
    model_free(&A);
    model_free(&P);
    return rc;
}

int cogaps_oracle_alpha_parameters(const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                                   const float *Amat, const float *Pmat, const float *uncertainty,
                                   uint32_t n, const int32_t *variant, const uint32_t *r1,
                                   const uint32_t *c1, const uint32_t *r2, const uint32_t *c2,
                                   const float *ch, float *s_out, float *smu_out, float *ap_out,
                                   const oracle_options *opt)
{
    return alpha_parameters_probe(0, data, nGenes, nSamples, k, Amat, Pmat, uncertainty, n, variant, r1, c1, r2, c2, ch,
                                  s_out, smu_out, ap_out, opt);
}

/* the same probe on SparseNormalModel (SparseNormalModel.cpp:153-292) */
int cogaps_oracle_alpha_parameters_sparse(const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                                          const float *Amat, const float *Pmat,
                                          uint32_t n, const int32_t *variant, const uint32_t *r1,
                                          const uint32_t *c1, const uint32_t *r2, const uint32_t *c2,
                                          const float *ch, float *s_out, float *smu_out,
                                          const oracle_options *opt)
{
    return alpha_parameters_probe(1, data, nGenes, nSamples, k, Amat, Pmat, NULL, n, variant, r1, c1, r2, c2, ch,
                                  s_out, smu_out, NULL, opt);
}

static int chisq_probe(int sparse, const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                       const float *Amat, const float *Pmat, const float *uncertainty, float *out)
{
    cgb_params p;
    memset(&p, 0, sizeof(p));
    p.useSparseOptimization = sparse;
    p.nPatterns = k;
    p.alphaA = p.alphaP = 0.01f;
    p.maxGibbsMassA = p.maxGibbsMassP = 100.f;
    model_t A, P;
    model_init(&A, data, nGenes, nSamples, 1, 1, &p, p.alphaA, p.maxGibbsMassA, NULL, 1);
    model_init(&P, data, nGenes, nSamples, 0, 0, &p, p.alphaP, p.maxGibbsMassP, NULL, 0);
    if (uncertainty)
    {
        model_set_uncertainty(&A, uncertainty, nGenes, nSamples, 1, 1, &p);
        model_set_uncertainty(&P, uncertainty, nGenes, nSamples, 0, 0, &p);
    }
    model_set_matrix(&A, Amat);
    model_set_matrix(&P, Pmat);
    model_sync(&A, &P);
    model_sync(&P, &A);
    model_extra_initialization(&A);
    model_extra_initialization(&P);
    out[0] = model_chisq(&A);
    out[1] = model_chisq(&P);
    out[2] = model_data_sparsity(&P);
    model_free(&A);
    model_free(&P);
    return 0;
}

int cogaps_oracle_chisq(const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                        const float *Amat, const float *Pmat, const float *uncertainty, float *out)
{
    return chisq_probe(0, data, nGenes, nSamples, k, Amat, Pmat, uncertainty, out);
}

/* SparseNormalModel::chiSq (SparseNormalModel.cpp:39-60) */
int cogaps_oracle_chisq_sparse(const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                               const float *Amat, const float *Pmat, float *out)
{
    return chisq_probe(1, data, nGenes, nSamples, k, Amat, Pmat, NULL, out);
}
