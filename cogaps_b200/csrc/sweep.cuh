// sweep.cuh — the row-parallel sweep: one CTA per row of the factor matrix, the whole update() in one launch.
//
// What it is (north star: "a conflict-partitioned commit step — atoms binned by row so non-conflicting updates apply in
// parallel within one sweep", "device-side RNG for proposal draws"): row r of the factor matrix owns the
// contiguous segment [r*k*binLength, (r+1)*k*binLength) of the atomic domain (ProposalQueue.cpp:172-173), the other factor
// is constant during update() (GapsRunner.cpp:201-222) and the scans of different rows share no element of D or AP
// (DenseNormalModel.cpp:162-240).  Given the other factor the rows are independent, so every row runs the reference's four
// proposal types — evaluated by the very device code of the exact path (decide<>, AsynchronousGibbsSampler.h:126-219) — on
// its own segment, sequentially, all rows at once.  The row's D and AP lines are staged ONCE into shared memory (TMA bulk
// copies) and every proposal of the row scans and commits them there: 12*L bytes of HBM traffic per row and update()
// instead of ~21*L per proposal.  It is a different chain from the reference's for the same seed (birth/death balance from
// the atom count frozen at the start of update(), moves bounded by the row's segment, Philox4x32-10 draws); the exact
// path stays the parity path.  oracle/cogaps_oracle.c `sweep_row` restates this file bit for bit.
#ifndef CGB_SWEEP_CUH
#define CGB_SWEEP_CUH

#include "kernels.cuh"

namespace cgb {

struct SweepCounters
{
    unsigned long long steps;     // proposals made (same-bin moves / exchanges and no-ops included, like nSteps of the reference)
    unsigned long long scans1;    // single-column scans (birth, death): 16*L algorithmic bytes each
    unsigned long long scans2;    // two-column same-row scans (move, exchange): 20*L
    unsigned long long scansX;    // two-row scans of the transport between adjacent rows (moves, exchanges): 32*L
    unsigned long long commits;   // AP lines rewritten: +4*L each
    unsigned long long overflow;  // births dropped because the row's atom store was full
    unsigned long long rowsActive; // rows that made at least one proposal (their D / AP lines were staged)
    unsigned long long visited;   // sparse model: common non-zeros the scans visited (8 + 4k algorithmic bytes each)
    long long atomDelta;          // change of the total atom count
    unsigned int maxCount;        // largest per-row atom count after the sweep
    unsigned int pad;
    unsigned long long phase[8];  // COGAPS_SWEEP_PROFILE: SM cycles thread 0 spent per phase of a proposal, summed over rows
};

struct SweepArgs
{
    ModelView mv;
    uint64_t *pos;        // [nRows][cap] atom positions relative to the row's segment, ascending
    float *mass;          // [nRows][cap]
    uint32_t *count;      // [nRows]
    const float *qgamma;  // truncGammaUpper's table (same-bin exchanges)
    SweepCounters *counters;
    const uint32_t *order; // rows in the order they are handed to CTAs (sweep_order_kernel), or NULL: by index
    uint64_t key;         // Philox key: one seeder value per update()
    uint64_t binLength;
    uint64_t binMagic;    // floor(2^64 / binLength): floor(x / binLength) = mulhi(x, binMagic) + {0,1,2}
    double birthRow, deathAtom, moveAtom, exchAtom, perAtom; // proposal weights, see sweepRates() in cogaps_b200.cu
    uint32_t cap, nSteps;
    uint32_t colour;      // transport kernel: pairs (r, r+1) with r = colour, colour + 2, ...
    uint32_t profile;     // debug: thread 0 clocks the phases of every proposal into counters->phase
};

// Philox4x32-10 (Salmon et al., SC'11) block function
__device__ __forceinline__ void philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t x0, uint32_t x1, uint32_t out[4])
{
#pragma unroll
    for (int round = 0; round < 10; ++round)
    {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ x0, n2 = hi0 ^ c3 ^ x1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        x0 += 0x9E3779B9u;
        x1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// The draws of one proposal (oracle: sweep_draws): two blocks countered by (row, step, block, stream), so a proposal's
// random words are a function of (key, row, step) alone — any lane can form them ahead of time.  Kept in shared memory
// together with the two candidate log(uniform()) of the accept test (PreLog, kernels.cuh), which depend on the seed only.
struct SweepDraw
{
    uint32_t typeWord;   // -> uniform that picks the proposal type
    uint32_t pick;       // -> atom pick
    uint64_t posDraw;    // -> birth position / move destination
    uint64_t seed;       // state of the proposal's own PCG stream
    float log0, log1;    // portable_logf of the stream's first and second uniform
};

static const uint32_t kSweepDrawRing = 64; // proposals whose draws are staged at a time (two halves of 32)

__device__ __forceinline__ void sweep_make_draw(uint64_t key, uint32_t row, uint32_t step, uint32_t stream, SweepDraw *out)
{
    uint32_t w[4], v[4];
    philox_block(row, step, 0u, stream, static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32), w);
    philox_block(row, step, 1u, stream, static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32), v);
    out->typeWord = w[0];
    out->pick = w[1];
    out->posDraw = (static_cast<uint64_t>(w[3]) << 32) | w[2];
    const uint64_t seed = (static_cast<uint64_t>(v[1]) << 32) | v[0];
    out->seed = seed;
    Pcg r;
    r.state = seed;
    out->log0 = portable_logf(r.uniform());
    out->log1 = portable_logf(r.uniform());
}

__device__ __forceinline__ float u32_uniform(uint32_t v) { return fmul(__uint2float_rn(v), 2.3283064365386963e-10f); }

// floor(x / binLength) with the precomputed reciprocal (FastDivU64, atomic_domain.h)
__device__ __forceinline__ uint32_t sweep_bin(uint64_t x, uint64_t binLength, uint64_t magic)
{
    uint64_t q = __umul64hi(x, magic);
    uint64_t r = x - q * binLength;
    while (r >= binLength) { ++q; r -= binLength; }
    return static_cast<uint32_t>(q);
}

// what thread 0 publishes for one proposal (double-buffered: thread 0 prepares proposal j+1 while the others commit j)
struct SweepCtl
{
    uint64_t seed;       // PCG state of the proposal's own stream
    uint32_t type;       // 'B','D','M','E', or 0: nothing to evaluate (same-bin move / exchange, no-op)
    uint32_t r1, c1;     // bin of atom1 (r1 = r2 = the CTA's row in the row sweep)
    uint32_t r2, c2;     // bin of the move destination / of atom2
    uint32_t scan;       // 0: birth into a pattern the other factor has no mass in (exponential draw, no scan)
    float m1, m2;
    float d1, d2;        // commit deltas for element (r1,c1) / (r2,c2)
    uint32_t flags;      // bit0: AP[r1] += d1 * other[:,c1]; bit1: then AP[r2] += d2 * other[:,c2]
    float log0, log1;    // PreLog of the proposal's stream
    uint32_t pad;
};

struct SweepSmem
{
    uint64_t bar;
    float warpS[32];
    float warpMu[32];
    SweepCtl ctl[2];
    uint32_t steps;
    uint32_t count;
    uint32_t dirty;
    uint32_t pad;
    float baseS, baseMu;  // sparse model: the Z-table terms of the scan in flight
    uint32_t warpCnt[kSparseGroup * (kSparseThreads / 32)]; // sparse model: compaction counters of sparse_scan_row
    unsigned long long phase[8]; // debug profile (SweepArgs::profile)
};

static const uint32_t kSweepHdrBytes = 768;
static_assert(sizeof(SweepSmem) <= kSweepHdrBytes, "SweepSmem outgrew its slot");
static const uint32_t kSweepDrawBytes = kSweepDrawRing * sizeof(SweepDraw);

// dynamic shared memory of one CTA: [SweepSmem | 768][draw ring][pos: cap u64][mass: cap f32][M row: k f32]
// [canUseGibbs: k i32] then, 128-byte aligned, the staged lines D, AP (, S) of `ld` floats each when the row is kept in
// shared memory
__host__ __device__ inline uint32_t sweepRowOffset(uint32_t cap, uint32_t k)
{
    const uint32_t bytes = kSweepHdrBytes + kSweepDrawBytes + cap * 12u + k * 8u;
    return (bytes + 127u) & ~127u;
}

// One scan of the staged row against one or two factor columns read through L2 (DenseNormalModel.cpp:162-240); same
// element arithmetic and lane order as scan_segment (kernels.cuh).  KEEP: the columns stay in registers for the commit.
// DG: the D (and S) line lies in global memory and is read past L1 (ld.global.cg) like the columns — the little L1 left beside
// the staged rows holds the erf / erfinv tables and the one-lane code's stack, which a row streaming through would evict
template <int T, int NV, bool HAS_S, bool USE_V2, bool WITH_CHANGE, bool DG = false>
__device__ __forceinline__ void sweep_scan(const float *bufD, const float *bufS, const float *bufAP, const float *gV1,
                                           const float *gV2, uint32_t len, float ch, float &accS, float &accMu,
                                           float4 (&keep1)[NV > 0 ? NV : 1], float4 (&keep2)[NV > 0 ? NV : 1])
{
    const uint32_t tid = threadIdx.x;
    const uint32_t nVec = (len + kVec - 1) / kVec;
    auto element = [&](uint32_t j, const float4 &v4, const float4 &w4)
    {
        const float4 d4 = DG ? __ldcg(reinterpret_cast<const float4*>(bufD) + j) : reinterpret_cast<const float4*>(bufD)[j];
        const float4 a4 = reinterpret_cast<const float4*>(bufAP)[j];
        float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (HAS_S) { s4 = DG ? __ldcg(reinterpret_cast<const float4*>(bufS) + j) : reinterpret_cast<const float4*>(bufS)[j]; }
        const float d[4] = {d4.x, d4.y, d4.z, d4.w};
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
        const float su[4] = {s4.x, s4.y, s4.z, s4.w};
        const uint32_t base = j * kVec;
        float ts[4], tmu[4];
#pragma unroll
        for (int c = 0; c < kVec; ++c)
        {
            const float mat = USE_V2 ? fsub(v[c], w[c]) : v[c];
            const bool live = (base + c < len) && (mat != 0.f);
            const float sd = HAS_S ? su[c] : derive_s(d[c]);
            const float ratio = fdiv(live ? mat : 1.f, fmul(sd, sd));
            const float resid = WITH_CHANGE ? fsub(d[c], fadd(a[c], fmul(ch, v[c]))) : fsub(d[c], a[c]);
            ts[c] = live ? fmul(mat, ratio) : 0.f;
            tmu[c] = live ? fmul(ratio, resid) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < kVec; ++c)
        {
            accS = fadd(accS, ts[c]);
            accMu = fadd(accMu, tmu[c]);
        }
    };
    if (NV > 0)
    {
        // all column loads first (they are the only trips to L2), then the arithmetic in index order
#pragma unroll
        for (int i = 0; i < NV; ++i)
        {
            const uint32_t j = tid + static_cast<uint32_t>(i) * T;
            if (j < nVec)
            {
                keep1[i] = __ldcg(reinterpret_cast<const float4*>(gV1) + j);
                if (USE_V2) { keep2[i] = __ldcg(reinterpret_cast<const float4*>(gV2) + j); }
            }
        }
#pragma unroll
        for (int i = 0; i < NV; ++i)
        {
            const uint32_t j = tid + static_cast<uint32_t>(i) * T;
            if (j < nVec) { element(j, keep1[i], USE_V2 ? keep2[i] : make_float4(0.f, 0.f, 0.f, 0.f)); }
        }
    }
    else
    {
#pragma unroll 2
        for (uint32_t j = tid; j < nVec; j += T)
        {
            const float4 v4 = __ldcg(reinterpret_cast<const float4*>(gV1) + j);
            float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (USE_V2) { w4 = __ldcg(reinterpret_cast<const float4*>(gV2) + j); }
            element(j, v4, w4);
        }
    }
}

// lanes -> warp -> CTA: the butterflies of the exact path (cgb_reduction_order with one segment); the total lands in
// lane 0 of warp 0.  Contains one __syncthreads.
template <int T>
__device__ __forceinline__ void sweep_reduce(SweepSmem *hdr, float &accS, float &accMu)
{
    const uint32_t tid = threadIdx.x;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1)
    {
        accS = fadd(accS, __shfl_xor_sync(0xffffffffu, accS, off));
        accMu = fadd(accMu, __shfl_xor_sync(0xffffffffu, accMu, off));
    }
    if ((tid & 31u) == 0u)
    {
        hdr->warpS[tid >> 5] = accS;
        hdr->warpMu[tid >> 5] = accMu;
    }
    __syncthreads();
    if (tid < 32)
    {
        accS = (tid < T / 32) ? hdr->warpS[tid] : 0.f;
        accMu = (tid < T / 32) ? hdr->warpMu[tid] : 0.f;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
        {
            accS = fadd(accS, __shfl_xor_sync(0xffffffffu, accS, off));
            accMu = fadd(accMu, __shfl_xor_sync(0xffffffffu, accMu, off));
        }
    }
}

// AP line += d * column (updateAPMatrix, DenseNormalModel.cpp:243-258); a thread rewrites exactly the elements it scans
template <int T>
__device__ __forceinline__ void sweep_axpy(float *ap, const float *gV, float d, uint32_t len)
{
    const uint32_t nVec = (len + kVec - 1) / kVec;
    for (uint32_t j = threadIdx.x; j < nVec; j += T)
    {
        float4 a = reinterpret_cast<float4*>(ap)[j];
        const float4 v = __ldcg(reinterpret_cast<const float4*>(gV) + j);
        a.x = fadd(a.x, fmul(d, v.x));
        a.y = fadd(a.y, fmul(d, v.y));
        a.z = fadd(a.z, fmul(d, v.z));
        a.w = fadd(a.w, fmul(d, v.w));
        reinterpret_cast<float4*>(ap)[j] = a;
    }
}

// GapsRng::truncGammaUpper (math/Random.cpp:194-200) with the portable exp of gaps_math.h
__device__ __forceinline__ float sweep_trunc_gamma_upper(Pcg &rng, const float *qgamma, float b, float scale)
{
    const float q = fdiv(b, scale);
    const float upper = fsub(1.f, fmul(portable_expf(-q), fadd(1.f, q)));
    const unsigned ndx = f2u(rng.uniform(0.f, fmul(upper, 5000.f)));
    return fmul(qgamma[ndx], scale);
}

// what apply needs to remember about the proposal thread 0 published
struct SweepPick
{
    uint32_t idx, idx2;
    uint64_t newPos;
};

// Thread 0: the next proposal of the row (ProposalQueue.cpp:129-283 restricted to the row's segment) from its staged
// draws.  Proposals that need no evaluation are applied on the spot and leave ctl->type == 0.
__device__ __forceinline__ void sweep_propose(const SweepArgs &a, uint32_t row, const SweepDraw &dr, uint64_t *sPos, float *sMass,
                                              const int *sCan, uint32_t cnt, SweepCtl *ctl, SweepPick *pick,
                                              unsigned long long *overflow)
{
    const uint64_t binLength = a.binLength;
    const uint64_t Lseg = binLength * a.mv.k;
    ctl->type = 0u;
    ctl->scan = 0u;
    ctl->flags = 0u;
    ctl->seed = dr.seed;
    ctl->log0 = dr.log0;
    ctl->log1 = dr.log1;
    ctl->r1 = ctl->r2 = row;
    const double dc = static_cast<double>(cnt);
    const double b = a.birthRow, d = dmul(dc, a.deathAtom), mv = dmul(dc, a.moveAtom);
    const double tot = dadd(b, dmul(dc, a.perAtom));
    const double x = dmul(static_cast<double>(u32_uniform(dr.typeWord)), tot);
    if (cnt == 0u || x < b)
    {
        const uint64_t p = 1ull + __umul64hi(dr.posDraw, Lseg - 1ull);
        uint32_t idx = 0u;
        while (idx < cnt && sPos[idx] < p) { ++idx; }
        if (idx < cnt && sPos[idx] == p) { return; }
        if (cnt == a.cap) { *overflow += 1ull; return; }
        const uint32_t col = sweep_bin(p, binLength, a.binMagic);
        ctl->type = 'B';
        ctl->c1 = ctl->c2 = col;
        ctl->m1 = ctl->m2 = 0.f;
        ctl->scan = sCan[col] != 0 ? 1u : 0u;
        pick->idx = idx;
        pick->newPos = p;
    }
    else if (x < dadd(b, d))
    {
        const uint32_t idx = __umulhi(dr.pick, cnt);
        ctl->type = 'D';
        ctl->c1 = ctl->c2 = sweep_bin(sPos[idx], binLength, a.binMagic);
        ctl->m1 = sMass[idx];
        ctl->m2 = 0.f;
        ctl->scan = 1u;
        pick->idx = idx;
    }
    else if (x < dadd(dadd(b, d), mv))
    {
        const uint32_t idx = __umulhi(dr.pick, cnt);
        const uint64_t lb = idx > 0u ? sPos[idx - 1u] : 0ull;
        const uint64_t rb = idx + 1u < cnt ? sPos[idx + 1u] : Lseg;
        if (rb - lb < 2ull) { return; }
        const uint64_t p = lb + 1ull + __umul64hi(dr.posDraw, rb - lb - 1ull);
        const uint32_t c1 = sweep_bin(sPos[idx], binLength, a.binMagic), c2 = sweep_bin(p, binLength, a.binMagic);
        if (c1 == c2)
        {
            sPos[idx] = p; // "automatically accept moves in same bin" (ProposalQueue.cpp:236-240)
            return;
        }
        ctl->type = 'M';
        ctl->c1 = c1;
        ctl->c2 = c2;
        ctl->m1 = sMass[idx];
        ctl->m2 = 0.f;
        ctl->scan = 1u;
        pick->idx = idx;
        pick->newPos = p;
    }
    else
    {
        const uint32_t idx = __umulhi(dr.pick, cnt);
        if (cnt < 2u) { return; }
        const uint32_t j = idx + 1u < cnt ? idx + 1u : 0u;
        const uint32_t c1 = sweep_bin(sPos[idx], binLength, a.binMagic), c2 = sweep_bin(sPos[j], binLength, a.binMagic);
        const float m1 = sMass[idx], m2 = sMass[j];
        if (c1 == c2)
        {
            // "automatically accept exchanges in same bin" (ProposalQueue.cpp:266-276)
            Pcg rng;
            rng.state = dr.seed;
            const float newMass = sweep_trunc_gamma_upper(rng, a.qgamma, fadd(m1, m2), fdiv(1.f, a.mv.lambda));
            const float delta = (m1 > m2) ? fsub(newMass, m1) : fsub(m2, newMass);
            if (fadd(m1, delta) > kEpsilon && fsub(m2, delta) > kEpsilon)
            {
                sMass[idx] = fadd(m1, delta);
                sMass[j] = fsub(m2, delta);
            }
            return;
        }
        if (sCan[c1] == 0 && sCan[c2] == 0) { return; }
        ctl->type = 'E';
        ctl->c1 = c1;
        ctl->c2 = c2;
        ctl->m1 = m1;
        ctl->m2 = m2;
        ctl->scan = 1u;
        pick->idx = idx;
        pick->idx2 = j;
    }
}

// Thread 0: gibbsMass / accept test of the published proposal (decide<>, the exact path's epilogue); leaves the commit
// deltas in ctl.  M1 / M2 are the current factor elements and are updated in place.
__device__ __forceinline__ bool sweep_decide(const ModelView &mv, SweepCtl *ctl, float s, float mu, float &M1, float &M2,
                                             int can1, int can2, DevOutcome *out)
{
    DevProposal pr;
    pr.rng = ctl->seed;
    pr.r1 = ctl->r1;
    pr.r2 = ctl->r2;
    pr.c1 = ctl->c1;
    pr.c2 = ctl->c2;
    pr.m1 = ctl->m1;
    pr.m2 = ctl->m2;
    pr.type = ctl->type;
    pr.variant = 0u;
    pr.ch = 0.f;
    pr.pad = 0u;
    PreLog pre;
    pre.state0 = pr.rng;
    pre.logFirst = ctl->log0;
    pre.logSecond = ctl->log1;
    Verdict v;
    decide_body<false, true>(mv, mv.erf, mv.erfinv, mv.annealingTemp, pr, 0u, false, s, mu, M1, M2, 0.f, 0.f, can1, can2, pre, &v);
    if (v.dec.flags & 1u) { M1 = v.newM1; }
    if (v.dec.flags & 2u) { M2 = v.newM2; }
    ctl->d1 = v.dec.dOwn1;
    ctl->d2 = v.dec.dOwn2;
    ctl->flags = v.dec.flags & 3u;
    *out = v.out;
    return v.out.accepted != 0u;
}

// STAGE: what a row keeps in shared memory for the length of the update — 1: its D and AP lines (and S), 2: its AP line only
// (the one that changes; D, read-only, comes through L2 like the factor columns, and twice as many rows fit on an SM),
// 0: nothing (rows beyond shared memory).  Same arithmetic, same bits in all three.
__host__ __device__ constexpr int sweepMinBlocks(int T, int STAGE)
{
    return STAGE == 2 ? (T <= 128 ? 8 : (T <= 256 ? 4 : (T <= 512 ? 2 : 1))) : (T <= 128 ? 5 : (T <= 256 ? 4 : 1));
}

template <int T, int NV, bool HAS_S, int STAGE>
__global__ void __launch_bounds__(T, sweepMinBlocks(T, STAGE)) sweep_kernel(const __grid_constant__ SweepArgs a)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const ModelView &mv = a.mv;
    SweepSmem *hdr = reinterpret_cast<SweepSmem*>(smemRaw);
    SweepDraw *draws = reinterpret_cast<SweepDraw*>(smemRaw + kSweepHdrBytes);
    uint64_t *sPos = reinterpret_cast<uint64_t*>(smemRaw + kSweepHdrBytes + kSweepDrawBytes);
    float *sMass = reinterpret_cast<float*>(sPos + a.cap);
    float *sM = sMass + a.cap;
    int *sCan = reinterpret_cast<int*>(sM + mv.k);
    const uint32_t rowPad = mv.ld; // floats per line, multiple of 32
    float *stage = reinterpret_cast<float*>(smemRaw + sweepRowOffset(a.cap, mv.k));
    const uint32_t tid = threadIdx.x;
    const uint32_t row = a.order != nullptr ? a.order[blockIdx.x] : blockIdx.x;
    const uint32_t L = mv.L;
    const size_t rowOff = static_cast<size_t>(row) * mv.ld;

    uint32_t cnt = 0u;
    if (tid == 0)
    {
        cnt = a.count[row];
        double lam = dmul(static_cast<double>(a.nSteps), dadd(a.birthRow, dmul(static_cast<double>(cnt), a.perAtom)));
        if (lam > 1.0e9) { lam = 1.0e9; }
        uint32_t steps = __double2uint_rz(lam);
        const float frac = __double2float_rn(dadd(lam, -static_cast<double>(steps)));
        uint32_t w[4];
        philox_block(row, 0xFFFFFFFFu, 0u, 0u, static_cast<uint32_t>(a.key), static_cast<uint32_t>(a.key >> 32), w);
        if (u32_uniform(w[0]) < frac) { steps += 1u; }
        hdr->steps = steps;
        hdr->count = cnt;
        hdr->dirty = 0u;
        if (STAGE != 0 && steps > 0u)
        {
            mbar_init(&hdr->bar, 1);
            fence_mbar_init();
        }
    }
    __syncthreads();
    const uint32_t steps = hdr->steps;
    if (steps == 0u) { return; }
    const uint32_t cnt0 = hdr->count;

    // ---- stage the row: D and AP lines (TMA bulk), its atoms, its factor elements, the canUseGibbs flags, and the
    //      draws of its first proposals ----
    const float *bufD, *bufS = nullptr;
    float *bufAP;
    if (STAGE == 1)
    {
        bufD = stage;
        bufAP = stage + rowPad;
        if (HAS_S) { bufS = stage + 2u * rowPad; }
        if (tid == 0)
        {
            const uint32_t bytes = ((L + 3u) & ~3u) * 4u;
            mbar_expect_tx(&hdr->bar, bytes * (HAS_S ? 3u : 2u));
            bulk_g2s(stage, mv.D + rowOff, bytes, &hdr->bar);
            bulk_g2s(stage + rowPad, mv.AP + rowOff, bytes, &hdr->bar);
            if (HAS_S) { bulk_g2s(stage + 2u * rowPad, mv.S + rowOff, bytes, &hdr->bar); }
        }
    }
    else if (STAGE == 2)
    {
        bufD = mv.D + rowOff;
        bufAP = stage;
        if (HAS_S) { bufS = mv.S + rowOff; }
        if (tid == 0)
        {
            const uint32_t bytes = ((L + 3u) & ~3u) * 4u;
            mbar_expect_tx(&hdr->bar, bytes);
            bulk_g2s(stage, mv.AP + rowOff, bytes, &hdr->bar);
        }
    }
    else
    {
        bufD = mv.D + rowOff;
        bufAP = mv.AP + rowOff;
        if (HAS_S) { bufS = mv.S + rowOff; }
    }
    for (uint32_t i = tid; i < kSweepDrawRing && i < steps; i += T) { sweep_make_draw(a.key, row, i, 0u, &draws[i]); }
    for (uint32_t i = tid; i < cnt0; i += T)
    {
        sPos[i] = a.pos[static_cast<size_t>(row) * a.cap + i];
        sMass[i] = a.mass[static_cast<size_t>(row) * a.cap + i];
    }
    for (uint32_t c = tid; c < mv.k; c += T)
    {
        sM[c] = mv.M[static_cast<size_t>(c) * mv.ldM + row];
        sCan[c] = mv.otherColNonzero[c];
    }
    __syncthreads();
    if (STAGE != 0) { mbar_wait(&hdr->bar, 0u); }

    unsigned long long nScan1 = 0ull, nScan2 = 0ull, nCommit = 0ull, nOverflow = 0ull;
    SweepPick pick;
    pick.idx = pick.idx2 = 0u;
    pick.newPos = 0ull;
    float4 keep1[NV > 0 ? NV : 1], keep2[NV > 0 ? NV : 1];
    const bool prof = a.profile != 0u && tid == 0;
    long long tPrev = 0;
    if (prof)
    {
        for (int i = 0; i < 8; ++i) { hdr->phase[i] = 0ull; }
        tPrev = clock64();
    }
    // phase p ends here: the cycles since the previous mark are its
    auto mark = [&](int p)
    {
        if (prof)
        {
            const long long now = clock64();
            hdr->phase[p] += static_cast<unsigned long long>(now - tPrev);
            tPrev = now;
        }
    };
    if (tid == 0) { sweep_propose(a, row, draws[0], sPos, sMass, sCan, cnt, &hdr->ctl[0], &pick, &nOverflow); }
    for (uint32_t step = 0; step < steps; ++step)
    {
        SweepCtl *ctl = &hdr->ctl[step & 1u];
        // the draws of proposals step+32 .. step+63 replace the half of the ring that was used up 32 proposals ago
        if ((step & 31u) == 0u && step > 0u)
        {
            const uint32_t base = step + 32u;
            if (tid >= 32u && tid < 64u && base + (tid - 32u) < steps) { sweep_make_draw(a.key, row, base + (tid - 32u), 0u, &draws[(base + (tid - 32u)) % kSweepDrawRing]); }
        }
        __syncthreads(); // proposal `step` is published
        mark(0);
        const uint32_t type = ctl->type;
        if (type == 0u)
        {
            if (tid == 0 && step + 1u < steps) { sweep_propose(a, row, draws[(step + 1u) % kSweepDrawRing], sPos, sMass, sCan, cnt, &hdr->ctl[(step + 1u) & 1u], &pick, &nOverflow); }
            mark(6);
            continue;
        }
        const uint32_t c1 = ctl->c1, c2 = ctl->c2;
        const bool pairType = (type == 'M') || (type == 'E');
        const float *gV1 = mv.otherM + static_cast<size_t>(c1) * mv.ldOther;
        const float *gV2 = mv.otherM + static_cast<size_t>(c2) * mv.ldOther;
        float accS = 0.f, accMu = 0.f;
        if (ctl->scan != 0u)
        {
            if (pairType) { sweep_scan<T, NV, HAS_S, true, false, STAGE == 2>(bufD, bufS, bufAP, gV1, gV2, L, 0.f, accS, accMu, keep1, keep2); }
            else if (type == 'D') { sweep_scan<T, NV, HAS_S, false, true, STAGE == 2>(bufD, bufS, bufAP, gV1, gV2, L, -ctl->m1, accS, accMu, keep1, keep2); }
            else { sweep_scan<T, NV, HAS_S, false, false, STAGE == 2>(bufD, bufS, bufAP, gV1, gV2, L, 0.f, accS, accMu, keep1, keep2); }
        }
        sweep_reduce<T>(hdr, accS, accMu);
        mark(1);
        if (tid == 0)
        {
            // ---- decision and the atom bookkeeping of AsynchronousGibbsSampler::birth/death/move/exchange ----
            DevOutcome out;
            const bool accepted = sweep_decide(mv, ctl, accS, accMu, sM[c1], sM[c2], sCan[c1], sCan[c2], &out);
            if (ctl->scan != 0u) { if (pairType) { ++nScan2; } else { ++nScan1; } }
            if (ctl->flags != 0u) { ++nCommit; hdr->dirty = 1u; }
            if (type == 'B')
            {
                if (accepted)
                {
                    for (uint32_t i = cnt; i > pick.idx; --i) { sPos[i] = sPos[i - 1u]; sMass[i] = sMass[i - 1u]; }
                    sPos[pick.idx] = pick.newPos;
                    sMass[pick.idx] = out.mass1;
                    cnt += 1u;
                }
            }
            else if (type == 'D')
            {
                if (accepted) { sMass[pick.idx] = out.mass1; }
                else
                {
                    for (uint32_t i = pick.idx; i + 1u < cnt; ++i) { sPos[i] = sPos[i + 1u]; sMass[i] = sMass[i + 1u]; }
                    cnt -= 1u;
                }
            }
            else if (type == 'M')
            {
                if (accepted) { sPos[pick.idx] = pick.newPos; }
            }
            else if (accepted)
            {
                sMass[pick.idx] = out.mass1;
                sMass[pick.idx2] = out.mass2;
            }
        }
        mark(2);
        __syncthreads(); // the decision is published
        mark(3);
        // ---- commit in place: AP[row,:] += d1 * other[:,c1] (+ d2 * other[:,c2]), updateAPMatrix
        //      (DenseNormalModel.cpp:243-258); a thread rewrites exactly the elements it scans, so no barrier follows ----
        const uint32_t flags = ctl->flags;
        if (flags != 0u)
        {
            const float d1 = ctl->d1, d2 = ctl->d2;
            if (NV > 0 && ctl->scan != 0u)
            {
                const uint32_t nVec = (L + kVec - 1) / kVec;
#pragma unroll
                for (int i = 0; i < (NV > 0 ? NV : 1); ++i)
                {
                    const uint32_t j = tid + static_cast<uint32_t>(i) * T;
                    if (j < nVec)
                    {
                        float4 ap = reinterpret_cast<float4*>(bufAP)[j];
                        if (flags & 1u)
                        {
                            ap.x = fadd(ap.x, fmul(d1, keep1[i].x));
                            ap.y = fadd(ap.y, fmul(d1, keep1[i].y));
                            ap.z = fadd(ap.z, fmul(d1, keep1[i].z));
                            ap.w = fadd(ap.w, fmul(d1, keep1[i].w));
                        }
                        if (flags & 2u)
                        {
                            ap.x = fadd(ap.x, fmul(d2, keep2[i].x));
                            ap.y = fadd(ap.y, fmul(d2, keep2[i].y));
                            ap.z = fadd(ap.z, fmul(d2, keep2[i].z));
                            ap.w = fadd(ap.w, fmul(d2, keep2[i].w));
                        }
                        reinterpret_cast<float4*>(bufAP)[j] = ap;
                    }
                }
            }
            else
            {
                if (flags & 1u) { sweep_axpy<T>(bufAP, gV1, d1, L); }
                if (flags & 2u) { sweep_axpy<T>(bufAP, gV2, d2, L); }
            }
        }
        mark(4);
        if (tid == 0 && step + 1u < steps) { sweep_propose(a, row, draws[(step + 1u) % kSweepDrawRing], sPos, sMass, sCan, cnt, &hdr->ctl[(step + 1u) & 1u], &pick, &nOverflow); }
        mark(5);
    }
    if (tid == 0) { hdr->count = cnt; }
    __syncthreads();
    if (prof)
    {
        mark(7);
        for (int i = 0; i < 8; ++i) { atomicAdd(&a.counters->phase[i], hdr->phase[i]); }
    }

    // ---- write the row back: its AP line (if any proposal changed it) and its atoms ----
    const uint32_t cntEnd = hdr->count;
    if (STAGE != 0 && hdr->dirty != 0u)
    {
        float *apRow = mv.AP + rowOff;
        const uint32_t nVec = (L + kVec - 1) / kVec;
        for (uint32_t j = tid; j < nVec; j += T) { reinterpret_cast<float4*>(apRow)[j] = reinterpret_cast<const float4*>(bufAP)[j]; }
    }
    for (uint32_t i = tid; i < cntEnd; i += T)
    {
        a.pos[static_cast<size_t>(row) * a.cap + i] = sPos[i];
        a.mass[static_cast<size_t>(row) * a.cap + i] = sMass[i];
    }
    if (tid == 0)
    {
        a.count[row] = cntEnd;
        SweepCounters *c = a.counters;
        atomicAdd(&c->steps, static_cast<unsigned long long>(steps));
        atomicAdd(&c->rowsActive, 1ull);
        if (nScan1) { atomicAdd(&c->scans1, nScan1); }
        if (nScan2) { atomicAdd(&c->scans2, nScan2); }
        if (nCommit) { atomicAdd(&c->commits, nCommit); }
        if (nOverflow) { atomicAdd(&c->overflow, nOverflow); }
        if (cntEnd != cnt0)
        {
            atomicAdd(reinterpret_cast<unsigned long long*>(&c->atomDelta),
                      static_cast<unsigned long long>(static_cast<long long>(cntEnd) - static_cast<long long>(cnt0)));
        }
    }
}


// ------------------------------------------------------------------------------------------------
// The sweep over the SparseNormalModel (SparseNormalModel.cpp:153-292): same row independence — a scan reads the data
// row's non-zeros, the row's own factor row and the other factor — same proposals, same draws, same transport.  There is
// no AP line to stage: what stays in shared memory across the row's proposals is its factor row in both copies (the row
// copy, and the column copy in which values below epsilon are stored as 0, data_structures/HybridMatrix.cpp:25-39);
// the scan itself is the exact path's (sparse_scan_row: compact the common non-zeros, gather the other factor's rows).
// dynamic shared memory: [SweepSmem | 768][draw ring][pos][mass][sRow: ldR f32][sCol: k f32][canUseGibbs: k i32], then
// 128-byte aligned [sIdx][sD][sV1][sV2], kSparseThreads * kSparseGroup entries each
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t sweepSparseScanOffset(uint32_t cap, uint32_t k, uint32_t ldR)
{
    const uint32_t bytes = kSweepHdrBytes + kSweepDrawBytes + cap * 12u + (ldR + 2u * k) * 4u;
    return (bytes + 127u) & ~127u;
}

__global__ void __launch_bounds__(kSparseThreads, 4) sweep_sparse_kernel(const __grid_constant__ SweepArgs a)
{
    constexpr int T = kSparseThreads;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const ModelView &mv = a.mv;
    SweepSmem *hdr = reinterpret_cast<SweepSmem*>(smemRaw);
    SweepDraw *draws = reinterpret_cast<SweepDraw*>(smemRaw + kSweepHdrBytes);
    uint64_t *sPos = reinterpret_cast<uint64_t*>(smemRaw + kSweepHdrBytes + kSweepDrawBytes);
    float *sMass = reinterpret_cast<float*>(sPos + a.cap);
    float *sRow = sMass + a.cap;
    float *sCol = sRow + mv.ldR;
    int *sCan = reinterpret_cast<int*>(sCol + mv.k);
    uint32_t *sIdx = reinterpret_cast<uint32_t*>(smemRaw + sweepSparseScanOffset(a.cap, mv.k, mv.ldR));
    float *sD = reinterpret_cast<float*>(sIdx + kSparseThreads * kSparseGroup);
    float *sV1 = sD + kSparseThreads * kSparseGroup;
    float *sV2 = sV1 + kSparseThreads * kSparseGroup;
    const uint32_t tid = threadIdx.x;
    const uint32_t row = a.order != nullptr ? a.order[blockIdx.x] : blockIdx.x;

    uint32_t cnt = 0u;
    if (tid == 0)
    {
        cnt = a.count[row];
        double lam = dmul(static_cast<double>(a.nSteps), dadd(a.birthRow, dmul(static_cast<double>(cnt), a.perAtom)));
        if (lam > 1.0e9) { lam = 1.0e9; }
        uint32_t steps = __double2uint_rz(lam);
        const float frac = __double2float_rn(dadd(lam, -static_cast<double>(steps)));
        uint32_t w[4];
        philox_block(row, 0xFFFFFFFFu, 0u, 0u, static_cast<uint32_t>(a.key), static_cast<uint32_t>(a.key >> 32), w);
        if (u32_uniform(w[0]) < frac) { steps += 1u; }
        hdr->steps = steps;
        hdr->count = cnt;
    }
    __syncthreads();
    const uint32_t steps = hdr->steps;
    if (steps == 0u) { return; }
    const uint32_t cnt0 = hdr->count;
    for (uint32_t i = tid; i < kSweepDrawRing && i < steps; i += T) { sweep_make_draw(a.key, row, i, 0u, &draws[i]); }
    for (uint32_t i = tid; i < cnt0; i += T)
    {
        sPos[i] = a.pos[static_cast<size_t>(row) * a.cap + i];
        sMass[i] = a.mass[static_cast<size_t>(row) * a.cap + i];
    }
    for (uint32_t i = tid; i < mv.ldR; i += T) { sRow[i] = __ldcg(mv.Mrows + static_cast<size_t>(row) * mv.ldR + i); }
    for (uint32_t c = tid; c < mv.k; c += T)
    {
        sCol[c] = __ldcg(mv.M + static_cast<size_t>(c) * mv.ldM + row);
        sCan[c] = mv.otherColNonzero[c];
    }
    __syncthreads();

    unsigned long long nScan1 = 0ull, nScan2 = 0ull, nOverflow = 0ull, nVisited = 0ull;
    SweepPick pick;
    pick.idx = pick.idx2 = 0u;
    pick.newPos = 0ull;
    if (tid == 0) { sweep_propose(a, row, draws[0], sPos, sMass, sCan, cnt, &hdr->ctl[0], &pick, &nOverflow); }
    for (uint32_t step = 0; step < steps; ++step)
    {
        SweepCtl *ctl = &hdr->ctl[step & 1u];
        if ((step & 31u) == 0u && step > 0u)
        {
            const uint32_t base = step + 32u;
            if (tid >= 32u && tid < 64u && base + (tid - 32u) < steps) { sweep_make_draw(a.key, row, base + (tid - 32u), 0u, &draws[(base + (tid - 32u)) % kSweepDrawRing]); }
        }
        __syncthreads(); // proposal `step` is published
        const uint32_t type = ctl->type;
        if (type == 0u)
        {
            if (tid == 0 && step + 1u < steps) { sweep_propose(a, row, draws[(step + 1u) % kSweepDrawRing], sPos, sMass, sCan, cnt, &hdr->ctl[(step + 1u) & 1u], &pick, &nOverflow); }
            continue;
        }
        const uint32_t c1 = ctl->c1, c2 = ctl->c2;
        const bool pairType = (type == 'M') || (type == 'E');
        const bool withChange = (type == 'D');
        const float ch = -ctl->m1;
        float accS = 0.f, accMu = 0.f;
        uint32_t visited = 0u;
        if (ctl->scan != 0u)
        {
            if (tid == 64)
            {
                float bs, bmu;
                sparse_table_terms(mv, sRow, c1, c1, c2, pairType, withChange, ch, bs, bmu);
                hdr->baseS = bs;
                hdr->baseMu = bmu;
            }
            sparse_scan_row(mv, hdr->warpCnt, sRow, sIdx, sD, sV1, sV2, row, c1, c2, pairType, withChange, ch, accS, accMu, visited);
        }
        sweep_reduce<T>(hdr, accS, accMu);
        if (tid == 0)
        {
            float sTot = 0.f, muTot = 0.f;
            if (ctl->scan != 0u)
            {
                sTot = fmul(pairType ? fsub(hdr->baseS, accS) : fadd(hdr->baseS, accS), mv.beta);
                muTot = fmul(fadd(hdr->baseMu, accMu), mv.beta);
                if (pairType) { ++nScan2; } else { ++nScan1; }
                nVisited += visited;
            }
            DevProposal pr;
            pr.rng = ctl->seed;
            pr.r1 = pr.r2 = row;
            pr.c1 = c1;
            pr.c2 = c2;
            pr.m1 = ctl->m1;
            pr.m2 = ctl->m2;
            pr.type = type;
            pr.variant = 0u;
            pr.ch = 0.f;
            pr.pad = 0u;
            PreLog pre;
            pre.state0 = pr.rng;
            pre.logFirst = ctl->log0;
            pre.logSecond = ctl->log1;
            Verdict v;
            decide_body<true, true>(mv, mv.erf, mv.erfinv, mv.annealingTemp, pr, 0u, false, sTot, muTot, sRow[c1], sRow[c2], sCol[c1], sCol[c2],
                                    sCan[c1], sCan[c2], pre, &v);
            if (v.dec.flags & 1u) { sRow[c1] = v.newM1; sCol[c1] = v.newC1; }
            if (v.dec.flags & 2u) { sRow[c2] = v.newM2; sCol[c2] = v.newC2; }
            const bool accepted = v.out.accepted != 0u;
            if (type == 'B')
            {
                if (accepted)
                {
                    for (uint32_t i = cnt; i > pick.idx; --i) { sPos[i] = sPos[i - 1u]; sMass[i] = sMass[i - 1u]; }
                    sPos[pick.idx] = pick.newPos;
                    sMass[pick.idx] = v.out.mass1;
                    cnt += 1u;
                }
            }
            else if (type == 'D')
            {
                if (accepted) { sMass[pick.idx] = v.out.mass1; }
                else
                {
                    for (uint32_t i = pick.idx; i + 1u < cnt; ++i) { sPos[i] = sPos[i + 1u]; sMass[i] = sMass[i + 1u]; }
                    cnt -= 1u;
                }
            }
            else if (type == 'M')
            {
                if (accepted) { sPos[pick.idx] = pick.newPos; }
            }
            else if (accepted)
            {
                sMass[pick.idx] = v.out.mass1;
                sMass[pick.idx2] = v.out.mass2;
            }
            if (step + 1u < steps) { sweep_propose(a, row, draws[(step + 1u) % kSweepDrawRing], sPos, sMass, sCan, cnt, &hdr->ctl[(step + 1u) & 1u], &pick, &nOverflow); }
        }
        // the factor row and the next proposal become visible at the next iteration's barrier
    }
    if (tid == 0) { hdr->count = cnt; }
    __syncthreads();
    const uint32_t cntEnd = hdr->count;
    for (uint32_t i = tid; i < cntEnd; i += T)
    {
        a.pos[static_cast<size_t>(row) * a.cap + i] = sPos[i];
        a.mass[static_cast<size_t>(row) * a.cap + i] = sMass[i];
    }
    if (tid == 0)
    {
        a.count[row] = cntEnd;
        SweepCounters *c = a.counters;
        atomicAdd(&c->steps, static_cast<unsigned long long>(steps));
        atomicAdd(&c->rowsActive, 1ull);
        if (nScan1) { atomicAdd(&c->scans1, nScan1); }
        if (nScan2) { atomicAdd(&c->scans2, nScan2); }
        if (nVisited) { atomicAdd(&c->visited, nVisited); }
        if (nOverflow) { atomicAdd(&c->overflow, nOverflow); }
        if (cntEnd != cnt0)
        {
            atomicAdd(reinterpret_cast<unsigned long long*>(&c->atomDelta),
                      static_cast<unsigned long long>(static_cast<long long>(cntEnd) - static_cast<long long>(cnt0)));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Transport between adjacent rows (oracle: sweep_pair).  What the row-local sweep cannot do is carry an atom from one row
// to another, which the reference's move does all the time (its bounds are the atom's neighbours in the WHOLE domain,
// ProposalQueue.cpp:213-214).  After the rows have run, the pairs of adjacent rows (r, r+1) — even r on even-numbered
// updates, odd r on odd-numbered ones, so that concurrently handled pairs never share a row and every boundary has its
// turn every other update — give the two atoms at their common boundary, a = the
// last atom of row r and b = the first atom of row r+1, the reference's own proposals on the two-row segment: move a or
// b to a uniform position between its neighbours there (across the boundary that is a two-row move: two scans combined
// as AlphaParameters::operator+ does, AlphaParameters.cpp:11-14), or exchange mass between a and b.  Few proposals per
// pair, no reuse: the lines are read and rewritten where they live (L2 / HBM).
// ------------------------------------------------------------------------------------------------
// SPARSE: the SparseNormalModel's scans (T = kSparseThreads; dynamic shared memory: [sRow: ldR f32] then, 128-byte aligned,
// [sIdx][sD][sV1][sV2] of kSparseThreads * kSparseGroup entries each)
template <int T, bool HAS_S, bool SPARSE>
__global__ void __launch_bounds__(T, (T <= 256 ? 4 : (T <= 512 ? 2 : 1))) sweep_transport_kernel(const __grid_constant__ SweepArgs a)
{
    extern __shared__ __align__(128) unsigned char transportDyn[];
    __shared__ SweepSmem hdrStore;
    __shared__ SweepDraw draws[32];
    SweepSmem *hdr = &hdrStore;
    const ModelView &mv = a.mv;
    const uint32_t tid = threadIdx.x;
    const uint32_t r = a.colour + 2u * blockIdx.x;
    if (r + 1u >= mv.nRows) { return; }
    const uint32_t L = mv.L;
    const uint64_t binLength = a.binLength;
    const uint64_t Lseg = binLength * mv.k;
    uint64_t *posL = a.pos + static_cast<size_t>(r) * a.cap, *posR = posL + a.cap;
    float *massL = a.mass + static_cast<size_t>(r) * a.cap, *massR = massL + a.cap;

    uint32_t cl = 0u, cr = 0u;
    if (tid == 0)
    {
        cl = a.count[r];
        cr = a.count[r + 1u];
        double lam = dmul(static_cast<double>(a.nSteps),
                          dadd(dadd(cl > 0u ? a.moveAtom : 0.0, cr > 0u ? a.moveAtom : 0.0), (cl > 0u && cr > 0u) ? a.exchAtom : 0.0));
        if (lam > 1.0e9) { lam = 1.0e9; }
        uint32_t steps = __double2uint_rz(lam);
        const float frac = __double2float_rn(dadd(lam, -static_cast<double>(steps)));
        uint32_t w[4];
        philox_block(r, 0xFFFFFFFFu, 0u, 1u, static_cast<uint32_t>(a.key), static_cast<uint32_t>(a.key >> 32), w);
        if (u32_uniform(w[0]) < frac) { steps += 1u; }
        hdr->steps = steps;
    }
    __syncthreads();
    const uint32_t steps = hdr->steps;
    if (steps == 0u) { return; }
    unsigned long long nScan2 = 0ull, nScanX = 0ull, nCommit = 0ull, nOverflow = 0ull, nVisited = 0ull;
    for (uint32_t step = 0; step < steps; ++step)
    {
        SweepCtl *ctl = &hdr->ctl[step & 1u]; // double-buffered: thread 0 may run one proposal ahead of a slow reader
        if ((step & 31u) == 0u)
        {
            __syncthreads(); // nobody still reads the previous 32 draws
            if (tid < 32u && step + tid < steps) { sweep_make_draw(a.key, r, step + tid, 1u, &draws[tid]); }
            __syncthreads();
        }
        // ---- thread 0: the proposal ----
        bool moveA = false, cross = false;
        uint64_t to = 0ull;
        if (tid == 0)
        {
            const SweepDraw &dr = draws[step & 31u];
            ctl->type = 0u;
            ctl->scan = 1u;
            ctl->flags = 0u;
            ctl->seed = dr.seed;
            ctl->log0 = dr.log0;
            ctl->log1 = dr.log1;
            const double wa = cl > 0u ? a.moveAtom : 0.0, wb = cr > 0u ? a.moveAtom : 0.0, we = (cl > 0u && cr > 0u) ? a.exchAtom : 0.0;
            const double x = dmul(static_cast<double>(u32_uniform(dr.typeWord)), dadd(dadd(wa, wb), we));
            if (cl == 0u && cr == 0u)
            {
            }
            else if (x < dadd(wa, wb))
            {
                moveA = (cr == 0u) || (cl > 0u && x < wa);
                uint64_t from, lb, rb;
                if (moveA)
                {
                    from = posL[cl - 1u];
                    lb = cl > 1u ? posL[cl - 2u] : 0ull;
                    rb = cr > 0u ? Lseg + posR[0] : 2ull * Lseg;
                }
                else
                {
                    from = Lseg + posR[0];
                    lb = cl > 0u ? posL[cl - 1u] : 0ull;
                    rb = cr > 1u ? Lseg + posR[1] : 2ull * Lseg;
                }
                if (rb - lb >= 2ull)
                {
                    to = lb + 1ull + __umul64hi(dr.posDraw, rb - lb - 1ull);
                    const uint32_t r1 = from < Lseg ? r : r + 1u, r2 = to < Lseg ? r : r + 1u;
                    const uint32_t c1 = sweep_bin(from < Lseg ? from : from - Lseg, binLength, a.binMagic);
                    const uint32_t c2 = sweep_bin(to < Lseg ? to : to - Lseg, binLength, a.binMagic);
                    if (r1 == r2 && c1 == c2)
                    {
                        if (moveA) { posL[cl - 1u] = to; } else { posR[0] = to - Lseg; }
                    }
                    else if (r1 != r2 && (moveA ? cr : cl) == a.cap) { nOverflow += 1ull; }
                    else
                    {
                        cross = r1 != r2;
                        ctl->type = 'M';
                        ctl->r1 = r1; ctl->c1 = c1; ctl->r2 = r2; ctl->c2 = c2;
                        ctl->m1 = moveA ? massL[cl - 1u] : massR[0];
                        ctl->m2 = 0.f;
                    }
                }
            }
            else
            {
                const uint32_t c1 = sweep_bin(posL[cl - 1u], binLength, a.binMagic), c2 = sweep_bin(posR[0], binLength, a.binMagic);
                if (mv.otherColNonzero[c1] != 0 || mv.otherColNonzero[c2] != 0)
                {
                    cross = true;
                    ctl->type = 'E';
                    ctl->r1 = r; ctl->c1 = c1; ctl->r2 = r + 1u; ctl->c2 = c2;
                    ctl->m1 = massL[cl - 1u];
                    ctl->m2 = massR[0];
                }
            }
        }
        __syncthreads(); // the proposal is published
        const uint32_t type = ctl->type;
        if (type == 0u) { continue; }
        const uint32_t r1 = ctl->r1, r2 = ctl->r2, c1 = ctl->c1, c2 = ctl->c2;
        const float *gV1 = mv.otherM + static_cast<size_t>(c1) * mv.ldOther;
        const float *gV2 = mv.otherM + static_cast<size_t>(c2) * mv.ldOther;
        float *ap1 = SPARSE ? nullptr : mv.AP + static_cast<size_t>(r1) * mv.ld, *ap2 = SPARSE ? nullptr : mv.AP + static_cast<size_t>(r2) * mv.ld;
        float4 unused1[1], unused2[1];
        float s = 0.f, mu = 0.f;
        uint32_t visitedHere = 0u;
        if (SPARSE)
        {
            // SparseNormalModel::alphaParameters(r1,c1,r2,c2) (SparseNormalModel.cpp:200-292): one two-column scan of the
            // row, or alphaParameters(r1,c1) + alphaParameters(r2,c2), each with its table terms and beta before the sum
            float *sRow = reinterpret_cast<float*>(transportDyn);
            uint32_t *sIdx = reinterpret_cast<uint32_t*>(transportDyn + ((mv.ldR * 4u + 127u) & ~127u));
            float *sD = reinterpret_cast<float*>(sIdx + kSparseThreads * kSparseGroup);
            float *sV1 = sD + kSparseThreads * kSparseGroup;
            float *sV2 = sV1 + kSparseThreads * kSparseGroup;
            const bool sameRow = (r1 == r2);
            float sPart[2] = {0.f, 0.f}, muPart[2] = {0.f, 0.f};
            for (uint32_t part = 0; part < (sameRow ? 1u : 2u); ++part)
            {
                const uint32_t rowP = part ? r2 : r1, colP = part ? c2 : c1;
                __syncthreads(); // the previous part's factor row and warp totals are no longer read
                for (uint32_t i = tid; i < mv.ldR; i += T) { sRow[i] = __ldcg(mv.Mrows + static_cast<size_t>(rowP) * mv.ldR + i); }
                __syncthreads();
                if (tid == 64)
                {
                    float bs, bmu;
                    sparse_table_terms(mv, sRow, colP, c1, c2, sameRow, false, 0.f, bs, bmu);
                    hdr->baseS = bs;
                    hdr->baseMu = bmu;
                }
                float accS = 0.f, accMu = 0.f;
                uint32_t visited = 0u;
                sparse_scan_row(mv, hdr->warpCnt, sRow, sIdx, sD, sV1, sV2, rowP, colP, c2, sameRow, false, 0.f, accS, accMu, visited);
                sweep_reduce<T>(hdr, accS, accMu);
                sPart[part] = fmul(sameRow ? fsub(hdr->baseS, accS) : fadd(hdr->baseS, accS), mv.beta);
                muPart[part] = fmul(fadd(hdr->baseMu, accMu), mv.beta);
                visitedHere += visited;
            }
            s = sameRow ? sPart[0] : fadd(sPart[0], sPart[1]);
            mu = sameRow ? muPart[0] : fsub(muPart[0], muPart[1]);
        }
        else if (r1 == r2)
        {
            sweep_scan<T, 0, HAS_S, true, false>(mv.D + static_cast<size_t>(r1) * mv.ld, HAS_S ? mv.S + static_cast<size_t>(r1) * mv.ld : nullptr,
                                                 ap1, gV1, gV2, L, 0.f, s, mu, unused1, unused2);
            sweep_reduce<T>(hdr, s, mu);
        }
        else
        {
            // alphaParameters(r1,c1) + alphaParameters(r2,c2): s = s1 + s2, s_mu = s_mu1 - s_mu2 (AlphaParameters.cpp:11-14)
            float s1 = 0.f, mu1 = 0.f, s2 = 0.f, mu2 = 0.f;
            sweep_scan<T, 0, HAS_S, false, false>(mv.D + static_cast<size_t>(r1) * mv.ld, HAS_S ? mv.S + static_cast<size_t>(r1) * mv.ld : nullptr,
                                                  ap1, gV1, gV1, L, 0.f, s1, mu1, unused1, unused2);
            sweep_reduce<T>(hdr, s1, mu1);
            __syncthreads(); // the warp totals are read before the second scan overwrites them
            sweep_scan<T, 0, HAS_S, false, false>(mv.D + static_cast<size_t>(r2) * mv.ld, HAS_S ? mv.S + static_cast<size_t>(r2) * mv.ld : nullptr,
                                                  ap2, gV2, gV2, L, 0.f, s2, mu2, unused1, unused2);
            sweep_reduce<T>(hdr, s2, mu2);
            s = fadd(s1, s2);
            mu = fsub(mu1, mu2);
        }
        if (tid == 0)
        {
            DevOutcome out;
            bool accepted;
            if (SPARSE)
            {
                DevProposal pr;
                pr.rng = ctl->seed;
                pr.r1 = r1; pr.c1 = c1; pr.r2 = r2; pr.c2 = c2;
                pr.m1 = ctl->m1;
                pr.m2 = ctl->m2;
                pr.type = type;
                pr.variant = 0u;
                pr.ch = 0.f;
                pr.pad = 0u;
                PreLog pre;
                pre.state0 = pr.rng;
                pre.logFirst = ctl->log0;
                pre.logSecond = ctl->log1;
                Verdict v;
                decide_body<true, true>(mv, mv.erf, mv.erfinv, mv.annealingTemp, pr, 0u, false, s, mu,
                                        ld_cg_f32(mv.Mrows + static_cast<size_t>(r1) * mv.ldR + c1), ld_cg_f32(mv.Mrows + static_cast<size_t>(r2) * mv.ldR + c2),
                                        ld_cg_f32(mv.M + static_cast<size_t>(c1) * mv.ldM + r1), ld_cg_f32(mv.M + static_cast<size_t>(c2) * mv.ldM + r2),
                                        mv.otherColNonzero[c1], mv.otherColNonzero[c2], pre, &v);
                ctl->flags = 0u; // no AP line to rewrite
                out = v.out;
                accepted = v.out.accepted != 0u;
                nVisited += visitedHere;
            }
            else
            {
                float M1 = ld_cg_f32(mv.M + static_cast<size_t>(c1) * mv.ldM + r1), M2 = ld_cg_f32(mv.M + static_cast<size_t>(c2) * mv.ldM + r2);
                accepted = sweep_decide(mv, ctl, s, mu, M1, M2, mv.otherColNonzero[c1], mv.otherColNonzero[c2], &out);
            }
            if (cross) { ++nScanX; } else { ++nScan2; }
            if (ctl->flags & 1u) { ++nCommit; }
            if (cross && (ctl->flags & 2u)) { ++nCommit; }
            if (accepted && type == 'M')
            {
                if (!cross)
                {
                    if (moveA) { posL[cl - 1u] = to; } else { posR[0] = to - Lseg; }
                }
                else if (moveA)
                {
                    // the new first atom of row r+1
                    for (uint32_t i = cr; i > 0u; --i) { posR[i] = posR[i - 1u]; massR[i] = massR[i - 1u]; }
                    posR[0] = to - Lseg;
                    massR[0] = ctl->m1;
                    cr += 1u;
                    cl -= 1u;
                }
                else
                {
                    // the new last atom of row r
                    posL[cl] = to;
                    massL[cl] = ctl->m1;
                    cl += 1u;
                    for (uint32_t i = 0u; i + 1u < cr; ++i) { posR[i] = posR[i + 1u]; massR[i] = massR[i + 1u]; }
                    cr -= 1u;
                }
            }
            else if (accepted && type == 'E')
            {
                massL[cl - 1u] = out.mass1;
                massR[0] = out.mass2;
            }
        }
        __syncthreads(); // the decision is published
        const uint32_t flags = ctl->flags;
        if (flags & 1u) { sweep_axpy<T>(ap1, gV1, ctl->d1, L); }
        if (flags & 2u) { sweep_axpy<T>(ap2, gV2, ctl->d2, L); } // same line twice when r1 == r2: same thread, same elements
        __syncthreads(); // the lines are rewritten before the next proposal of this pair scans them
    }
    if (tid == 0)
    {
        a.count[r] = cl;
        a.count[r + 1u] = cr;
        SweepCounters *c = a.counters;
        atomicAdd(&c->steps, static_cast<unsigned long long>(steps));
        if (nScan2) { atomicAdd(&c->scans2, nScan2); }
        if (nScanX) { atomicAdd(&c->scansX, nScanX); }
        if (nVisited) { atomicAdd(&c->visited, nVisited); }
        if (nCommit) { atomicAdd(&c->commits, nCommit); }
        if (nOverflow) { atomicAdd(&c->overflow, nOverflow); }
    }
}

// Rows by decreasing atom count.  A row's chain is sequential and its length is proportional to its atom count, so a long
// row that starts late is what the whole launch ends up waiting for; CTAs are handed out in blockIdx order, so the longest
// chains go first.  Which CTA runs a row does not enter its result (the draws are countered by row, not by CTA).
// One CTA: histogram of min(count, bins - 1), offsets from the top bin down, scatter.
static const uint32_t kSweepOrderBins = 1024;

__global__ void __launch_bounds__(1024) sweep_order_kernel(const uint32_t *count, uint32_t nRows, uint32_t *order)
{
    __shared__ uint32_t hist[kSweepOrderBins];
    __shared__ uint32_t warpTot[32];
    const uint32_t tid = threadIdx.x;
    hist[tid] = 0u;
    __syncthreads();
    for (uint32_t r = tid; r < nRows; r += 1024u) { atomicAdd(&hist[min(count[r], kSweepOrderBins - 1u)], 1u); }
    __syncthreads();
    // thread t owns bin (bins - 1 - t): its rows start after those of all higher bins
    const uint32_t mine = hist[kSweepOrderBins - 1u - tid];
    uint32_t incl = mine;
    for (int off = 1; off < 32; off <<= 1)
    {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, off);
        if ((tid & 31u) >= static_cast<uint32_t>(off)) { incl += up; }
    }
    if ((tid & 31u) == 31u) { warpTot[tid >> 5] = incl; }
    __syncthreads();
    if (tid < 32u)
    {
        uint32_t w = warpTot[tid];
        for (int off = 1; off < 32; off <<= 1)
        {
            const uint32_t up = __shfl_up_sync(0xffffffffu, w, off);
            if (tid >= static_cast<uint32_t>(off)) { w += up; }
        }
        warpTot[tid] = w;
    }
    __syncthreads();
    const uint32_t before = (tid >= 32u ? warpTot[(tid >> 5) - 1u] : 0u) + incl - mine;
    __syncthreads();
    hist[kSweepOrderBins - 1u - tid] = before;
    __syncthreads();
    for (uint32_t r = tid; r < nRows; r += 1024u) { order[atomicAdd(&hist[min(count[r], kSweepOrderBins - 1u)], 1u)] = r; }
}

// largest per-row atom count (rows the sweep did not visit keep theirs): the host sizes the store from it
__global__ void sweep_max_count_kernel(const uint32_t *count, uint32_t nRows, unsigned int *out)
{
    unsigned int m = 0u;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nRows; r += gridDim.x * blockDim.x) { m = max(m, count[r]); }
    for (int off = 16; off >= 1; off >>= 1) { m = max(m, __shfl_xor_sync(0xffffffffu, m, off)); }
    if ((threadIdx.x & 31u) == 0u && m > 0u) { atomicMax(out, m); }
}

// re-lay the per-row atom store for a larger capacity
__global__ void sweep_regrow_kernel(const uint64_t *posIn, const float *massIn, const uint32_t *count, uint32_t capIn,
                                    uint64_t *posOut, float *massOut, uint32_t capOut, uint32_t nRows)
{
    const uint32_t row = blockIdx.x;
    if (row >= nRows) { return; }
    const uint32_t n = count[row];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    {
        posOut[static_cast<size_t>(row) * capOut + i] = posIn[static_cast<size_t>(row) * capIn + i];
        massOut[static_cast<size_t>(row) * capOut + i] = massIn[static_cast<size_t>(row) * capIn + i];
    }
}

} // namespace cgb

#endif // CGB_SWEEP_CUH
