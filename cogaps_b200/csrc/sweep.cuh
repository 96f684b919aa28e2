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
    unsigned long long commits;   // proposals that rewrote the row's AP line: +4*L
    unsigned long long overflow;  // births dropped because the row's atom store was full
    unsigned long long rowsActive; // rows that made at least one proposal (their D / AP lines were staged)
    long long atomDelta;          // change of the total atom count
    unsigned int maxCount;        // largest per-row atom count after the sweep
    unsigned int pad;
};

struct SweepArgs
{
    ModelView mv;
    uint64_t *pos;        // [nRows][cap] atom positions relative to the row's segment, ascending
    float *mass;          // [nRows][cap]
    uint32_t *count;      // [nRows]
    const float *qgamma;  // truncGammaUpper's table (same-bin exchanges)
    SweepCounters *counters;
    uint64_t key;         // Philox key: one seeder value per update()
    uint64_t binLength;
    double birthRow, deathAtom, moveAtom, exchAtom, perAtom; // proposal weights, see sweepRates() in cogaps_b200.cu
    uint32_t cap, nSteps;
};

// Philox4x32-10 (Salmon et al., SC'11), key = (key lo, key hi), counter = (row, block, 0, 0); words are handed out in
// order, blocks in order — the stream of a row is a function of (key, row) only
struct Philox
{
    uint32_t k0, k1, row, blk;
    uint32_t b0, b1, b2, b3;
    uint32_t have;
    __device__ __forceinline__ void init(uint64_t key, uint32_t r)
    {
        k0 = static_cast<uint32_t>(key);
        k1 = static_cast<uint32_t>(key >> 32);
        row = r;
        blk = 0u;
        have = 0u;
        b0 = b1 = b2 = b3 = 0u;
    }
    __device__ __forceinline__ void refill()
    {
        uint32_t c0 = row, c1 = blk, c2 = 0u, c3 = 0u, x0 = k0, x1 = k1;
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
        b0 = c0; b1 = c1; b2 = c2; b3 = c3;
        blk += 1u;
        have = 4u;
    }
    __device__ __forceinline__ uint32_t u32()
    {
        if (have == 0u) { refill(); }
        const uint32_t i = 4u - have;
        have -= 1u;
        return i == 0u ? b0 : (i == 1u ? b1 : (i == 2u ? b2 : b3));
    }
    __device__ __forceinline__ uint64_t u64()
    {
        const uint64_t lo = u32();
        const uint64_t hi = u32();
        return (hi << 32) | lo;
    }
    __device__ __forceinline__ float uniform() { return fmul(__uint2float_rn(u32()), 2.3283064365386963e-10f); }
};

// what thread 0 publishes for one proposal (double-buffered: thread 0 prepares proposal j+1 while the others commit j)
struct SweepCtl
{
    uint64_t seed;       // PCG state of the proposal's own stream
    uint32_t type;       // 'B','D','M','E', or 0: nothing to evaluate (same-bin move / exchange, no-op)
    uint32_t c1, c2;
    uint32_t scan;       // 0: birth into a pattern the other factor has no mass in (exponential draw, no scan)
    float m1, m2;
    float d1, d2;        // commit deltas for column c1 / c2
    uint32_t flags;      // bit0: AP += d1 * other[:,c1]; bit1: then AP += d2 * other[:,c2]
    uint32_t pad;
};

struct SweepSmem
{
    uint64_t bar;
    float warpS[32];
    float warpMu[32];
    SweepCtl ctl[2];
    float preLog[2];
    uint32_t steps;
    uint32_t count;
    uint32_t dirty;
    uint32_t pad;
};

static const uint32_t kSweepHdrBytes = 512;
static_assert(sizeof(SweepSmem) <= kSweepHdrBytes, "SweepSmem outgrew its slot");

// dynamic shared memory of one CTA: [SweepSmem | 512][pos: cap u64][mass: cap f32][M row: k f32][canUseGibbs: k i32] then,
// 128-byte aligned, the staged lines D, AP (, S) of rowPad floats each when the row is kept in shared memory
__host__ __device__ inline uint32_t sweepRowOffset(uint32_t cap, uint32_t k)
{
    const uint32_t bytes = kSweepHdrBytes + cap * 12u + k * 8u;
    return (bytes + 127u) & ~127u;
}

template <int T, bool HAS_S, bool USE_V2, bool WITH_CHANGE>
__device__ __forceinline__ void sweep_scan(const float *bufD, const float *bufS, const float *bufAP, const float *gV1,
                                           const float *gV2, uint32_t len, float ch, float &accS, float &accMu)
{
    const uint32_t tid = threadIdx.x;
    const uint32_t nVec = (len + kVec - 1) / kVec;
#pragma unroll 2
    for (uint32_t j = tid; j < nVec; j += T)
    {
        const float4 v4 = __ldcg(reinterpret_cast<const float4*>(gV1) + j);
        float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (USE_V2) { w4 = __ldcg(reinterpret_cast<const float4*>(gV2) + j); }
        const float4 d4 = reinterpret_cast<const float4*>(bufD)[j];
        const float4 a4 = reinterpret_cast<const float4*>(bufAP)[j];
        float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (HAS_S) { s4 = reinterpret_cast<const float4*>(bufS)[j]; }
        const float d[4] = {d4.x, d4.y, d4.z, d4.w};
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
        const float su[4] = {s4.x, s4.y, s4.z, s4.w};
        const uint32_t base = j * kVec;
        float ts[4], tmu[4];
        // same element arithmetic as scan_segment (kernels.cuh): DenseNormalModel.cpp:162-240
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

// Thread 0: draw the next proposal of the row (ProposalQueue.cpp:129-283 restricted to the row's segment) and publish
// what the CTA has to evaluate.  Proposals that need no evaluation are applied on the spot.  idx / idx2 / newPos keep
// what apply_outcome needs.
struct SweepPick
{
    uint32_t idx, idx2;
    uint64_t newPos;
};

__device__ __noinline__ void sweep_propose(const SweepArgs &a, Philox &g, uint64_t *sPos, float *sMass, const int *sCan,
                                           uint32_t &cnt, SweepCtl *ctl, SweepPick *pick, unsigned long long *overflow)
{
    const uint64_t binLength = a.binLength;
    const uint64_t Lseg = binLength * a.mv.k;
    ctl->type = 0u;
    ctl->scan = 0u;
    ctl->flags = 0u;
    const double dc = static_cast<double>(cnt);
    const double b = a.birthRow, d = dmul(dc, a.deathAtom), mv = dmul(dc, a.moveAtom);
    const double tot = dadd(b, dmul(dc, a.perAtom));
    const double x = dmul(static_cast<double>(g.uniform()), tot);
    if (cnt == 0u || x < b)
    {
        const uint64_t p = 1ull + __umul64hi(g.u64(), Lseg - 1ull);
        ctl->seed = g.u64();
        uint32_t idx = 0u;
        while (idx < cnt && sPos[idx] < p) { ++idx; }
        if (idx < cnt && sPos[idx] == p) { return; }
        if (cnt == a.cap) { *overflow += 1ull; return; }
        const uint32_t col = static_cast<uint32_t>(p / binLength);
        ctl->type = 'B';
        ctl->c1 = ctl->c2 = col;
        ctl->m1 = ctl->m2 = 0.f;
        ctl->scan = sCan[col] != 0 ? 1u : 0u;
        pick->idx = idx;
        pick->newPos = p;
    }
    else if (x < dadd(b, d))
    {
        const uint32_t idx = __umulhi(g.u32(), cnt);
        ctl->seed = g.u64();
        ctl->type = 'D';
        ctl->c1 = ctl->c2 = static_cast<uint32_t>(sPos[idx] / binLength);
        ctl->m1 = sMass[idx];
        ctl->m2 = 0.f;
        ctl->scan = 1u;
        pick->idx = idx;
    }
    else if (x < dadd(dadd(b, d), mv))
    {
        const uint32_t idx = __umulhi(g.u32(), cnt);
        const uint64_t draw = g.u64();
        ctl->seed = g.u64();
        const uint64_t lb = idx > 0u ? sPos[idx - 1u] : 0ull;
        const uint64_t rb = idx + 1u < cnt ? sPos[idx + 1u] : Lseg;
        if (rb - lb < 2ull) { return; }
        const uint64_t p = lb + 1ull + __umul64hi(draw, rb - lb - 1ull);
        const uint32_t c1 = static_cast<uint32_t>(sPos[idx] / binLength), c2 = static_cast<uint32_t>(p / binLength);
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
        const uint32_t idx = __umulhi(g.u32(), cnt);
        ctl->seed = g.u64();
        if (cnt < 2u) { return; }
        const uint32_t j = idx + 1u < cnt ? idx + 1u : 0u;
        const uint32_t c1 = static_cast<uint32_t>(sPos[idx] / binLength), c2 = static_cast<uint32_t>(sPos[j] / binLength);
        const float m1 = sMass[idx], m2 = sMass[j];
        if (c1 == c2)
        {
            // "automatically accept exchanges in same bin" (ProposalQueue.cpp:266-276)
            Pcg rng;
            rng.state = ctl->seed;
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

template <int T, bool HAS_S, bool ROW_SMEM>
__global__ void __launch_bounds__(T) sweep_kernel(const __grid_constant__ SweepArgs a)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const ModelView &mv = a.mv;
    SweepSmem *hdr = reinterpret_cast<SweepSmem*>(smemRaw);
    uint64_t *sPos = reinterpret_cast<uint64_t*>(smemRaw + kSweepHdrBytes);
    float *sMass = reinterpret_cast<float*>(sPos + a.cap);
    float *sM = sMass + a.cap;
    int *sCan = reinterpret_cast<int*>(sM + mv.k);
    const uint32_t rowPad = mv.ld; // floats per line, multiple of 32
    float *stage = reinterpret_cast<float*>(smemRaw + sweepRowOffset(a.cap, mv.k));
    const uint32_t tid = threadIdx.x;
    const uint32_t row = blockIdx.x;
    const uint32_t L = mv.L;
    const size_t rowOff = static_cast<size_t>(row) * mv.ld;

    Philox g;
    uint32_t cnt = 0u;
    if (tid == 0)
    {
        cnt = a.count[row];
        g.init(a.key, row);
        double lam = dmul(static_cast<double>(a.nSteps), dadd(a.birthRow, dmul(static_cast<double>(cnt), a.perAtom)));
        if (lam > 1.0e9) { lam = 1.0e9; }
        uint32_t steps = __double2uint_rz(lam);
        const float frac = __double2float_rn(dadd(lam, -static_cast<double>(steps)));
        if (g.uniform() < frac) { steps += 1u; }
        hdr->steps = steps;
        hdr->count = cnt;
        hdr->dirty = 0u;
        if (ROW_SMEM && steps > 0u)
        {
            mbar_init(&hdr->bar, 1);
            fence_mbar_init();
        }
    }
    __syncthreads();
    const uint32_t steps = hdr->steps;
    if (steps == 0u) { return; }
    const uint32_t cnt0 = hdr->count;

    // ---- stage the row: D and AP lines (TMA bulk), its atoms, its factor elements, the canUseGibbs flags ----
    const float *bufD, *bufS = nullptr;
    float *bufAP;
    if (ROW_SMEM)
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
    else
    {
        bufD = mv.D + rowOff;
        bufAP = mv.AP + rowOff;
        if (HAS_S) { bufS = mv.S + rowOff; }
    }
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
    if (ROW_SMEM) { mbar_wait(&hdr->bar, 0u); }

    unsigned long long nScan1 = 0ull, nScan2 = 0ull, nCommit = 0ull, nOverflow = 0ull;
    SweepPick pick;
    pick.idx = pick.idx2 = 0u;
    pick.newPos = 0ull;
    if (tid == 0) { sweep_propose(a, g, sPos, sMass, sCan, cnt, &hdr->ctl[0], &pick, &nOverflow); }
    for (uint32_t step = 0; step < steps; ++step)
    {
        SweepCtl *ctl = &hdr->ctl[step & 1u];
        __syncthreads(); // proposal `step` is published
        const uint32_t type = ctl->type;
        if (type == 0u)
        {
            if (tid == 0 && step + 1u < steps) { sweep_propose(a, g, sPos, sMass, sCan, cnt, &hdr->ctl[(step + 1u) & 1u], &pick, &nOverflow); }
            continue;
        }
        const uint32_t c1 = ctl->c1, c2 = ctl->c2;
        const bool pairType = (type == 'M') || (type == 'E');
        const float *gV1 = mv.otherM + static_cast<size_t>(c1) * mv.ldOther;
        const float *gV2 = mv.otherM + static_cast<size_t>(c2) * mv.ldOther;
        if (tid == 32 && type != 'E')
        {
            // both candidate log(uniform()) of the accept test, off thread 0's serial tail (PreLog, kernels.cuh)
            Pcg r;
            r.state = ctl->seed;
            hdr->preLog[0] = portable_logf(r.uniform());
            hdr->preLog[1] = portable_logf(r.uniform());
        }
        float accS = 0.f, accMu = 0.f;
        if (ctl->scan != 0u)
        {
            if (pairType) { sweep_scan<T, HAS_S, true, false>(bufD, bufS, bufAP, gV1, gV2, L, 0.f, accS, accMu); }
            else if (type == 'D') { sweep_scan<T, HAS_S, false, true>(bufD, bufS, bufAP, gV1, gV2, L, -ctl->m1, accS, accMu); }
            else { sweep_scan<T, HAS_S, false, false>(bufD, bufS, bufAP, gV1, gV2, L, 0.f, accS, accMu); }
        }
        // lanes -> warp -> CTA: the butterflies of the exact path (cgb_reduction_order with one segment)
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
            float sS = (tid < T / 32) ? hdr->warpS[tid] : 0.f;
            float sMu = (tid < T / 32) ? hdr->warpMu[tid] : 0.f;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1)
            {
                sS = fadd(sS, __shfl_xor_sync(0xffffffffu, sS, off));
                sMu = fadd(sMu, __shfl_xor_sync(0xffffffffu, sMu, off));
            }
            if (tid == 0)
            {
                // ---- decision and the atom bookkeeping of AsynchronousGibbsSampler::birth/death/move/exchange ----
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
                pre.logFirst = hdr->preLog[0];
                pre.logSecond = hdr->preLog[1];
                Verdict v;
                decide<false>(mv, mv.erf, mv.erfinv, mv.annealingTemp, pr, 0u, false, sS, sMu, sM[c1], sM[c2], 0.f, 0.f,
                              sCan[c1], sCan[c2], pre, &v);
                if (v.dec.flags & 1u) { sM[c1] = v.newM1; }
                if (v.dec.flags & 2u) { sM[c2] = v.newM2; }
                ctl->d1 = v.dec.dOwn1;
                ctl->d2 = v.dec.dOwn2;
                ctl->flags = v.dec.flags & 3u;
                if (ctl->scan != 0u) { if (pairType) { ++nScan2; } else { ++nScan1; } }
                if (ctl->flags != 0u) { ++nCommit; hdr->dirty = 1u; }
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
            }
        }
        __syncthreads(); // the decision is published
        // ---- commit: AP[row,:] += d1 * other[:,c1] (+ d2 * other[:,c2]) where the row lives (updateAPMatrix,
        //      DenseNormalModel.cpp:243-258); a thread rewrites exactly the elements it scans, so no barrier follows ----
        const uint32_t flags = ctl->flags;
        if (flags != 0u)
        {
            const float d1 = ctl->d1, d2 = ctl->d2;
            const uint32_t nVec = (L + kVec - 1) / kVec;
            for (uint32_t j = tid; j < nVec; j += T)
            {
                float4 ap = reinterpret_cast<float4*>(bufAP)[j];
                if (flags & 1u)
                {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(gV1) + j);
                    ap.x = fadd(ap.x, fmul(d1, v.x));
                    ap.y = fadd(ap.y, fmul(d1, v.y));
                    ap.z = fadd(ap.z, fmul(d1, v.z));
                    ap.w = fadd(ap.w, fmul(d1, v.w));
                }
                if (flags & 2u)
                {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(gV2) + j);
                    ap.x = fadd(ap.x, fmul(d2, v.x));
                    ap.y = fadd(ap.y, fmul(d2, v.y));
                    ap.z = fadd(ap.z, fmul(d2, v.z));
                    ap.w = fadd(ap.w, fmul(d2, v.w));
                }
                reinterpret_cast<float4*>(bufAP)[j] = ap;
            }
        }
        if (tid == 0 && step + 1u < steps) { sweep_propose(a, g, sPos, sMass, sCan, cnt, &hdr->ctl[(step + 1u) & 1u], &pick, &nOverflow); }
    }
    if (tid == 0) { hdr->count = cnt; }
    __syncthreads();

    // ---- write the row back: its AP line (if any proposal changed it) and its atoms ----
    const uint32_t cntEnd = hdr->count;
    if (ROW_SMEM && hdr->dirty != 0u)
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
