// sampler.h — host side of one Gibbs sampler (one factor matrix): the atomic domain, the sequential
// proposal generator, and the handle that owns the device buffers.
//
// Reference restated: atomic/ProposalQueue.cpp:19-283 (generator, conflict rules, seed roll-back),
// gibbs_sampler/AsynchronousGibbsSampler.h:63-122 (ctor, update loop, queue-length diagnostics),
// gibbs_sampler/DenseNormalModel.h:66-95 (ctor: data orientation, default uncertainty, lambda).
#ifndef CGB_SAMPLER_H
#define CGB_SAMPLER_H

#include "atomic_domain.h"
#include "checkpoint.h"
#include "device_types.h"
#include "host_rng.h"

#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include <cstring>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace cgb {

// AtomicProposal (atomic/ProposalQueue.h:15-28) as the host keeps it
struct HostProposal
{
    HostRng rng;
    uint64_t pos;
    uint32_t atom1, atom2;
    uint32_t r1, c1, r2, c2;
    char type;
};

// static_cast<uint64_t>(mDomainLength) as the reference's default build evaluates it
// (ProposalQueue.cpp:214).  The domain length in f64 is usually exactly 2^64, which is out of range
// for the cast; the SSE2 code path (cvttsd2si + fix-up) yields 0.  Spelled out here so the behaviour
// does not depend on our own compiler.
inline uint64_t referenceDoubleToU64(double x)
{
    if (x >= 18446744073709551616.0) { return 0; }
    return static_cast<uint64_t>(x);
}

// SingleThreadedGibbsSampler's own state (gibbs_sampler/SingleThreadedGibbsSampler.h:52-81): ONE rng stream
// shared by proposal generation and evaluation, no queue
struct SequentialState
{
    HostRng rng;
    uint64_t binLength, numPatterns;
    FastDivU64 binDiv, colDiv;
    double numBins, domainLength, alpha;
    void init(uint64_t nElements, uint64_t nPatterns, cgb_randstate *rs, float a)
    {
        rng = HostRng(rs->seeder);
        numBins = static_cast<double>(nElements);
        binLength = 0xFFFFFFFFFFFFFFFFull / nElements;
        numPatterns = nPatterns;
        binDiv.init(binLength);
        colDiv.init(nPatterns);
        domainLength = static_cast<double>(binLength * nElements);
        alpha = static_cast<double>(a);
    }
    void binOf(uint64_t pos, uint32_t &row, uint32_t &col) const
    {
        const uint64_t bin = binDiv.div(pos);
        const uint64_t r = colDiv.div(bin);
        row = static_cast<uint32_t>(r);
        col = static_cast<uint32_t>(bin - r * numPatterns);
    }
};

// ProposalQueue (atomic/ProposalQueue.h:30-72)
class ProposalQueue
{
public:
    void init(uint64_t nElements, uint64_t nPatterns, cgb_randstate *rs, float alpha, float lambda);
    // called for every proposal the moment it is queued, so the device can start on it while the rest of
    // the batch is still being generated
    typedef void (*SinkFn)(void *ctx, const HostProposal &prop, size_t index);
    void populate(AtomicDomain &domain, unsigned limit, SinkFn sink = nullptr, void *sinkCtx = nullptr); // ProposalQueue.cpp:53-76
    void clear();                                           // :78-85
    unsigned nProcessed() const { return mNumProcessed; }
    std::vector<HostProposal> &entries() { return mQueue; }
    void acceptDeath() { --mMaxAtoms; }                     // :97-119
    void rejectDeath() { ++mMinAtoms; }
    void acceptBirth() { ++mMinAtoms; }
    void rejectBirth() { --mMaxAtoms; }
    uint64_t minAtoms() const { return mMinAtoms; }
    uint64_t maxAtoms() const { return mMaxAtoms; }
    // the atomic domain was replaced wholesale (set_atoms, a checkpoint, the sweep store): the window collapses onto its size
    void setAtomCount(uint64_t n) { mMinAtoms = mMaxAtoms = n; }
    // the members operator<< archives (ProposalQueue.cpp:285-299); restore() refuses a state whose bin geometry
    // is not this queue's (the file was made from a matrix of another shape)
    void save(QueueState &out) const;
    bool restore(const QueueState &in);

private:
    float deathProb(double nAtoms) const;                   // :123-127
    bool makeProposal(AtomicDomain &domain);                // :129-160
    bool birth(AtomicDomain &domain);                       // :162-187
    bool death(AtomicDomain &domain);                       // :189-207
    bool move(AtomicDomain &domain);                        // :209-248
    bool exchange(AtomicDomain &domain);                    // :250-283
    bool rowUsed(uint32_t r) const { return mUsedRows[r] == mEpoch; }
    void useRow(uint32_t r) { mUsedRows[r] = mEpoch; }
    bool moveOverlap(uint64_t pos) const;
    // (row, col) of the matrix element a position falls in: (pos / binLength) / nPatterns, % nPatterns
    void binOf(uint64_t pos, uint32_t &row, uint32_t &col) const
    {
        const uint64_t bin = mBinDiv.div(pos);
        const uint64_t r = mColDiv.div(bin);
        row = static_cast<uint32_t>(r);
        col = static_cast<uint32_t>(bin - r * mNumCols);
    }

    std::vector<HostProposal> mQueue;
    std::vector<uint32_t> mUsedRows;                        // FixedHashSetU32 (HashSets.cpp:5-37)
    std::vector<uint64_t> mMoveLo, mMoveHi;                 // SmallPairedHashSetU64 (:71-113)
    uint32_t mEpoch;
    cgb_randstate *mRandState;
    HostRng mRng;
    uint64_t mMinAtoms, mMaxAtoms, mBinLength, mNumCols;
    FastDivU64 mBinDiv, mColDiv;
    uint64_t mBirthIPart;                                   // UINT64_MAX / domainLength: uniform64(1, domainLength)
    double mAlpha, mDomainLength, mNumBins;
    float mLambda, mU1, mU2;
    unsigned mNumProcessed;
    bool mUseCachedRng;
};

// n contiguous elements added, in order, into an fp32 running sum — the bits of `for (i) sum += x[i]` without its chain of
// dependent floating-point adds.  While the sum stays inside one binade [2^e, 2^(e+1)) it is an integer multiple m * u of
// its ulp u = 2^(e-23), and RN(m*u + x) = (m + q + r) * u with q = floor(x / u) and r = 1 when the fraction of x / u is above
// one half, 0 when below — neither depends on m.  Only an exact tie (fraction == 1/2) looks at the sum (round half to
// even).  So a chunk is summed as integers: one pass without any loop-carried floating-point dependence (four elements at a
// time) forms the increments and counts the ties, and with no tie the increments just add up.  A chunk with a tie, one
// whose adds would leave the binade, or one that holds a negative, huge or NaN element is added the plain way — as are the
// first elements, until the sum is a normal number; after a tie the next 64 chunks go the plain way untried (count data
// lies on a grid where ties are everywhere).  gaps::nonZeroMean sums 10^8 elements twice per cgb_run at BASELINE configs[2].
inline void accumulateRun(const float *x, size_t n, float &sum, unsigned &nnz)
{
    const size_t kChunk = 1024;
    float s = sum;
    unsigned count = nnz;
    size_t i = 0;
    uint32_t backoff = 0u; // chunks to add the plain way before the integer route is tried again
    while (i < n)
    {
        const size_t len = (n - i < kChunk) ? n - i : kChunk;
        const float *c = x + i;
        bool done = false;
        if (backoff > 0u) { --backoff; }
        else if (s >= 1.0e-30f && s < 1.0e30f)
        {
            uint32_t bits;
            std::memcpy(&bits, &s, sizeof(bits));
            const uint32_t expField = bits >> 23;                       // s in [2^e, 2^(e+1)), e = expField - 127
            const uint32_t m0 = (bits & 0x007fffffu) | 0x00800000u;     // s = m0 * 2^(e-23)
            const uint32_t scaleBits = (127u + 23u + 127u - expField) << 23; // 2^(23-e); expField is within 27..227 here
            float scale;
            std::memcpy(&scale, &scaleBits, sizeof(scale));
            uint32_t incSum = 0u, ties = 0u, bad = 0u, pos = 0u;
            size_t j = 0;
#if defined(__SSE2__)
            {
                // four elements at a time; compare masks are all-ones, so subtracting them counts
                const __m128 vScale = _mm_set1_ps(scale), vHalf = _mm_set1_ps(0.5f), vTop = _mm_set1_ps(16384.f), vZero = _mm_setzero_ps();
                __m128i vInc = _mm_setzero_si128(), vTies = _mm_setzero_si128(), vBad = _mm_setzero_si128(), vPos = _mm_setzero_si128();
                for (; j + 4 <= len; j += 4)
                {
                    const __m128 xv = _mm_loadu_ps(c + j);
                    const __m128 y = _mm_mul_ps(xv, vScale);
                    const __m128 ok = _mm_and_ps(_mm_cmpge_ps(y, vZero), _mm_cmplt_ps(y, vTop));
                    const __m128 yy = _mm_and_ps(y, ok);
                    const __m128i q = _mm_cvttps_epi32(yy);
                    const __m128 fr = _mm_sub_ps(yy, _mm_cvtepi32_ps(q));
                    vInc = _mm_sub_epi32(_mm_add_epi32(vInc, q), _mm_castps_si128(_mm_cmpgt_ps(fr, vHalf)));
                    vTies = _mm_sub_epi32(vTies, _mm_castps_si128(_mm_cmpeq_ps(fr, vHalf)));
                    vBad = _mm_or_si128(vBad, _mm_castps_si128(_mm_cmpeq_ps(ok, vZero))); // lanes whose mask is all zeros
                    vPos = _mm_sub_epi32(vPos, _mm_castps_si128(_mm_cmpgt_ps(xv, vZero)));
                }
                uint32_t lanes[4];
                _mm_storeu_si128(reinterpret_cast<__m128i*>(lanes), vInc);
                incSum = lanes[0] + lanes[1] + lanes[2] + lanes[3];
                _mm_storeu_si128(reinterpret_cast<__m128i*>(lanes), vTies);
                ties = lanes[0] + lanes[1] + lanes[2] + lanes[3];
                _mm_storeu_si128(reinterpret_cast<__m128i*>(lanes), vBad);
                bad = lanes[0] | lanes[1] | lanes[2] | lanes[3];
                _mm_storeu_si128(reinterpret_cast<__m128i*>(lanes), vPos);
                pos = lanes[0] + lanes[1] + lanes[2] + lanes[3];
            }
#endif
            for (; j < len; ++j)
            {
                const float xv = c[j];
                const float y = xv * scale;                             // exact (a power of two), or far below 1/2 if it underflows
                const bool ok = (y >= 0.f) && (y < 16384.f);            // also false for NaN; 1024 * 16384 = 2^24: no overflow below
                const float yy = ok ? y : 0.f;
                const int32_t q = static_cast<int32_t>(yy);
                const float fr = yy - static_cast<float>(q);
                incSum += static_cast<uint32_t>(q) + (fr > 0.5f ? 1u : 0u);
                ties += (fr == 0.5f) ? 1u : 0u;
                bad |= ok ? 0u : 1u;
                pos += (xv > 0.f) ? 1u : 0u;
            }
            if (bad == 0u && ties == 0u && m0 + incSum < 0x01000000u)
            {
                const uint32_t out = (expField << 23) | ((m0 + incSum) & 0x007fffffu);
                std::memcpy(&s, &out, sizeof(s));
                count += pos;
                done = true;
            }
            else if (ties != 0u) { backoff = 64u; } // data on a coarse grid (counts): ties everywhere, the plain adds are the faster way
        }
        if (!done)
        {
            for (size_t j = 0; j < len; ++j)
            {
                s += c[j];
                if (c[j] > 0.f) { ++count; }
            }
        }
        i += len;
    }
    sum = s;
    nnz = count;
}

// w sampler rows that run down columns r0 .. r0+w-1 of a row-major matrix, gathered into w contiguous rows of L floats:
// 16 x 16 tiles, so that both the reads and the writes move whole cache lines; a wide strip (64 rows: 256 bytes of every
// matrix row) keeps the page walks down — every matrix row lies on a page of its own
inline void gatherStrip(const float *base, uint32_t L, size_t strideL, uint32_t r0, uint32_t w, float *dst)
{
    uint32_t j0 = 0;
    for (; j0 + 16u <= w; j0 += 16u)
    {
        float tile[16][16];
        uint32_t l0 = 0;
        for (; l0 + 16u <= L; l0 += 16u)
        {
            for (uint32_t dl = 0; dl < 16u; ++dl)
            {
                const float *src = base + static_cast<size_t>(l0 + dl) * strideL + r0 + j0;
                for (uint32_t j = 0; j < 16u; ++j) { tile[dl][j] = src[j]; }
            }
            for (uint32_t j = 0; j < 16u; ++j)
            {
                float *out = dst + static_cast<size_t>(j0 + j) * L + l0;
                for (uint32_t dl = 0; dl < 16u; ++dl) { out[dl] = tile[dl][j]; }
            }
        }
        for (; l0 < L; ++l0)
        {
            const float *src = base + static_cast<size_t>(l0) * strideL + r0 + j0;
            for (uint32_t j = 0; j < 16u; ++j) { dst[static_cast<size_t>(j0 + j) * L + l0] = src[j]; }
        }
    }
    for (uint32_t l0 = 0; j0 < w && l0 < L; ++l0)
    {
        const float *src = base + static_cast<size_t>(l0) * strideL + r0;
        for (uint32_t j = j0; j < w; ++j) { dst[static_cast<size_t>(j) * L + l0] = src[j]; }
    }
}

// The column-walking case of runningSum (below) with the memory side taken off the chain of adds: helper threads gather
// strips of 16 sampler rows (one cache line's width of the row-major matrix) into contiguous buffers a few strips ahead,
// the calling thread only adds — same elements, same order, same bits; on a large matrix the call then costs what the
// dependent adds cost (about 1 ns each) instead of three to four times that.  Returns false (nothing summed) when the
// matrix is small or threads / buffers cannot be had; the caller then walks the strips itself.
inline bool runningSumPipelined(const float *base, uint32_t nRows, uint32_t L, size_t strideL, float &sum, unsigned &nnz)
{
    const uint32_t B = 16, kSlots = 6, kGatherers = 3;
    if (static_cast<uint64_t>(nRows) * L < (1ull << 24)) { return false; }
    {
        const char *knob = std::getenv("COGAPS_SUM_PIPELINE"); // 0: the calling thread gathers the strips itself
        if (knob && knob[0] == '0') { return false; }
    }
    const uint32_t nStrips = (nRows + B - 1) / B;
    float *slots = new (std::nothrow) float[static_cast<size_t>(kSlots) * B * L];
    if (slots == nullptr) { return false; }
    std::atomic<uint32_t> ready[kSlots];      // strip number + 1 held by the slot
    for (uint32_t i = 0; i < kSlots; ++i) { ready[i].store(0u); }
    std::atomic<uint32_t> consumed(0u);       // strips the adder is done with
    std::atomic<bool> stop(false);
    auto gather = [&](uint32_t first)
    {
        for (uint32_t sIdx = first; sIdx < nStrips && !stop.load(std::memory_order_relaxed); sIdx += kGatherers)
        {
            while (sIdx >= consumed.load(std::memory_order_acquire) + kSlots)   // the slot still holds strip sIdx - kSlots
            {
                if (stop.load(std::memory_order_relaxed)) { return; }
                std::this_thread::yield();
            }
            float *dst = slots + static_cast<size_t>(sIdx % kSlots) * B * L;
            const uint32_t r0 = sIdx * B;
            const uint32_t w = (nRows - r0 < B) ? nRows - r0 : B;
            gatherStrip(base, L, strideL, r0, w, dst);
            ready[sIdx % kSlots].store(sIdx + 1u, std::memory_order_release);
        }
    };
    std::thread workers[kGatherers];
    uint32_t started = 0;
    try
    {
        for (; started < kGatherers; ++started) { workers[started] = std::thread(gather, started); }
    }
    catch (...)
    {
        stop.store(true);
        for (uint32_t i = 0; i < started; ++i) { workers[i].join(); }
        delete[] slots;
        return false;
    }
    float acc = sum;
    unsigned count = nnz;
    for (uint32_t sIdx = 0; sIdx < nStrips; ++sIdx)
    {
        while (ready[sIdx % kSlots].load(std::memory_order_acquire) != sIdx + 1u) { std::this_thread::yield(); }
        const float *src = slots + static_cast<size_t>(sIdx % kSlots) * B * L;
        const uint32_t r0 = sIdx * B;
        const uint32_t w = (nRows - r0 < B) ? nRows - r0 : B;
        accumulateRun(src, static_cast<size_t>(w) * L, acc, count);          // the strip's rows lie one after the other
        consumed.store(sIdx + 1u, std::memory_order_release);
    }
    for (uint32_t i = 0; i < kGatherers; ++i) { workers[i].join(); }
    delete[] slots;
    sum = acc;
    nnz = count;
    return true;
}

// gaps::nonZeroMean's numerator and denominator (MatrixMath.cpp:39-55): ONE fp32 running sum over the sampler's rows in
// order, and the count of positive elements.  The order of the additions is the result, so the chain of dependent
// adds cannot be split; what can be helped is the memory side.  Row r starts at base + r * strideR, its elements are
// strideL floats apart.  When the rows run down the columns of a row-major matrix (strideR == 1) the plain walk
// touches a new cache line per element and thrashes any host whose L2 does not hold one full sweep (a 9x slowdown was
// measured on such a box); 16 sampler rows at a time are therefore gathered into a contiguous strip first — a
// streaming transpose of one cache line's width — and summed from there in exactly the same order.
inline void runningSum(const float *base, uint32_t nRows, uint32_t L, size_t strideR, size_t strideL, float &sumOut, unsigned &nnzOut)
{
    // runs on a helper thread: no exception may start here, so the strip is allocated nothrow and the plain walk is
    // the fallback
    struct Strip
    {
        enum { kRows = 16 };
        float *data;
        Strip() : data(nullptr) {}
        ~Strip() { delete[] data; }
        bool reserveFor(uint32_t rowLength)
        {
            data = new (std::nothrow) float[static_cast<size_t>(kRows) * rowLength];
            return data != nullptr;
        }
    } strip;
    float sum = 0.f;
    unsigned nnz = 0;
    if (strideL == 1)
    {
        for (uint32_t r = 0; r < nRows; ++r) { accumulateRun(base + static_cast<size_t>(r) * strideR, L, sum, nnz); }
    }
    else if (strideR == 1 && runningSumPipelined(base, nRows, L, strideL, sum, nnz))
    {
        // strips gathered by helper threads ahead of the chain of adds (below)
    }
    else if (strideR == 1 && strip.reserveFor(L))
    {
        const uint32_t B = Strip::kRows;
        for (uint32_t r0 = 0; r0 < nRows; r0 += B)
        {
            const uint32_t w = (nRows - r0 < B) ? nRows - r0 : B;
            gatherStrip(base, L, strideL, r0, w, strip.data);
            accumulateRun(strip.data, static_cast<size_t>(w) * L, sum, nnz);
        }
    }
    else
    {
        for (uint32_t r = 0; r < nRows; ++r)
        {
            const float *row = base + static_cast<size_t>(r) * strideR;
            for (uint32_t l = 0; l < L; ++l)
            {
                const float v = row[static_cast<size_t>(l) * strideL];
                sum += v;
                if (v > 0.f) { ++nnz; }
            }
        }
    }
    sumOut = sum;
    nnzOut = nnz;
}

// What one evaluated proposal does to the host's state once its outcome is known — the domain / queue half of
// AsynchronousGibbsSampler::birth/death/move/exchange (AsynchronousGibbsSampler.h:126-219).  `accepted` means: B the
// atom is born with mass1; D the atom survives with mass1; M the atom moves to hp.pos; E both masses change.
// postedMass1 is atom1's mass as the evaluator was given it.  Returns true when the factor matrix changed, i.e. the
// device committed a row for this proposal; type is left for the caller to reject when it is none of B/D/M/E.
inline bool applyToDomain(AtomicDomain &domain, ProposalQueue &queue, const HostProposal &hp, bool accepted,
                          float mass1, float mass2, float postedMass1)
{
    bool commit = false;
    switch (hp.type)
    {
        case 'B':
            if (accepted)
            {
                queue.acceptBirth();
                domain.atom(hp.atom1).mass = mass1;
                commit = true;
            }
            else
            {
                queue.rejectBirth();
                domain.cacheErase(hp.atom1);
            }
            break;
        case 'D':
            if (accepted)
            {
                queue.rejectDeath();
                domain.atom(hp.atom1).mass = mass1;
                commit = (mass1 != postedMass1);
            }
            else
            {
                queue.acceptDeath();
                domain.cacheErase(hp.atom1);
                commit = true;
            }
            break;
        case 'M':
            if (accepted) { domain.move(hp.atom1, hp.pos); }
            commit = accepted;
            break;
        case 'E':
            if (accepted)
            {
                domain.atom(hp.atom1).mass = mass1;
                domain.atom(hp.atom2).mass = mass2;
            }
            commit = accepted;
            break;
        default: break;
    }
    return commit;
}

} // namespace cgb

struct cgb_sampler
{
    // shape: nRows x k factor matrix; every row owns a length-L slice of D / S / AP
    uint32_t nRows, L, k;
    uint32_t ld, ldM;
    float alpha, lambda, maxGibbsMass, annealingTemp;
    float dataSparsity;
    bool hasS;
    cgb_randstate *rs;
    const cgb_sampler *other;
    int device;
    cudaStream_t stream;

    // SparseNormalModel: no AP, no S; D also as CSR over sampler rows, the factor in two copies, Z tables
    bool sparse;
    uint32_t *dSpRowPtr, *dSpIdx;
    float *dSpVal, *dMrows, *dZ1, *dZ2;
    uint32_t ldR;

    // device buffers
    float *dD, *dS, *dAP, *dM;
    int *dColNonzero;
    cgb::AlphaPair *dPartials;
    uint32_t *dTickets;
    double *dReducePartials;
    cgb::DevOutcome *hOutcomes;   // pinned + mapped; the kernel writes results straight to host memory
    double *hReducePartials;      // pinned
    uint32_t seg, nSeg, segPad;
    size_t smemBytes;
    bool tablesInSmem;            // resident grid: erf / erfinv tables copied to shared memory (else read through L2)

    // host generator
    cgb::AtomicDomain domain;
    cgb::ProposalQueue queue;
    bool sequential;              // asynchronousUpdates == 0: SingleThreadedGibbsSampler semantics
    cgb::SequentialState seq;
    float avgQueueLength, numQueueSamples;

    // row-parallel sweep (sweep.cuh): the atoms live on the device, one sorted run per row of the factor matrix
    int updateMode;               // CGB_UPDATE_EXACT / CGB_UPDATE_SWEEP
    uint64_t *dSwPos;             // [nRows][swCap] positions relative to the row's segment of the atomic domain
    float *dSwMass;               // [nRows][swCap]
    uint32_t *dSwCount;           // [nRows]
    uint32_t *dSwOrder;           // [nRows] rows by decreasing atom count: the order CTAs take them in (sweep_order_kernel)
    uint32_t swCap;
    void *dSwCounters;            // cgb::SweepCounters
    void *hSwCounters;            // pinned copy
    uint64_t swTotalAtoms;
    uint64_t swUpdates;           // update() calls made in sweep mode: its parity picks which adjacent-row pairs get their turn
    uint64_t swOverflow;          // births dropped because a row's store was full (expected 0; the store regrows between updates)

    // bench counters
    cgb_sampler_counters counters;
    bool timeKernels;
    unsigned long long *dPhaseClocks; // debug phase profile
    double phaseSum[cgb::kPhaseSlots];
    uint64_t phaseTasks;
    cudaEvent_t evStart, evStop;

    // per-row commit counts: the device's counter and what the host knows it will reach
    uint32_t *dRowVersion;
    std::vector<uint32_t> rowVersion;
    // commit tracking.  The device counts completed CTA-commits per parity of the chunk tag, one extra CTA mirrors
    // both counts into host memory.  rowPending[r] = tag (mailSeq) of the chunk whose proposal last rewrote row r,
    // 0 = none.  commitsExpected[q] = commits of every APPLIED outcome from chunks of parity q.  While chunk c is
    // being posted its own commits already land in counter c & 1, so only the other parity can be proven complete
    // by "mirror == expected" at any time; the own parity is proven once, before the chunk's first post
    // (provenThrough[q]: every chunk of parity q up to this tag is complete).
    volatile unsigned long long *hCommitsMirror; // [2], pinned + mapped
    std::vector<uint64_t> rowPending;
    uint64_t commitsExpected[2], provenThrough[2];

    // resident-kernel mode: task records streamed to the grid through pinned host memory
    bool usePersistent;
    bool persistentRunning;
    void *hSlots;                 // cgb::StreamRecord[nClusters][kStreamRing][nSeg] + the doorbell line, pinned + mapped
    size_t nSlotRecords;
    void *hStreamOutcomes;        // HostOutcome[kMaxPersistentBatch], pinned + mapped
    void *dStreamStats;           // cgb::StreamStats
    unsigned long long mailSeq;   // tag of the last chunk posted
    // Work only goes to clusters that have reported in: a worker cluster writes the launch epoch into hAlive[c]
    // when it starts, so a grid that is not (yet) fully resident — another chain's grid, another process — can
    // never be handed a task it cannot run.  Tickets are per cluster, consecutive from ticketBase.
    volatile uint32_t *hAlive;            // [nClusters], pinned + mapped
    uint32_t launchEpoch;                 // value the clusters of the current launch report
    uint32_t ticketBase;                  // ticket0 of the current launch
    std::vector<uint32_t> clusterTicket;  // last ticket handed to each cluster
    std::vector<uint32_t> aliveList;      // clusters known to be running, in the order they reported
    std::vector<uint8_t> aliveSeen;
    size_t aliveNext;                     // round-robin cursor into aliveList
    uint32_t aliveLooks;                  // posts since the launch: cadence of the look for newly started clusters
    int persistentGrid;           // CTAs of the resident grid (0 until first launch)
    int residentShare;            // this sampler's grid takes 1/share of the device (0: the process default)
    uint32_t nClusters;           // worker clusters of the resident grid (one more cluster mirrors the commit count)
    double lastPostTime;
    uint32_t chunkTag;            // low 31 bits of mailSeq for the chunk being posted
    uint32_t chunkPosted;         // proposals posted in this chunk
    size_t chunkBase;             // queue index of the chunk's first proposal
    size_t chunkCap;              // proposals per chunk (kMaxPersistentBatch; smaller in tests of the chunked path)
    bool forceRowWait;            // tests: never skip the rowVersion check of a row with a recorded commit
    std::vector<uint64_t> slotOwner;      // per (cluster, ring slot): (mailSeq << 32) | proposal index of the last record
    std::vector<cgb::DevProposal> posted; // what was sent for each proposal of the chunk (masses as posted)
    std::vector<uint8_t> arrived;         // outcome of proposal i of the chunk already collected
    std::vector<cgb::DevOutcome> collected;
};

#endif // CGB_SAMPLER_H
