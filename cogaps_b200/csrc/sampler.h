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
#include <new>
#include <string>
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
        for (uint32_t r = 0; r < nRows; ++r)
        {
            const float *row = base + static_cast<size_t>(r) * strideR;
            for (uint32_t l = 0; l < L; ++l)
            {
                sum += row[l];
                if (row[l] > 0.f) { ++nnz; }
            }
        }
    }
    else if (strideR == 1 && strip.reserveFor(L))
    {
        const uint32_t B = Strip::kRows;
        for (uint32_t r0 = 0; r0 < nRows; r0 += B)
        {
            const uint32_t w = (nRows - r0 < B) ? nRows - r0 : B;
            for (uint32_t l = 0; l < L; ++l)
            {
                const float *src = base + static_cast<size_t>(l) * strideL + r0;
                for (uint32_t j = 0; j < w; ++j) { strip.data[static_cast<size_t>(j) * L + l] = src[j]; }
            }
            for (uint32_t j = 0; j < w; ++j)
            {
                const float *row = strip.data + static_cast<size_t>(j) * L;
                for (uint32_t l = 0; l < L; ++l)
                {
                    sum += row[l];
                    if (row[l] > 0.f) { ++nnz; }
                }
            }
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
