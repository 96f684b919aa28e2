// sampler.h — host side of one Gibbs sampler (one factor matrix): the atomic domain, the sequential
// proposal generator, and the handle that owns the device buffers.
//
// Reference restated: atomic/ProposalQueue.cpp:19-283 (generator, conflict rules, seed roll-back),
// gibbs_sampler/AsynchronousGibbsSampler.h:63-122 (ctor, update loop, queue-length diagnostics),
// gibbs_sampler/DenseNormalModel.h:66-95 (ctor: data orientation, default uncertainty, lambda).
#ifndef CGB_SAMPLER_H
#define CGB_SAMPLER_H

#include "atomic_domain.h"
#include "device_types.h"
#include "host_rng.h"

#include <cuda_runtime.h>
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

// ProposalQueue (atomic/ProposalQueue.h:30-72)
class ProposalQueue
{
public:
    void init(uint64_t nElements, uint64_t nPatterns, cgb_randstate *rs, float alpha, float lambda);
    void populate(AtomicDomain &domain, unsigned limit);   // ProposalQueue.cpp:53-76
    void clear();                                           // :78-85
    unsigned nProcessed() const { return mNumProcessed; }
    std::vector<HostProposal> &entries() { return mQueue; }
    void acceptDeath() { --mMaxAtoms; }                     // :97-119
    void rejectDeath() { ++mMinAtoms; }
    void acceptBirth() { ++mMinAtoms; }
    void rejectBirth() { --mMaxAtoms; }
    uint64_t minAtoms() const { return mMinAtoms; }
    uint64_t maxAtoms() const { return mMaxAtoms; }

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

    std::vector<HostProposal> mQueue;
    std::vector<uint32_t> mUsedRows;                        // FixedHashSetU32 (HashSets.cpp:5-37)
    std::vector<uint64_t> mMoveLo, mMoveHi;                 // SmallPairedHashSetU64 (:71-113)
    uint32_t mEpoch;
    cgb_randstate *mRandState;
    HostRng mRng;
    uint64_t mMinAtoms, mMaxAtoms, mBinLength, mNumCols;
    double mAlpha, mDomainLength, mNumBins;
    float mLambda, mU1, mU2;
    unsigned mNumProcessed;
    bool mUseCachedRng;
};

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

    // host generator
    cgb::AtomicDomain domain;
    cgb::ProposalQueue queue;
    float avgQueueLength, numQueueSamples;

    // bench counters
    cgb_sampler_counters counters;
    bool timeKernels;
    unsigned long long *dPhaseClocks; // debug phase profile
    double phaseSum[cgb::kPhaseSlots];
    uint64_t phaseTasks;
    cudaEvent_t evStart, evStop;

    // persistent-kernel mode: mailbox between the host generator and the resident grid
    bool usePersistent;
    bool persistentRunning;
    void *hMailbox;               // cgb::HostMailbox, pinned + mapped
    void *dMailbox;               // cgb::DeviceMailbox
    unsigned long long mailSeq;   // id of the last batch posted
    int persistentGrid;
    double lastPostTime;
    uint32_t lastPostedTasks;
};

#endif // CGB_SAMPLER_H
