// device_types.h — plain structs that cross the host/device line (kernel parameters and results).
#ifndef CGB_DEVICE_TYPES_H
#define CGB_DEVICE_TYPES_H

#include <stdint.h>

namespace cgb {

static const int kThreads = 512;      // threads per CTA of the eval kernel (16 warps)
static const int kVec = 4;            // floats per vector access (16 B)
static const int kMaxBatch = 384;     // proposals per launch carried in kernel-parameter space
static const int kMaxCluster = 8;     // portable cluster limit
static const int kMaxPersistentBatch = 1024; // proposals per batch through the persistent kernel's mailbox
static const unsigned long long kExitSeq = ~0ull;
static const int kPhaseSlots = 12;    // debug phase timestamps per task
static const uint32_t kProbe = 'P';   // lock-step probe pseudo-proposal

struct AlphaPair { float s, s_mu; };

// One queued proposal as the device sees it (AtomicProposal, atomic/ProposalQueue.h:15-28, with the
// atom pointers replaced by the two masses the evaluators read).
struct DevProposal
{
    uint64_t rng;      // PCG state after the generator's own draws
    uint32_t r1, c1;   // bin of atom1
    uint32_t r2, c2;   // bin of the move destination / of atom2
    float m1, m2;      // atom masses
    uint32_t type;     // 'B','D','M','E' or kProbe
    uint32_t variant;  // probe only: 0 alphaParameters(r1,c1); 1 (r1,c1,r2,c2); 2 WithChange
    float ch;          // probe only: change for variant 2
    uint32_t pad;
};

struct DevOutcome
{
    float mass1;       // B: new mass; D: surviving mass; E: new mass of atom1
    float mass2;       // E: new mass of atom2
    uint32_t accepted; // B born / D survives / M moved / E exchanged
    float s;           // alpha parameters (annealed) as used by the decision; raw sums for probes
    float s_mu;
    uint32_t pad[3];
};

// Everything the eval kernel needs about one sampler.
struct ModelView
{
    const float *D;          // [nRows][ld]  data, one sampler row per line
    const float *S;          // [nRows][ld]  uncertainty, or nullptr when it is max(0.1 D, 0.1)
    float *AP;               // [nRows][ld]  cached product A*P for this orientation
    float *M;                // [k][ldM]     this sampler's factor matrix, pattern-major
    const float *otherM;     // [k][ldOther] the other sampler's factor matrix (length-L columns)
    const int *otherColNonzero; // [k]       canUseGibbs(col) of the other matrix
    const float *erf;        // lookup tables
    const float *erfinv;
    DevOutcome *outcomes;    // [kMaxBatch] pinned host memory, written by the kernel
    AlphaPair *partials;     // [kMaxPersistentBatch][2] cross-cluster (s, s_mu) of two-row proposals
    uint32_t *tickets;       // [kMaxPersistentBatch]
    unsigned long long *phaseClocks; // debug: [kMaxBatch][kPhaseSlots] SM clock at each phase, or nullptr
    uint32_t nRows, L, k;
    uint32_t ld, ldM, ldOther;
    uint32_t seg;            // floats per segment (multiple of 4)
    uint32_t nSeg;           // segments per row = cluster size
    uint32_t segPad;         // floats reserved per stream in shared memory
    float lambda, maxGibbsMass, annealingTemp;
};

struct EvalParams
{
    ModelView mv;
    uint32_t nProps;
    uint32_t nTasks;                 // nProps + number of two-row proposals
    uint16_t extra[kMaxBatch];       // task nProps+j is the second row of proposal extra[j]
    DevProposal props[kMaxBatch];
};

} // namespace cgb

#endif
