// device_types.h — plain structs that cross the host/device line (kernel parameters and results).
#ifndef CGB_DEVICE_TYPES_H
#define CGB_DEVICE_TYPES_H

#include <stdint.h>

namespace cgb {

static const int kThreads = 512;      // threads per CTA of the dense eval kernel (16 warps)
static const int kSparseThreads = 256; // threads per CTA of the sparse eval kernel: visited element e belongs to lane e % 256
static const int kSparseGroup = 4;     // non-zeros per thread and pass of the sparse scan (1024 per CTA and pass)
static const int kVec = 4;            // floats per vector access (16 B)
static const int kMaxBatch = 384;     // proposals per launch carried in kernel-parameter space
static const int kMaxCluster = 8;     // portable cluster limit
static const int kMaxPersistentBatch = 1024; // proposals per chunk through the resident kernel's mailbox
static const int kStreamRing = 8;     // task-record slots per resident cluster (ring, host pinned memory)
static const uint32_t kStreamExit = 0xFFFFFFFFu; // record type that retires a resident CTA
static const uint32_t kStreamWait1 = 0x100u;     // record type bits: the host has no proof yet that the last commit
static const uint32_t kStreamWait2 = 0x200u;     //   to row r1 / r2 has landed, so the task must check rowVersion
static const uint32_t kStreamSeq = 0x400u;       // record type bit: proposal of the sequential sampler (DevProposal::pad bit 0)
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
    uint32_t pad;      // bit 0: proposal of the sequential sampler (birth threshold, shared rng stream)
};

struct DevOutcome
{
    float mass1;       // B: new mass; D: surviving mass; E: new mass of atom1
    float mass2;       // E: new mass of atom2
    uint32_t accepted; // B born / D survives / M moved / E exchanged
    float s;           // alpha parameters (annealed) as used by the decision; raw sums for probes
    float s_mu;
    uint32_t pad[3];   // pad[0]: draws the epilogue took from the proposal's rng stream (0..2)
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
    uint32_t *rowVersion;    // [nRows] CTA-commits applied to each row (AP row + its factor elements)
    // SparseNormalModel only (all nullptr / 0 for the dense model)
    const uint32_t *spRowPtr;   // [nRows + 1] non-zeros of each sampler row of D ...
    const uint32_t *spIdx;      // [nnz]      ... their scan index, ascending within a row
    const float *spVal;         // [nnz]      ... and value (> 0)
    float *Mrows;               // [nRows][ldR] row copy of the factor (HybridMatrix::mRows: never zeroed below epsilon;
                                //             M above is mCols: values below epsilon stored as 0)
    const float *otherMrows;    // [L][ldR]
    const float *Z1;            // [k]     lookup tables of the other factor (generateLookupTables)
    const float *Z2;            // [k][k]
    uint32_t ldR;               // floats per row of Mrows (k rounded up to 4)
    float beta;
    uint32_t nRows, L, k;
    uint32_t ld, ldM, ldOther;
    uint32_t seg;            // floats per segment (multiple of 4)
    uint32_t nSeg;           // segments per row = cluster size
    uint32_t segPad;         // floats reserved per stream in shared memory
    uint32_t tablesInSmem;   // resident grid keeps the erf / erfinv tables in shared memory (short segments only)
    float lambda, maxGibbsMass, annealingTemp;
};

// One (proposal, row) work item as the host streams it to the resident grid: 64 bytes = one PCIe read.
// Word layout matters (the poller validates 16-byte chunks): w12..w15 = batch, ticket, check, pad.
struct StreamRecord
{
    uint64_t rng;         // PCG state after the generator's own draws
    uint32_t r1, c1;
    uint32_t r2, c2;
    float m1, m2;
    uint32_t type;        // 'B','D','M','E' (| kStreamWait1/2) or kStreamExit
    uint32_t piPart;      // index of the proposal in its chunk | (which of its rows this task is) << 31
    uint32_t ver1, ver2;  // rowVersion[r1] / rowVersion[r2] this task must see before it reads those rows
    uint32_t batch;       // chunk tag, echoed in the outcome record
    uint32_t ticket;      // 1-based serial of this record in its cluster's stream: what the poller waits for
    uint32_t check;       // stream_check of the other 15 words: tells a complete record from a torn one
    uint32_t pad;
};

// device-side diagnostics of the resident grid
struct StreamStats
{
    unsigned long long taskNs;     // sum over tasks: record seen -> commit done (leader CTA, globaltimer)
    unsigned long long tasks;
    unsigned long long maxTaskNs;
    unsigned long long verWaitNs;  // sum over tasks: time spent waiting for a row version
    unsigned long long decideNs;   // sum over tasks: record seen -> outcome posted (deciding CTA)
    unsigned long long outcomes;
    unsigned long long visited;     // sparse model: elements the scans visited (data non-zero and factor non-zero)
    unsigned long long pad;
    // CTA-commits completed (AP row written, fenced), one counter per parity of the chunk tag that carried the
    // proposal; the mirror CTA copies both to the host.  Two counters because commits of the chunk being posted
    // must not be mistaken for the stragglers of the one before (see rowSettled in cogaps_b200.cu).
    unsigned long long commitsDone[2];
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
