// checkpoint.h — the reference's checkpoint wire format, host side only (no CUDA in here).
//
// Reference restated: utils/Archive.h:16-87 (raw little-endian scalars behind a u32 magic number) and the
// operator<< / operator>> pairs createCheckpoint / processCheckpoint stream through it (GapsRunner.cpp:237-240,
// 264-268): GapsParameters.cpp:82-96, math/Random.cpp:202-212,250-260,347-357, Matrix.cpp:182-204,
// Vector.cpp:90-111, HybridMatrix.cpp:85-116, HybridVector.cpp:103-131, DenseNormalModel.cpp:260-270,
// SparseNormalModel.cpp:313-323, ConcurrentAtomicDomain.cpp:134-155, ConcurrentAtom.cpp:98-108,
// ProposalQueue.cpp:285-299, AsynchronousGibbsSampler.h:221-233, GapsStatistics.cpp:164-176.
//
// An *Image is the host-memory form of one archived object; cogaps_b200.cu moves images to and from the
// device-resident samplers and statistics.  Files we write are byte-identical to the reference's for the same
// state, and either side can resume from the other's (tests/test_checkpoint.py).
//
// Asynchronous sampler only: the reference's SingleThreadedGibbsSampler does not archive its rng stream and its
// operator>> writes instead of reading (SingleThreadedGibbsSampler.h:260-273), so there is no format to follow.
#ifndef CGB_CHECKPOINT_H
#define CGB_CHECKPOINT_H

#include <stdint.h>
#include <string>
#include <vector>

namespace cgb {

static const uint32_t kArchiveMagic = 0xB123AA4Du; // utils/Archive.h:16

// ProposalQueue's archived members (ProposalQueue.cpp:285-291)
struct QueueState
{
    uint64_t rng, minAtoms, maxAtoms, binLength, numCols;
    double alpha, domainLength, numBins;
    float lambda;
    bool useCachedRng;
    float u1, u2;
};

// AsynchronousGibbsSampler<DenseNormalModel | SparseNormalModel>
struct SamplerImage
{
    bool sparse;
    uint32_t nRows, k;
    std::vector<float> cols;   // [k][nRows]: Matrix::mCols (dense) / HybridMatrix::mCols (sparse: below epsilon stored as 0)
    std::vector<float> rows;   // sparse only, [nRows][k]: HybridMatrix::mRows
    float beta;                // sparse only (SparseNormalModel::mBeta)
    uint64_t domainLength;
    std::vector<uint64_t> pos; // atoms in pick-vector order (ConcurrentAtomicDomain::mAtoms)
    std::vector<float> mass;
    QueueState queue;
    SamplerImage() : sparse(false), nRows(0), k(0), beta(0.f), domainLength(0), queue() {}
};

// GapsStatistics: running sums, pattern-major like the factor matrices
struct StatsImage
{
    uint32_t nGenes, nSamples, k;
    std::vector<float> aMean, aSq; // [k][nGenes]
    std::vector<float> pMean, pSq; // [k][nSamples]
    uint32_t statUpdates, numPatterns;
    StatsImage() : nGenes(0), nSamples(0), k(0), statUpdates(0), numPatterns(0) {}
};

// the members of GapsParameters that are archived (GapsParameters.cpp:82-88)
struct ParamsImage
{
    uint32_t seed, nGenes, nSamples, nPatterns, nIterations;
    float alphaA, alphaP, maxGibbsMassA, maxGibbsMassP;
    bool useSparseOptimization;
    uint32_t checkpointInterval;
};

struct CheckpointImage
{
    ParamsImage params;
    uint64_t seeder[2];  // GapsRandomState: the xoroshiro128+ state
    SamplerImage A, P;
    StatsImage stats;
    int32_t phase;       // GapsAlgorithmPhase as int (GapsParameters.h:19-24)
    uint32_t iter;
    uint64_t rng;        // the run loop's GapsRng
};

class ByteWriter
{
public:
    template <class T> void put(T v)
    {
        const uint8_t *p = reinterpret_cast<const uint8_t*>(&v);
        mBytes.insert(mBytes.end(), p, p + sizeof(T));
    }
    void putFloats(const float *v, size_t n)
    {
        const uint8_t *p = reinterpret_cast<const uint8_t*>(v);
        mBytes.insert(mBytes.end(), p, p + n * sizeof(float));
    }
    std::vector<uint8_t> &bytes() { return mBytes; }
private:
    std::vector<uint8_t> mBytes;
};

class ByteReader
{
public:
    ByteReader(const uint8_t *data, size_t size) : mData(data), mSize(size), mOff(0), mOk(true) {}
    template <class T> T get()
    {
        T v = T();
        take(&v, sizeof(T));
        return v;
    }
    void getFloats(float *out, size_t n) { take(out, n * sizeof(float)); }
    void skip(size_t n) { if (mOk && mSize - mOff >= n) { mOff += n; } else { mOk = false; } }
    bool ok() const { return mOk; }
    void fail() { mOk = false; }
    size_t offset() const { return mOff; }
    size_t remaining() const { return mSize - mOff; }
private:
    void take(void *out, size_t n);
    const uint8_t *mData;
    size_t mSize, mOff;
    bool mOk;
};

void putParams(ByteWriter &w, const ParamsImage &p);
bool getParams(ByteReader &r, ParamsImage &p);
void putSampler(ByteWriter &w, const SamplerImage &s);
bool getSampler(ByteReader &r, bool sparse, SamplerImage &s, std::string &err);
void putStats(ByteWriter &w, const StatsImage &st);
bool getStats(ByteReader &r, StatsImage &st, std::string &err);
void putCheckpoint(ByteWriter &w, const CheckpointImage &c);
bool getCheckpoint(ByteReader &r, CheckpointImage &c, std::string &err);

// run_helper reads only this much before the samplers exist (GapsRunner.cpp:99-105)
bool readCheckpointHeader(const char *path, ParamsImage &p, uint64_t seeder[2], std::string &err);
bool readCheckpointFile(const char *path, CheckpointImage &c, std::string &err);
// createCheckpoint's file protocol (GapsRunner.cpp:233-244): the previous file is kept as <path>.backup while the
// new one is written, and removed afterwards
bool writeCheckpointFile(const char *path, const CheckpointImage &c, std::string &err);
bool readWholeFile(const char *path, std::vector<uint8_t> &out, std::string &err);
bool writeWholeFile(const char *path, const std::vector<uint8_t> &bytes, std::string &err);

} // namespace cgb

#endif // CGB_CHECKPOINT_H
