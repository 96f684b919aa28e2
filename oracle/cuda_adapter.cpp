// cuda_adapter.cpp — the reference-side binding, COMPILED: `CudaGibbsSampler` satisfies the Sampler concept of the
// reference's own run loop (runCoGAPSAlgorithm<Sampler>, src/GapsRunner.cpp:381-503; GapsStatistics.h:129-202) by
// forwarding every member to the C ABI of include/cogaps_b200.h.  This is the file INTEGRATION.md tells a maintainer to
// add next to chooseSampler (src/GapsRunner.cpp:65-78).
//
// TEST INFRASTRUCTURE in this repository (it lives under oracle/ because it compiles the reference where it lies):
// oracle/Makefile builds it into oracle/_ref/libcogaps_ref_adapter.so = the reference's sources, UNMODIFIED and not
// copied — src/GapsRunner.cpp is pulled into this translation unit with #include so that its file-local templates
// (runCoGAPSAlgorithm, runOnePhase, updateSampler, createCheckpoint, ...) can be instantiated with the new sampler type;
// every other reference source is compiled as its own object — plus oracle/ref_driver.cpp for the C entry points, linked
// against cogaps_b200/libcogaps_b200.so.  tests/test_gpu_parity.py::test_reference_run_loop_drives_the_c_abi checks that
// the reference's loop driving the device through this class gives, bit for bit, what cgb_run gives.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include <stdint.h>

// the adapter reads the random state's seeder and the archive's stream; access specifiers do not change layout
#define private public
#define protected public
#include "GapsRunner.cpp"
#undef private
#undef protected

#include "../include/cogaps_b200.h"

namespace {

void adapterCheck(int rc, const char *what)
{
    if (rc != CGB_OK) { GAPS_ERROR(what << ": " << cgb_last_error() << "\n"); }
}

// one device-side random state per reference GapsRandomState: the reference's object stays the master copy of the
// seeder (the run loop's own GapsRng and `ar << *randState` use it), ours is brought in line around every call that draws
std::map<const GapsRandomState*, cgb_randstate*> g_states;

cgb_randstate *deviceState(GapsRandomState *rs, unsigned seed)
{
    std::map<const GapsRandomState*, cgb_randstate*>::iterator it = g_states.find(rs);
    if (it != g_states.end()) { return it->second; }
    cgb_randstate *out = NULL;
    adapterCheck(cgb_randstate_create(seed, &out), "cgb_randstate_create");
    // the reference's own lookup tables (math/Random.cpp:269-295; with Boost linked these are Boost's)
    adapterCheck(cgb_randstate_set_tables(out, rs->mErfLookupTable, rs->mErfinvLookupTable, rs->mQgammaLookupTable), "cgb_randstate_set_tables");
    g_states[rs] = out;
    return out;
}

void seederIn(GapsRandomState *rs, cgb_randstate *dev)
{
    uint64_t st[2] = {rs->mSeeder.mState[0], rs->mSeeder.mState[1]};
    adapterCheck(cgb_randstate_set_state(dev, st), "cgb_randstate_set_state");
}

void seederOut(GapsRandomState *rs, cgb_randstate *dev)
{
    uint64_t st[2];
    adapterCheck(cgb_randstate_get_state(dev, st), "cgb_randstate_get_state");
    rs->mSeeder.mState[0] = st[0];
    rs->mSeeder.mState[1] = st[1];
}

void toParams(const GapsParameters &params, cgb_params *p)
{
    cgb_params_default(p);
    p->seed = params.seed;
    p->nPatterns = params.nPatterns;
    p->nIterations = params.nIterations;
    p->maxThreads = params.maxThreads;
    p->transposeData = params.transposeData ? 1 : 0;
    p->useSparseOptimization = 0;
    p->asynchronousUpdates = params.asynchronousUpdates ? 1 : 0;
    p->whichMatrixFixed = params.whichMatrixFixed;
    p->subsetGenes = params.subsetGenes ? 1 : 0;
    p->nSubsetIndices = params.subsetData ? static_cast<uint32_t>(params.dataIndicesSubset.size()) : 0;
    p->subsetIndices = p->nSubsetIndices ? &params.dataIndicesSubset[0] : NULL;
    const char *mode = std::getenv("COGAPS_UPDATE_MODE"); // the switch arrives through the environment
    p->updateMode = (mode != NULL && mode[0] == '1') ? CGB_UPDATE_SWEEP : CGB_UPDATE_EXACT;
}

} // namespace

class CudaGibbsSampler;
Archive& operator<<(Archive &ar, const CudaGibbsSampler &s);
Archive& operator>>(Archive &ar, CudaGibbsSampler &s);

// DenseNormalModel stays the base class: GapsStatistics reads mMatrix (update / updateA / updateP / takeSnapshot,
// GapsStatistics.h:129-202) and meanChiSq takes a `const DenseNormalModel&` for mDMatrix / mSMatrix
// (GapsStatistics.cpp:63-87).  Its host copies of D and S are kept for exactly that; mMatrix mirrors the device's factor
// matrix after every call that changes it; its AP matrix and its scans are never used.
class CudaGibbsSampler : public DenseNormalModel
{
public:
    CudaGibbsSampler(const Matrix &data, bool transpose, bool subsetRows, float alpha, float maxGibbsMass,
                     const GapsParameters &params, GapsRandomState *randState)
        : DenseNormalModel(data, transpose, subsetRows, params, alpha, maxGibbsMass), mHandle(NULL), mRandState(randState),
          mDevState(deviceState(randState, params.seed)), mUpdateMode(CGB_UPDATE_EXACT)
    {
        cgb_params p;
        toParams(params, &p);
        mUpdateMode = p.updateMode;
        std::vector<float> rowMajor(static_cast<size_t>(data.nRow()) * data.nCol());
        for (unsigned i = 0; i < data.nRow(); ++i)
        {
            for (unsigned j = 0; j < data.nCol(); ++j) { rowMajor[static_cast<size_t>(i) * data.nCol() + j] = data(i, j); }
        }
        seederIn(mRandState, mDevState);
        adapterCheck(cgb_sampler_create(&rowMajor[0], data.nRow(), data.nCol(), 0, transpose ? 1 : 0, subsetRows ? 1 : 0, alpha,
                                        maxGibbsMass, &p, mDevState, &mHandle), "cgb_sampler_create");
        seederOut(mRandState, mDevState);
    }
    // the path overload of gaps::run hands the samplers a file name (GapsRunner.cpp:119-159)
    CudaGibbsSampler(const std::string &, bool, bool, float, float, const GapsParameters &params, GapsRandomState *)
        : DenseNormalModel(Matrix(1, 1), false, false, params, 1.f, 1.f), mHandle(NULL), mRandState(NULL), mDevState(NULL),
          mUpdateMode(CGB_UPDATE_EXACT)
    {
        GAPS_ERROR("CudaGibbsSampler: read the file with the reference's loader and pass the matrix (or call cgb_run_file)\n");
    }
    ~CudaGibbsSampler() { cgb_sampler_destroy(mHandle); }

    unsigned nAtoms() const
    {
        uint64_t n = 0;
        adapterCheck(cgb_sampler_n_atoms(mHandle, &n), "cgb_sampler_n_atoms");
        return static_cast<unsigned>(n);
    }
    float getAverageQueueLength() const
    {
        float q = 0.f;
        adapterCheck(cgb_sampler_average_queue_length(mHandle, &q), "cgb_sampler_average_queue_length");
        return q;
    }
    void update(unsigned nSteps, unsigned nThreads)
    {
        seederIn(mRandState, mDevState);
        adapterCheck(cgb_sampler_update(mHandle, nSteps, nThreads), "cgb_sampler_update");
        seederOut(mRandState, mDevState);
        pullMatrix();
    }
    void sync(const CudaGibbsSampler &other, unsigned nThreads = 1)
    {
        (void)nThreads;
        adapterCheck(cgb_sampler_sync(mHandle, other.mHandle), "cgb_sampler_sync");
        if (mUpdateMode != CGB_UPDATE_EXACT) { adapterCheck(cgb_sampler_set_update_mode(mHandle, mUpdateMode), "cgb_sampler_set_update_mode"); }
    }
    void extraInitialization() { adapterCheck(cgb_sampler_extra_initialization(mHandle), "cgb_sampler_extra_initialization"); }
    void setAnnealingTemp(float temp)
    {
        DenseNormalModel::setAnnealingTemp(temp);
        adapterCheck(cgb_sampler_set_annealing_temp(mHandle, temp), "cgb_sampler_set_annealing_temp");
    }
    void setMatrix(const Matrix &mat)
    {
        DenseNormalModel::setMatrix(mat);
        std::vector<float> rowMajor(static_cast<size_t>(mat.nRow()) * mat.nCol());
        for (unsigned i = 0; i < mat.nRow(); ++i)
        {
            for (unsigned j = 0; j < mat.nCol(); ++j) { rowMajor[static_cast<size_t>(i) * mat.nCol() + j] = mat(i, j); }
        }
        adapterCheck(cgb_sampler_set_matrix(mHandle, &rowMajor[0]), "cgb_sampler_set_matrix");
    }
    void setUncertainty(const Matrix &unc, bool transpose, bool subsetRows, const GapsParameters &params)
    {
        DenseNormalModel::setUncertainty(unc, transpose, subsetRows, params); // meanChiSq reads the host copy
        cgb_params p;
        toParams(params, &p);
        std::vector<float> rowMajor(static_cast<size_t>(unc.nRow()) * unc.nCol());
        for (unsigned i = 0; i < unc.nRow(); ++i)
        {
            for (unsigned j = 0; j < unc.nCol(); ++j) { rowMajor[static_cast<size_t>(i) * unc.nCol() + j] = unc(i, j); }
        }
        adapterCheck(cgb_sampler_set_uncertainty(mHandle, &rowMajor[0], unc.nRow(), unc.nCol(), 0, transpose ? 1 : 0,
                                                 subsetRows ? 1 : 0, &p), "cgb_sampler_set_uncertainty");
    }
    void setUncertainty(const std::string &, bool, bool, const GapsParameters &) { GAPS_ERROR("CudaGibbsSampler: uncertainty by file name is not bound\n"); }
    float chiSq() const
    {
        float cs = 0.f;
        adapterCheck(cgb_sampler_chisq(mHandle, &cs), "cgb_sampler_chisq");
        return cs;
    }
    friend Archive& operator<<(Archive &ar, const CudaGibbsSampler &s);
    friend Archive& operator>>(Archive &ar, CudaGibbsSampler &s);

private:
    // device factor matrix -> mMatrix, which GapsStatistics reads
    void pullMatrix()
    {
        const unsigned rows = mMatrix.nRow(), k = mMatrix.nCol();
        std::vector<float> rowMajor(static_cast<size_t>(rows) * k);
        adapterCheck(cgb_sampler_get_matrix(mHandle, &rowMajor[0]), "cgb_sampler_get_matrix");
        for (unsigned i = 0; i < rows; ++i)
        {
            for (unsigned j = 0; j < k; ++j) { mMatrix(i, j) = rowMajor[static_cast<size_t>(i) * k + j]; }
        }
    }

    cgb_sampler *mHandle;
    GapsRandomState *mRandState;
    cgb_randstate *mDevState;
    int mUpdateMode;
};

// `Archive << sampler` (AsynchronousGibbsSampler.h:221-226): cgb_sampler_serialize returns exactly the bytes the reference
// streams for its own sampler (tests/test_checkpoint.py), so they go out one char at a time through the Archive
Archive& operator<<(Archive &ar, const CudaGibbsSampler &s)
{
    uint64_t n = 0;
    adapterCheck(cgb_sampler_serialize(s.mHandle, NULL, 0, &n), "cgb_sampler_serialize");
    std::vector<char> bytes(n);
    adapterCheck(cgb_sampler_serialize(s.mHandle, &bytes[0], n, &n), "cgb_sampler_serialize");
    for (uint64_t i = 0; i < n; ++i) { ar << bytes[i]; }
    return ar;
}

// `Archive >> sampler` (:228-233): the length of a sampler's record follows from its own header fields — the factor
// matrix (nRow, nCol, then per column its length and values), the atomic domain (length, count, 12 bytes per atom) and
// the 77 bytes of the proposal queue — so it is measured on the stream, read in one piece and handed to the library.
// The random state was read just before the samplers (processCheckpoint, GapsRunner.cpp:258-270): bring ours in line.
Archive& operator>>(Archive &ar, CudaGibbsSampler &s)
{
    std::fstream &f = ar.mStream;
    const std::streampos start = f.tellg();
    uint32_t nRow = 0, nCol = 0;
    f.read(reinterpret_cast<char*>(&nRow), 4);
    f.read(reinterpret_cast<char*>(&nCol), 4);
    const uint64_t matrixBytes = 8ull + static_cast<uint64_t>(nCol) * (4ull + 4ull * nRow);
    f.seekg(start + static_cast<std::streamoff>(matrixBytes + 8ull));
    uint64_t nAtoms = 0;
    f.read(reinterpret_cast<char*>(&nAtoms), 8);
    const uint64_t total = matrixBytes + 16ull + 12ull * nAtoms + 77ull;
    std::vector<char> bytes(total);
    f.seekg(start);
    f.read(&bytes[0], static_cast<std::streamsize>(total));
    if (!f) { GAPS_ERROR("CudaGibbsSampler: checkpoint ends inside a sampler\n"); }
    adapterCheck(cgb_sampler_deserialize(s.mHandle, &bytes[0], total), "cgb_sampler_deserialize");
    seederIn(s.mRandState, s.mDevState);
    s.pullMatrix();
    return ar;
}

// gaps::run with the CUDA sampler: what the dispatch next to chooseSampler (GapsRunner.cpp:65-78) would call
static GapsResult gaps_run_cuda(const Matrix &data, GapsParameters &params, const Matrix &uncertainty, GapsRandomState *randState)
{
    if (params.useSparseOptimization) { GAPS_ERROR("CudaGibbsSampler binds the dense model; the sparse model runs through cgb_run\n"); }
    if (params.useCheckPoint)
    {
        // run_helper, GapsRunner.cpp:96-107
        Archive ar(params.checkpointFile, ARCHIVE_READ);
        ar >> params;
        ar >> *randState;
    }
    GapsResult result = runCoGAPSAlgorithm<CudaGibbsSampler>(data, params, uncertainty, randState);
    std::map<const GapsRandomState*, cgb_randstate*>::iterator it = g_states.find(randState);
    if (it != g_states.end())
    {
        cgb_randstate_destroy(it->second);
        g_states.erase(it);
    }
    return result;
}

// the C entry points of ref_driver.cpp (cogaps_ref_run, ...), with the run loop's sampler replaced
#define COGAPS_REF_RUN gaps_run_cuda
#include "ref_driver.cpp"
