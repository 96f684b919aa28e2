// ref_driver.cpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// A C-callable wrapper around the UNMODIFIED reference C++ core, compiled in place from
// /root/reference/src by oracle/Makefile into oracle/_ref/libcogaps_ref_<variant>.so.  It is used
//   (1) to pin the C restatement in oracle/cogaps_oracle.c against the reference itself, and
//   (2) as the CPU baseline (`bench.py --impl reference`, cpu_baseline.kind = "reference").
// Nothing under cogaps_b200/ links or loads it.
//
// Only this translation unit is ours; every other object in the .so is a reference source file
// compiled where it lies.  Private members are reached by re-declaring access in THIS TU only
// (the reference TUs are untouched; access specifiers do not change layout).

#include <cstdio>
#include <cstring>
#include <cmath>
#include <chrono>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include <algorithm>
#include <stdint.h>

#define private public
#define protected public
#include "math/Random.h"
#include "gibbs_sampler/DenseNormalModel.h"
#include "gibbs_sampler/SparseNormalModel.h"
#include "gibbs_sampler/AsynchronousGibbsSampler.h"
#undef private
#undef protected

#include "GapsParameters.h"
#include "GapsResult.h"
#include "GapsRunner.h"
#include "data_structures/Matrix.h"
#include "file_parser/FileParser.h"
#include "utils/GlobalConfig.h"
#include <boost/date_time/posix_time/posix_time.hpp> // the shim: its clock_marks, see cogaps_ref_last_run_seconds

#include "../include/cogaps_b200.h"

// what runs the algorithm: the reference's own gaps::run, or (oracle/cuda_adapter.cpp, which includes this file) the
// reference's run loop instantiated with the sampler that forwards to the C ABI
#ifndef COGAPS_REF_RUN
#define COGAPS_REF_RUN gaps::run
#endif

#ifdef _OPENMP
#include <omp.h>
#endif

static Matrix toMatrix(const float *data, uint32_t nrow, uint32_t ncol)
{
    Matrix m(nrow, ncol);
    for (uint32_t i = 0; i < nrow; ++i)
    {
        for (uint32_t j = 0; j < ncol; ++j)
        {
            m(i, j) = data[static_cast<size_t>(i) * ncol + j];
        }
    }
    return m;
}

static void fromMatrix(const Matrix &m, float *out)
{
    if (out == NULL) { return; }
    for (unsigned i = 0; i < m.nRow(); ++i)
    {
        for (unsigned j = 0; j < m.nCol(); ++j)
        {
            out[static_cast<size_t>(i) * m.nCol() + j] = m(i, j);
        }
    }
}

static double g_lastLoadSeconds = -1.0, g_lastLoopSeconds = -1.0;

extern "C" {

// Phases of the last cogaps_ref_run* call by the reference's own clock readings: everything before the sampler loop
// (Matrix copies, both samplers' constructors, sync, extraInitialization) and the loop itself (both phases) — the
// interval GapsResult::totalRunningTime covers, in whole seconds there.  -1 when unknown.
int cogaps_ref_last_run_seconds(double *load, double *loop)
{
    *load = g_lastLoadSeconds;
    *loop = g_lastLoopSeconds;
    return 0;
}

const char *cogaps_ref_build_report(void)
{
    static std::string s;
    s = buildReport();
    return s.c_str();
}

int cogaps_ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// gaps::run on an in-memory matrix (src/GapsRunner.cpp:113-117), with the reference's own checkpointing
// (createCheckpoint / processCheckpoint, GapsRunner.cpp:224-270): interval 0 = never write one;
// inFile non-empty = resume from that file (params.useCheckPoint, run_helper :96-107).
int cogaps_ref_run_checkpointed(const float *data, uint32_t nrow, uint32_t ncol, const float *uncertainty,
                                const cgb_params *p, cgb_result *r, uint32_t checkpointInterval,
                                const char *checkpointOutFile, const char *checkpointInFile)
{
    Matrix D = toMatrix(data, nrow, ncol);
    Matrix U = (uncertainty != NULL) ? toMatrix(uncertainty, nrow, ncol) : Matrix();

    std::vector<unsigned> idx;
    for (uint32_t i = 0; i < p->nSubsetIndices; ++i) { idx.push_back(p->subsetIndices[i]); }
    GapsParameters params(D, p->transposeData != 0, p->nSubsetIndices > 0, p->subsetGenes != 0, idx);
    params.seed = p->seed;
    params.nPatterns = p->nPatterns;
    params.nIterations = p->nIterations;
    params.maxThreads = p->maxThreads;
    params.outputFrequency = p->outputFrequency;
    params.checkpointInterval = checkpointInterval;
    if (checkpointOutFile != NULL && checkpointOutFile[0] != 0) { params.checkpointOutFile = checkpointOutFile; }
    if (checkpointInFile != NULL && checkpointInFile[0] != 0)
    {
        params.checkpointFile = checkpointInFile;
        params.useCheckPoint = true;
    }
    params.snapshotFrequency = p->snapshotFrequency;
    params.snapshotPhase = static_cast<GapsAlgorithmPhase>(p->snapshotPhase);
    params.alphaA = p->alphaA;
    params.alphaP = p->alphaP;
    params.maxGibbsMassA = p->maxGibbsMassA;
    params.maxGibbsMassP = p->maxGibbsMassP;
    params.useSparseOptimization = p->useSparseOptimization != 0;
    params.asynchronousUpdates = p->asynchronousUpdates != 0;
    params.takePumpSamples = p->takePumpSamples != 0;
    params.printMessages = p->printMessages != 0;
    params.printThreadUsage = false;
    params.workerID = p->workerID;
    params.runningDistributed = p->runningDistributed != 0;
    params.whichMatrixFixed = static_cast<char>(p->whichMatrixFixed);
    if (p->fixedPatterns != NULL && p->whichMatrixFixed != 'N')
    {
        unsigned rows = (p->whichMatrixFixed == 'A') ? params.nGenes : params.nSamples;
        params.fixedPatterns = toMatrix(p->fixedPatterns, rows, p->nPatterns);
        params.useFixedPatterns = true;
    }

    GapsRandomState randState(params.seed);
    boost::posix_time::marks().n = 0;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    GapsResult res = COGAPS_REF_RUN(D, params, U, &randState);
    std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
    {
        // the reference's own clock readings (see ref_shim/.../posix_time.hpp): [0] loading starts, [1] loading done,
        // [2] sampler loop starts (GapsRunner.cpp:450), latest = loop done (:473; a distributed run reads it once more
        // for its "finished" line, a few microseconds later)
        const boost::posix_time::clock_marks &m = boost::posix_time::marks();
        g_lastLoadSeconds = g_lastLoopSeconds = -1.0;
        if (m.n >= 4)
        {
            g_lastLoadSeconds = std::chrono::duration<double>(m.t[2] - t0).count();
            g_lastLoopSeconds = std::chrono::duration<double>(m.t[63] - m.t[2]).count();
        }
    }

    fromMatrix(res.Amean, r->Amean);
    fromMatrix(res.Asd, r->Asd);
    fromMatrix(res.Pmean, r->Pmean);
    fromMatrix(res.Psd, r->Psd);
    if (p->takePumpSamples)
    {
        fromMatrix(res.pumpMatrix, r->pumpMatrix);
        fromMatrix(res.meanPatternAssignment, r->meanPatternAssignment);
    }
    r->nHistory = 0;
    for (size_t i = 0; i < res.chisqHistory.size() && i < r->historyCapacity; ++i)
    {
        if (r->chisqHistory) { r->chisqHistory[i] = res.chisqHistory[i]; }
        if (r->atomHistoryA) { r->atomHistoryA[i] = res.atomHistoryA[i]; }
        if (r->atomHistoryP) { r->atomHistoryP[i] = res.atomHistoryP[i]; }
        r->nHistory = static_cast<uint32_t>(i + 1);
    }
    r->nSnapshotsEquilibration = static_cast<uint32_t>(res.equilibrationSnapshotsA.size());
    r->nSnapshotsSampling = static_cast<uint32_t>(res.samplingSnapshotsA.size());
    uint32_t slot = 0;
    for (int phase = 0; phase < 2; ++phase)
    {
        const std::vector<Matrix> &sa = phase == 0 ? res.equilibrationSnapshotsA : res.samplingSnapshotsA;
        const std::vector<Matrix> &sp = phase == 0 ? res.equilibrationSnapshotsP : res.samplingSnapshotsP;
        for (size_t i = 0; i < sa.size() && slot < r->snapshotCapacity; ++i, ++slot)
        {
            if (r->snapshotsA) { fromMatrix(sa[i], r->snapshotsA + static_cast<size_t>(slot) * sa[i].nRow() * sa[i].nCol()); }
            if (r->snapshotsP) { fromMatrix(sp[i], r->snapshotsP + static_cast<size_t>(slot) * sp[i].nRow() * sp[i].nCol()); }
        }
    }
    r->seed = params.seed;
    r->totalUpdates = res.totalUpdates;
    r->totalRunningTime = std::chrono::duration<double>(t1 - t0).count();
    r->meanChiSq = res.meanChiSq;
    r->averageQueueLengthA = res.averageQueueLengthA;
    r->averageQueueLengthP = res.averageQueueLengthP;
    return 0;
}

int cogaps_ref_run(const float *data, uint32_t nrow, uint32_t ncol, const float *uncertainty,
                   const cgb_params *p, cgb_result *r)
{
    return cogaps_ref_run_checkpointed(data, nrow, ncol, uncertainty, p, r, 0, NULL, NULL);
}

// Matrix(const std::string &path, ...) through the reference's own FileParser (src/data_structures/Matrix.cpp:72-134,
// src/file_parser/*): what the path overload of gaps::run loads.  out: nrow x ncol row-major, or NULL for dimensions.
int cogaps_ref_read_file(const char *path, float *out, uint32_t *nrow, uint32_t *ncol)
{
    Matrix m(std::string(path), false, false, std::vector<unsigned>());
    *nrow = m.nRow();
    *ncol = m.nCol();
    fromMatrix(m, out);
    return 0;
}

// getFileInfo_cpp's three fields through the reference's own FileParser (src/Cogaps.cpp:245-256): names are packed
// NUL-separated into buf; counts[0] = rowNames, counts[1] = colNames
int cogaps_ref_file_info(const char *path, uint32_t *nrow, uint32_t *ncol, char *buf, uint64_t capacity, uint32_t *counts)
{
    FileParser fp((std::string(path)));
    *nrow = fp.nRow();
    *ncol = fp.nCol();
    std::vector<std::string> rows = fp.rowNames(), cols = fp.colNames();
    counts[0] = static_cast<uint32_t>(rows.size());
    counts[1] = static_cast<uint32_t>(cols.size());
    uint64_t used = 0;
    for (size_t pass = 0; pass < 2; ++pass)
    {
        const std::vector<std::string> &v = pass == 0 ? rows : cols;
        for (size_t i = 0; i < v.size(); ++i)
        {
            if (used + v[i].size() + 1 > capacity) { return -1; }
            std::memcpy(buf + used, v[i].c_str(), v[i].size() + 1);
            used += v[i].size() + 1;
        }
    }
    return 0;
}

// FileParser::writeToCsv through the reference itself (src/file_parser/FileParser.h:59-89)
int cogaps_ref_write_csv(const char *path, const float *data, uint32_t nrow, uint32_t ncol)
{
    FileParser::writeToCsv(std::string(path), toMatrix(data, nrow, ncol));
    return 0;
}

// The three lookup tables of GapsRandomState (src/math/Random.cpp:269-295) as this build makes them.
int cogaps_ref_tables(float *erf, float *erfinv, float *qgamma)
{
    GapsRandomState rs(1);
    std::memcpy(erf, rs.mErfLookupTable, sizeof(float) * ERF_LOOKUP_TABLE_SIZE);
    std::memcpy(erfinv, rs.mErfinvLookupTable, sizeof(float) * ERF_INV_LOOKUP_TABLE_SIZE);
    std::memcpy(qgamma, rs.mQgammaLookupTable, sizeof(float) * Q_GAMMA_LOOKUP_TABLE_SIZE);
    return 0;
}

// Known-answer streams from GapsRandomState/GapsRng (src/math/Random.cpp:32-200,216-260).
// kind: 0 nextSeed (u64), 1 uniform32, 2 uniform32(a,b), 3 uniform64(a,b), 4 uniform() bits,
//       5 poisson(lambda), 6 exponential(lambda) bits, 7 truncNormal(a,b,mean,sd) bits (NaN bits = none),
//       8 truncGammaUpper(b, scale) bits
int cogaps_ref_rng_stream(uint32_t seed, int kind, uint32_t n, uint64_t a, uint64_t b,
                          double lambda, float f0, float f1, float f2, float f3, uint64_t *out)
{
    GapsRandomState rs(seed);
    if (kind == 0)
    {
        for (uint32_t i = 0; i < n; ++i) { out[i] = rs.nextSeed(); }
        return 0;
    }
    GapsRng rng(&rs);
    for (uint32_t i = 0; i < n; ++i)
    {
        float f = 0.f;
        uint32_t bits = 0;
        switch (kind)
        {
            case 1: out[i] = rng.uniform32(); break;
            case 2: out[i] = rng.uniform32(static_cast<uint32_t>(a), static_cast<uint32_t>(b)); break;
            case 3: out[i] = rng.uniform64(a, b); break;
            case 4: f = rng.uniform(); std::memcpy(&bits, &f, 4); out[i] = bits; break;
            case 5: out[i] = static_cast<uint64_t>(static_cast<int64_t>(rng.poisson(lambda))); break;
            case 6: f = rng.exponential(f0); std::memcpy(&bits, &f, 4); out[i] = bits; break;
            case 7:
            {
                OptionalFloat o = rng.truncNormal(f0, f1, f2, f3);
                f = o.value();
                std::memcpy(&bits, &f, 4);
                out[i] = o.hasValue() ? bits : 0xFFFFFFFFFFFFFFFFull;
                break;
            }
            case 8: f = rng.truncGammaUpper(f0, f1); std::memcpy(&bits, &f, 4); out[i] = bits; break;
            default: return -1;
        }
    }
    return 0;
}

// Lock-step probe of the PERFORMANCE CRITICAL scans (src/gibbs_sampler/DenseNormalModel.cpp:162-240)
// on a model whose factor matrices are given: `data` is nGenes x nSamples; the probed model is the
// A-side one (rows = genes, row length = nSamples) with matrix = A (nGenes x k) and other = P
// (nSamples x k); AP is rebuilt by extraInitialization.  variant as in cgb_sampler_alpha_parameters.
int cogaps_ref_alpha_parameters(const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                                const float *A, const float *P, const float *uncertainty,
                                uint32_t n, const int32_t *variant, const uint32_t *r1,
                                const uint32_t *c1, const uint32_t *r2, const uint32_t *c2,
                                const float *ch, float *s_out, float *smu_out, float *ap_out)
{
    Matrix D = toMatrix(data, nGenes, nSamples);
    GapsParameters params(D);
    params.nPatterns = k;
    params.printMessages = false;
    // A-side model: transpose = true (GapsRunner.cpp:402), P-side: transpose = false (:405)
    DenseNormalModel AModel(D, true, true, params, params.alphaA, params.maxGibbsMassA);
    DenseNormalModel PModel(D, false, false, params, params.alphaP, params.maxGibbsMassP);
    if (uncertainty != NULL)
    {
        Matrix U = toMatrix(uncertainty, nGenes, nSamples);
        AModel.setUncertainty(U, true, true, params);
        PModel.setUncertainty(U, false, false, params);
    }
    AModel.setMatrix(toMatrix(A, nGenes, k));
    PModel.setMatrix(toMatrix(P, nSamples, k));
    AModel.sync(PModel);
    PModel.sync(AModel);
    AModel.extraInitialization();
    PModel.extraInitialization();
    for (uint32_t i = 0; i < n; ++i)
    {
        AlphaParameters a(0.f, 0.f);
        switch (variant[i])
        {
            case 0: a = AModel.alphaParameters(r1[i], c1[i]); break;
            case 1: a = AModel.alphaParameters(r1[i], c1[i], r2[i], c2[i]); break;
            case 2: a = AModel.alphaParametersWithChange(r1[i], c1[i], ch[i]); break;
            default: return -1;
        }
        s_out[i] = a.s;
        smu_out[i] = a.s_mu;
    }
    if (ap_out != NULL)
    {
        for (uint32_t g = 0; g < nGenes; ++g)
        {
            for (uint32_t s = 0; s < nSamples; ++s)
            {
                ap_out[static_cast<size_t>(g) * nSamples + s] = AModel.mAPMatrix(s, g);
            }
        }
    }
    return 0;
}

// DenseNormalModel::chiSq (src/gibbs_sampler/DenseNormalModel.cpp:56-68) of both models for given
// factor matrices; out[0] = A-side model, out[1] = P-side model.
int cogaps_ref_chisq(const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                     const float *A, const float *P, const float *uncertainty, float *out)
{
    Matrix D = toMatrix(data, nGenes, nSamples);
    GapsParameters params(D);
    params.nPatterns = k;
    params.printMessages = false;
    DenseNormalModel AModel(D, true, true, params, params.alphaA, params.maxGibbsMassA);
    DenseNormalModel PModel(D, false, false, params, params.alphaP, params.maxGibbsMassP);
    if (uncertainty != NULL)
    {
        Matrix U = toMatrix(uncertainty, nGenes, nSamples);
        AModel.setUncertainty(U, true, true, params);
        PModel.setUncertainty(U, false, false, params);
    }
    AModel.setMatrix(toMatrix(A, nGenes, k));
    PModel.setMatrix(toMatrix(P, nSamples, k));
    AModel.sync(PModel);
    PModel.sync(AModel);
    AModel.extraInitialization();
    PModel.extraInitialization();
    out[0] = AModel.chiSq();
    out[1] = PModel.chiSq();
    return 0;
}

// The same two probes on SparseNormalModel (src/gibbs_sampler/SparseNormalModel.cpp:39-60,153-311): the three
// alphaParameters variants of the A-side model, and chiSq of both models, for given factor matrices.
int cogaps_ref_alpha_parameters_sparse(const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                                       const float *A, const float *P,
                                       uint32_t n, const int32_t *variant, const uint32_t *r1,
                                       const uint32_t *c1, const uint32_t *r2, const uint32_t *c2,
                                       const float *ch, float *s_out, float *smu_out)
{
    Matrix D = toMatrix(data, nGenes, nSamples);
    GapsParameters params(D);
    params.nPatterns = k;
    params.printMessages = false;
    params.useSparseOptimization = true;
    SparseNormalModel AModel(D, true, true, params, params.alphaA, params.maxGibbsMassA);
    SparseNormalModel PModel(D, false, false, params, params.alphaP, params.maxGibbsMassP);
    AModel.setMatrix(toMatrix(A, nGenes, k));
    PModel.setMatrix(toMatrix(P, nSamples, k));
    AModel.sync(PModel);
    PModel.sync(AModel);
    for (uint32_t i = 0; i < n; ++i)
    {
        AlphaParameters a(0.f, 0.f);
        switch (variant[i])
        {
            case 0: a = AModel.alphaParameters(r1[i], c1[i]); break;
            case 1: a = AModel.alphaParameters(r1[i], c1[i], r2[i], c2[i]); break;
            case 2: a = AModel.alphaParametersWithChange(r1[i], c1[i], ch[i]); break;
            default: return -1;
        }
        s_out[i] = a.s;
        smu_out[i] = a.s_mu;
    }
    return 0;
}

int cogaps_ref_chisq_sparse(const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                            const float *A, const float *P, float *out)
{
    Matrix D = toMatrix(data, nGenes, nSamples);
    GapsParameters params(D);
    params.nPatterns = k;
    params.printMessages = false;
    params.useSparseOptimization = true;
    SparseNormalModel AModel(D, true, true, params, params.alphaA, params.maxGibbsMassA);
    SparseNormalModel PModel(D, false, false, params, params.alphaP, params.maxGibbsMassP);
    AModel.setMatrix(toMatrix(A, nGenes, k));
    PModel.setMatrix(toMatrix(P, nSamples, k));
    AModel.sync(PModel);
    PModel.sync(AModel);
    out[0] = AModel.chiSq();
    out[1] = PModel.chiSq();
    out[2] = PModel.dataSparsity();
    return 0;
}

} // extern "C"
