/*
 * cogaps_b200.h — C ABI of the B200-native CoGAPS hot path.
 *
 * This is the drop-in boundary for ONE path of FertigLab/CoGAPS: the atomic-domain Gibbs
 * sampler that alternately updates the A and P factor matrices.  Every entry point below
 * replaces one member of the reference's implicit "Sampler concept" (what
 * runCoGAPSAlgorithm<Sampler> and GapsStatistics call) or the run loop itself; the
 * reference interface each one stands in for is cited as file:line relative to the
 * reference tree.  INTEGRATION.md shows the C++ adapter a maintainer would add next to
 * chooseSampler (src/GapsRunner.cpp:65-78) to bind these.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, opaque handles, no C++/torch types.
 *   - Every call returns an int status: CGB_OK (0) or a negative CGB_E* code;
 *     cgb_last_error() returns a thread-local message for the last failure.
 *     (The reference prints and calls Rcpp::stop / exit — utils/GapsAssert.h:18-25; the
 *     C++ adapter turns a non-zero status back into GAPS_ERROR.)
 *   - Host pointers are borrowed for the duration of the call only; the library owns all
 *     device memory.  Matrices cross the boundary as fp32 row-major (element (i,j) at
 *     ptr[i*ncol + j]) unless a `colmajor` flag says otherwise.
 *   - A handle is not thread-safe; one host thread drives one handle (the reference calls
 *     the sampler from one thread and fans out with OpenMP inside — here the fan-out is
 *     the GPU).
 *   - There is no CPU fallback: if no sm_100-class device is usable the calls that need
 *     one fail with CGB_ENODEVICE.
 */
#ifndef COGAPS_B200_H
#define COGAPS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGB_OK            0
#define CGB_EINVAL       -1   /* bad argument */
#define CGB_ENODEVICE    -2   /* no usable CUDA device */
#define CGB_ECUDA        -3   /* CUDA runtime / kernel failure */
#define CGB_ENOMEM       -4
#define CGB_EUNSUPPORTED -5   /* feature of the reference not on this path yet */
#define CGB_EINTERNAL    -6
#define CGB_EINTERRUPTED -7   /* cgb_run_ex: the interrupt callback asked to stop (Rcpp::checkUserInterrupt, utils/GlobalConfig.h:16-20) */

#define CGB_ERF_TABLE_SIZE     3001  /* math/Random.h:13 */
#define CGB_ERFINV_TABLE_SIZE  5001  /* math/Random.h:14 */
#define CGB_QGAMMA_TABLE_SIZE  5001  /* math/Random.h:15 */

#define CGB_UPDATE_EXACT 0
#define CGB_UPDATE_SWEEP 1

#define CGB_PHASE_EQUILIBRATION 1    /* GapsParameters.h:19-24 */
#define CGB_PHASE_SAMPLING      2
#define CGB_PHASE_ALL           3

const char *cgb_last_error(void);
/* library / kernel build facts: "sm_100a;threads=256;..." (utils/GlobalConfig.h:26-54 buildReport) */
const char *cgb_build_report(void);
/* select the CUDA device used by handles created afterwards (default: COGAPS_DEVICE env or 0) */
int cgb_set_device(int device);
/* Several chains on one device, one host thread each (what distributed CoGAPS does with its worker processes,
 * R/DistributedCogaps.R:60-68): handles are independent, but every sampler's resident grid must fit on the device
 * beside the others.  Call with the number of chains that will run at once BEFORE creating their samplers; each
 * resident grid then takes 1/parts of the device.  Default 1. */
int cgb_set_resident_share(int32_t parts);
/* (the per-handle form, cgb_sampler_set_resident_share below, needs no process-wide setting) */
/* Matrix-sized device buffers of destroyed samplers are kept for the next sampler of the same shape (a cudaFree /
 * cudaMalloc pair of 400 MB costs up to a second on some calls); this hands them back to the device.  The cache holds
 * at most COGAPS_DEVICE_CACHE_MB (environment, default 16384; 0 disables it). */
int cgb_release_device_cache(void);
/* number of kernel launches issued by this library since process start (bench.py gpu_launches) */
uint64_t cgb_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Run parameters — mirrors struct GapsParameters (src/GapsParameters.h:25-70, defaults :79-114)
 * ---------------------------------------------------------------------------------------- */
typedef struct cgb_params
{
    uint32_t struct_size;          /* = sizeof(cgb_params); ABI guard */
    uint32_t seed;
    uint32_t nPatterns;            /* default 3 */
    uint32_t nIterations;          /* per phase, default 1000 */
    uint32_t maxThreads;           /* accepted, unused on the device path */
    uint32_t outputFrequency;      /* default 500; 0 = never */
    uint32_t snapshotFrequency;    /* default 0 */
    uint32_t snapshotPhase;        /* CGB_PHASE_* ; default ALL */
    float alphaA;                  /* 0.01 */
    float alphaP;                  /* 0.01 */
    float maxGibbsMassA;           /* 100 */
    float maxGibbsMassP;           /* 100 */
    int32_t transposeData;
    int32_t useSparseOptimization; /* SparseNormalModel (gibbs_sampler/SparseNormalModel.h); uncertainty is ignored */
    int32_t asynchronousUpdates;   /* 1: AsynchronousGibbsSampler; 0: SingleThreadedGibbsSampler (one proposal at a time) */
    int32_t takePumpSamples;
    int32_t printMessages;
    int32_t whichMatrixFixed;      /* 'N', 'A' or 'P' */
    int32_t subsetGenes;           /* meaning of subsetIndices: 1 = genes, 0 = samples */
    uint32_t nSubsetIndices;       /* 0 = no subsetting */
    const uint32_t *subsetIndices; /* 1-based, like R (data_structures/Matrix.cpp:30-69) */
    const float *fixedPatterns;    /* rows(of the fixed matrix) x nPatterns row-major, or NULL */
    uint32_t workerID;             /* default 1 */
    int32_t runningDistributed;
    int32_t updateMode;            /* CGB_UPDATE_EXACT (default): the reference's chain, proposal for proposal.
                                    * CGB_UPDATE_SWEEP: row-parallel sweep (dense model) — every row of the factor runs
                                    * the same four proposal types on its own segment of the atomic domain, all rows at
                                    * once, draws from Philox4x32-10.  A different chain for the same seed, validated
                                    * statistically against the reference and bit for bit against oracle/ (see
                                    * cgb_sampler_set_update_mode). */
} cgb_params;

/* fill *p with the reference defaults (GapsParameters.h:79-114) */
void cgb_params_default(cgb_params *p);

/* ------------------------------------------------------------------------------------------
 * Run result — mirrors struct GapsResult (src/GapsResult.h:11-36).  All arrays are
 * caller-allocated; capacities are given, counts are written back.
 * ---------------------------------------------------------------------------------------- */
typedef struct cgb_result
{
    uint32_t struct_size;
    uint32_t historyCapacity;      /* in: room in the three history arrays */
    float *Amean;                  /* nGenes   x nPatterns row-major */
    float *Asd;
    float *Pmean;                  /* nSamples x nPatterns row-major */
    float *Psd;
    float *chisqHistory;           /* one entry per outputFrequency iterations, both phases */
    uint32_t *atomHistoryA;
    uint32_t *atomHistoryP;
    float *pumpMatrix;             /* optional (NULL ok): nGenes x nPatterns */
    float *meanPatternAssignment;  /* optional */
    float *snapshotsA;             /* optional: snapshotCapacity x nGenes x nPatterns, equilibration then sampling */
    float *snapshotsP;             /* optional: snapshotCapacity x nSamples x nPatterns */
    uint32_t snapshotCapacity;
    uint32_t nSnapshotsEquilibration; /* out */
    uint32_t nSnapshotsSampling;      /* out */
    uint32_t nHistory;             /* out */
    uint32_t seed;                 /* out */
    uint64_t totalUpdates;         /* out */
    double totalRunningTime;       /* out, seconds in the sampler loop (reference: whole seconds) */
    float meanChiSq;               /* out */
    float averageQueueLengthA;     /* out */
    float averageQueueLengthP;     /* out */
    /* extras the reference does not report; used by bench.py */
    uint64_t nBatchesA;            /* out: proposal batches evaluated (= eval-kernel launches) */
    uint64_t nBatchesP;
    double secondsUpdateA;         /* out: host wall time inside A.update / P.update */
    double secondsUpdateP;
    double secondsDevice;          /* out: CUDA-event time of all eval kernels */
    double algorithmicBytes;       /* out: SURVEY 8(d) reference-formulation bytes of all evaluated proposals */
} cgb_result;

/* ------------------------------------------------------------------------------------------
 * gaps::run (src/GapsRunner.h:14-24, src/GapsRunner.cpp:113-117,381-503)
 * data: nrow x ncol fp32 (row-major unless colmajor != 0); uncertainty: same shape or NULL.
 * ---------------------------------------------------------------------------------------- */
int cgb_run(const float *data, uint32_t nrow, uint32_t ncol, int32_t colmajor,
            const float *uncertainty, const cgb_params *params, cgb_result *result);

/* The path overload of gaps::run (src/GapsRunner.h:19-24, GapsRunner.cpp:119-159): data / uncertainty read from
 * .mtx, .csv, .tsv or .gct files with the reference's parsing rules (src/file_parser/ — FileParser.cpp:69-86,
 * MtxParser.cpp, CharacterDelimitedParser.cpp, MatrixElement.cpp).  uncertaintyPath may be NULL or "".  Subset
 * indices are applied in increasing order, as the reference does when it reads a subset from a file
 * (data_structures/Matrix.cpp:113-131).  Size the result arrays with cgb_read_matrix_file(path, NULL, 0, &nrow, &ncol). */
int cgb_run_file(const char *dataPath, const char *uncertaintyPath, const cgb_params *params, cgb_result *result);
/* Reads a matrix file into out (row-major, nrow*ncol floats; NULL: dimensions only).  Host only, no device needed. */
int cgb_read_matrix_file(const char *path, float *out, uint64_t capacity, uint32_t *nrow, uint32_t *ncol);

/* getFileInfo_cpp (src/Cogaps.cpp:245-256) = cgb_read_matrix_file(path, NULL, 0, &nrow, &ncol) for the dimensions plus
 * this for colNames: the header cells of a .csv / .tsv file as FileParser::colNames returns them, each followed by a NUL,
 * packed into buf (size it with buf == NULL).  .mtx and .gct files have none, and rowNames are empty for every format in
 * the reference (no parser fills them), so there is nothing to export for those.  Host only. */
int cgb_file_col_names(const char *path, char *buf, uint64_t capacity, uint64_t *needed, uint32_t *count);

/* FileParser::writeToCsv (file_parser/FileParser.h:59-89) and GapsResult::writeToFile (GapsResult.cpp:27-35): a matrix
 * (row-major) as the reference writes it — "" and "Col<j>" headers, "Row<i>" names, values in the default stream format
 * (%g, six significant digits) — and the four result matrices to <prefix>_<nPatterns>_{Amean,Pmean,Asd,Psd}.csv.  Host only. */
int cgb_write_matrix_csv(const char *path, const float *mat, uint32_t nrow, uint32_t ncol);
int cgb_result_write_files(const char *pathPrefix, uint32_t nGenes, uint32_t nSamples, uint32_t nPatterns, const cgb_result *result);

/* The compressed rows (byRows != 0) or compressed columns of a Matrix-Market file exactly as cgb_run_file hands them to
 * the sparse model's two samplers (SparseMatrix(path, ...), data_structures/SparseMatrix.cpp:52-107; SparseVector keeps
 * the positive entries in ascending index order, SparseVector.cpp:20-35): with useSparseOptimization, a whole-matrix
 * .mtx input never exists as a dense matrix on the host.  A later entry for a cell overwrites an earlier one, like the
 * dense reader.  ptr gets nMajor + 1 offsets; idx / val are filled when non-NULL (size them with a first call).  A file
 * with a negative value is refused with CGB_EUNSUPPORTED (cgb_run_file then takes the dense route).  Host only. */
int cgb_read_matrix_csr(const char *path, int32_t byRows, uint32_t *nrow, uint32_t *ncol, uint32_t *ptr,
                        uint64_t ptrCapacity, uint32_t *idx, float *val, uint64_t capacity, uint64_t *nnz);

/* ------------------------------------------------------------------------------------------
 * GapsRandomState (src/math/Random.h:79-98): the xoroshiro128+ seeder every sampler and every
 * proposal pulls its PCG seed from, plus the three lookup tables.
 * ---------------------------------------------------------------------------------------- */
typedef struct cgb_randstate cgb_randstate;
int cgb_randstate_create(uint32_t seed, cgb_randstate **out);          /* Random.cpp:264-267 */
/* Replace the built-in tables (a host that links Boost passes the reference's own:
 * Random.cpp:269-295).  Must be called before samplers are created. */
int cgb_randstate_set_tables(cgb_randstate *rs, const float *erf, const float *erfinv,
                             const float *qgamma);
int cgb_randstate_get_tables(const cgb_randstate *rs, float *erf, float *erfinv, float *qgamma);
int cgb_randstate_next_seed(cgb_randstate *rs, uint64_t *out);         /* Random.cpp:297-300 */
void cgb_randstate_destroy(cgb_randstate *rs);

/* A GapsRng stream (Random.cpp:32-200), host side; used by the run loop for the Poisson
 * update counts (GapsRunner.cpp:294-295) and exposed for known-answer tests. */
typedef struct cgb_rng cgb_rng;
int cgb_rng_create(cgb_randstate *rs, cgb_rng **out);
int cgb_rng_uniform32(cgb_rng *r, uint32_t *out);
int cgb_rng_uniform32_range(cgb_rng *r, uint32_t a, uint32_t b, uint32_t *out);
int cgb_rng_uniform64_range(cgb_rng *r, uint64_t a, uint64_t b, uint64_t *out);
int cgb_rng_uniform(cgb_rng *r, float *out);
int cgb_rng_poisson(cgb_rng *r, double lambda, int32_t *out);
int cgb_rng_exponential(cgb_rng *r, float lambda, float *out);
int cgb_rng_trunc_normal(cgb_rng *r, float a, float b, float mean, float sd, float *out,
                         int32_t *has_value);
int cgb_rng_trunc_gamma_upper(cgb_rng *r, float b, float scale, float *out);
void cgb_rng_destroy(cgb_rng *r);

/* ------------------------------------------------------------------------------------------
 * The Sampler concept: AsynchronousGibbsSampler<DenseNormalModel>
 * (src/gibbs_sampler/AsynchronousGibbsSampler.h:31-56, DenseNormalModel.h:15-64).
 * One handle per factor matrix, exactly as the reference builds ASampler and PSampler
 * (GapsRunner.cpp:402-406).
 * ---------------------------------------------------------------------------------------- */
typedef struct cgb_sampler cgb_sampler;

/* ctor (AsynchronousGibbsSampler.h:63-76 + DenseNormalModel.h:66-88).  `transpose` and
 * `subsetRows` have the reference's meaning (Matrix.cpp:30-69: genesInCols / subsetGenes);
 * the A sampler is created with transpose = !transposeData, the P sampler with transposeData. */
int cgb_sampler_create(const float *data, uint32_t nrow, uint32_t ncol, int32_t colmajor,
                       int32_t transpose, int32_t subsetRows, float alpha, float maxGibbsMass,
                       const cgb_params *params, cgb_randstate *rs, cgb_sampler **out);
void cgb_sampler_destroy(cgb_sampler *s);
/* setUncertainty (DenseNormalModel.h:90-95) */
int cgb_sampler_set_uncertainty(cgb_sampler *s, const float *unc, uint32_t nrow, uint32_t ncol,
                                int32_t colmajor, int32_t transpose, int32_t subsetRows,
                                const cgb_params *params);
/* setMatrix (DenseNormalModel.cpp:9-12): rows x nPatterns row-major */
int cgb_sampler_set_matrix(cgb_sampler *s, const float *mat);
/* setAnnealingTemp (DenseNormalModel.cpp:14-17) */
int cgb_sampler_set_annealing_temp(cgb_sampler *s, float temp);
/* sync (DenseNormalModel.cpp:20-36): AP <- transpose(other.AP), repoint the other factor */
int cgb_sampler_sync(cgb_sampler *s, const cgb_sampler *other);
/* extraInitialization (DenseNormalModel.cpp:38-54): AP <- Other * Matrix^T */
int cgb_sampler_extra_initialization(cgb_sampler *s);
/* update (AsynchronousGibbsSampler.h:88-122): nSteps birth/death/move/exchange proposals */
int cgb_sampler_update(cgb_sampler *s, uint32_t nSteps, uint32_t nThreads);
/* chiSq (DenseNormalModel.cpp:56-68) */
int cgb_sampler_chisq(const cgb_sampler *s, float *out);
/* nAtoms (AsynchronousGibbsSampler.h:78-82) */
int cgb_sampler_n_atoms(const cgb_sampler *s, uint64_t *out);
/* dataSparsity (DenseNormalModel.cpp:70-73) */
int cgb_sampler_data_sparsity(const cgb_sampler *s, float *out);
/* getAverageQueueLength (AsynchronousGibbsSampler.h:84-88) */
int cgb_sampler_average_queue_length(const cgb_sampler *s, float *out);
/* mMatrix.getMatrix() (read by GapsStatistics.h:129-202): rows x nPatterns row-major */
int cgb_sampler_get_matrix(const cgb_sampler *s, float *out);
/* shape of the factor matrix this sampler owns */
int cgb_sampler_shape(const cgb_sampler *s, uint32_t *rows, uint32_t *nPatterns, uint32_t *rowLength);
/* lambda / maxGibbsMass after the ctor's rescaling (DenseNormalModel.h:79-81) */
int cgb_sampler_lambda(const cgb_sampler *s, float *lambda, float *maxGibbsMass);
/* atomic domain contents (operator<< of ConcurrentAtomicDomain, ConcurrentAtomicDomain.cpp:133-141):
 * atoms in the reference's vector order; pass NULLs to query the count */
int cgb_sampler_get_atoms(const cgb_sampler *s, uint64_t *pos, float *mass, uint64_t capacity,
                          uint64_t *count);
/* cached AP row (debug / tests): row of the sampler's AP matrix, rowLength floats */
int cgb_sampler_get_ap_row(const cgb_sampler *s, uint32_t row, float *out);

/* Lock-step probes of the three scan variants (DenseNormalModel.cpp:162-240), evaluated by the
 * same device code the update uses.  n queries; r2/c2 may equal r1/c1.  variant: 0 =
 * alphaParameters(r1,c1); 1 = alphaParameters(r1,c1,r2,c2); 2 = alphaParametersWithChange(r1,c1,ch). */
int cgb_sampler_alpha_parameters(cgb_sampler *s, uint32_t n, const int32_t *variant,
                                 const uint32_t *r1, const uint32_t *c1, const uint32_t *r2,
                                 const uint32_t *c2, const float *ch, float *s_out, float *smu_out);

/* Per-sampler counters accumulated over update() calls (bench.py). */
typedef struct cgb_sampler_counters
{
    uint64_t nBatches;          /* conflict-free batches evaluated (= eval-kernel launches in launch-per-batch mode) */
    uint64_t nProposalsQueued;  /* proposals evaluated on the device */
    uint64_t nProposalsTotal;   /* nSteps processed (includes same-bin moves/exchanges) */
    double algorithmicBytes;    /* SURVEY 8(d) bytes of the queued proposals */
    double secondsHostGenerate; /* wall time in proposal generation + bookkeeping */
    double secondsDeviceWait;   /* wall time launching + waiting for the device */
    double secondsKernel;       /* CUDA-event time of eval kernels: every resident-grid launch; per-batch launches
                                   only when timing is enabled */
} cgb_sampler_counters;
int cgb_sampler_get_counters(const cgb_sampler *s, cgb_sampler_counters *out);
int cgb_sampler_reset_counters(cgb_sampler *s);
/* enable CUDA-event timing of every eval-kernel launch (costs a sync each; bench roofline leg) */
int cgb_sampler_set_kernel_timing(cgb_sampler *s, int32_t enabled);
/* How update() reaches the device.  1 (default, or COGAPS_PERSISTENT=1): one resident grid per update(); every
 * proposal is written to a pinned-memory record ring the moment it is generated and evaluated while the rest of
 * its batch is still being generated.  0: one eval-kernel launch per conflict-free batch.  Both run the same
 * device code and give bit-identical chains. */
int cgb_sampler_set_persistent(cgb_sampler *s, int32_t enabled);
/* This sampler's resident grid takes 1/parts of the device (several chains on one GPU, one host thread each).  Per
 * handle, no process-wide state; call before the sampler's first update().  Default: cgb_set_resident_share's value. */
int cgb_sampler_set_resident_share(cgb_sampler *s, int32_t parts);

/* The device reduction order of the scan, so a checker can reproduce it bit-for-bit:
 * a row of length L is cut into `nSegments` contiguous segments of `segmentLength` floats;
 * within a segment element e belongs to lane ((e / vectorWidth) % threadsPerSegment); lanes
 * sum their elements in increasing index order; 32 consecutive lanes (a warp) combine by an xor
 * butterfly (offsets 16,8,4,2,1); the threadsPerSegment/32 warp totals, zero-padded to 32 lanes,
 * combine by the same butterfly; segment totals add in increasing order. */
typedef struct cgb_reduction_order
{
    uint32_t threadsPerSegment;
    uint32_t vectorWidth;
    uint32_t nSegments;
    uint32_t segmentLength;
} cgb_reduction_order;
int cgb_sampler_reduction_order(const cgb_sampler *s, cgb_reduction_order *out);
/* the same, for any sampler whose rows have this length (no device needed) */
int cgb_reduction_order_for_length(uint32_t rowLength, cgb_reduction_order *out);

/* ------------------------------------------------------------------------------------------
 * Row-parallel sweep (no counterpart in the reference; BASELINE north star: "a conflict-partitioned commit step — atoms
 * binned by row so non-conflicting updates apply in parallel within one sweep", device-side counter-based draws).
 * The loop it stands in for is AsynchronousGibbsSampler::update (AsynchronousGibbsSampler.h:88-122).  Row r of the factor
 * matrix owns the segment [r*k*binLength, (r+1)*k*binLength) of the atomic domain (ProposalQueue.cpp:172-173) and, given
 * the other factor, the rows are independent; in this mode every row runs the reference's four proposal types on its own
 * segment — same evaluation code — one CTA per row, the whole update() in one launch, the row's D / AP lines staged once
 * in shared memory.  Dense model only.  A different chain from the reference's for the same seed: validated against it
 * statistically, and bit for bit against oracle/cogaps_oracle.c (sweep_row).  Switching converts the atoms between the
 * host's atomic domain and the device's per-row store; cgb_sampler_update then takes the mode's path.
 * ---------------------------------------------------------------------------------------- */
int cgb_sampler_set_update_mode(cgb_sampler *s, int32_t mode);   /* CGB_UPDATE_EXACT / CGB_UPDATE_SWEEP */
/* reduction order of the sweep's scans (one segment per row, thread count chosen from the row length; no device needed) */
int cgb_sweep_reduction_order_for_length(uint32_t rowLength, cgb_reduction_order *out);

/* ------------------------------------------------------------------------------------------
 * GapsStatistics (src/GapsStatistics.h:17-64): running sums for Amean/Asd/Pmean/Psd kept on
 * the device next to the factor matrices.
 * ---------------------------------------------------------------------------------------- */
typedef struct cgb_stats cgb_stats;
int cgb_stats_create(uint32_t nGenes, uint32_t nSamples, uint32_t nPatterns, cgb_stats **out);
void cgb_stats_destroy(cgb_stats *st);
int cgb_stats_update(cgb_stats *st, const cgb_sampler *A, const cgb_sampler *P);   /* GapsStatistics.h:129-149 */
int cgb_stats_update_a(cgb_stats *st, const cgb_sampler *A, const cgb_sampler *P); /* :151-167 */
int cgb_stats_update_p(cgb_stats *st, const cgb_sampler *A, const cgb_sampler *P); /* :169-185 */
int cgb_stats_update_pump(cgb_stats *st, const cgb_sampler *A);                    /* :119-127 */
int cgb_stats_amean(const cgb_stats *st, float *out);   /* GapsStatistics.cpp:13-21 */
int cgb_stats_asd(const cgb_stats *st, float *out);     /* :23-36 */
int cgb_stats_pmean(const cgb_stats *st, float *out);   /* :38-46 */
int cgb_stats_psd(const cgb_stats *st, float *out);     /* :48-61 */
int cgb_stats_pump_matrix(const cgb_stats *st, float *out);          /* :113-117 */
int cgb_stats_mean_pattern(const cgb_stats *st, float *out);         /* :119-131 */
/* meanChiSq (GapsStatistics.cpp:63-87) against the P sampler's D and S */
int cgb_stats_mean_chisq(const cgb_stats *st, const cgb_sampler *P, float *out);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, each running an independent chain on its own
 * subset (R/DistributedCogaps.R:48-119).  The only exchange is the concatenation of per-shard
 * factor rows (stitchTogether, R/DistributedCogaps.R:226-278), done by the caller with NCCL on
 * device buffers obtained here.
 * ---------------------------------------------------------------------------------------- */
/* device pointer + stride of the sampler's factor matrix, column (pattern) major:
 * element (row r, pattern p) at dev[p * ld + r] */
int cgb_sampler_device_matrix(const cgb_sampler *s, void **dev, uint64_t *ld);
/* device pointers of the statistics sums (same layout) */
int cgb_stats_device_sums(const cgb_stats *st, void **AmeanSum, void **AsqSum, void **PmeanSum,
                          void **PsqSum, uint64_t *ldA, uint64_t *ldP, uint32_t *nUpdates);

/* The exchange itself, for a C / C++ caller (the reference is C++ and has no torch): an NCCL communicator over the
 * ranks of the job and the all-gather of per-shard factor rows — what stitchTogether's rbind does on the host
 * (R/DistributedCogaps.R:251-272) — straight from device memory.  NCCL is bound at run time (libnccl.so.2, or the
 * library named by COGAPS_NCCL_LIB); without one these calls fail with CGB_EUNSUPPORTED.  Rank 0 makes the id
 * (ncclGetUniqueId) and hands it to the other ranks by whatever means the launcher has (MPI, a file, a socket);
 * every rank then calls cgb_comm_init on the device it selected with cgb_set_device. */
#define CGB_COMM_UNIQUE_ID_BYTES 128
typedef struct cgb_comm cgb_comm;
int cgb_comm_get_unique_id(uint8_t *id /* [CGB_COMM_UNIQUE_ID_BYTES] */);
int cgb_comm_init(const uint8_t *id, int32_t rank, int32_t nRanks, cgb_comm **out);
void cgb_comm_destroy(cgb_comm *c);
/* Gathers every rank's factor matrix (rowsPerRank[r] rows x nPatterns; this rank's is the sampler's) into `out`,
 * (sum of rowsPerRank) x nPatterns row-major on the host, ranks in order.  Collective: every rank calls it with the same
 * rowsPerRank.  deviceMs (optional) receives the CUDA-event time of the ncclAllGather. */
int cgb_allgather_rows(cgb_comm *c, const cgb_sampler *s, const uint32_t *rowsPerRank, float *out, double *deviceMs);
/* the same for any pattern-major device block (element (row r, pattern p) at dev[p * ld + r]), e.g. the statistics sums
 * of cgb_stats_device_sums */
int cgb_allgather_device_rows(cgb_comm *c, const void *dev, uint64_t ld, uint32_t nPatterns, const uint32_t *rowsPerRank,
                              float *out, double *deviceMs);

/* ------------------------------------------------------------------------------------------
 * Checkpoints and interrupts (GapsRunner.cpp:224-270,280; utils/Archive.h:16-87).
 *
 * The wire format is the reference's own: a file written here is byte-identical to the one the reference
 * writes for the same state, and either side resumes from the other's.  Asynchronous sampler only — the
 * reference's SingleThreadedGibbsSampler does not archive its rng stream and its operator>> writes instead of
 * reading (SingleThreadedGibbsSampler.h:260-273), so there is nothing to be compatible with: asking for a
 * checkpoint with asynchronousUpdates == 0 fails with CGB_EUNSUPPORTED.
 * ---------------------------------------------------------------------------------------- */
typedef struct cgb_run_options
{
    uint32_t struct_size;          /* = sizeof(cgb_run_options) */
    uint32_t checkpointInterval;   /* GapsParameters.h:46; 0 = never.  As in createCheckpoint (GapsRunner.cpp:226-256) a
                                    * checkpoint is taken at the top of every iteration with (iter+1) % interval == 0, in
                                    * both phases, never when subsetting; taking one rebuilds AP from the factors
                                    * (extraInitialization), so the chain differs from a run without checkpoints exactly
                                    * as the reference's does. */
    const char *checkpointOutFile; /* :38; NULL or "" = "gaps_checkpoint.out" (:83).  A checkpoint that cannot be written
                                    * ends the run with CGB_EINVAL; the reference's Archive ignores a stream that failed
                                    * to open and carries on without one. */
    const char *checkpointInFile;  /* :37,56 (useCheckPoint); NULL or "" = start from scratch.  seed, nPatterns,
                                    * nIterations, alpha*, maxGibbsMass*, useSparseOptimization and checkpointInterval
                                    * are then taken from the file (run_helper, GapsRunner.cpp:99-105). */
    int32_t (*interrupt)(void *user); /* polled once per iteration on the calling thread (GapsRunner.cpp:280);
                                       * non-zero return stops the run with CGB_EINTERRUPTED.  NULL = never. */
    void *interruptUser;
} cgb_run_options;

/* cgb_run with checkpoints / interrupt polling; options == NULL is exactly cgb_run. */
int cgb_run_ex(const float *data, uint32_t nrow, uint32_t ncol, int32_t colmajor, const float *uncertainty,
               const cgb_params *params, const cgb_run_options *options, cgb_result *result);

/* the path overload with the same options (gaps::run(const std::string&, ...) reads checkpoints too, GapsRunner.cpp:119-159) */
int cgb_run_file_ex(const char *dataPath, const char *uncertaintyPath, const cgb_params *params,
                    const cgb_run_options *options, cgb_result *result);

/* `Archive << sampler` / `Archive >> sampler` of the Sampler concept (AsynchronousGibbsSampler.h:221-233): the bytes the
 * reference streams for the factor matrix, the atomic domain (atoms in pick order) and the proposal queue.  Size the
 * buffer with buf == NULL.  Deserialising replaces the factor matrix, the atoms and the queue state of a sampler of the
 * same shape and model; call sync() / extraInitialization() afterwards, as runCoGAPSAlgorithm does (GapsRunner.cpp:444-447). */
int cgb_sampler_serialize(const cgb_sampler *s, void *buf, uint64_t capacity, uint64_t *size);
int cgb_sampler_deserialize(cgb_sampler *s, const void *buf, uint64_t size);
/* Replace the atomic domain: atoms are inserted in the order given, which becomes the pick order
 * (ConcurrentAtomicDomain.cpp:144-155).  The factor matrix is not touched (cgb_sampler_set_matrix). */
int cgb_sampler_set_atoms(cgb_sampler *s, const uint64_t *pos, const float *mass, uint64_t n);
/* `Archive << stats` / `>>` (GapsStatistics.cpp:164-176): the four running sums, statUpdates, nPatterns */
int cgb_stats_serialize(const cgb_stats *st, void *buf, uint64_t capacity, uint64_t *size);
int cgb_stats_deserialize(cgb_stats *st, const void *buf, uint64_t size);
/* seeder state of a GapsRandomState (math/Random.cpp:347-357) */
int cgb_randstate_get_state(const cgb_randstate *rs, uint64_t state[2]);
int cgb_randstate_set_state(cgb_randstate *rs, const uint64_t state[2]);
/* PCG state of a GapsRng (math/Random.cpp:202-212) */
int cgb_rng_get_state(const cgb_rng *r, uint64_t *state);
int cgb_rng_set_state(cgb_rng *r, uint64_t state);

/* What a checkpoint file holds.  Host only, no device needed. */
typedef struct cgb_checkpoint_info
{
    uint32_t struct_size;
    uint32_t seed, nGenes, nSamples, nPatterns, nIterations;
    float alphaA, alphaP, maxGibbsMassA, maxGibbsMassP;
    int32_t useSparseOptimization;
    uint32_t checkpointInterval;
    int32_t phase;                 /* CGB_PHASE_EQUILIBRATION or CGB_PHASE_SAMPLING */
    uint32_t iter;                 /* the iteration the resumed run starts with */
    uint64_t nAtomsA, nAtomsP;
    uint32_t statUpdates;
    uint64_t fileBytes;
} cgb_checkpoint_info;
int cgb_checkpoint_info_read(const char *path, cgb_checkpoint_info *out);
/* Parses inPath completely (every structural check a resume would make) and writes it back out through the library's
 * own writer; the result is byte-identical for any file the reference wrote.  Host only. */
int cgb_checkpoint_rewrite(const char *inPath, const char *outPath);

/* ------------------------------------------------------------------------------------------
 * Test hooks (no reference counterpart)
 * ---------------------------------------------------------------------------------------- */
/* lookup tables cgb_run hands to the GapsRandomState it creates (NULLs restore the built-ins) */
int cgb_run_set_tables(const float *erf, const float *erfinv, const float *qgamma);
/* the portable log of csrc/gaps_math.h evaluated on the device / on the host */
int cgb_debug_logf(const float *in, float *out, uint32_t n);
float cgb_debug_host_logf(float x);
/* the fp32 running sum and positive count behind lambda (gaps::nonZeroMean, MatrixMath.cpp:39-55) as the samplers take
 * them: row by row over a row-major nrow x ncol matrix, or (byColumns != 0) column by column */
int cgb_debug_running_sum(const float *data, uint32_t nrow, uint32_t ncol, int32_t byColumns, float *sum, uint32_t *nnz);
/* floor(x / divisor) as the host generator computes it for its per-sampler divisors (tests: == x / divisor) */
uint64_t cgb_debug_fastdiv(uint64_t divisor, uint64_t x);

/* One evaluated proposal of a recorded run (layout of the oracle's trace records, oracle/cogaps_oracle.h) */
typedef struct cgb_trace_record
{
    uint32_t phase;     /* CGB_PHASE_EQUILIBRATION / CGB_PHASE_SAMPLING */
    uint32_t iter;
    uint32_t side;      /* 'A' or 'P' */
    uint32_t batch;     /* batch index within this update() call */
    uint32_t type;      /* 'B', 'D', 'M', 'E' */
    uint32_t r1, c1, r2, c2;
    uint32_t accepted;  /* B: born; D: the atom survives; M: moved; E: masses changed */
    uint64_t pos;       /* birth position / move destination */
    uint64_t atom1Pos;
    uint64_t atom2Pos;
    uint64_t rngState;  /* PCG state handed to the evaluator */
    float mass1, mass2; /* atom masses before evaluation */
    float newMass1, newMass2;
    float s, s_mu;
} cgb_trace_record;
/* Host only, no device: runs the library's own proposal generator and atomic domain (the code cgb_sampler_update
 * drives; atomic/ProposalQueue.cpp:53-283, ConcurrentAtomicDomain.cpp:14-132) through a whole gaps::run on `data`
 * (nrow x ncol row-major; asynchronous sampler, whole untransposed matrix) with the OUTCOME of every proposal taken
 * from `trace`, and compares every proposal it queues — type, bins, positions, atom masses, PCG state, batch
 * boundaries — with the trace.  CGB_OK when all n records match and none is left over; otherwise CGB_EINTERNAL and
 * cgb_debug_replay_message() says which field of which record differs.  *checked = records matched. */
int cgb_debug_replay_generator(const float *data, uint32_t nrow, uint32_t ncol, const cgb_params *params,
                               const cgb_trace_record *trace, uint64_t n, uint64_t *checked);
const char *cgb_debug_replay_message(void);
/* Host only: the library's bin-indexed atomic domain against a naive restatement of the reference's structures (a sorted
 * map for order and neighbours, a vector with swap-with-last erases for the pick order — ConcurrentAtomicDomain.cpp:
 * 14-132) under nOps random inserts, batched erases and in-gap moves over nBins bins.  CGB_OK, or CGB_EINTERNAL with
 * cgb_debug_replay_message() and *opsDone = the operation that diverged. */
int cgb_debug_domain_fuzz(uint64_t seed, uint64_t nBins, uint32_t nOps, uint64_t *opsDone);

#ifdef __cplusplus
}
#endif

#endif /* COGAPS_B200_H */
