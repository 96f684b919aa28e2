/*
 * cogaps_oracle.h — CPU restatement of the CoGAPS Gibbs-sampler hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker the CUDA path is compared against; it is never
 * linked, loaded or called by anything under cogaps_b200/.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg use it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this restatement bit-for-bit
 * (atom histories, chi-square histories, snapshots of A and P, posterior means/sds, lookup
 * tables, RNG streams, the alphaParameters scans) against the reference itself, compiled
 * unmodified by oracle/Makefile into oracle/_ref/ (scalar and AVX variants), and against golden
 * vectors generated from those builds (tests/golden/).
 */
#ifndef COGAPS_ORACLE_H
#define COGAPS_ORACLE_H

#include <stdint.h>
#include "../include/cogaps_b200.h" /* cgb_params / cgb_result / cgb_reduction_order layouts only */

#ifdef __cplusplus
extern "C" {
#endif

/* how the two sums of an alphaParameters scan are associated */
#define ORACLE_REDUCE_SCALAR 0 /* reference, math/SIMD.h scalar path: one running sum          */
#define ORACLE_REDUCE_AVX8   1 /* reference, math/SIMD.h:8-19,101-114: 8 lanes + hadd tree      */
#define ORACLE_REDUCE_DEVICE 2 /* the CUDA kernel's order, described by cgb_reduction_order     */

/* which log() the accept tests use */
#define ORACLE_MATH_LIBM     0 /* glibc logf, what the reference build calls                    */
#define ORACLE_MATH_PORTABLE 1 /* the f64 atanh-series log shared bit-for-bit with the device   */

typedef struct oracle_options
{
    int32_t reduceMode;
    int32_t mathMode;
    cgb_reduction_order orderA; /* used when reduceMode == ORACLE_REDUCE_DEVICE */
    cgb_reduction_order orderP;
    const float *erf;           /* optional table overrides (NULL = built in)   */
    const float *erfinv;
    const float *qgamma;
    /* checkpoints (GapsParameters.h:37-38,46,56; GapsRunner.cpp:224-270), asynchronous sampler only */
    uint32_t checkpointInterval;   /* 0 = never */
    const char *checkpointOutFile; /* NULL = "gaps_checkpoint.out" (GapsParameters.h:83) */
    const char *checkpointInFile;  /* non-NULL = resume (useCheckPoint) */
    uint32_t stopAfterCheckpoints; /* tests: return -7 right after writing this many checkpoints (a user interrupt,
                                    * GapsRunner.cpp:280, at a known place); 0 = run to the end */
} oracle_options;

/* one evaluated proposal, for lock-step comparison and host-logic replay tests */
typedef struct oracle_trace_record
{
    uint32_t phase;     /* 1 equilibration, 2 sampling */
    uint32_t iter;
    uint32_t side;      /* 'A' or 'P' */
    uint32_t batch;     /* batch index within this update() call */
    uint32_t type;      /* 'B','D','M','E' */
    uint32_t r1, c1, r2, c2;
    uint32_t accepted;  /* B: born; D: atom survives; M: moved; E: masses changed */
    uint64_t pos;       /* move destination / birth position */
    uint64_t atom1Pos;
    uint64_t atom2Pos;
    uint64_t rngState;  /* PCG state handed to the evaluator */
    float mass1, mass2; /* atom masses before evaluation */
    float newMass1, newMass2;
    float s, s_mu;      /* alpha parameters after annealing */
} oracle_trace_record;

int cogaps_oracle_run(const float *data, uint32_t nrow, uint32_t ncol, const float *uncertainty,
                      const cgb_params *params, cgb_result *result, const oracle_options *opt);

/* as cogaps_oracle_run, additionally recording up to `capacity` evaluated proposals */
int cogaps_oracle_run_trace(const float *data, uint32_t nrow, uint32_t ncol, const float *uncertainty,
                            const cgb_params *params, cgb_result *result, const oracle_options *opt,
                            oracle_trace_record *trace, uint64_t capacity, uint64_t *count);

int cogaps_oracle_tables(float *erf, float *erfinv, float *qgamma);

/* same `kind` codes as cogaps_ref_rng_stream in ref_driver.cpp */
int cogaps_oracle_rng_stream(uint32_t seed, int kind, uint32_t n, uint64_t a, uint64_t b, double lambda,
                             float f0, float f1, float f2, float f3, uint64_t *out);

/* same contract as cogaps_ref_alpha_parameters, plus the reduction order */
int cogaps_oracle_alpha_parameters(const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                                   const float *A, const float *P, const float *uncertainty,
                                   uint32_t n, const int32_t *variant, const uint32_t *r1,
                                   const uint32_t *c1, const uint32_t *r2, const uint32_t *c2,
                                   const float *ch, float *s_out, float *smu_out, float *ap_out,
                                   const oracle_options *opt);

/* chiSq of the A-side (out[0]) and P-side (out[1]) models for given factors; out[2] = dataSparsity */
int cogaps_oracle_chisq(const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                        const float *A, const float *P, const float *uncertainty, float *out);

/* the two probes above on SparseNormalModel (default uncertainty only) */
int cogaps_oracle_alpha_parameters_sparse(const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                                          const float *A, const float *P,
                                          uint32_t n, const int32_t *variant, const uint32_t *r1,
                                          const uint32_t *c1, const uint32_t *r2, const uint32_t *c2,
                                          const float *ch, float *s_out, float *smu_out,
                                          const oracle_options *opt);
int cogaps_oracle_chisq_sparse(const float *data, uint32_t nGenes, uint32_t nSamples, uint32_t k,
                               const float *A, const float *P, float *out);

float cogaps_oracle_portable_logf(float x);
/* the sweep's building blocks (cogaps_b200/csrc/sweep.cuh, gaps_math.h): Philox4x32-10 block function, portable exp */
void cogaps_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float cogaps_oracle_portable_expf(float x);

#ifdef __cplusplus
}
#endif
#endif
