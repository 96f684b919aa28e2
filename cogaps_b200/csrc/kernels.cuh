// kernels.cuh — sm_100a kernels of the Gibbs-sampler hot path.
//
//   eval_stream_kernel the resident grid (one launch per update()): worker clusters poll rings of task records in
//                      pinned host memory, run process_task / commit_task (or sparse_task) on each, post outcome
//                      records; one extra CTA mirrors the count of finished commits to the host
//   eval_kernel        the same device functions, one launch per conflict-free batch.  process_task: one cluster per
//                      (proposal, row): TMA bulk-stage the touched D/S/AP row segment and the 1-2 factor columns
//                      into shared memory, the alphaParameters scan (DenseNormalModel.cpp:162-240), gibbsMass /
//                      accept epilogue on one lane (AsynchronousGibbsSampler.h:126-219, SingleThreadedGibbsSampler.h:
//                      130-257); commit_task: AP commit from shared memory (DenseNormalModel.cpp:243-258), row
//                      version + done count.  12-20 B per row element.
//   probe_kernel       any number of single-row alphaParameters queries in one launch (the scan at saturation)
//   eval_sparse_kernel / eval_stream_sparse_kernel, sparse_tables_kernel, sparse_chisq_kernel
//                      SparseNormalModel (SparseNormalModel.cpp:39-60,153-311)
//   transpose_kernel   DenseNormalModel::sync (DenseNormalModel.cpp:20-36)
//   csr_scatter_kernel compressed rows -> dense D (Matrix-Market input of the sparse model)
//   rebuild_ap_kernel  extraInitialization (:38-54)
//   chisq_kernel       chiSq (:56-68)
//   col_nonzero_kernel canUseGibbs precompute (:100-108)
//   stats kernels      GapsStatistics::update/updateA/updateP, meanChiSq (GapsStatistics.h:129-185, .cpp:63-87)
#ifndef CGB_KERNELS_CUH
#define CGB_KERNELS_CUH

#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "../../include/cogaps_b200.h"
#include "device_types.h"
#include "gaps_math.h"

namespace cgb {

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA without a tensor map; SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// global -> this CTA's shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dstSmem, const void *srcGmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dstSmem)), "l"(srcGmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------
// eval kernel
// ------------------------------------------------------------------------------------------------
struct Decision
{
    float dOwn1;        // delta applied to the own row with stream V1
    float dOwn2;        // delta applied afterwards with stream V2 (same-row move / exchange)
    float dOther;       // delta for the other row of a two-row proposal
    uint32_t otherRow;
    uint32_t otherCol;
    uint32_t flags;     // bit0 own1, bit1 own2, bit2 other
};

struct EvalSmem
{
    uint64_t bar;
    float warpS[kThreads / 32];
    float warpMu[kThreads / 32];
    float segS[kMaxCluster];
    float segMu[kMaxCluster];
    Decision dec;
    float preLog[2];    // log of the proposal stream's next and next-but-one uniform (see PreLog)
};

__device__ __forceinline__ float derive_s(float d)
{
    // gaps::pmax(D, 0.1f): max(D * 0.1, 0.1) (MatrixMath.cpp:74-84, DenseNormalModel.h:73)
    const float a = fmul(d, 0.1f);
    return a < 0.1f ? 0.1f : a;
}

template <bool HAS_S, bool USE_V2, bool WITH_CHANGE>
__device__ __forceinline__ void scan_segment(const float *bufD, const float *bufS, const float *bufAP,
                                             const float *bufV1, const float *bufV2, uint32_t len, float ch,
                                             float &accS, float &accMu)
{
    const uint32_t tid = threadIdx.x;
    const uint32_t nVec = (len + kVec - 1) / kVec;
#pragma unroll 2
    for (uint32_t j = tid; j < nVec; j += kThreads)
    {
        const float4 d4 = reinterpret_cast<const float4*>(bufD)[j];
        const float4 a4 = reinterpret_cast<const float4*>(bufAP)[j];
        const float4 v4 = reinterpret_cast<const float4*>(bufV1)[j];
        float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (USE_V2) { w4 = reinterpret_cast<const float4*>(bufV2)[j]; }
        if (HAS_S) { s4 = reinterpret_cast<const float4*>(bufS)[j]; }
        const float d[4] = {d4.x, d4.y, d4.z, d4.w};
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
        const float su[4] = {s4.x, s4.y, s4.z, s4.w};
        const uint32_t base = j * kVec;
        float ts[4], tmu[4];
        // Four independent element pipelines, no branches: an element with mat == 0 (or past the end of
        // the row) contributes exactly +0 to both sums, so it is replaced by a select instead of being
        // computed — which also keeps the IEEE division off the slow path a zero numerator would take.
#pragma unroll
        for (int c = 0; c < kVec; ++c)
        {
            const float mat = USE_V2 ? fsub(v[c], w[c]) : v[c];
            const bool live = (base + c < len) && (mat != 0.f);
            const float sd = HAS_S ? su[c] : derive_s(d[c]);
            const float ratio = fdiv(live ? mat : 1.f, fmul(sd, sd));
            const float resid = WITH_CHANGE ? fsub(d[c], fadd(a[c], fmul(ch, v[c]))) : fsub(d[c], a[c]);
            ts[c] = live ? fmul(mat, ratio) : 0.f;
            tmu[c] = live ? fmul(ratio, resid) : 0.f;
        }
        // the lane's running sums take the elements in increasing index order
#pragma unroll
        for (int c = 0; c < kVec; ++c)
        {
            accS = fadd(accS, ts[c]);
            accMu = fadd(accMu, tmu[c]);
        }
    }
}

// debug phase profile: SM clock of the leader CTA's lane 0 at each phase boundary (off unless requested)
__device__ __forceinline__ void stamp(const ModelView &mv, uint32_t task, uint32_t rank, int slot)
{
    if (mv.phaseClocks != nullptr && rank == 0 && threadIdx.x == 0)
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        mv.phaseClocks[static_cast<size_t>(task) * kPhaseSlots + slot] = (slot == 0) ? t : static_cast<unsigned long long>(clock64());
        if (slot == 0) { mv.phaseClocks[static_cast<size_t>(task) * kPhaseSlots + 1] = static_cast<unsigned long long>(clock64()); }
    }
}

__device__ __forceinline__ float ld_cg_f32(const float *p)
{
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// What the deciding lane hands back to its cluster.
struct Verdict
{
    Decision dec;
    DevOutcome out;
    float newM1, newM2; // the factor elements (r1,c1) / (r2,c2) after the proposal (sparse model: the row copy)
    float newC1, newC2; // sparse model: the column copy (values below epsilon stored as 0)
};

// The accept tests take log(uniform()) of the proposal's own PCG stream, as its first draw (move, death
// without a Gibbs draw) or its second (death after gibbsMass drew).  Both candidates depend only on the
// stream state the host sent, so another warp computes them while the scan runs and the f64 log is off
// the serial tail; which one applies is known once gibbsMass has or has not advanced the stream.
struct PreLog
{
    uint64_t state0;   // stream state as posted
    float logFirst;    // portable_logf(uniform()) of the first draw
    float logSecond;   // ... of the second draw
    __device__ __forceinline__ float take(Pcg &rng) const
    {
        const bool first = (rng.state == state0);
        rng.advance();
        return first ? logFirst : logSecond;
    }
};

// The serial tail of one proposal: alpha parameters -> gibbsMass / accept test -> deltas
// (AsynchronousGibbsSampler.h:126-219).  One lane runs it; kept out of line so its registers (f64 log,
// divisions) do not inflate the allocation of the 255 lanes that only scan.
// SPARSE: M1/M2 are the row-copy values, C1/C2 the column-copy values of the two elements; both copies are
// rewritten with HybridMatrix::add / set semantics (data_structures/HybridMatrix.cpp:25-39).
// LEAN (the sweep, which inlines it): asynchronous sampler's proposals only, no probes, no count of the draws taken.
template <bool SPARSE, bool LEAN>
__device__ __forceinline__ void decide_body(const ModelView &mv, const float *erfT, const float *erfinvT, float T, const DevProposal &pr, uint32_t part, bool twoRow,
                                            float s, float mu, float M1, float M2, float C1, float C2, int can1, int can2, const PreLog &pre, Verdict *v)
{
    bool add1 = false, add2 = false; // element changed by changeMatrix (add) rather than safelyChangeMatrix (set)
    const bool sequential = !LEAN && (pr.pad & 1u) != 0u; // proposal of the SingleThreadedGibbsSampler (one shared stream)
    const uint32_t type = pr.type;
    const uint32_t r1 = pr.r1, c1 = pr.c1, r2 = pr.r2, c2 = pr.c2;
    const float m1 = pr.m1, m2 = pr.m2;
    Decision dec;
    dec.dOwn1 = dec.dOwn2 = dec.dOther = 0.f;
    dec.otherRow = dec.otherCol = 0u;
    dec.flags = 0u;
    DevOutcome out;
    out.mass1 = 0.f;
    out.mass2 = 0.f;
    out.accepted = 0u;
    out.pad[0] = out.pad[1] = out.pad[2] = 0u;
    const float as = fmul(s, T), amu = fmul(mu, T);
    out.s = as;
    out.s_mu = amu;
    Pcg rng;
    rng.state = pr.rng;
    float d1 = 0.f, d2 = 0.f;   // deltas of element (r1,c1) and (r2,c2)
    bool ch1 = false, ch2 = false;
    if (!LEAN && type == kProbe)
    {
        out.s = s;
        out.s_mu = mu;
    }
    else if (type == 'B')
    {
        // AsynchronousGibbsSampler::birth, AsynchronousGibbsSampler.h:126-144
        float mass = 0.f;
        bool has;
        if (can1 != 0)
        {
            has = gibbs_mass(rng, erfT, erfinvT, as, amu, 0.f, mv.maxGibbsMass, true, mv.lambda, &mass);
        }
        else
        {
            mass = fdiv(fmul(-1.f, pre.take(rng)), mv.lambda);
            has = true;
        }
        // AsynchronousGibbsSampler.h:139 accepts mass >= epsilon, SingleThreadedGibbsSampler.h:144 mass > epsilon
        if (has && (sequential ? (mass > kEpsilon) : (mass >= kEpsilon)))
        {
            out.accepted = 1u;
            out.mass1 = mass;
            d1 = mass;                      // changeMatrix: no clamp
            M1 = fadd(M1, mass);
            ch1 = true;
            add1 = true;
        }
    }
    else if (type == 'D')
    {
        // AsynchronousGibbsSampler::death, :147-180
        float rebirth = m1;
        if (can1 != 0)
        {
            float g;
            if (gibbs_mass(rng, erfT, erfinvT, as, amu, 0.f, mv.maxGibbsMass, true, mv.lambda, &g)) { rebirth = g; }
        }
        const float dLL = fmul(rebirth, fsub(amu, fmul(fmul(as, rebirth), 0.5f)));
        if (pre.take(rng) < dLL)
        {
            out.accepted = 1u;
            out.mass1 = rebirth;
            if (rebirth != m1)
            {
                const float nv = gmax(fadd(M1, fsub(rebirth, m1)), 0.f); // safelyChangeMatrix
                d1 = fsub(nv, M1);
                M1 = nv;
                ch1 = true;
            }
        }
        else
        {
            const float nv = gmax(fadd(M1, fmul(-1.f, m1)), 0.f);
            d1 = fsub(nv, M1);
            M1 = nv;
            ch1 = true;
        }
    }
    else if (type == 'M')
    {
        // AsynchronousGibbsSampler::move, :183-196; deltaLogLikelihood DenseNormalModel.cpp:125-130
        const float dLL = fmul(fmul(-1.f, m1), fadd(amu, fmul(fmul(as, m1), 0.5f)));
        if (pre.take(rng) < dLL)
        {
            out.accepted = 1u;
            out.mass1 = m1;
            const float nv = gmax(fadd(M1, -m1), 0.f);
            d1 = fsub(nv, M1);
            M1 = nv;
            ch1 = true;
            d2 = m1;                        // changeMatrix(r2, c2, mass)
            M2 = fadd(M2, m1);
            ch2 = true;
            add2 = true;
        }
    }
    else if (type == 'E')
    {
        // AsynchronousGibbsSampler::exchange, :200-219; sampleExchange DenseNormalModel.cpp:154-159
        if (can1 != 0 || can2 != 0)
        {
            float g;
            const bool has = gibbs_mass(rng, erfT, erfinvT, as, amu, -m1, m2, false, 0.f, &g);
            const float n1 = fadd(m1, g), n2 = fsub(m2, g);
            if (has && n1 > kEpsilon && n2 > kEpsilon)
            {
                out.accepted = 1u;
                out.mass1 = n1;
                out.mass2 = n2;
                const float nv1 = gmax(fadd(M1, fsub(n1, m1)), 0.f);
                d1 = fsub(nv1, M1);
                M1 = nv1;
                ch1 = true;
                const float nv2 = gmax(fadd(M2, fsub(n2, m2)), 0.f);
                d2 = fsub(nv2, M2);
                M2 = nv2;
                ch2 = true;
            }
        }
    }
    if (SPARSE)
    {
        if (ch1)
        {
            mv.Mrows[static_cast<size_t>(r1) * mv.ldR + c1] = M1;
            const float cv = add1 ? fadd(C1, d1) : M1;
            C1 = (cv < kEpsilon) ? 0.f : cv;
            mv.M[static_cast<size_t>(c1) * mv.ldM + r1] = C1;
        }
        if (ch2)
        {
            mv.Mrows[static_cast<size_t>(r2) * mv.ldR + c2] = M2;
            const float cv = add2 ? fadd(C2, d2) : M2;
            C2 = (cv < kEpsilon) ? 0.f : cv;
            mv.M[static_cast<size_t>(c2) * mv.ldM + r2] = C2;
        }
    }
    else
    {
        if (ch1) { mv.M[static_cast<size_t>(c1) * mv.ldM + r1] = M1; }
        if (ch2) { mv.M[static_cast<size_t>(c2) * mv.ldM + r2] = M2; }
    }
    if (twoRow)
    {
        // own row is row `part`; the other row is committed through global memory
        const float dOwn = part ? d2 : d1, dOth = part ? d1 : d2;
        const bool cOwn = part ? ch2 : ch1, cOth = part ? ch1 : ch2;
        dec.dOwn1 = dOwn;
        dec.dOther = dOth;
        dec.otherRow = part ? r1 : r2;
        dec.otherCol = part ? c1 : c2;
        dec.flags = (cOwn ? 1u : 0u) | (cOth ? 4u : 0u);
    }
    else
    {
        dec.dOwn1 = d1;
        dec.dOwn2 = d2;
        dec.flags = (ch1 ? 1u : 0u) | (ch2 ? 2u : 0u);
    }
    // the sequential sampler continues ITS stream after us: tell it how many draws we made (0, 1 or 2)
    if (!LEAN)
    {
        Pcg probe;
        probe.state = pr.rng;
        uint32_t draws = 0u;
        if (rng.state != probe.state)
        {
            probe.advance();
            draws = (rng.state == probe.state) ? 1u : 2u;
        }
        out.pad[0] = draws;
    }
    v->dec = dec;
    v->out = out;
    v->newM1 = M1;
    v->newM2 = M2;
    v->newC1 = C1;
    v->newC2 = C2;
}

template <bool SPARSE>
__device__ __noinline__ void decide(const ModelView &mv, const float *erfT, const float *erfinvT, float T, const DevProposal &pr, uint32_t part, bool twoRow,
                                    float s, float mu, float M1, float M2, float C1, float C2, int can1, int can2, const PreLog &pre, Verdict *v)
{
    decide_body<SPARSE, false>(mv, erfT, erfinvT, T, pr, part, twoRow, s, mu, M1, M2, C1, C2, can1, can2, pre, v);
}

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// One (proposal, row) work item as a CTA sees it.
struct TaskIn
{
    DevProposal pr;
    uint32_t pi;       // index of the proposal in its batch (where its outcome goes)
    uint32_t part;     // 0: row r1 (or the only row); 1: row r2 of a two-row move / exchange
    uint32_t ver1;     // streaming mode: rowVersion[r1] / rowVersion[r2] to wait for ...
    uint32_t ver2;
    uint32_t waitMask; // ... when bit 0 / bit 1 is set (the host cannot yet prove the row's last commit has landed)
};

// One (proposal, row) task, executed by one cluster of nSeg CTAs.  `parity` is the phase of the
// staging mbarrier (flips every task in the resident kernel).  Returns true on the lane that owns the
// proposal's outcome (leader CTA, lane 0, deciding cluster), with the outcome in *outp.
//
// STREAM: tasks of successive batches are in flight at once, so a row (its AP line and its factor
// elements) may only be read once every commit the host knows about has landed: the host sends the
// row's expected commit count, committers bump rowVersion[row] after a fence.  D and the factor columns
// are constant for the whole update(), so their copies are issued before that wait.
template <bool HAS_S, bool STREAM>
__device__ __forceinline__ bool process_task(const ModelView &mv, const float *erfT, const float *erfinvT, float annealingTemp, const TaskIn &in,
                                             uint32_t task, unsigned char *smemRaw, uint32_t parity,
                                             cg::cluster_group &cluster, uint32_t rank, DevOutcome *outp,
                                             unsigned long long *verWaitNs)
{
    EvalSmem *hdr = reinterpret_cast<EvalSmem*>(smemRaw);
    float *bufD = reinterpret_cast<float*>(smemRaw + 256);
    float *bufAP = bufD + mv.segPad;
    float *bufV1 = bufAP + mv.segPad;
    float *bufV2 = bufV1 + mv.segPad;
    float *bufS = bufV2 + mv.segPad;
    const uint32_t nSeg = mv.nSeg;
    const uint32_t tid = threadIdx.x;
    const DevProposal &pr = in.pr;
    const uint32_t pi = in.pi, part = in.part;

    const uint32_t type = pr.type;
    const uint32_t r1 = pr.r1, c1 = pr.c1, r2 = pr.r2, c2 = pr.c2;
    const uint32_t variant = pr.variant;
    const bool pairType = (type == 'M') || (type == 'E') || (type == kProbe && variant == 1);
    const bool twoRow = pairType && (r1 != r2);
    const bool useV2 = pairType && (r1 == r2);
    const bool withChange = (type == 'D') || (type == kProbe && variant == 2);
    const float ch = (type == 'D') ? -pr.m1 : pr.ch;
    const uint32_t row = part ? r2 : r1;
    const uint32_t colA = part ? c2 : c1;

    const uint32_t segStart = rank * mv.seg;
    const uint32_t len = segStart >= mv.L ? 0u : min(mv.seg, mv.L - segStart);
    const uint32_t lenPad = (len + 3u) & ~3u;

    stamp(mv, task, rank, 0); // slots 0 (globaltimer), 1 (clock)
    // ---- stage the touched row segment and factor columns with bulk async copies ----
    float M1 = 0.f, M2 = 0.f;
    int can1 = 0, can2 = 0;
    if (tid == 0)
    {
        const uint32_t bytes = lenPad * 4u;
        const size_t rowOff = static_cast<size_t>(row) * mv.ld + segStart;
        if (len > 0)
        {
            const uint32_t nStreams = 3u + (useV2 ? 1u : 0u) + (HAS_S ? 1u : 0u);
            mbar_expect_tx(&hdr->bar, bytes * nStreams);
            bulk_g2s(bufD, mv.D + rowOff, bytes, &hdr->bar);
            bulk_g2s(bufV1, mv.otherM + static_cast<size_t>(colA) * mv.ldOther + segStart, bytes, &hdr->bar);
            if (useV2) { bulk_g2s(bufV2, mv.otherM + static_cast<size_t>(c2) * mv.ldOther + segStart, bytes, &hdr->bar); }
            if (HAS_S) { bulk_g2s(bufS, mv.S + rowOff, bytes, &hdr->bar); }
        }
        if (rank == 0 && type != kProbe)
        {
            can1 = mv.otherColNonzero[c1];
            if (pairType) { can2 = mv.otherColNonzero[c2]; }
        }
        if (STREAM)
        {
            // both rows of a two-row proposal: the leader also reads the other row's factor element
            if (in.waitMask != 0u)
            {
                const unsigned long long t0 = global_timer_ns();
                if (in.waitMask & 1u) { while (ld_acquire_gpu_u32(mv.rowVersion + r1) != in.ver1) { } }
                if (in.waitMask & 2u) { while (ld_acquire_gpu_u32(mv.rowVersion + r2) != in.ver2) { } }
                asm volatile("fence.proxy.async;" ::: "memory");
                if (verWaitNs) { *verWaitNs = global_timer_ns() - t0; }
            }
        }
        if (len > 0) { bulk_g2s(bufAP, mv.AP + rowOff, bytes, &hdr->bar); }
        if (rank == 0 && type != kProbe)
        {
            // current factor-matrix elements (safelyChangeMatrix); their latency hides under the copies.
            // L2 loads: an earlier batch may have written M.
            M1 = ld_cg_f32(mv.M + static_cast<size_t>(c1) * mv.ldM + r1);
            if (pairType) { M2 = ld_cg_f32(mv.M + static_cast<size_t>(c2) * mv.ldM + r2); }
        }
    }
    else if (tid == 32 && rank == 0 && (type == 'D' || type == 'M' || type == 'B'))
    {
        // PreLog: both candidate log(uniform()) values of the accept test, off the serial tail
        Pcg r;
        r.state = pr.rng;
        hdr->preLog[0] = portable_logf(r.uniform());
        hdr->preLog[1] = portable_logf(r.uniform());
    }
    stamp(mv, task, rank, 2); // copies issued

    // ---- the scan ----
    float accS = 0.f, accMu = 0.f;
    if (len > 0)
    {
        mbar_wait(&hdr->bar, parity);
        stamp(mv, task, rank, 3); // data landed
        if (useV2)
        {
            scan_segment<HAS_S, true, false>(bufD, bufS, bufAP, bufV1, bufV2, len, 0.f, accS, accMu);
        }
        else if (withChange)
        {
            scan_segment<HAS_S, false, true>(bufD, bufS, bufAP, bufV1, bufV2, len, ch, accS, accMu);
        }
        else
        {
            scan_segment<HAS_S, false, false>(bufD, bufS, bufAP, bufV1, bufV2, len, 0.f, accS, accMu);
        }
    }
    stamp(mv, task, rank, 4); // scan done
    // lanes -> warp: xor butterfly, offsets 16,8,4,2,1 (the order the oracle reproduces)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1)
    {
        accS = fadd(accS, __shfl_xor_sync(0xffffffffu, accS, off));
        accMu = fadd(accMu, __shfl_xor_sync(0xffffffffu, accMu, off));
    }
    if ((tid & 31) == 0)
    {
        hdr->warpS[tid >> 5] = accS;
        hdr->warpMu[tid >> 5] = accMu;
    }
    __syncthreads();
    // warps -> segment: the warp totals sit in the low lanes of warp 0 (zeros above) and combine by
    // the same xor butterfly; segment totals go to the cluster leader
    if (tid < 32)
    {
        float sS = (tid < kThreads / 32) ? hdr->warpS[tid] : 0.f;
        float sMu = (tid < kThreads / 32) ? hdr->warpMu[tid] : 0.f;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
        {
            sS = fadd(sS, __shfl_xor_sync(0xffffffffu, sS, off));
            sMu = fadd(sMu, __shfl_xor_sync(0xffffffffu, sMu, off));
        }
        if (tid == 0)
        {
            EvalSmem *lead = (nSeg > 1) ? cluster.map_shared_rank(hdr, 0) : hdr;
            lead->segS[rank] = sS;
            lead->segMu[rank] = sMu;
        }
    }
    if (nSeg > 1) { cluster.sync(); } // single-CTA rows: the same lane wrote and reads the totals
    stamp(mv, task, rank, 5); // cluster reduce done

    // ---- decision: one lane of the leader CTA ----
    bool owner = false;
    if (rank == 0 && tid == 0)
    {
        float s = hdr->segS[0], mu = hdr->segMu[0];
        for (uint32_t q = 1; q < nSeg; ++q)
        {
            s = fadd(s, hdr->segS[q]);
            mu = fadd(mu, hdr->segMu[q]);
        }
        bool decideHere = true;
        if (twoRow)
        {
            // the cluster that arrives second owns the decision; sums combine in (row1,row2) order
            // whatever the arrival order: s = s1 + s2, s_mu = s_mu1 - s_mu2 (AlphaParameters.cpp:11-14)
            AlphaPair mine;
            mine.s = s;
            mine.s_mu = mu;
            mv.partials[pi * 2 + part] = mine;
            __threadfence();
            const uint32_t ticket = atomicAdd(&mv.tickets[pi], 1u);
            decideHere = (ticket == 1u);
            if (decideHere)
            {
                __threadfence();
                const volatile AlphaPair *o = &mv.partials[pi * 2 + (1u - part)];
                const float os = o->s, omu = o->s_mu;
                mv.tickets[pi] = 0u;
                const float s1 = part ? os : s, s2 = part ? s : os;
                const float mu1 = part ? omu : mu, mu2 = part ? mu : omu;
                s = fadd(s1, s2);
                mu = fsub(mu1, mu2);
            }
        }
        Verdict v;
        v.dec.dOwn1 = v.dec.dOwn2 = v.dec.dOther = 0.f;
        v.dec.otherRow = v.dec.otherCol = 0u;
        v.dec.flags = 0u;
        if (decideHere)
        {
            PreLog pre;
            pre.state0 = pr.rng;
            pre.logFirst = hdr->preLog[0];
            pre.logSecond = hdr->preLog[1];
            decide<false>(mv, erfT, erfinvT, annealingTemp, pr, part, twoRow, s, mu, M1, M2, 0.f, 0.f, can1, can2, pre, &v);
            *outp = v.out;
            owner = true;
        }
        // push the decision into every CTA of the cluster (DSMEM stores), so nobody reads our shared
        // memory after the barrier and no third cluster barrier is needed before exit
        hdr->dec = v.dec;
        for (uint32_t q = 1; q < nSeg; ++q) { cluster.map_shared_rank(hdr, q)->dec = v.dec; }
        stamp(mv, task, rank, 6); // decision made
    }
    return owner;
}

// Second half of a task: AP[row,:] += delta * other[:,col] (updateAPMatrix, DenseNormalModel.cpp:243-258)
// from the staged copies, then the row-version bump that lets a later batch read the row.  Separate
// from process_task so the resident kernel can post the outcome to the host in between.
template <bool HAS_S>
__device__ __forceinline__ void commit_task(const ModelView &mv, const TaskIn &in, uint32_t task, unsigned char *smemRaw,
                                            cg::cluster_group &cluster, uint32_t rank, unsigned long long *commitsDone)
{
    EvalSmem *hdr = reinterpret_cast<EvalSmem*>(smemRaw);
    float *bufD = reinterpret_cast<float*>(smemRaw + 256);
    float *bufAP = bufD + mv.segPad;
    float *bufV1 = bufAP + mv.segPad;
    float *bufV2 = bufV1 + mv.segPad;
    const uint32_t nSeg = mv.nSeg;
    const uint32_t tid = threadIdx.x;
    const uint32_t row = in.part ? in.pr.r2 : in.pr.r1;
    const uint32_t segStart = rank * mv.seg;
    const uint32_t len = segStart >= mv.L ? 0u : min(mv.seg, mv.L - segStart);
    const uint32_t lenPad = (len + 3u) & ~3u;

    if (nSeg > 1) { cluster.sync(); } else { __syncthreads(); }
    stamp(mv, task, rank, 7); // decision broadcast

    const Decision dec = hdr->dec;
    if ((dec.flags & 3u) != 0u && len > 0)
    {
        float *apRow = mv.AP + static_cast<size_t>(row) * mv.ld + segStart;
        const uint32_t nVec = lenPad / kVec;
        const bool own1 = (dec.flags & 1u) != 0u, own2 = (dec.flags & 2u) != 0u;
        for (uint32_t j = tid; j < nVec; j += kThreads)
        {
            float4 a = reinterpret_cast<const float4*>(bufAP)[j];
            if (own1)
            {
                const float4 v = reinterpret_cast<const float4*>(bufV1)[j];
                a.x = fadd(a.x, fmul(dec.dOwn1, v.x));
                a.y = fadd(a.y, fmul(dec.dOwn1, v.y));
                a.z = fadd(a.z, fmul(dec.dOwn1, v.z));
                a.w = fadd(a.w, fmul(dec.dOwn1, v.w));
            }
            if (own2)
            {
                const float4 v = reinterpret_cast<const float4*>(bufV2)[j];
                a.x = fadd(a.x, fmul(dec.dOwn2, v.x));
                a.y = fadd(a.y, fmul(dec.dOwn2, v.y));
                a.z = fadd(a.z, fmul(dec.dOwn2, v.z));
                a.w = fadd(a.w, fmul(dec.dOwn2, v.w));
            }
            reinterpret_cast<float4*>(apRow)[j] = a;
        }
    }
    if ((dec.flags & 4u) != 0u && len > 0)
    {
        float *apRow = mv.AP + static_cast<size_t>(dec.otherRow) * mv.ld + segStart;
        const float *vCol = mv.otherM + static_cast<size_t>(dec.otherCol) * mv.ldOther + segStart;
        const uint32_t nVec = lenPad / kVec;
        for (uint32_t j = tid; j < nVec; j += kThreads)
        {
            float4 a = __ldcg(reinterpret_cast<const float4*>(apRow) + j);
            const float4 v = __ldcg(reinterpret_cast<const float4*>(vCol) + j);
            a.x = fadd(a.x, fmul(dec.dOther, v.x));
            a.y = fadd(a.y, fmul(dec.dOther, v.y));
            a.z = fadd(a.z, fmul(dec.dOther, v.z));
            a.w = fadd(a.w, fmul(dec.dOther, v.w));
            reinterpret_cast<float4*>(apRow)[j] = a;
        }
    }
    // every CTA of the cluster counts once per row it changed (the host expects nSeg per commit); the
    // leader's store of the factor element precedes its fence, so the bump publishes that too
    if ((dec.flags & 7u) != 0u)
    {
        __syncthreads();
        if (tid == 0)
        {
            // bar.sync ordered every thread's AP stores before this fence; the fence makes them visible
            // device-wide (L2) before the bumps.  The done count travels on to the host through the mirror
            // CTA, which acquires it at gpu scope and releases at system scope (mirror_loop), so gpu scope
            // is enough here and the committer is not held up by a system-scope fence.
            __threadfence();
            if ((dec.flags & 3u) != 0u) { atomicAdd(mv.rowVersion + row, 1u); }
            if ((dec.flags & 4u) != 0u) { atomicAdd(mv.rowVersion + dec.otherRow, 1u); }
            if (commitsDone != nullptr) { atomicAdd(commitsDone, 1ull); }
        }
    }
    stamp(mv, task, rank, 8); // commit done
}

// ------------------------------------------------------------------------------------------------
// SparseNormalModel (gibbs_sampler/SparseNormalModel.cpp:153-292): no AP matrix; the scan visits the
// non-zeros of the data row whose factor-column entry is non-zero too and needs, for each, the dot of two
// factor ROWS (k floats, a gather).  One CTA of 256 threads per (proposal, row).
//
// Order of the sums (what oracle mode "device" restates): the visited elements, in ascending scan index,
// are numbered e = 0,1,...; element e belongs to lane e % 256; a lane adds its elements in order; lanes
// combine by the xor butterflies of the dense kernel.  The reference walks the same elements in the same
// order but adds them into one running sum (fp32 re-association only).
// ------------------------------------------------------------------------------------------------
struct SparseSmem
{
    float warpS[kSparseThreads / 32];
    float warpMu[kSparseThreads / 32];
    uint32_t warpCnt[kSparseGroup * (kSparseThreads / 32)];
    float baseS, baseMu;      // Z-table terms of (s, s_mu)
    Decision dec;
    float preLog[2];
    uint32_t pad[2];
};

// dot of two factor rows, ascending, mul and add rounded separately (gaps::dot, oracle mode "device")
__device__ __forceinline__ float row_dot(const float *a, const float *b, uint32_t k)
{
    float acc = 0.f;
    for (uint32_t i = 0; i < k; ++i) { acc = fadd(acc, fmul(a[i], b[i])); }
    return acc;
}

// The Z-table terms of one sparse scan (SparseNormalModel.cpp:153-292): s = Z1[c] (or Z1[c1] - 2 Z2[c2][c1] + Z1[c2]);
// s_mu = -(A[row,:] . Z2[:,c]) ...  One lane computes them while the others gather.
__device__ __forceinline__ void sparse_table_terms(const ModelView &mv, const float *sRow, uint32_t colA, uint32_t c1, uint32_t c2,
                                                   bool useV2, bool withChange, float ch, float &bsOut, float &bmuOut)
{
    const uint32_t k = mv.k;
    float bs, bmu;
    if (useV2)
    {
        bs = fadd(fsub(mv.Z1[c1], fmul(2.f, mv.Z2[static_cast<size_t>(c2) * k + c1])), mv.Z1[c2]);
        const float *za = mv.Z2 + static_cast<size_t>(c1) * k, *zb = mv.Z2 + static_cast<size_t>(c2) * k;
        float acc = 0.f;
        for (uint32_t i = 0; i < k; ++i) { acc = fadd(acc, fmul(sRow[i], fsub(za[i], zb[i]))); } // gaps::dot_diff
        bmu = fmul(-1.f, acc);
    }
    else
    {
        bs = mv.Z1[colA];
        bmu = fmul(-1.f, row_dot(sRow, mv.Z2 + static_cast<size_t>(colA) * k, k));
        if (withChange) { bmu = fsub(bmu, fmul(ch, mv.Z2[static_cast<size_t>(colA) * k + colA])); }
    }
    bsOut = bs;
    bmuOut = bmu;
}

// The scan over one data row's non-zeros (the whole CTA of kSparseThreads; contains __syncthreads): per-lane partial
// sums in accS / accMu, elements visited in `visited` (same on every thread).  warpCnt: kSparseGroup * 8 counters.
__device__ __forceinline__ void sparse_scan_row(const ModelView &mv, uint32_t *warpCnt, const float *sRow, uint32_t *sIdx, float *sD,
                                                float *sV1, float *sV2, uint32_t row, uint32_t colA, uint32_t c2, bool useV2,
                                                bool withChange, float ch, float &accS, float &accMu, uint32_t &visited)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t k = mv.k;
    // ---- the scan over the row's non-zeros, kSparseGroup x 256 at a time: compact the common ones, deal them
    // out to the lanes.  Every load of a group is issued before any is consumed (index/value, then the factor
    // column gather, then the factor row gathers), so a row of <= 1024 non-zeros pays each dependent latency once.
    const uint32_t start = mv.spRowPtr[row], nnz = mv.spRowPtr[row + 1] - start;
    const float *V1 = mv.otherM + static_cast<size_t>(colA) * mv.ldOther;
    const float *V2 = mv.otherM + static_cast<size_t>(c2) * mv.ldOther;
    accS = 0.f;
    accMu = 0.f;
    visited = 0; // elements visited so far (same on every thread)
    for (uint32_t base = 0; base < nnz; base += kSparseThreads * kSparseGroup)
    {
        uint32_t l[kSparseGroup], ballot[kSparseGroup];
        float d[kSparseGroup], v1[kSparseGroup], v2[kSparseGroup];
        bool pred[kSparseGroup];
#pragma unroll
        for (int g = 0; g < kSparseGroup; ++g)
        {
            // sub-group g holds elements base + g * 256 + tid: ascending index is (g, tid) order
            const uint32_t j = base + g * kSparseThreads + tid;
            l[g] = 0u;
            d[g] = 0.f;
            if (j < nnz)
            {
                l[g] = mv.spIdx[start + j];
                d[g] = mv.spVal[start + j];
            }
        }
#pragma unroll
        for (int g = 0; g < kSparseGroup; ++g)
        {
            const uint32_t j = base + g * kSparseThreads + tid;
            v1[g] = 0.f;
            v2[g] = 0.f;
            if (j < nnz)
            {
                v1[g] = V1[l[g]];
                if (useV2) { v2[g] = V2[l[g]]; }
            }
        }
#pragma unroll
        for (int g = 0; g < kSparseGroup; ++g)
        {
            const uint32_t j = base + g * kSparseThreads + tid;
            pred[g] = (j < nnz) && (useV2 ? (v1[g] != 0.f || v2[g] != 0.f) : (v1[g] != 0.f));
            ballot[g] = __ballot_sync(0xffffffffu, pred[g]);
            if (lane == 0) { warpCnt[g * (kSparseThreads / 32) + warp] = __popc(ballot[g]); }
        }
        __syncthreads();
        uint32_t before[kSparseGroup], total = 0;
#pragma unroll
        for (int g = 0; g < kSparseGroup; ++g)
        {
            before[g] = total;
#pragma unroll
            for (uint32_t w = 0; w < kSparseThreads / 32; ++w)
            {
                const uint32_t c = warpCnt[g * (kSparseThreads / 32) + w];
                before[g] += (w < warp) ? c : 0u;
                total += c;
            }
        }
#pragma unroll
        for (int g = 0; g < kSparseGroup; ++g)
        {
            if (pred[g])
            {
                const uint32_t p = before[g] + __popc(ballot[g] & ((1u << lane) - 1u));
                sIdx[p] = l[g];
                sD[p] = d[g];
                sV1[p] = v1[g];
                sV2[p] = v2[g];
            }
        }
        __syncthreads();
        // element e = visited + i goes to lane e % 256: a lane takes its elements in increasing e
        const uint32_t first = (tid + kSparseThreads - (visited % kSparseThreads)) % kSparseThreads;
#pragma unroll
        for (int m = 0; m < kSparseGroup; ++m)
        {
            const uint32_t i = first + m * kSparseThreads;
            if (i < total)
            {
                const uint32_t el = sIdx[i];
                const float ed = sD[i], ev1 = sV1[i];
                const float *orow = mv.otherMrows + static_cast<size_t>(el) * mv.ldR;
                float dotv = 0.f;
                // ascending k, mul and add rounded separately; 16-byte gathers of the other factor's row
                for (uint32_t q = 0; q < k; q += 4)
                {
                    const float4 o4 = *reinterpret_cast<const float4*>(orow + q);
                    dotv = fadd(dotv, fmul(sRow[q], o4.x));
                    if (q + 1 < k) { dotv = fadd(dotv, fmul(sRow[q + 1], o4.y)); }
                    if (q + 2 < k) { dotv = fadd(dotv, fmul(sRow[q + 2], o4.z)); }
                    if (q + 3 < k) { dotv = fadd(dotv, fmul(sRow[q + 3], o4.w)); }
                }
                if (useV2)
                {
                    const float dRecip = fdiv(1.f, ed);
                    const float term1 = fsub(1.f, fmul(dRecip, dRecip));
                    const float vDiff = fsub(ev1, sV2[i]);
                    accS = fadd(accS, fmul(fmul(vDiff, vDiff), term1));
                    accMu = fadd(accMu, fmul(vDiff, fadd(fmul(dotv, term1), dRecip)));
                }
                else
                {
                    const float term1 = fdiv(ev1, ed);
                    const float term2 = fsub(ev1, fdiv(term1, ed));
                    accS = fadd(accS, fsub(fmul(term1, term1), fmul(ev1, ev1)));
                    accMu = fadd(accMu, fadd(term1, fmul(term2, dotv)));
                    if (withChange) { accMu = fadd(accMu, fmul(fmul(term2, orow[colA]), ch)); }
                }
            }
        }
        visited += total;
        __syncthreads(); // the compaction buffers are rewritten by the next group
    }
}

// smemRaw: [SparseSmem | pad to 256 B][sRow: ldR floats][sIdx][sD][sV1][sV2], 256 * kSparseGroup entries each
template <bool STREAM>
__device__ __forceinline__ bool sparse_task(const ModelView &mv, const float *erfT, const float *erfinvT, float annealingTemp, const TaskIn &in,
                                            unsigned char *smemRaw, DevOutcome *outp, unsigned long long *verWaitNs,
                                            uint32_t *commitFlags, unsigned long long *visitedTotal)
{
    SparseSmem *hdr = reinterpret_cast<SparseSmem*>(smemRaw);
    float *sRow = reinterpret_cast<float*>(smemRaw + 256);
    uint32_t *sIdx = reinterpret_cast<uint32_t*>(sRow + mv.ldR);
    float *sD = reinterpret_cast<float*>(sIdx + kSparseThreads * kSparseGroup);
    float *sV1 = sD + kSparseThreads * kSparseGroup;
    float *sV2 = sV1 + kSparseThreads * kSparseGroup;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const DevProposal &pr = in.pr;
    const uint32_t pi = in.pi, part = in.part;
    const uint32_t type = pr.type;
    const uint32_t r1 = pr.r1, c1 = pr.c1, r2 = pr.r2, c2 = pr.c2;
    const uint32_t variant = pr.variant;
    const uint32_t k = mv.k;
    const bool pairType = (type == 'M') || (type == 'E') || (type == kProbe && variant == 1);
    const bool twoRow = pairType && (r1 != r2);
    const bool useV2 = pairType && (r1 == r2);
    const bool withChange = (type == 'D') || (type == kProbe && variant == 2);
    const float ch = (type == 'D') ? -pr.m1 : pr.ch;
    const uint32_t row = part ? r2 : r1;
    const uint32_t colA = part ? c2 : c1;

    if (STREAM && tid == 0 && in.waitMask != 0u)
    {
        const unsigned long long t0 = global_timer_ns();
        if (in.waitMask & 1u) { while (ld_acquire_gpu_u32(mv.rowVersion + r1) != in.ver1) { } }
        if (in.waitMask & 2u) { while (ld_acquire_gpu_u32(mv.rowVersion + r2) != in.ver2) { } }
        if (verWaitNs) { *verWaitNs = global_timer_ns() - t0; }
    }
    if (STREAM) { __syncthreads(); }
    // own factor row (L2 loads: an earlier batch of this kernel may have written it)
    for (uint32_t i = tid; i < mv.ldR; i += kSparseThreads) { sRow[i] = __ldcg(mv.Mrows + static_cast<size_t>(row) * mv.ldR + i); }
    float M1 = 0.f, M2 = 0.f, C1 = 0.f, C2 = 0.f;
    int can1 = 0, can2 = 0;
    if (tid == 0 && type != kProbe)
    {
        M1 = ld_cg_f32(mv.Mrows + static_cast<size_t>(r1) * mv.ldR + c1);
        C1 = ld_cg_f32(mv.M + static_cast<size_t>(c1) * mv.ldM + r1);
        can1 = mv.otherColNonzero[c1];
        if (pairType)
        {
            M2 = ld_cg_f32(mv.Mrows + static_cast<size_t>(r2) * mv.ldR + c2);
            C2 = ld_cg_f32(mv.M + static_cast<size_t>(c2) * mv.ldM + r2);
            can2 = mv.otherColNonzero[c2];
        }
    }
    else if (tid == 32 && (type == 'D' || type == 'M' || type == 'B'))
    {
        Pcg r;
        r.state = pr.rng;
        hdr->preLog[0] = portable_logf(r.uniform());
        hdr->preLog[1] = portable_logf(r.uniform());
    }
    __syncthreads();
    if (tid == 64)
    {
        float bs, bmu;
        sparse_table_terms(mv, sRow, colA, c1, c2, useV2, withChange, ch, bs, bmu);
        hdr->baseS = bs;
        hdr->baseMu = bmu;
    }

    float accS = 0.f, accMu = 0.f;
    uint32_t visited = 0;
    sparse_scan_row(mv, hdr->warpCnt, sRow, sIdx, sD, sV1, sV2, row, colA, c2, useV2, withChange, ch, accS, accMu, visited);
    if (visitedTotal != nullptr && tid == 0) { atomicAdd(visitedTotal, static_cast<unsigned long long>(visited)); }
    // lanes -> warp -> CTA: the butterflies of the dense kernel
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1)
    {
        accS = fadd(accS, __shfl_xor_sync(0xffffffffu, accS, off));
        accMu = fadd(accMu, __shfl_xor_sync(0xffffffffu, accMu, off));
    }
    if (lane == 0)
    {
        hdr->warpS[warp] = accS;
        hdr->warpMu[warp] = accMu;
    }
    __syncthreads();
    bool owner = false;
    if (tid < 32)
    {
        float sS = (tid < kSparseThreads / 32) ? hdr->warpS[tid] : 0.f;
        float sMu = (tid < kSparseThreads / 32) ? hdr->warpMu[tid] : 0.f;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
        {
            sS = fadd(sS, __shfl_xor_sync(0xffffffffu, sS, off));
            sMu = fadd(sMu, __shfl_xor_sync(0xffffffffu, sMu, off));
        }
        if (tid == 0)
        {
            float s = useV2 ? fsub(hdr->baseS, sS) : fadd(hdr->baseS, sS);
            float mu = fadd(hdr->baseMu, sMu);
            s = fmul(s, mv.beta);
            mu = fmul(mu, mv.beta);
            bool decideHere = true;
            if (twoRow)
            {
                AlphaPair mine;
                mine.s = s;
                mine.s_mu = mu;
                mv.partials[pi * 2 + part] = mine;
                __threadfence();
                const uint32_t ticket = atomicAdd(&mv.tickets[pi], 1u);
                decideHere = (ticket == 1u);
                if (decideHere)
                {
                    __threadfence();
                    const volatile AlphaPair *o = &mv.partials[pi * 2 + (1u - part)];
                    const float os = o->s, omu = o->s_mu;
                    mv.tickets[pi] = 0u;
                    const float s1 = part ? os : s, s2 = part ? s : os;
                    const float mu1 = part ? omu : mu, mu2 = part ? mu : omu;
                    s = fadd(s1, s2);
                    mu = fsub(mu1, mu2);
                }
            }
            if (decideHere)
            {
                Verdict v;
                PreLog pre;
                pre.state0 = pr.rng;
                pre.logFirst = hdr->preLog[0];
                pre.logSecond = hdr->preLog[1];
                decide<true>(mv, erfT, erfinvT, annealingTemp, pr, part, twoRow, s, mu, M1, M2, C1, C2, can1, can2, pre, &v);
                *outp = v.out;
                owner = true;
                *commitFlags = v.dec.flags;
            }
        }
    }
    return owner;
}

// The sparse model's commit is the two factor elements decide() stored; this publishes them (deciding lane
// only, after the outcome has gone to the host): fence, then the row versions and the done count.
__device__ __forceinline__ void sparse_publish(const ModelView &mv, const TaskIn &in, uint32_t flags, unsigned long long *commitsDone)
{
    if ((flags & 7u) == 0u) { return; }
    const uint32_t row = in.part ? in.pr.r2 : in.pr.r1, otherRow = in.part ? in.pr.r1 : in.pr.r2;
    __threadfence(); // gpu scope: see commit_task
    if ((flags & 3u) != 0u) { atomicAdd(mv.rowVersion + row, 1u); }
    if ((flags & 4u) != 0u) { atomicAdd(mv.rowVersion + otherRow, 1u); }
    if (commitsDone != nullptr) { atomicAdd(commitsDone, 1ull); }
}

// One launch per conflict-free batch, sparse model.
__global__ void __launch_bounds__(kSparseThreads, 4) eval_sparse_kernel(const __grid_constant__ EvalParams P)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const uint32_t task = blockIdx.x;
    TaskIn in;
    in.pi = task < P.nProps ? task : static_cast<uint32_t>(P.extra[task - P.nProps]);
    in.part = task < P.nProps ? 0u : 1u;
    in.ver1 = in.ver2 = in.waitMask = 0u;
    in.pr = P.props[in.pi];
    DevOutcome out;
    uint32_t flags = 0u;
    if (sparse_task<false>(P.mv, P.mv.erf, P.mv.erfinv, P.mv.annealingTemp, in, smemRaw, &out, nullptr, &flags, nullptr))
    {
        P.mv.outcomes[in.pi] = out;
        sparse_publish(P.mv, in, flags, nullptr);
    }
}

// One launch per conflict-free batch; proposals travel in kernel-parameter space.
template <bool HAS_S>
__global__ void __launch_bounds__(kThreads, 2) eval_kernel(const __grid_constant__ EvalParams P)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    EvalSmem *hdr = reinterpret_cast<EvalSmem*>(smemRaw);
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    const uint32_t task = blockIdx.y;
    TaskIn in;
    in.pi = task < P.nProps ? task : static_cast<uint32_t>(P.extra[task - P.nProps]);
    in.part = task < P.nProps ? 0u : 1u;
    in.ver1 = in.ver2 = in.waitMask = 0u;
    in.pr = P.props[in.pi];
    if (threadIdx.x == 0)
    {
        mbar_init(&hdr->bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    DevOutcome out;
    if (process_task<HAS_S, false>(P.mv, P.mv.erf, P.mv.erfinv, P.mv.annealingTemp, in, task, smemRaw, 0u, cluster, rank, &out, nullptr))
    {
        P.mv.outcomes[in.pi] = out;
    }
    commit_task<HAS_S>(P.mv, in, task, smemRaw, cluster, rank, nullptr);
}

// Bulk probes: any number of single-row alphaParameters queries in ONE launch, proposals and results in device
// memory (cgb_sampler_alpha_parameters).  Same staging + scan + reduce as every proposal; with thousands of tasks
// queued behind each SM this is the launch that shows what the scan sustains when work is not the limit.
struct ProbeParams
{
    ModelView mv;
    const DevProposal *props;
    DevOutcome *outs;
};

template <bool HAS_S>
__global__ void __launch_bounds__(kThreads, 2) probe_kernel(const __grid_constant__ ProbeParams P)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    EvalSmem *hdr = reinterpret_cast<EvalSmem*>(smemRaw);
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    TaskIn in;
    in.pi = blockIdx.y;
    in.part = 0u;
    in.ver1 = in.ver2 = in.waitMask = 0u;
    in.pr = P.props[in.pi];
    if (threadIdx.x == 0)
    {
        mbar_init(&hdr->bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    DevOutcome out;
    if (process_task<HAS_S, false>(P.mv, P.mv.erf, P.mv.erfinv, P.mv.annealingTemp, in, 0u, smemRaw, 0u, cluster, rank, &out, nullptr))
    {
        P.outs[in.pi] = out;
    }
    if (P.mv.nSeg > 1) { cluster.sync(); } // nobody leaves while a peer may still write its shared memory
}

// ------------------------------------------------------------------------------------------------
// Resident ("streaming") variant: launched once per update().  The host generator writes each proposal
// into its cluster's ring of 64-byte task records in pinned memory THE MOMENT it is generated; every
// worker cluster polls its own ring across PCIe (one round trip from "posted" to "running", no grid-wide
// release), so the evaluation of a batch overlaps its generation.  Outcomes go back as self-tagged
// 16-byte records the host spins on, posted BEFORE the AP commit; cross-batch ordering is per row
// (rowVersion + a count of finished commits that one extra CTA mirrors to the host), not a grid barrier.
// Records of cluster c: slots[(c * kStreamRing + (ticket - 1) % kStreamRing) * nSeg + rank].
// ------------------------------------------------------------------------------------------------
struct HostOutcome   // one 16-byte store, self-validating for the polling host
{
    uint32_t mass1Bits, mass2Bits;
    uint32_t seqAndAccepted; // (chunk tag << 3) | (rng draws made << 1) | accepted
    uint32_t check;          // outcome_check of the three words above
};

__host__ __device__ __forceinline__ uint32_t outcome_check(uint32_t w0, uint32_t w1, uint32_t w2)
{
    return (w0 * 0x9E3779B1u) ^ (w1 * 0x85EBCA77u) ^ (w2 * 0xC2B2AE3Du) ^ 0x27D4EB2Fu;
}

// checksum of a StreamRecord: word i weighted by an odd constant, word 14 (the checksum itself) skipped
__host__ __device__ __forceinline__ uint32_t stream_weight(uint32_t i) { return 0x9E3779B1u * (2u * i + 1u); }
__host__ __device__ __forceinline__ uint32_t stream_check(const uint32_t *w)
{
    uint32_t h = 0x27D4EB2Fu;
    for (uint32_t i = 0; i < 16; ++i)
    {
        if (i != 14) { h ^= w[i] * stream_weight(i); }
    }
    return h;
}

// Mailbox reads must never be served from L1: the same addresses carry a new record every few
// microseconds.  Host memory: ld.volatile (system scope, uncached).
__device__ __forceinline__ uint4 ld_host_u4(const void *p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

static const unsigned long long kDoorbellExit = 1ull << 63;

struct StreamParams
{
    const StreamRecord *slots;                 // pinned host memory: record rings of the worker clusters
    const volatile unsigned long long *doorbell; // pinned host memory: kDoorbellExit retires the mirror CTA
    HostOutcome *outcomes;                     // pinned host memory
    volatile unsigned long long *commitsMirror; // pinned host memory [2]: stats->commitsDone as last seen by the mirror CTA
    StreamStats *stats;
    uint32_t *alive;                           // pinned host memory [nWorkers]: a cluster writes `epoch` here when it starts
    uint32_t ticket0;                          // every cluster's first ticket is ticket0 + 1
    uint32_t epoch;
    unsigned long long idleTimeoutNs;
    uint32_t nWorkers;                         // worker clusters; cluster nWorkers only mirrors the commit count
    uint32_t pollSleepNs;
};

// The mirror CTA tells the host how many CTA-commits are complete.  acquire on the device counter,
// release towards the host: whoever learns the count from the host may read the rows those commits wrote.
__device__ __forceinline__ void mirror_loop(const StreamParams &sp)
{
    if (threadIdx.x != 0) { return; }
    unsigned long long mirrored[2] = {0ull, 0ull};
    unsigned long long tLast = global_timer_ns(); // last time a counter moved
    for (uint32_t it = 0;; ++it)
    {
        bool changed = false;
#pragma unroll
        for (int q = 0; q < 2; ++q)
        {
            unsigned long long done;
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(done) : "l"(&sp.stats->commitsDone[q]) : "memory");
            if (done != mirrored[q])
            {
                __threadfence_system();
                asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(sp.commitsMirror + q), "l"(done) : "memory");
                mirrored[q] = done;
                changed = true;
            }
        }
        if (!changed) { __nanosleep(100); }
        if ((it & 15u) == 0u)
        {
            // the host retires us at the end of update().  Without the host (a profiler or CUDA_LAUNCH_BLOCKING holds
            // it inside the launch call until the grid has ended) the workers leave after their idle timeout and
            // nothing will ever commit again: leave shortly after them instead of holding the device.
            unsigned long long bell;
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(bell) : "l"(sp.doorbell) : "memory");
            if (bell & kDoorbellExit) { break; }
            const unsigned long long now = global_timer_ns();
            if (changed) { tLast = now; }
            if (now - tLast > sp.idleTimeoutNs + sp.idleTimeoutNs / 2ull) { break; }
        }
    }
}

// MODE 0: dense model, default uncertainty; 1: dense with an S matrix; 2: sparse model
template <int MODE>
__device__ __forceinline__ void stream_worker(const ModelView &mv, const StreamParams &sp, unsigned char *smemRaw)
{
    constexpr bool HAS_S = (MODE == 1);
    constexpr bool SPARSE = (MODE == 2);
    EvalSmem *hdr = reinterpret_cast<EvalSmem*>(smemRaw);
    __shared__ __align__(16) StreamRecord sRec;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = SPARSE ? 0u : cluster.block_rank();
    const uint32_t nSeg = mv.nSeg;
    const uint32_t clusterId = blockIdx.x / nSeg;
    const uint32_t tid = threadIdx.x;
    if (clusterId == sp.nWorkers)
    {
        if (rank == 0) { mirror_loop(sp); }
        return;
    }
    // lookup tables of the epilogue live in shared memory for the whole update(), behind the staging area
    const size_t stageFloats = SPARSE ? (static_cast<size_t>(mv.ldR) + 4u * kSparseThreads * kSparseGroup)
                                      : static_cast<size_t>(HAS_S ? 5 : 4) * mv.segPad;
    const float *erfS = mv.erf, *erfinvS = mv.erfinv;
    if (mv.tablesInSmem != 0u)
    {
        float *e = reinterpret_cast<float*>(smemRaw + 256) + stageFloats;
        float *ei = e + ((CGB_ERF_TABLE_SIZE + 3) & ~3);
        for (uint32_t i = tid; i < CGB_ERF_TABLE_SIZE; i += blockDim.x) { e[i] = mv.erf[i]; }
        for (uint32_t i = tid; i < CGB_ERFINV_TABLE_SIZE; i += blockDim.x) { ei[i] = mv.erfinv[i]; }
        erfS = e;
        erfinvS = ei;
    }
    if (!SPARSE && tid == 0)
    {
        mbar_init(&hdr->bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const uint32_t ticket0 = sp.ticket0;
    // report in: the host hands work only to clusters it knows to be running (all CTAs of a cluster are scheduled
    // together, so the leader speaks for them)
    if (rank == 0 && tid == 0)
    {
        asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(sp.alive + clusterId), "r"(sp.epoch) : "memory");
    }
    uint32_t parity = 0;
    for (uint32_t j = 0;; ++j)
    {
        const uint32_t want = ticket0 + j + 1u;
        const StreamRecord *slot = sp.slots + (static_cast<size_t>(clusterId) * kStreamRing + ((want - 1u) % kStreamRing)) * nSeg + rank;
        // ---- warp 0 polls the slot: 4 lanes x 16 B, valid when the ticket matches and the checksum holds ----
        unsigned long long tSeen = 0;
        if (tid < 32)
        {
            const unsigned long long t0 = global_timer_ns();
            bool ok = false, dead = false;
            uint4 w = make_uint4(0u, 0u, 0u, 0u);
            while (!ok && !dead)
            {
                if (tid < 4) { w = ld_host_u4(reinterpret_cast<const uint4*>(slot) + tid); }
                uint32_t h = 0u;
                if (tid < 4)
                {
                    h = (w.x * stream_weight(4u * tid)) ^ (w.y * stream_weight(4u * tid + 1u)) ^ (w.w * stream_weight(4u * tid + 3u));
                    if (tid != 3) { h ^= w.z * stream_weight(4u * tid + 2u); }
                }
                h ^= __shfl_xor_sync(0xffffffffu, h, 1);
                h ^= __shfl_xor_sync(0xffffffffu, h, 2);
                h = __shfl_sync(0xffffffffu, h, 0) ^ 0x27D4EB2Fu;
                const uint32_t ticket = __shfl_sync(0xffffffffu, w.y, 3);
                const uint32_t check = __shfl_sync(0xffffffffu, w.z, 3);
                ok = (ticket == want) && (check == h);
                if (!ok)
                {
                    dead = __shfl_sync(0xffffffffu, (global_timer_ns() - t0 > sp.idleTimeoutNs) ? 1 : 0, 0) != 0;
                    if (sp.pollSleepNs) { __nanosleep(sp.pollSleepNs); }
                }
            }
            if (tid < 4)
            {
                if (dead) { w.x = kStreamExit; } // lane 2's w.x is the type word
                reinterpret_cast<uint4*>(&sRec)[tid] = w;
            }
            if (tid == 0) { tSeen = global_timer_ns(); }
        }
        __syncthreads();
        if (sRec.type == kStreamExit) { break; }
        TaskIn in;
        in.pr.rng = sRec.rng;
        in.pr.r1 = sRec.r1; in.pr.c1 = sRec.c1; in.pr.r2 = sRec.r2; in.pr.c2 = sRec.c2;
        in.pr.m1 = sRec.m1; in.pr.m2 = sRec.m2;
        in.pr.type = sRec.type & 0xffu;
        in.waitMask = (sRec.type >> 8) & 3u;
        const uint32_t seqFlag = (sRec.type >> 10) & 1u;
        in.pr.variant = 0u;
        in.pr.ch = 0.f;
        in.pr.pad = seqFlag;
        in.pi = sRec.piPart & 0x7fffffffu;
        in.part = sRec.piPart >> 31;
        in.ver1 = sRec.ver1;
        in.ver2 = sRec.ver2;
        const uint32_t batch = sRec.batch;
        const uint32_t task = (in.pi % kMaxBatch) + in.part * kMaxBatch; // debug phase-clock slot
        DevOutcome out;
        unsigned long long verWait = 0;
        uint32_t sparseFlags = 0u;
        bool owner;
        if (SPARSE) { owner = sparse_task<true>(mv, erfS, erfinvS, mv.annealingTemp, in, smemRaw, &out, &verWait, &sparseFlags, &sp.stats->visited); }
        else { owner = process_task<HAS_S, true>(mv, erfS, erfinvS, mv.annealingTemp, in, task, smemRaw, parity, cluster, rank, &out, &verWait); }
        unsigned long long tPosted = 0;
        if (owner)
        {
            // one 16-byte store across PCIe; word 2 carries the chunk tag, word 3 a checksum of the
            // other three, so the polling host can tell a complete record from a stale or torn one
            const uint32_t w0 = __float_as_uint(out.mass1), w1 = __float_as_uint(out.mass2);
            const uint32_t w2 = (batch << 3) | ((out.pad[0] & 3u) << 1) | (out.accepted & 1u);
            const uint32_t w3 = outcome_check(w0, w1, w2);
            asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};"
                         ::"l"(sp.outcomes + in.pi), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
            tPosted = global_timer_ns();
        }
        if (SPARSE)
        {
            if (owner) { sparse_publish(mv, in, sparseFlags, &sp.stats->commitsDone[batch & 1u]); }
        }
        else
        {
            commit_task<HAS_S>(mv, in, task, smemRaw, cluster, rank, &sp.stats->commitsDone[batch & 1u]);
        }
        parity ^= 1u;
        __syncthreads(); // staging buffers and sRec are free for the next task
        if (tid == 0 && rank == 0)
        {
            const unsigned long long dt = global_timer_ns() - tSeen;
            atomicAdd(&sp.stats->taskNs, dt);
            atomicAdd(&sp.stats->tasks, 1ull);
            atomicMax(&sp.stats->maxTaskNs, dt);
            atomicAdd(&sp.stats->verWaitNs, verWait);
            if (owner)
            {
                atomicAdd(&sp.stats->decideNs, tPosted - tSeen);
                atomicAdd(&sp.stats->outcomes, 1ull);
            }
        }
    }
}

template <bool HAS_S>
__global__ void __launch_bounds__(kThreads, 2) eval_stream_kernel(const __grid_constant__ ModelView mv,
                                                               const __grid_constant__ StreamParams sp)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    if (HAS_S) { stream_worker<1>(mv, sp, smemRaw); } else { stream_worker<0>(mv, sp, smemRaw); }
}

__global__ void __launch_bounds__(kSparseThreads, 4) eval_stream_sparse_kernel(const __grid_constant__ ModelView mv,
                                                                            const __grid_constant__ StreamParams sp)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    stream_worker<2>(mv, sp, smemRaw);
}

// ------------------------------------------------------------------------------------------------
// compressed rows -> the zero-filled dense copy (one CTA per sampler row; D was cleared beforehand).  Used when the
// data arrives as Matrix-Market triplets and never exists as a dense matrix on the host.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) csr_scatter_kernel(const uint32_t *__restrict__ rowPtr, const uint32_t *__restrict__ idx,
                                                          const float *__restrict__ val, uint32_t ld, float *__restrict__ D)
{
    const uint32_t r = blockIdx.x;
    const uint32_t b = rowPtr[r], e = rowPtr[r + 1];
    float *row = D + static_cast<size_t>(r) * ld;
    for (uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) { row[idx[i]] = val[i]; }
}

// ------------------------------------------------------------------------------------------------
// sync: dst[r][l] = src[l][r]  (DenseNormalModel::sync, DenseNormalModel.cpp:20-36)
// ------------------------------------------------------------------------------------------------
// 64 x 64 tiles, 16-byte loads and stores on both sides (rows are padded to 32 floats, so every tile row starts on a
// 16-byte boundary); tiles that reach over an edge of the matrix go element by element and never write a pad column.
static const int kTransposeTile = 64;

__global__ void __launch_bounds__(256) transpose_kernel(float *__restrict__ dst, const float *__restrict__ src,
                                                        uint32_t dstRows, uint32_t dstCols, uint32_t ldDst,
                                                        uint32_t ldSrc)
{
    __shared__ float tile[kTransposeTile][kTransposeTile + 1]; // [src row in tile][src col in tile]
    const uint32_t bx = blockIdx.x * kTransposeTile, by = blockIdx.y * kTransposeTile; // bx: dst col block, by: dst row block
    const uint32_t t = threadIdx.x;
    const bool interior = (bx + kTransposeTile <= dstCols) && (by + kTransposeTile <= dstRows);
    if (interior)
    {
        // src is [dstCols][dstRows]: tile row q = src row bx + q, 16 vectors of 4 along src columns by ..
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const uint32_t idx = t + 256u * i, q = idx >> 4, c4 = idx & 15u;
            const float4 v = __ldcs(reinterpret_cast<const float4*>(src + static_cast<size_t>(bx + q) * ldSrc + by) + c4);
            tile[q][c4 * 4u + 0u] = v.x;
            tile[q][c4 * 4u + 1u] = v.y;
            tile[q][c4 * 4u + 2u] = v.z;
            tile[q][c4 * 4u + 3u] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const uint32_t idx = t + 256u * i, r = idx >> 4, c4 = idx & 15u; // dst row by + r, dst cols bx + 4 c4 ..
            float4 v;
            v.x = tile[c4 * 4u + 0u][r];
            v.y = tile[c4 * 4u + 1u][r];
            v.z = tile[c4 * 4u + 2u][r];
            v.w = tile[c4 * 4u + 3u][r];
            reinterpret_cast<float4*>(dst + static_cast<size_t>(by + r) * ldDst + bx)[c4] = v;
        }
        return;
    }
    for (uint32_t idx = t; idx < kTransposeTile * kTransposeTile; idx += 256u)
    {
        const uint32_t q = idx / kTransposeTile, c = idx % kTransposeTile;
        const uint32_t srcRow = bx + q, srcCol = by + c;
        tile[q][c] = (srcRow < dstCols && srcCol < dstRows) ? src[static_cast<size_t>(srcRow) * ldSrc + srcCol] : 0.f;
    }
    __syncthreads();
    for (uint32_t idx = t; idx < kTransposeTile * kTransposeTile; idx += 256u)
    {
        const uint32_t r = idx / kTransposeTile, c = idx % kTransposeTile;
        const uint32_t dRow = by + r, dCol = bx + c;
        if (dRow < dstRows && dCol < dstCols) { dst[static_cast<size_t>(dRow) * ldDst + dCol] = tile[c][r]; }
    }
}

// ------------------------------------------------------------------------------------------------
// extraInitialization: AP[r][l] = sum_c other[c][l] * M[c][r], c ascending, mul and add rounded
// separately exactly like the reference's scalar triple loop (DenseNormalModel.cpp:38-54)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rebuild_ap_kernel(float *__restrict__ AP, const float *__restrict__ M,
                                                         const float *__restrict__ otherM, uint32_t nRows,
                                                         uint32_t L, uint32_t k, uint32_t ld, uint32_t ldM,
                                                         uint32_t ldOther)
{
    // one sampler row per blockIdx.x (rows can exceed the 65535 limit of the other grid dimensions), 1024 elements of
    // it per block: 16-byte loads of the other factor's columns, the row's k factor elements broadcast from registers
    const uint32_t r = blockIdx.x;
    const uint32_t l = (blockIdx.y * blockDim.x + threadIdx.x) * 4u;
    if (r >= nRows || l >= L) { return; }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t c = 0; c < k; ++c)
    {
        const float m = __ldg(M + static_cast<size_t>(c) * ldM + r);
        const float4 o = __ldg(reinterpret_cast<const float4*>(otherM + static_cast<size_t>(c) * ldOther + l)); // padded with zeros
        acc.x = fadd(acc.x, fmul(o.x, m));
        acc.y = fadd(acc.y, fmul(o.y, m));
        acc.z = fadd(acc.z, fmul(o.z, m));
        acc.w = fadd(acc.w, fmul(o.w, m));
    }
    reinterpret_cast<float4*>(AP + static_cast<size_t>(r) * ld + l)[0] = acc; // the pad of the line gets exact zeros
}

// ------------------------------------------------------------------------------------------------
// chiSq = sum ((D - AP) / S)^2 (DenseNormalModel.cpp:56-68).  Terms in fp32 exactly as the reference
// forms them; accumulation in f64 in a fixed order (per-thread, warp tree, block, then the host adds
// the per-block partials in index order) — deterministic and more accurate than the reference's
// single fp32 running sum.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chisq_kernel(const float *__restrict__ D, const float *__restrict__ S,
                                                    const float *__restrict__ AP, uint32_t nRows, uint32_t L,
                                                    uint32_t ld, double *__restrict__ partials)
{
    // 16-byte loads; the lines are padded (D and AP with zeros, a term of exactly 0) so whole vectors are read.  A block
    // walks rows blockIdx.x, + gridDim.x, ...; a thread's f64 sum takes its elements in that fixed order.
    __shared__ double warpSum[8];
    double acc = 0.0;
    const uint32_t nVec = (L + 3u) / 4u;
    for (uint32_t r = blockIdx.x; r < nRows; r += gridDim.x)
    {
        const float4 *d = reinterpret_cast<const float4*>(D + static_cast<size_t>(r) * ld);
        const float4 *a = reinterpret_cast<const float4*>(AP + static_cast<size_t>(r) * ld);
        const float4 *s = S ? reinterpret_cast<const float4*>(S + static_cast<size_t>(r) * ld) : nullptr;
#pragma unroll 2
        for (uint32_t j = threadIdx.x; j < nVec; j += blockDim.x)
        {
            const float4 dv = __ldcs(d + j), av = __ldcs(a + j);
            float4 sv;
            if (s) { sv = __ldcs(s + j); }
            else { sv = make_float4(derive_s(dv.x), derive_s(dv.y), derive_s(dv.z), derive_s(dv.w)); }
            const uint32_t base = j * 4u;
            const float t0 = fdiv(fsub(dv.x, av.x), sv.x), t1 = fdiv(fsub(dv.y, av.y), sv.y);
            const float t2 = fdiv(fsub(dv.z, av.z), sv.z), t3 = fdiv(fsub(dv.w, av.w), sv.w);
            acc += static_cast<double>(fmul(t0, t0));
            if (base + 1u < L) { acc += static_cast<double>(fmul(t1, t1)); }
            if (base + 2u < L) { acc += static_cast<double>(fmul(t2, t2)); }
            if (base + 3u < L) { acc += static_cast<double>(fmul(t3, t3)); }
        }
    }
    for (int off = 16; off >= 1; off >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, off); }
    if ((threadIdx.x & 31) == 0) { warpSum[threadIdx.x >> 5] = acc; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) { t += warpSum[w]; }
        partials[blockIdx.x] = t;
    }
}

// ------------------------------------------------------------------------------------------------
// SparseNormalModel::generateLookupTables (SparseNormalModel.cpp:294-311): Z1[i] = sum_r rows[r][i]^2
// (row copy), Z2[i][j] = sum_r cols[i][r] cols[j][r] (column copy) of the OTHER factor.  Element r goes to
// lane r % 256, lanes add in order, butterflies combine them (oracle mode "device").  Block b < k gives
// Z1[b]; the others give the pairs (i <= j) in row-major order of the upper triangle.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSparseThreads) sparse_tables_kernel(const float *__restrict__ Mrows, uint32_t ldR,
                                                                       const float *__restrict__ M, uint32_t ldM,
                                                                       uint32_t nRows, uint32_t k,
                                                                       float *__restrict__ Z1, float *__restrict__ Z2)
{
    __shared__ float warpSum[kSparseThreads / 32];
    const uint32_t tid = threadIdx.x;
    float acc = 0.f;
    uint32_t zi = 0, zj = 0;
    const bool isZ1 = blockIdx.x < k;
    if (isZ1)
    {
        zi = blockIdx.x;
        for (uint32_t r = tid; r < nRows; r += kSparseThreads)
        {
            const float v = Mrows[static_cast<size_t>(r) * ldR + zi];
            acc = fadd(acc, fmul(v, v));
        }
    }
    else
    {
        uint32_t p = blockIdx.x - k; // index into the upper triangle, rows of length k, k-1, ...
        while (p >= k - zi)
        {
            p -= k - zi;
            ++zi;
        }
        zj = zi + p;
        const float *a = M + static_cast<size_t>(zi) * ldM, *b = M + static_cast<size_t>(zj) * ldM;
        for (uint32_t r = tid; r < nRows; r += kSparseThreads) { acc = fadd(acc, fmul(a[r], b[r])); }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { acc = fadd(acc, __shfl_xor_sync(0xffffffffu, acc, off)); }
    if ((tid & 31) == 0) { warpSum[tid >> 5] = acc; }
    __syncthreads();
    if (tid < 32)
    {
        float t = (tid < kSparseThreads / 32) ? warpSum[tid] : 0.f;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) { t = fadd(t, __shfl_xor_sync(0xffffffffu, t, off)); }
        if (tid == 0)
        {
            if (isZ1) { Z1[zi] = t; }
            else
            {
                Z2[static_cast<size_t>(zi) * k + zj] = t;
                Z2[static_cast<size_t>(zj) * k + zi] = t;
            }
        }
    }
}

// SparseNormalModel::chiSq (SparseNormalModel.cpp:39-60): sum over every (row j, scan index i) of dot^2,
// plus 1 + dot (dot - 2 d - d^2 dot) / d^2 where d > 0, dot = A[j,:] . P[i,:]; times beta on the host.
// Terms in fp32 as the reference forms them, f64 accumulation in a fixed order.
__global__ void __launch_bounds__(kSparseThreads) sparse_chisq_kernel(const float *__restrict__ D, uint32_t ld,
                                                                      const float *__restrict__ Mrows,
                                                                      const float *__restrict__ otherMrows, uint32_t ldR,
                                                                      uint32_t nRows, uint32_t L, uint32_t k,
                                                                      double *__restrict__ partials)
{
    extern __shared__ float sOwn[];
    __shared__ double warpSum[kSparseThreads / 32];
    double acc = 0.0;
    for (uint32_t j = blockIdx.x; j < nRows; j += gridDim.x)
    {
        __syncthreads();
        for (uint32_t q = threadIdx.x; q < ldR; q += kSparseThreads) { sOwn[q] = Mrows[static_cast<size_t>(j) * ldR + q]; }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < L; i += kSparseThreads)
        {
            const float dot = row_dot(sOwn, otherMrows + static_cast<size_t>(i) * ldR, k);
            acc += static_cast<double>(fmul(dot, dot));
            const float d = D[static_cast<size_t>(j) * ld + i];
            if (d > 0.f)
            {
                const float dsq = fmul(d, d);
                const float t = fdiv(fmul(dot, fsub(fsub(dot, fmul(2.f, d)), fmul(dsq, dot))), dsq);
                acc += static_cast<double>(fadd(1.f, t));
            }
        }
    }
    for (int off = 16; off >= 1; off >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, off); }
    if ((threadIdx.x & 31) == 0) { warpSum[threadIdx.x >> 5] = acc; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double t = 0.0;
        for (int w = 0; w < kSparseThreads / 32; ++w) { t += warpSum[w]; }
        partials[blockIdx.x] = t;
    }
}

// canUseGibbs(col) == "column col of the factor matrix has a positive entry" (DenseNormalModel.cpp:100-108)
__global__ void __launch_bounds__(256) col_nonzero_kernel(const float *__restrict__ M, uint32_t nRows, uint32_t ldM,
                                                          int *__restrict__ flags)
{
    const float *col = M + static_cast<size_t>(blockIdx.x) * ldM;
    int any = 0;
    for (uint32_t i = threadIdx.x; i < nRows; i += blockDim.x) { any |= (col[i] > 0.f) ? 1 : 0; }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) { flags[blockIdx.x] = any; }
}

// gaps::max(Vector) per pattern (VectorMath.cpp; starts from 0), then norm = (max == 0) ? 1 : max
__global__ void __launch_bounds__(256) col_max_kernel(const float *__restrict__ M, uint32_t nRows, uint32_t ldM,
                                                      float *__restrict__ norms, int forceOne)
{
    __shared__ float warpMax[8];
    const float *col = M + static_cast<size_t>(blockIdx.x) * ldM;
    float mx = 0.f;
    for (uint32_t i = threadIdx.x; i < nRows; i += blockDim.x) { mx = (col[i] > mx) ? col[i] : mx; }
    for (int off = 16; off >= 1; off >>= 1)
    {
        const float o = __shfl_xor_sync(0xffffffffu, mx, off);
        mx = (o > mx) ? o : mx;
    }
    if ((threadIdx.x & 31) == 0) { warpMax[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 0; w < 8; ++w) { mx = (warpMax[w] > mx) ? warpMax[w] : mx; }
        mx = (mx == 0.f) ? 1.f : mx;
        norms[blockIdx.x] = forceOne ? 1.f : mx;
    }
}

// GapsStatistics::update (GapsStatistics.h:129-149): mean += x, sq += x*x with x = M/norm (divide != 0)
// or x = M*norm
__global__ void __launch_bounds__(256) stats_accumulate_kernel(const float *__restrict__ M, uint32_t nRows,
                                                               uint32_t ldM, const float *__restrict__ norms,
                                                               int divide, float *__restrict__ meanSum,
                                                               float *__restrict__ sqSum)
{
    const uint32_t c = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRows) { return; }
    const size_t idx = static_cast<size_t>(c) * ldM + i;
    const float nrm = norms[c];
    const float x = divide ? fdiv(M[idx], nrm) : fmul(M[idx], nrm);
    meanSum[idx] = fadd(meanSum[idx], x);
    sqSum[idx] = fadd(sqSum[idx], fmul(x, x));
}

// pumpMatrix*Threshold (GapsStatistics.h:66-117): per row, the first pattern holding the maximum
__global__ void __launch_bounds__(256) pump_kernel(const float *__restrict__ M, uint32_t nRows, uint32_t ldM,
                                                   uint32_t k, float divisor, float *__restrict__ pump)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRows) { return; }
    float mx = 0.f;
    uint32_t arg = 0;
    for (uint32_t c = 0; c < k; ++c)
    {
        const float v = fdiv(M[static_cast<size_t>(c) * ldM + i], divisor);
        if (mx < v)
        {
            mx = v;
            arg = c;
        }
    }
    pump[static_cast<size_t>(arg) * ldM + i] += 1.f;
}

// meanChiSq (GapsStatistics.cpp:63-87): m = (sum_c Asum[i,c] Psum[j,c]) / n^2 in fp32, c ascending;
// terms (d-m)^2 / s^2 in fp32; f64 accumulation as in chisq_kernel.  D is the P sampler's copy: [S][ld].
__global__ void __launch_bounds__(256) mean_chisq_kernel(const float *__restrict__ D, const float *__restrict__ S,
                                                         uint32_t nSamples, uint32_t nGenes, uint32_t ld,
                                                         const float *__restrict__ Asum, uint32_t ldA,
                                                         const float *__restrict__ Psum, uint32_t ldP, uint32_t k,
                                                         float nSq, double *__restrict__ partials)
{
    __shared__ double warpSum[8];
    double acc = 0.0;
    for (uint32_t j = blockIdx.x; j < nSamples; j += gridDim.x)
    {
        const float *d = D + static_cast<size_t>(j) * ld;
        const float *s = S ? S + static_cast<size_t>(j) * ld : nullptr;
        for (uint32_t i = threadIdx.x; i < nGenes; i += blockDim.x)
        {
            float m = 0.f;
            for (uint32_t c = 0; c < k; ++c)
            {
                m = fadd(m, fmul(Asum[static_cast<size_t>(c) * ldA + i], Psum[static_cast<size_t>(c) * ldP + j]));
            }
            m = fdiv(m, nSq);
            const float dv = d[i];
            const float sv = s ? s[i] : derive_s(dv);
            const float diff = fsub(dv, m);
            acc += static_cast<double>(fdiv(fmul(diff, diff), fmul(sv, sv)));
        }
    }
    for (int off = 16; off >= 1; off >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, off); }
    if ((threadIdx.x & 31) == 0) { warpSum[threadIdx.x >> 5] = acc; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) { t += warpSum[w]; }
        partials[blockIdx.x] = t;
    }
}

// known-answer probe of the device portable log
__global__ void logf_probe_kernel(const float *__restrict__ in, float *__restrict__ out, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[i] = portable_logf(in[i]); }
}

} // namespace cgb

#endif // CGB_KERNELS_CUH
