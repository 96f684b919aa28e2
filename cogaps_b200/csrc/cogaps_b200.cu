// cogaps_b200.cu — the C ABI (include/cogaps_b200.h): device-resident sampler state, kernel launches,
// statistics and the gaps::run loop.  There is no CPU fallback anywhere in this file: every entry
// point that computes needs an sm_100-class device and fails with CGB_ENODEVICE / CGB_ECUDA otherwise.
#include "../../include/cogaps_b200.h"
#include "kernels.cuh"
#include "sweep.cuh"
#include "sampler.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>

using namespace cgb;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_lastError;
static std::atomic<uint64_t> g_kernelLaunches(0);
static std::atomic<int> g_residentShare(1); // chains that share the device: each resident grid takes 1/share of it
static int g_device = -1;
// Set once a tool that serialises kernel launches (Nsight Compute, compute-sanitizer, CUDA_LAUNCH_BLOCKING=1) is known to
// be attached: a resident grid that converses with the host cannot run under one, so samplers use one launch per batch.
static std::atomic<int> g_serialisedLaunches(-1); // -1 unknown, 0 no, 1 yes

static int fail(int code, const std::string &msg)
{
    g_lastError = msg;
    return code;
}

// Nothing may leave the library as a C++ exception (include/cogaps_b200.h: status codes only): every entry point that
// can allocate or start a thread runs its body through this.
template <class Body>
static int guarded(const char *who, Body body) noexcept
{
    try
    {
        return body();
    }
    catch (const std::bad_alloc &)
    {
        return fail(CGB_ENOMEM, std::string(who) + ": out of host memory");
    }
    catch (const std::exception &e)
    {
        return fail(CGB_EINTERNAL, std::string(who) + ": " + e.what());
    }
    catch (...)
    {
        return fail(CGB_EINTERNAL, std::string(who) + ": unknown exception");
    }
}

#define CGB_CUDA(call)                                                                                  \
    do                                                                                                  \
    {                                                                                                   \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
        {                                                                                               \
            const int code__ = (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver)         \
                ? CGB_ENODEVICE : (e__ == cudaErrorMemoryAllocation ? CGB_ENOMEM : CGB_ECUDA);          \
            return fail(code__, std::string(#call) + ": " + cudaGetErrorString(e__));                   \
        }                                                                                               \
    } while (0)

#define CGB_CHECK(cond, msg)                                       \
    do                                                             \
    {                                                              \
        if (!(cond)) { return fail(CGB_EINVAL, msg); }             \
    } while (0)

#define CGB_TRY(expr)                      \
    do                                     \
    {                                      \
        int rc__ = (expr);                 \
        if (rc__ != CGB_OK) { return rc__; } \
    } while (0)

// optional host-side split of update() (COGAPS_HOST_PROFILE=1): TSC ticks in posting, applying, flushing
static int g_hostProfile = -1;   // env value; g_prof is set for the sampled batches only (every 16th)
static bool g_prof = false;
static unsigned long long g_profBatches = 0, g_profCounter = 0, g_tscWaitTail = 0;
static unsigned long long g_tscPost = 0, g_tscApply = 0, g_tscFlush = 0, g_tscPopulate = 0, g_nPosts = 0;
static inline unsigned long long tsc() { return __builtin_ia32_rdtsc(); }

static inline double nowSeconds()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int envInt(const char *name, int dflt)
{
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}


// A resident grid needs the host to keep running while it is on the device.  Tools that hold the host inside the launch
// call until the kernel has ended make that impossible; they announce themselves through the environment.
static bool serialisingToolAnnounced()
{
    const char *blocking = std::getenv("CUDA_LAUNCH_BLOCKING");
    if (blocking && blocking[0] == '1') { return true; }
    static const char *const names[] = {"CUDA_INJECTION64_PATH", "CUDA_INJECTION32_PATH", "NV_COMPUTE_PROFILER_PERFWORKS_DIR",
                                        "NV_NSIGHT_INJECTION_PORT_BASE", "NV_NSIGHT_INJECTION_TRANSPORT_TYPE", "COMPUTE_SANITIZER_INJECTION",
                                        "NV_SANITIZER_INJECTION_PORT_BASE", "NSYS_PROFILING_SESSION_ID"};
    for (size_t i = 0; i < sizeof(names) / sizeof(names[0]); ++i)
    {
        const char *v = std::getenv(names[i]);
        if (v && v[0]) { return true; }
    }
    return false;
}

static bool residentGridAllowed()
{
    int known = g_serialisedLaunches.load();
    if (known < 0)
    {
        known = serialisingToolAnnounced() ? 1 : 0;
        g_serialisedLaunches.store(known);
    }
    return known == 0;
}

static int ensureDevice()
{
    if (g_device < 0) { g_device = envInt("COGAPS_DEVICE", 0); }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
    {
        return fail(CGB_ENODEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count is 0"));
    }
    if (g_device >= n) { return fail(CGB_EINVAL, "COGAPS_DEVICE / cgb_set_device out of range"); }
    CGB_CUDA(cudaSetDevice(g_device));
    cudaDeviceProp prop;
    CGB_CUDA(cudaGetDeviceProperties(&prop, g_device));
    if (prop.major < 10)
    {
        return fail(CGB_ENODEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor)
            + "; this library is built for sm_100a only");
    }
    return CGB_OK;
}

extern "C" const char *cgb_last_error(void) { return g_lastError.c_str(); }

extern "C" const char *cgb_build_report(void)
{
    static std::string s;
    s = "cogaps_b200: sm_100a; eval kernel: " + std::to_string(kThreads) + " threads/CTA, "
        + std::to_string(kVec) + "-float vectors, clusters <= " + std::to_string(kMaxCluster)
        + ", <= " + std::to_string(kMaxBatch) + " proposals/launch; TMA bulk staging; sparse model: "
        + std::to_string(kSparseThreads) + " threads/CTA; no CPU fallback";
    return s.c_str();
}

static int cgb_set_device_body(int device)
{
    g_device = device;
    return ensureDevice();
}

extern "C" int cgb_set_device(int device)
{
    return guarded("cgb_set_device", [&]() { return cgb_set_device_body(device); });
}

extern "C" uint64_t cgb_kernel_launch_count(void) { return g_kernelLaunches.load(); }

static int cgb_set_resident_share_body(int32_t parts)
{
    CGB_CHECK(parts >= 1 && parts <= 64, "cgb_set_resident_share: parts must be in 1..64");
    g_residentShare.store(parts);
    return CGB_OK;
}

extern "C" int cgb_set_resident_share(int32_t parts)
{
    return guarded("cgb_set_resident_share", [&]() { return cgb_set_resident_share_body(parts); });
}

extern "C" void cgb_params_default(cgb_params *p)
{
    std::memset(p, 0, sizeof(*p));
    p->struct_size = sizeof(cgb_params);
    p->nPatterns = 3;
    p->nIterations = 1000;
    p->maxThreads = 1;
    p->outputFrequency = 500;
    p->snapshotPhase = CGB_PHASE_ALL;
    p->alphaA = 0.01f;
    p->alphaP = 0.01f;
    p->maxGibbsMassA = 100.f;
    p->maxGibbsMassP = 100.f;
    p->asynchronousUpdates = 1;
    p->whichMatrixFixed = 'N';
    p->workerID = 1;
}

// ------------------------------------------------------------------------------------------------
// GapsRandomState / GapsRng
// ------------------------------------------------------------------------------------------------
static int cgb_randstate_create_body(uint32_t seed, cgb_randstate **out)
{
    CGB_CHECK(out != nullptr, "cgb_randstate_create: out is NULL");
    *out = new (std::nothrow) cgb_randstate(seed);
    return *out ? CGB_OK : fail(CGB_ENOMEM, "cgb_randstate_create: out of memory");
}

extern "C" int cgb_randstate_create(uint32_t seed, cgb_randstate **out)
{
    return guarded("cgb_randstate_create", [&]() { return cgb_randstate_create_body(seed, out); });
}

static int cgb_randstate_set_tables_body(cgb_randstate *rs, const float *erf, const float *erfinv, const float *qgamma)
{
    CGB_CHECK(rs && erf && erfinv && qgamma, "cgb_randstate_set_tables: NULL argument");
    CGB_CHECK(rs->dErf == nullptr, "cgb_randstate_set_tables: tables already uploaded; set them before creating samplers");
    std::memcpy(rs->tables.erf, erf, sizeof(rs->tables.erf));
    std::memcpy(rs->tables.erfinv, erfinv, sizeof(rs->tables.erfinv));
    std::memcpy(rs->tables.qgamma, qgamma, sizeof(rs->tables.qgamma));
    return CGB_OK;
}

extern "C" int cgb_randstate_set_tables(cgb_randstate *rs, const float *erf, const float *erfinv, const float *qgamma)
{
    return guarded("cgb_randstate_set_tables", [&]() { return cgb_randstate_set_tables_body(rs, erf, erfinv, qgamma); });
}

static int cgb_randstate_get_tables_body(const cgb_randstate *rs, float *erf, float *erfinv, float *qgamma)
{
    CGB_CHECK(rs && erf && erfinv && qgamma, "cgb_randstate_get_tables: NULL argument");
    std::memcpy(erf, rs->tables.erf, sizeof(rs->tables.erf));
    std::memcpy(erfinv, rs->tables.erfinv, sizeof(rs->tables.erfinv));
    std::memcpy(qgamma, rs->tables.qgamma, sizeof(rs->tables.qgamma));
    return CGB_OK;
}

extern "C" int cgb_randstate_get_tables(const cgb_randstate *rs, float *erf, float *erfinv, float *qgamma)
{
    return guarded("cgb_randstate_get_tables", [&]() { return cgb_randstate_get_tables_body(rs, erf, erfinv, qgamma); });
}

static int cgb_randstate_next_seed_body(cgb_randstate *rs, uint64_t *out)
{
    CGB_CHECK(rs && out, "cgb_randstate_next_seed: NULL argument");
    *out = rs->seeder.next();
    return CGB_OK;
}

extern "C" int cgb_randstate_next_seed(cgb_randstate *rs, uint64_t *out)
{
    return guarded("cgb_randstate_next_seed", [&]() { return cgb_randstate_next_seed_body(rs, out); });
}

extern "C" void cgb_randstate_destroy(cgb_randstate *rs)
{
    if (!rs) { return; }
    if (rs->dErf) { cudaFree(rs->dErf); }
    if (rs->dErfinv) { cudaFree(rs->dErfinv); }
    if (rs->dQgamma) { cudaFree(rs->dQgamma); }
    delete rs;
}

static int uploadTables(cgb_randstate *rs)
{
    if (rs->dErf) { return CGB_OK; }
    CGB_CUDA(cudaMalloc(&rs->dErf, sizeof(rs->tables.erf)));
    CGB_CUDA(cudaMalloc(&rs->dErfinv, sizeof(rs->tables.erfinv)));
    CGB_CUDA(cudaMemcpy(rs->dErf, rs->tables.erf, sizeof(rs->tables.erf), cudaMemcpyHostToDevice));
    CGB_CUDA(cudaMemcpy(rs->dErfinv, rs->tables.erfinv, sizeof(rs->tables.erfinv), cudaMemcpyHostToDevice));
    CGB_CUDA(cudaMalloc(&rs->dQgamma, sizeof(rs->tables.qgamma)));
    CGB_CUDA(cudaMemcpy(rs->dQgamma, rs->tables.qgamma, sizeof(rs->tables.qgamma), cudaMemcpyHostToDevice));
    CGB_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    return CGB_OK;
}

static int cgb_rng_create_body(cgb_randstate *rs, cgb_rng **out)
{
    CGB_CHECK(rs && out, "cgb_rng_create: NULL argument");
    cgb_rng *r = new (std::nothrow) cgb_rng();
    if (!r) { return fail(CGB_ENOMEM, "cgb_rng_create: out of memory"); }
    r->rs = rs;
    r->rng = HostRng(rs->seeder);
    *out = r;
    return CGB_OK;
}

extern "C" int cgb_rng_create(cgb_randstate *rs, cgb_rng **out)
{
    return guarded("cgb_rng_create", [&]() { return cgb_rng_create_body(rs, out); });
}
extern "C" int cgb_rng_uniform32(cgb_rng *r, uint32_t *out) { *out = r->rng.next(); return CGB_OK; }
extern "C" int cgb_rng_uniform32_range(cgb_rng *r, uint32_t a, uint32_t b, uint32_t *out) { *out = r->rng.uniform32(a, b); return CGB_OK; }
extern "C" int cgb_rng_uniform64_range(cgb_rng *r, uint64_t a, uint64_t b, uint64_t *out) { *out = r->rng.uniform64(a, b); return CGB_OK; }
extern "C" int cgb_rng_uniform(cgb_rng *r, float *out) { *out = r->rng.uniform(); return CGB_OK; }
extern "C" int cgb_rng_poisson(cgb_rng *r, double lambda, int32_t *out) { *out = r->rng.poisson(lambda); return CGB_OK; }
extern "C" int cgb_rng_exponential(cgb_rng *r, float lambda, float *out) { *out = r->rng.exponential(lambda); return CGB_OK; }
static int cgb_rng_trunc_normal_body(cgb_rng *r, float a, float b, float mean, float sd, float *out, int32_t *has)
{
    *has = trunc_normal(r->rng, r->rs->tables.erf, r->rs->tables.erfinv, a, b, mean, sd, out) ? 1 : 0;
    return CGB_OK;
}

extern "C" int cgb_rng_trunc_normal(cgb_rng *r, float a, float b, float mean, float sd, float *out, int32_t *has)
{
    return guarded("cgb_rng_trunc_normal", [&]() { return cgb_rng_trunc_normal_body(r, a, b, mean, sd, out, has); });
}
static int cgb_rng_trunc_gamma_upper_body(cgb_rng *r, float b, float scale, float *out)
{
    *out = r->rng.truncGammaUpper(r->rs->tables.qgamma, b, scale);
    return CGB_OK;
}

extern "C" int cgb_rng_trunc_gamma_upper(cgb_rng *r, float b, float scale, float *out)
{
    return guarded("cgb_rng_trunc_gamma_upper", [&]() { return cgb_rng_trunc_gamma_upper_body(r, b, scale, out); });
}
extern "C" void cgb_rng_destroy(cgb_rng *r) { delete r; }

// ------------------------------------------------------------------------------------------------
// sampler construction
// ------------------------------------------------------------------------------------------------
static uint32_t roundUp(uint32_t x, uint32_t m) { return (x + m - 1) / m * m; }

// Matrix(const Matrix&, genesInCols, subsetGenes, indices) (data_structures/Matrix.cpp:30-69), written
// straight into the padded one-sampler-row-per-line layout: out[j * ld + i] = result(i, j)
static void orientData(const float *data, uint32_t nrow, uint32_t ncol, bool colmajor, bool genesInCols,
                       bool subsetGenes, const uint32_t *indices, uint32_t nIdx, uint32_t &nRows,
                       uint32_t &L, uint32_t &ld, std::vector<float> &out, float padValue)
{
    const bool subsetData = nIdx > 0;
    const uint32_t nGenes = (subsetData && subsetGenes) ? nIdx : (genesInCols ? ncol : nrow);
    const uint32_t nSamples = (subsetData && !subsetGenes) ? nIdx : (genesInCols ? nrow : ncol);
    L = nGenes;
    nRows = nSamples;
    ld = roundUp(L, 32);
    out.assign(static_cast<size_t>(nRows) * ld, padValue);
    if (!subsetData)
    {
        // whole matrix: result(i, j) is data(j, i) when genes are in columns, data(i, j) otherwise; in memory
        // that is either row-by-row copies or a cache-blocked transpose
        const bool straight = (genesInCols != colmajor); // out row j is contiguous in the input
        if (straight)
        {
            for (uint32_t j = 0; j < nSamples; ++j)
            {
                std::memcpy(&out[static_cast<size_t>(j) * ld], data + static_cast<size_t>(j) * nGenes, sizeof(float) * nGenes);
            }
        }
        else
        {
            const uint32_t B = 32;
            for (uint32_t j0 = 0; j0 < nSamples; j0 += B)
            {
                const uint32_t j1 = std::min(nSamples, j0 + B);
                for (uint32_t i0 = 0; i0 < nGenes; i0 += B)
                {
                    const uint32_t i1 = std::min(nGenes, i0 + B);
                    for (uint32_t i = i0; i < i1; ++i)
                    {
                        const float *src = data + static_cast<size_t>(i) * nSamples;
                        for (uint32_t j = j0; j < j1; ++j) { out[static_cast<size_t>(j) * ld + i] = src[j]; }
                    }
                }
            }
        }
        return;
    }
    for (uint32_t j = 0; j < nSamples; ++j)
    {
        for (uint32_t i = 0; i < nGenes; ++i)
        {
            const uint32_t dataRow = (subsetData && (subsetGenes != genesInCols))
                ? indices[genesInCols ? j : i] - 1 : (genesInCols ? j : i);
            const uint32_t dataCol = (subsetData && (subsetGenes == genesInCols))
                ? indices[genesInCols ? i : j] - 1 : (genesInCols ? i : j);
            const size_t src = colmajor ? static_cast<size_t>(dataCol) * nrow + dataRow
                                        : static_cast<size_t>(dataRow) * ncol + dataCol;
            out[static_cast<size_t>(j) * ld + i] = data[src];
        }
    }
}

// The reference indexes the data with whatever it is given (data_structures/Matrix.cpp:30-69); an index past the
// end is undefined behaviour there and an error here.  Indices address rows of the caller's matrix when
// subsetGenes != genesInCols, columns otherwise (orientData below).
static int checkSubset(const cgb_params *p, uint32_t nrow, uint32_t ncol, bool genesInCols, bool subsetGenes)
{
    const uint32_t bound = (subsetGenes != genesInCols) ? nrow : ncol;
    for (uint32_t i = 0; i < p->nSubsetIndices; ++i)
    {
        CGB_CHECK(p->subsetIndices && p->subsetIndices[i] >= 1, "subset indices are 1-based (R convention)");
        CGB_CHECK(p->subsetIndices[i] <= bound, "subset index past the end of the data");
    }
    return CGB_OK;
}

// one cluster per row scan: nSeg CTAs of kThreads threads, each staging `seg` floats per stream.  The fewest
// segments whose staging (4 streams) still lets two CTAs share an SM: 6700 floats -> 107 KB per CTA.  Registers cap
// an SM at two 512-thread CTAs anyway, so fewer CTAs per row means more rows in flight at once (L = 20000: three
// segments, 98 resident clusters; four would give 74 and a P-side batch of ~85 tasks would need a second round).
static const uint32_t kTableSegFloats = 5120; // up to here the epilogue's lookup tables fit in shared memory too

static void segmentsForLength(uint32_t L, uint32_t &nSeg, uint32_t &seg)
{
    const uint32_t target = static_cast<uint32_t>(envInt("COGAPS_SEG_FLOATS", 6700));
    const uint32_t maxCluster = static_cast<uint32_t>(envInt("COGAPS_MAX_CLUSTER", 8));
    uint32_t n = 1;
    while (n < maxCluster && n < static_cast<uint32_t>(kMaxCluster) && (L + n - 1) / n > target) { ++n; }
    nSeg = n;
    seg = roundUp((L + n - 1) / n, 4);
}

static void chooseSegments(cgb_sampler *s)
{
    segmentsForLength(s->L, s->nSeg, s->seg);
    s->segPad = roundUp(s->seg, 32);
    s->smemBytes = 256 + static_cast<size_t>(5) * s->segPad * sizeof(float);
    s->tablesInSmem = s->sparse || s->seg <= kTableSegFloats;
}

static int cgb_reduction_order_for_length_body(uint32_t rowLength, cgb_reduction_order *out)
{
    CGB_CHECK(out && rowLength, "cgb_reduction_order_for_length: bad argument");
    out->threadsPerSegment = kThreads;
    out->vectorWidth = kVec;
    segmentsForLength(rowLength, out->nSegments, out->segmentLength);
    return CGB_OK;
}

extern "C" int cgb_reduction_order_for_length(uint32_t rowLength, cgb_reduction_order *out)
{
    return guarded("cgb_reduction_order_for_length", [&]() { return cgb_reduction_order_for_length_body(rowLength, out); });
}

static int stopPersistent(cgb_sampler *s);

// Matrix-sized device buffers (D, AP, S: 400 MB each at BASELINE configs[2]) are parked here when a sampler is destroyed and
// handed to the next sampler that asks for the same size on the same device.  cudaFree of such a buffer hands the memory
// back to the system and the next cudaMalloc maps it afresh: measured on the B200 box, a cgb_run's teardown took anything
// from 5 ms to 0.9 s and its allocations from 3 to 95 ms, call to call — and callers like distributed CoGAPS run one
// factorisation after the other.  At most COGAPS_DEVICE_CACHE_MB (default 16384, 0 disables) stay parked;
// cgb_release_device_cache() frees them.
struct ParkedBuffer { void *ptr; size_t bytes; int device; };
static std::mutex g_parkLock;
static std::vector<ParkedBuffer> g_parked;
static size_t g_parkedBytes = 0;

static cudaError_t matrixAlloc(void **out, size_t bytes, int device)
{
    {
        std::lock_guard<std::mutex> hold(g_parkLock);
        for (size_t i = 0; i < g_parked.size(); ++i)
        {
            if (g_parked[i].bytes == bytes && g_parked[i].device == device)
            {
                *out = g_parked[i].ptr;
                g_parkedBytes -= bytes;
                g_parked.erase(g_parked.begin() + static_cast<long>(i));
                return cudaSuccess;
            }
        }
    }
    return cudaMalloc(out, bytes);
}

// the caller has synchronised every stream that touched the buffer
static void matrixFree(void *ptr, size_t bytes, int device)
{
    if (ptr == nullptr) { return; }
    const size_t limit = static_cast<size_t>(std::max(envInt("COGAPS_DEVICE_CACHE_MB", 16384), 0)) << 20;
    std::vector<ParkedBuffer> evicted; // the oldest parked buffers make room (freed outside the lock)
    bool parked = false;
    {
        std::lock_guard<std::mutex> hold(g_parkLock);
        if (bytes >= (1u << 20) && bytes <= limit)
        {
            try
            {
                while (g_parkedBytes + bytes > limit && !g_parked.empty())
                {
                    evicted.push_back(g_parked.front());
                    g_parkedBytes -= g_parked.front().bytes;
                    g_parked.erase(g_parked.begin());
                }
                g_parked.push_back(ParkedBuffer{ptr, bytes, device});
                g_parkedBytes += bytes;
                parked = true;
            }
            catch (...) { }
        }
    }
    for (size_t i = 0; i < evicted.size(); ++i)
    {
        cudaSetDevice(evicted[i].device);
        cudaFree(evicted[i].ptr);
    }
    if (!evicted.empty()) { cudaSetDevice(device); }
    if (!parked) { cudaFree(ptr); }
}

extern "C" int cgb_release_device_cache(void)
{
    std::vector<ParkedBuffer> drop;
    {
        std::lock_guard<std::mutex> hold(g_parkLock);
        drop.swap(g_parked);
        g_parkedBytes = 0;
    }
    int current = 0;
    cudaGetDevice(&current);
    for (size_t i = 0; i < drop.size(); ++i)
    {
        cudaSetDevice(drop[i].device);
        cudaFree(drop[i].ptr);
    }
    cudaSetDevice(current);
    return CGB_OK;
}

extern "C" void cgb_sampler_destroy(cgb_sampler *s)
{
    if (!s) { return; }
    cudaSetDevice(s->device);
    const bool lapOn = envInt("COGAPS_HOST_PROFILE", 0) != 0;
    double lapT = nowSeconds();
    auto lap = [&](const char *what)
    {
        if (lapOn)
        {
            const double now = nowSeconds();
            if (now - lapT > 2.0e-3) { std::printf("[sampler destroy] %s %.1f ms\n", what, (now - lapT) * 1e3); }
            lapT = now;
        }
    };
    if (s->persistentRunning) { stopPersistent(s); } // never free what a resident grid may still touch
    if (s->stream) { cudaStreamSynchronize(s->stream); }
    cudaStreamSynchronize(cudaStreamLegacy);
    lap("waiting for the streams");
    {
        const size_t matBytes = static_cast<size_t>(s->nRows) * s->ld * sizeof(float);
        matrixFree(s->dD, matBytes, s->device);
        matrixFree(s->dS, matBytes, s->device);
        matrixFree(s->dAP, matBytes, s->device);
    }
    lap("matrix buffers");
    cudaFree(s->dM); cudaFree(s->dColNonzero);
    cudaFree(s->dPartials); cudaFree(s->dTickets); cudaFree(s->dReducePartials); cudaFree(s->dPhaseClocks);
    cudaFree(s->dRowVersion); cudaFree(s->dStreamStats);
    cudaFree(s->dSwPos); cudaFree(s->dSwMass); cudaFree(s->dSwCount); cudaFree(s->dSwCounters); cudaFree(s->dSwOrder);
    lap("small device buffers, sweep store");
    if (s->hSwCounters) { cudaFreeHost(s->hSwCounters); }
    cudaFree(s->dSpRowPtr); cudaFree(s->dSpIdx); cudaFree(s->dSpVal); cudaFree(s->dMrows); cudaFree(s->dZ1); cudaFree(s->dZ2);
    if (s->hCommitsMirror) { cudaFreeHost(const_cast<unsigned long long*>(s->hCommitsMirror)); }
    if (s->hAlive) { cudaFreeHost(const_cast<uint32_t*>(s->hAlive)); }
    if (s->hSlots) { cudaFreeHost(s->hSlots); }
    if (s->hStreamOutcomes) { cudaFreeHost(s->hStreamOutcomes); }
    if (s->hOutcomes) { cudaFreeHost(s->hOutcomes); }
    if (s->hReducePartials) { cudaFreeHost(s->hReducePartials); }
    lap("pinned host buffers");
    if (s->evStart) { cudaEventDestroy(s->evStart); }
    if (s->evStop) { cudaEventDestroy(s->evStop); }
    if (s->stream) { cudaStreamDestroy(s->stream); }
    lap("events and stream");
    delete s;
    lap("host object");
}

static const int kReduceBlocks = 592; // 148 SMs x 4
// erf + erfinv tables the resident kernel keeps in shared memory behind the staging buffers
static const size_t kStreamTableBytes = (((CGB_ERF_TABLE_SIZE + 3) & ~3) + ((CGB_ERFINV_TABLE_SIZE + 3) & ~3)) * sizeof(float);

// The sampler's generator (domain + queue / sequential state).  Separate from the data preparation because
// this is where seeds are taken from the shared random state, in construction order (A first, then P).
static void initGenerator(cgb_sampler *s, const cgb_params *params, cgb_randstate *rs)
{
    // AsynchronousGibbsSampler ctor, AsynchronousGibbsSampler.h:63-76: domain over nRows*k bins, queue
    // rng seeded from the shared state
    const uint64_t nElements = static_cast<uint64_t>(s->nRows) * s->k;
    s->domain.init(nElements);
    s->sequential = params->asynchronousUpdates == 0;
    if (s->sequential) { s->seq.init(nElements, s->k, rs, s->alpha); } // SingleThreadedGibbsSampler.h:66-81
    else { s->queue.init(nElements, s->k, rs, s->alpha, s->lambda); }
}

// Hand-over between the two samplers of one run: the one whose orientation is contiguous in the caller's matrix
// uploads it straight from there; the other takes its copy of D as a device transpose of that upload instead of a
// second pass over the matrix on the host.
struct TwinLink
{
    std::atomic<int> state;     // 0 pending, 1 uploaded, -1 failed
    const cgb_sampler *sampler;
    TwinLink() : state(0), sampler(nullptr) {}
};

// The data of ONE sampler already in compressed rows (sampler rows x L, ascending index, positive entries only):
// what the Matrix-Market loader hands over instead of a dense matrix (sparse model, whole matrix only)
struct CsrView
{
    uint32_t nRows, L;
    const std::vector<uint32_t> *ptr, *idx;
    const std::vector<float> *val;
};

static int samplerCreateImpl(const float *data, uint32_t nrow, uint32_t ncol, int32_t colmajor,
                             int32_t transpose, int32_t subsetRows, float alpha, float maxGibbsMass,
                             const cgb_params *params, cgb_randstate *rs, bool withGenerator, cgb_sampler **out,
                             TwinLink *publish = nullptr, TwinLink *twin = nullptr, const CsrView *csr = nullptr)
{
    CGB_CHECK((data || csr) && params && rs && out, "cgb_sampler_create: NULL argument");
    CGB_CHECK(csr == nullptr || (params->useSparseOptimization != 0 && params->nSubsetIndices == 0 && twin == nullptr && publish == nullptr),
              "cgb_sampler_create: compressed-row input is for the sparse model on the whole matrix");
    CGB_CHECK(params->struct_size == sizeof(cgb_params), "cgb_sampler_create: cgb_params ABI mismatch");
    CGB_CHECK(params->nPatterns >= 1, "cgb_sampler_create: nPatterns must be >= 1");
    CGB_TRY(checkSubset(params, nrow, ncol, transpose != 0, subsetRows != 0));
    CGB_TRY(ensureDevice());

    cgb_sampler *s = new (std::nothrow) cgb_sampler();
    if (!s) { return fail(CGB_ENOMEM, "cgb_sampler_create: out of memory"); }
    std::memset(static_cast<void*>(&s->counters), 0, sizeof(s->counters));
    s->dD = s->dS = s->dAP = s->dM = nullptr;
    s->sparse = params->useSparseOptimization != 0;
    s->dSpRowPtr = s->dSpIdx = nullptr; s->dSpVal = s->dMrows = s->dZ1 = s->dZ2 = nullptr; s->ldR = 0;
    s->dColNonzero = nullptr; s->dPartials = nullptr; s->dTickets = nullptr; s->dReducePartials = nullptr;
    s->usePersistent = envInt("COGAPS_PERSISTENT", residentGridAllowed() ? 1 : 0) != 0; s->persistentRunning = false;
    // test knobs for the rarely taken paths: every row read behind a rowVersion check; small chunks
    s->forceRowWait = envInt("COGAPS_FORCE_ROW_WAIT", 0) != 0;
    {
        const int cap = envInt("COGAPS_CHUNK_PROPOSALS", kMaxPersistentBatch);
        s->chunkCap = static_cast<size_t>(cap >= 1 && cap <= kMaxPersistentBatch ? cap : kMaxPersistentBatch);
    }
    s->hSlots = nullptr; s->nSlotRecords = 0; s->hStreamOutcomes = nullptr; s->dStreamStats = nullptr; s->dRowVersion = nullptr;
    s->mailSeq = 0; s->persistentGrid = 0; s->nClusters = 0; s->lastPostTime = 0.0;
    s->hAlive = nullptr; s->launchEpoch = 0; s->ticketBase = 0; s->aliveNext = 0; s->aliveLooks = 0;
    s->chunkTag = 0; s->chunkPosted = 0; s->chunkBase = 0;
    s->hCommitsMirror = nullptr;
    s->commitsExpected[0] = s->commitsExpected[1] = s->provenThrough[0] = s->provenThrough[1] = 0;
    s->dPhaseClocks = nullptr; s->phaseTasks = 0;
    for (int i = 0; i < kPhaseSlots; ++i) { s->phaseSum[i] = 0.0; }
    s->hOutcomes = nullptr; s->hReducePartials = nullptr; s->stream = nullptr; s->evStart = s->evStop = nullptr;
    s->other = nullptr;
    s->rs = rs;
    s->device = g_device;
    s->hasS = false;
    s->timeKernels = false;
    s->avgQueueLength = s->numQueueSamples = 0.f;
    s->k = params->nPatterns;
    s->alpha = alpha;
    s->annealingTemp = 1.f;

    // DenseNormalModel ctor, DenseNormalModel.h:66-88
    std::vector<float> host;
    const bool wholeMatrix = params->nSubsetIndices == 0;
    const bool straight = wholeMatrix && ((transpose != 0) != (colmajor != 0)); // a sampler row is contiguous in `data`
    // direct: no oriented copy on the host at all (dense model; the sparse model builds its CSR from one)
    const bool direct = wholeMatrix && !s->sparse && (straight || twin != nullptr);
    const float *meanBase;
    size_t meanStrideR, meanStrideL;
    if (csr)
    {
        s->L = csr->L;
        s->nRows = csr->nRows;
        s->ld = roundUp(s->L, 32);
        meanBase = nullptr;
        meanStrideR = meanStrideL = 0;
    }
    else if (direct)
    {
        const uint32_t nGenes = transpose ? ncol : nrow, nSamples = transpose ? nrow : ncol;
        s->L = nGenes;
        s->nRows = nSamples;
        s->ld = roundUp(s->L, 32);
        meanBase = data;
        meanStrideR = straight ? nGenes : 1;
        meanStrideL = straight ? 1 : nSamples;
    }
    else
    {
        orientData(data, nrow, ncol, colmajor != 0, transpose != 0, subsetRows != 0, params->subsetIndices,
            params->nSubsetIndices, s->nRows, s->L, s->ld, host, 0.f);
        meanBase = host.data();
        meanStrideR = s->ld;
        meanStrideL = 1;
    }
    if (!(s->nRows >= 1 && s->L >= 1))
    {
        if (publish) { publish->state.store(-1); }
        delete s;
        return fail(CGB_EINVAL, "cgb_sampler_create: empty data");
    }
    s->ldM = roundUp(s->nRows, 32);
    // gaps::nonZeroMean (MatrixMath.cpp:39-55): ONE fp32 running sum over everything / count of positives.  The
    // order of the additions is the result, so this is a serial pass over the whole matrix; it runs on its own
    // thread beside the allocations and the upload below and is joined before anything needs lambda.
    struct MeanJob { float sum; unsigned nnz; } meanJob = {0.f, 0u};
    std::thread meanThread([meanBase, meanStrideR, meanStrideL, &meanJob, s, csr]()
    {
        float sum = 0.f;
        unsigned nnz = 0;
        if (csr)
        {
            // the same running sum: the elements that are not stored are zeros and add nothing
            const std::vector<float> &v = *csr->val;
            for (size_t i = 0; i < v.size(); ++i) { sum += v[i]; }
            meanJob.sum = sum;
            meanJob.nnz = static_cast<unsigned>(v.size());
            return;
        }
        const double t0 = nowSeconds();
        runningSum(meanBase, s->nRows, s->L, meanStrideR, meanStrideL, sum, nnz);
        if (envInt("COGAPS_HOST_PROFILE", 0)) { std::printf("[sampler create] lambda's running sum over %u x %u (%s): %.3f s\n", s->nRows, s->L, meanStrideL == 1 ? "along rows" : "down columns", nowSeconds() - t0); }
        meanJob.sum = sum;
        meanJob.nnz = nnz;
    });
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) { t.join(); } } } meanJoiner = {meanThread};
    chooseSegments(s);
    std::vector<uint32_t> spPtrOwn, spIdxOwn;
    std::vector<float> spValOwn;
    const std::vector<uint32_t> &spPtr = csr ? *csr->ptr : spPtrOwn, &spIdx = csr ? *csr->idx : spIdxOwn;
    const std::vector<float> &spVal = csr ? *csr->val : spValOwn;
    if (csr)
    {
        s->nSeg = 1;
        s->ldR = roundUp(s->k, 4);
    }
    else if (s->sparse)
    {
        std::vector<uint32_t> &spPtr = spPtrOwn, &spIdx = spIdxOwn;
        std::vector<float> &spVal = spValOwn;
        // SparseMatrix (data_structures/SparseMatrix.cpp, SparseVector.cpp:20-35): the positive entries of every
        // sampler row, ascending index
        s->nSeg = 1;
        s->ldR = roundUp(s->k, 4);
        spPtr.assign(s->nRows + 1, 0u);
        for (uint32_t r = 0; r < s->nRows; ++r)
        {
            const float *row = host.data() + static_cast<size_t>(r) * s->ld;
            for (uint32_t l = 0; l < s->L; ++l)
            {
                if (row[l] > 0.f)
                {
                    spIdx.push_back(l);
                    spVal.push_back(row[l]);
                }
            }
            if (spIdx.size() > 0xFFFFFFF0ull) { meanThread.join(); delete s; return fail(CGB_EUNSUPPORTED, "cgb_sampler_create: more than 2^32 non-zeros"); }
            spPtr[r + 1] = static_cast<uint32_t>(spIdx.size());
        }
    }

    int rc = CGB_OK;
    do
    {
        const size_t matBytes = static_cast<size_t>(s->nRows) * s->ld * sizeof(float);
        const size_t facBytes = static_cast<size_t>(s->k) * s->ldM * sizeof(float);
#define CGB_CUDA_BREAK(call) { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = fail(e__ == cudaErrorMemoryAllocation ? CGB_ENOMEM : CGB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); break; } }
        const bool lapOn = envInt("COGAPS_HOST_PROFILE", 0) != 0;
        double lapT = nowSeconds();
        auto lap = [&](const char *what)
        {
            if (lapOn)
            {
                const double now = nowSeconds();
                std::printf("[sampler create %ux%u] %s %.1f ms\n", s->nRows, s->L, what, (now - lapT) * 1e3);
                lapT = now;
            }
        };
        CGB_CUDA_BREAK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        CGB_CUDA_BREAK(matrixAlloc(reinterpret_cast<void**>(&s->dD), matBytes, s->device));
        if (!s->sparse) { CGB_CUDA_BREAK(matrixAlloc(reinterpret_cast<void**>(&s->dAP), matBytes, s->device)); }
        if (s->sparse)
        {
            const size_t nnz = spIdx.size();
            CGB_CUDA_BREAK(cudaMalloc(&s->dSpRowPtr, sizeof(uint32_t) * (s->nRows + 1)));
            CGB_CUDA_BREAK(cudaMalloc(&s->dSpIdx, sizeof(uint32_t) * (nnz ? nnz : 1)));
            CGB_CUDA_BREAK(cudaMalloc(&s->dSpVal, sizeof(float) * (nnz ? nnz : 1)));
            CGB_CUDA_BREAK(cudaMalloc(&s->dMrows, sizeof(float) * s->nRows * s->ldR));
            CGB_CUDA_BREAK(cudaMalloc(&s->dZ1, sizeof(float) * s->k));
            CGB_CUDA_BREAK(cudaMalloc(&s->dZ2, sizeof(float) * s->k * s->k));
            CGB_CUDA_BREAK(cudaMemcpy(s->dSpRowPtr, spPtr.data(), sizeof(uint32_t) * (s->nRows + 1), cudaMemcpyHostToDevice));
            if (nnz)
            {
                CGB_CUDA_BREAK(cudaMemcpy(s->dSpIdx, spIdx.data(), sizeof(uint32_t) * nnz, cudaMemcpyHostToDevice));
                CGB_CUDA_BREAK(cudaMemcpy(s->dSpVal, spVal.data(), sizeof(float) * nnz, cudaMemcpyHostToDevice));
            }
            CGB_CUDA_BREAK(cudaMemset(s->dMrows, 0, sizeof(float) * s->nRows * s->ldR));
            CGB_CUDA_BREAK(cudaMemset(s->dZ1, 0, sizeof(float) * s->k));
            CGB_CUDA_BREAK(cudaMemset(s->dZ2, 0, sizeof(float) * s->k * s->k));
        }
        CGB_CUDA_BREAK(cudaMalloc(&s->dM, facBytes));
        CGB_CUDA_BREAK(cudaMalloc(&s->dColNonzero, sizeof(int) * s->k));
        CGB_CUDA_BREAK(cudaMalloc(&s->dPartials, sizeof(AlphaPair) * 2 * kMaxPersistentBatch));
        CGB_CUDA_BREAK(cudaMalloc(&s->dTickets, sizeof(uint32_t) * kMaxPersistentBatch));
        CGB_CUDA_BREAK(cudaMalloc(&s->dReducePartials, sizeof(double) * kReduceBlocks));
        CGB_CUDA_BREAK(cudaHostAlloc(&s->hOutcomes, sizeof(DevOutcome) * kMaxBatch, cudaHostAllocMapped));
        CGB_CUDA_BREAK(cudaHostAlloc(&s->hReducePartials, sizeof(double) * kReduceBlocks, cudaHostAllocDefault));
        CGB_CUDA_BREAK(cudaHostAlloc(&s->hStreamOutcomes, sizeof(HostOutcome) * kMaxPersistentBatch, cudaHostAllocMapped));
        std::memset(s->hStreamOutcomes, 0, sizeof(HostOutcome) * kMaxPersistentBatch);
        CGB_CUDA_BREAK(cudaMalloc(&s->dStreamStats, sizeof(StreamStats)));
        CGB_CUDA_BREAK(cudaMalloc(&s->dRowVersion, sizeof(uint32_t) * s->nRows));
        CGB_CUDA_BREAK(cudaMemset(s->dRowVersion, 0, sizeof(uint32_t) * s->nRows));
        s->rowVersion.assign(s->nRows, 0u);
        s->rowPending.assign(s->nRows, 0ull);
        {
            void *mirror = nullptr;
            CGB_CUDA_BREAK(cudaHostAlloc(&mirror, 64, cudaHostAllocMapped));
            std::memset(mirror, 0, 64);
            s->hCommitsMirror = static_cast<volatile unsigned long long*>(mirror);
        }
        CGB_CUDA_BREAK(cudaEventCreate(&s->evStart));
        CGB_CUDA_BREAK(cudaEventCreate(&s->evStop));
        lap("allocations (device, pinned host)");
        if (csr)
        {
            // the dense copy the chi-square kernels walk is rebuilt on the device from the rows just uploaded (pageable
            // copies through the legacy stream: staged, not necessarily landed — wait before our own stream reads them)
            CGB_CUDA_BREAK(cudaStreamSynchronize(cudaStreamLegacy));
            CGB_CUDA_BREAK(cudaMemsetAsync(s->dD, 0, matBytes, s->stream));
            csr_scatter_kernel<<<s->nRows, 256, 0, s->stream>>>(s->dSpRowPtr, s->dSpIdx, s->dSpVal, s->ld, s->dD);
            ++g_kernelLaunches;
            CGB_CUDA_BREAK(cudaGetLastError());
            CGB_CUDA_BREAK(cudaStreamSynchronize(s->stream));
        }
        else if (!direct)
        {
            CGB_CUDA_BREAK(cudaMemcpy(s->dD, host.data(), matBytes, cudaMemcpyHostToDevice));
        }
        else if (straight)
        {
            CGB_CUDA_BREAK(cudaMemset(s->dD, 0, matBytes)); // the pad columns
            CGB_CUDA_BREAK(cudaMemcpy2D(s->dD, sizeof(float) * s->ld, data, sizeof(float) * s->L, sizeof(float) * s->L, s->nRows,
                cudaMemcpyHostToDevice));
        }
        else
        {
            // the twin holds the other orientation: ours is its transpose (pads stay zero)
            while (twin->state.load() == 0) { __builtin_ia32_pause(); }
            if (twin->state.load() < 0)
            {
                rc = fail(CGB_EINTERNAL, "cgb_sampler_create: the sampler holding the other orientation failed");
                break;
            }
            const cgb_sampler *t = twin->sampler;
            CGB_CUDA_BREAK(cudaMemsetAsync(s->dD, 0, matBytes, s->stream));
            dim3 grid((s->L + kTransposeTile - 1) / kTransposeTile, (s->nRows + kTransposeTile - 1) / kTransposeTile);
            transpose_kernel<<<grid, 256, 0, s->stream>>>(s->dD, t->dD, s->nRows, s->L, s->ld, t->ld);
            ++g_kernelLaunches;
            CGB_CUDA_BREAK(cudaGetLastError());
            CGB_CUDA_BREAK(cudaStreamSynchronize(s->stream));
        }
        // cudaMemcpy from pageable memory and cudaMemset return before the device has finished (the copy is only
        // staged, the memset only queued) and the samplers' own streams are non-blocking, i.e. NOT ordered behind
        // the legacy stream these calls use: wait here, before anybody (the twin's transpose first of all) reads D
        CGB_CUDA_BREAK(cudaStreamSynchronize(cudaStreamLegacy));
        lap("D on the device (upload or transpose of the twin's, incl. waiting for the twin)");
        if (publish)
        {
            publish->sampler = s;
            publish->state.store(1);
        }
        if (!s->sparse) { CGB_CUDA_BREAK(cudaMemset(s->dAP, 0, matBytes)); }
        CGB_CUDA_BREAK(cudaMemset(s->dM, 0, facBytes));
        CGB_CUDA_BREAK(cudaMemset(s->dColNonzero, 0, sizeof(int) * s->k));
        CGB_CUDA_BREAK(cudaMemset(s->dTickets, 0, sizeof(uint32_t) * kMaxPersistentBatch));
        CGB_CUDA_BREAK(cudaFuncSetAttribute(eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CGB_CUDA_BREAK(cudaFuncSetAttribute(eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CGB_CUDA_BREAK(cudaFuncSetAttribute(probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CGB_CUDA_BREAK(cudaFuncSetAttribute(probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CGB_CUDA_BREAK(cudaFuncSetAttribute(eval_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        CGB_CUDA_BREAK(cudaFuncSetAttribute(eval_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
        CGB_CUDA_BREAK(cudaFuncSetAttribute(eval_stream_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        if (!s->sparse && s->smemBytes + (s->tablesInSmem ? kStreamTableBytes : 0) > 226u * 1024u)
        {
            rc = fail(CGB_EUNSUPPORTED, "row length too large for one cluster of staged segments (raise COGAPS_MAX_CLUSTER)");
            break;
        }
        rc = uploadTables(rs);
        // every memset / copy above went through the legacy stream (see the note at the upload of D)
        if (rc == CGB_OK) { CGB_CUDA_BREAK(cudaStreamSynchronize(cudaStreamLegacy)); }
        lap("memsets, kernel attributes, tables");
    } while (0);
    {
        const double tj = nowSeconds();
        meanThread.join();
        if (envInt("COGAPS_HOST_PROFILE", 0)) { std::printf("[sampler create %ux%u] waited %.1f ms more for lambda's running sum\n", s->nRows, s->L, (nowSeconds() - tj) * 1e3); }
    }
    if (rc != CGB_OK)
    {
        if (publish && publish->state.load() == 0) { publish->state.store(-1); }
        cgb_sampler_destroy(s);
        return rc;
    }
    {
        const float meanD = meanJob.sum / static_cast<float>(meanJob.nnz);
        s->lambda = alpha * std::sqrt(static_cast<float>(static_cast<uint64_t>(s->k)) / meanD);
        s->maxGibbsMass = maxGibbsMass / s->lambda;
        const float size = static_cast<float>(s->nRows * s->L);                 // gaps::sparsity, MatrixMath.cpp:6-21
        s->dataSparsity = 1.f - static_cast<float>(meanJob.nnz) / size;
    }

    if (withGenerator) { initGenerator(s, params, rs); }
    *out = s;
    return CGB_OK;
}

static int cgb_sampler_create_body(const float *data, uint32_t nrow, uint32_t ncol, int32_t colmajor,
                                  int32_t transpose, int32_t subsetRows, float alpha, float maxGibbsMass,
                                  const cgb_params *params, cgb_randstate *rs, cgb_sampler **out)
{
    return samplerCreateImpl(data, nrow, ncol, colmajor, transpose, subsetRows, alpha, maxGibbsMass, params, rs, true, out);
}

extern "C" int cgb_sampler_create(const float *data, uint32_t nrow, uint32_t ncol, int32_t colmajor,
                                  int32_t transpose, int32_t subsetRows, float alpha, float maxGibbsMass,
                                  const cgb_params *params, cgb_randstate *rs, cgb_sampler **out)
{
    return guarded("cgb_sampler_create", [&]() { return cgb_sampler_create_body(data, nrow, ncol, colmajor, transpose, subsetRows, alpha, maxGibbsMass, params, rs, out); });
}

static int cgb_sampler_set_uncertainty_body(cgb_sampler *s, const float *unc, uint32_t nrow, uint32_t ncol,
                                           int32_t colmajor, int32_t transpose, int32_t subsetRows,
                                           const cgb_params *params)
{
    CGB_CHECK(s && unc && params, "cgb_sampler_set_uncertainty: NULL argument");
    if (s->sparse) { return CGB_OK; } // SparseNormalModel::setUncertainty is a nop (SparseNormalModel.h:92-98)
    CGB_CUDA(cudaSetDevice(s->device));
    CGB_TRY(checkSubset(params, nrow, ncol, transpose != 0, subsetRows != 0));
    std::vector<float> host;
    uint32_t nRows, L, ld;
    // pads are 1 like mSMatrix.pad(1.f) (DenseNormalModel.h:87,94); the kernels never read them as data
    orientData(unc, nrow, ncol, colmajor != 0, transpose != 0, subsetRows != 0, params->subsetIndices,
        params->nSubsetIndices, nRows, L, ld, host, 1.f);
    CGB_CHECK(nRows == s->nRows && L == s->L, "cgb_sampler_set_uncertainty: shape differs from the data");
    const size_t matBytes = static_cast<size_t>(s->nRows) * s->ld * sizeof(float);
    if (!s->dS) { CGB_CUDA(matrixAlloc(reinterpret_cast<void**>(&s->dS), matBytes, s->device)); }
    CGB_CUDA(cudaMemcpy(s->dS, host.data(), matBytes, cudaMemcpyHostToDevice));
    CGB_CUDA(cudaStreamSynchronize(cudaStreamLegacy)); // pageable copy: staged, not necessarily landed
    s->hasS = true;
    return CGB_OK;
}

extern "C" int cgb_sampler_set_uncertainty(cgb_sampler *s, const float *unc, uint32_t nrow, uint32_t ncol,
                                           int32_t colmajor, int32_t transpose, int32_t subsetRows,
                                           const cgb_params *params)
{
    return guarded("cgb_sampler_set_uncertainty", [&]() { return cgb_sampler_set_uncertainty_body(s, unc, nrow, ncol, colmajor, transpose, subsetRows, params); });
}

static int refreshColNonzero(cgb_sampler *s)
{
    col_nonzero_kernel<<<s->k, 256, 0, s->stream>>>(s->dM, s->nRows, s->ldM, s->dColNonzero);
    ++g_kernelLaunches;
    CGB_CUDA(cudaGetLastError());
    return CGB_OK;
}

static int cgb_sampler_set_matrix_body(cgb_sampler *s, const float *mat)
{
    CGB_CHECK(s && mat, "cgb_sampler_set_matrix: NULL argument");
    CGB_CUDA(cudaSetDevice(s->device));
    std::vector<float> host(static_cast<size_t>(s->k) * s->ldM, 0.f);
    for (uint32_t r = 0; r < s->nRows; ++r)
    {
        for (uint32_t c = 0; c < s->k; ++c) { host[static_cast<size_t>(c) * s->ldM + r] = mat[static_cast<size_t>(r) * s->k + c]; }
    }
    if (s->sparse)
    {
        // HybridMatrix::operator=(Matrix): the row copy keeps the value, the column copy drops it below epsilon
        std::vector<float> rows(static_cast<size_t>(s->nRows) * s->ldR, 0.f);
        for (uint32_t r = 0; r < s->nRows; ++r)
        {
            for (uint32_t c = 0; c < s->k; ++c)
            {
                const float v = mat[static_cast<size_t>(r) * s->k + c];
                rows[static_cast<size_t>(r) * s->ldR + c] = v;
                if (v < kEpsilon) { host[static_cast<size_t>(c) * s->ldM + r] = 0.f; }
            }
        }
        CGB_CUDA(cudaMemcpyAsync(s->dMrows, rows.data(), rows.size() * sizeof(float), cudaMemcpyHostToDevice, s->stream));
        CGB_CUDA(cudaStreamSynchronize(s->stream));
    }
    CGB_CUDA(cudaMemcpyAsync(s->dM, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    CGB_CUDA(cudaStreamSynchronize(s->stream));
    return refreshColNonzero(s);
}

extern "C" int cgb_sampler_set_matrix(cgb_sampler *s, const float *mat)
{
    return guarded("cgb_sampler_set_matrix", [&]() { return cgb_sampler_set_matrix_body(s, mat); });
}

static int cgb_sampler_set_annealing_temp_body(cgb_sampler *s, float temp)
{
    CGB_CHECK(s != nullptr, "cgb_sampler_set_annealing_temp: NULL sampler");
    s->annealingTemp = temp;
    return CGB_OK;
}

extern "C" int cgb_sampler_set_annealing_temp(cgb_sampler *s, float temp)
{
    return guarded("cgb_sampler_set_annealing_temp", [&]() { return cgb_sampler_set_annealing_temp_body(s, temp); });
}

static int cgb_sampler_sync_body(cgb_sampler *s, const cgb_sampler *other)
{
    CGB_CHECK(s && other, "cgb_sampler_sync: NULL argument");
    CGB_CHECK(other->nRows == s->L && other->L == s->nRows && other->k == s->k, "cgb_sampler_sync: shapes do not transpose");
    CGB_CUDA(cudaSetDevice(s->device));
    // the other sampler's stream may still hold its last commit
    CGB_CHECK(other->sparse == s->sparse, "cgb_sampler_sync: one sampler is sparse, the other dense");
    CGB_CUDA(cudaStreamSynchronize(other->stream));
    if (s->sparse)
    {
        // SparseNormalModel::sync (SparseNormalModel.cpp:27-31): point at the other factor, rebuild Z1 / Z2
        const uint32_t blocks = s->k + s->k * (s->k + 1) / 2;
        sparse_tables_kernel<<<blocks, kSparseThreads, 0, s->stream>>>(other->dMrows, other->ldR, other->dM, other->ldM,
            other->nRows, s->k, s->dZ1, s->dZ2);
    }
    else
    {
        dim3 grid((s->L + kTransposeTile - 1) / kTransposeTile, (s->nRows + kTransposeTile - 1) / kTransposeTile);
        transpose_kernel<<<grid, 256, 0, s->stream>>>(s->dAP, other->dAP, s->nRows, s->L, s->ld, other->ld);
    }
    ++g_kernelLaunches;
    CGB_CUDA(cudaGetLastError());
    // canUseGibbs flags of the other factor, constant for the whole of our next update()
    col_nonzero_kernel<<<other->k, 256, 0, s->stream>>>(other->dM, other->nRows, other->ldM, other->dColNonzero);
    ++g_kernelLaunches;
    CGB_CUDA(cudaGetLastError());
    s->other = other;
    return CGB_OK;
}

extern "C" int cgb_sampler_sync(cgb_sampler *s, const cgb_sampler *other)
{
    return guarded("cgb_sampler_sync", [&]() { return cgb_sampler_sync_body(s, other); });
}

static int cgb_sampler_extra_initialization_body(cgb_sampler *s)
{
    CGB_CHECK(s && s->other, "cgb_sampler_extra_initialization: sync() has not been called");
    if (s->sparse) { return CGB_OK; } // SparseNormalModel::extraInitialization is a nop (SparseNormalModel.cpp:33-37)
    CGB_CUDA(cudaSetDevice(s->device));
    CGB_CUDA(cudaStreamSynchronize(s->other->stream));
    dim3 grid(s->nRows, (s->L + 1023) / 1024);
    rebuild_ap_kernel<<<grid, 256, 0, s->stream>>>(s->dAP, s->dM, s->other->dM, s->nRows, s->L, s->k, s->ld, s->ldM, s->other->ldM);
    ++g_kernelLaunches;
    CGB_CUDA(cudaGetLastError());
    return CGB_OK;
}

extern "C" int cgb_sampler_extra_initialization(cgb_sampler *s)
{
    return guarded("cgb_sampler_extra_initialization", [&]() { return cgb_sampler_extra_initialization_body(s); });
}

// ------------------------------------------------------------------------------------------------
// evaluation of one conflict-free batch on the device
// ------------------------------------------------------------------------------------------------
static void fillModelView(const cgb_sampler *s, ModelView &mv)
{
    mv.D = s->dD;
    mv.S = s->hasS ? s->dS : nullptr;
    mv.AP = s->dAP;
    mv.M = s->dM;
    mv.otherM = s->other->dM;
    mv.otherColNonzero = s->other->dColNonzero;
    mv.erf = s->rs->dErf;
    mv.erfinv = s->rs->dErfinv;
    mv.outcomes = s->hOutcomes;
    mv.partials = s->dPartials;
    mv.tickets = s->dTickets;
    mv.phaseClocks = s->dPhaseClocks;
    mv.rowVersion = s->dRowVersion;
    mv.tablesInSmem = s->tablesInSmem ? 1u : 0u;
    mv.spRowPtr = s->dSpRowPtr;
    mv.spIdx = s->dSpIdx;
    mv.spVal = s->dSpVal;
    mv.Mrows = s->dMrows;
    mv.otherMrows = s->other->dMrows;
    mv.Z1 = s->dZ1;
    mv.Z2 = s->dZ2;
    mv.ldR = s->ldR;
    mv.beta = 100.f; // SparseNormalModel.h:77
    mv.nRows = s->nRows;
    mv.L = s->L;
    mv.k = s->k;
    mv.ld = s->ld;
    mv.ldM = s->ldM;
    mv.ldOther = s->other->ldM;
    mv.seg = s->seg;
    mv.nSeg = s->nSeg;
    mv.segPad = s->segPad;
    mv.lambda = s->lambda;
    mv.maxGibbsMass = s->maxGibbsMass;
    mv.annealingTemp = s->annealingTemp;
}

static size_t evalSmemBytes(const cgb_sampler *s)
{
    if (s->sparse) { return 256 + (static_cast<size_t>(s->ldR) + 4 * kSparseThreads * kSparseGroup) * sizeof(float); }
    return 256 + static_cast<size_t>(s->hasS ? 5 : 4) * s->segPad * sizeof(float);
}

// launches the eval kernel for params.nProps proposals already written to params.props; blocks until
// the outcomes are visible in s->hOutcomes
static int launchEval(cgb_sampler *s, EvalParams &params)
{
    uint32_t nExtra = 0;
    for (uint32_t i = 0; i < params.nProps; ++i)
    {
        const DevProposal &p = params.props[i];
        const bool pairType = p.type == 'M' || p.type == 'E' || (p.type == kProbe && p.variant == 1);
        if (pairType && p.r1 != p.r2) { params.extra[nExtra++] = static_cast<uint16_t>(i); }
    }
    params.nTasks = params.nProps + nExtra;

    if (s->sparse)
    {
        const double t0s = nowSeconds();
        if (s->timeKernels) { CGB_CUDA(cudaEventRecord(s->evStart, s->stream)); }
        eval_sparse_kernel<<<params.nTasks, kSparseThreads, evalSmemBytes(s), s->stream>>>(params);
        ++g_kernelLaunches;
        CGB_CUDA(cudaGetLastError());
        if (s->timeKernels) { CGB_CUDA(cudaEventRecord(s->evStop, s->stream)); }
        CGB_CUDA(cudaStreamSynchronize(s->stream));
        if (s->timeKernels)
        {
            float ms = 0.f;
            CGB_CUDA(cudaEventElapsedTime(&ms, s->evStart, s->evStop));
            s->counters.secondsKernel += static_cast<double>(ms) * 1e-3;
        }
        s->counters.secondsDeviceWait += nowSeconds() - t0s;
        s->counters.nBatches += 1;
        return CGB_OK;
    }
    cudaLaunchConfig_t cfg;
    cfg = cudaLaunchConfig_t();
    cfg.gridDim = dim3(s->nSeg, params.nTasks, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = evalSmemBytes(s);
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = s->nSeg;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;

    const double t0 = nowSeconds();
    if (s->timeKernels) { CGB_CUDA(cudaEventRecord(s->evStart, s->stream)); }
    if (s->hasS) { CGB_CUDA(cudaLaunchKernelEx(&cfg, eval_kernel<true>, params)); }
    else { CGB_CUDA(cudaLaunchKernelEx(&cfg, eval_kernel<false>, params)); }
    ++g_kernelLaunches;
    if (s->timeKernels) { CGB_CUDA(cudaEventRecord(s->evStop, s->stream)); }
    CGB_CUDA(cudaStreamSynchronize(s->stream));
    if (s->timeKernels)
    {
        float ms = 0.f;
        CGB_CUDA(cudaEventElapsedTime(&ms, s->evStart, s->evStop));
        s->counters.secondsKernel += static_cast<double>(ms) * 1e-3;
    }
    s->counters.secondsDeviceWait += nowSeconds() - t0;
    s->counters.nBatches += 1;
    if (s->dPhaseClocks)
    {
        std::vector<unsigned long long> h(static_cast<size_t>(params.nTasks) * kPhaseSlots);
        CGB_CUDA(cudaMemcpy(h.data(), s->dPhaseClocks, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        unsigned long long firstStart = ~0ull, lastStart = 0;
        for (uint32_t t = 0; t < params.nTasks; ++t)
        {
            const unsigned long long *p = h.data() + static_cast<size_t>(t) * kPhaseSlots;
            firstStart = std::min(firstStart, p[0]);
            lastStart = std::max(lastStart, p[0]);
            for (int i = 2; i < 9; ++i) { s->phaseSum[i] += static_cast<double>(p[i] - p[1]); }
        }
        s->phaseSum[0] += static_cast<double>(lastStart - firstStart); // ns between first and last CTA start
        s->phaseSum[1] += 1.0;                                          // launches
        s->phaseTasks += params.nTasks;
    }
    return CGB_OK;
}

// debug: per-phase SM-clock offsets of the leader CTA of every task, averaged (COGAPS phase profile)
static int cgb_sampler_debug_phase_clocks_body(cgb_sampler *s, int32_t enable, double *out, uint64_t *nTasks)
{
    CGB_CHECK(s != nullptr, "cgb_sampler_debug_phase_clocks: NULL sampler");
    CGB_CUDA(cudaSetDevice(s->device));
    if (out)
    {
        for (int i = 0; i < kPhaseSlots; ++i) { out[i] = s->phaseSum[i]; }
    }
    if (nTasks) { *nTasks = s->phaseTasks; }
    if (enable && !s->dPhaseClocks)
    {
        CGB_CUDA(cudaMalloc(&s->dPhaseClocks, sizeof(unsigned long long) * kPhaseSlots * 2 * kMaxBatch));
        CGB_CUDA(cudaMemset(s->dPhaseClocks, 0, sizeof(unsigned long long) * kPhaseSlots * 2 * kMaxBatch));
        CGB_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    }
    if (!enable && s->dPhaseClocks)
    {
        cudaFree(s->dPhaseClocks);
        s->dPhaseClocks = nullptr;
    }
    for (int i = 0; i < kPhaseSlots; ++i) { s->phaseSum[i] = 0.0; }
    s->phaseTasks = 0;
    return CGB_OK;
}

extern "C" int cgb_sampler_debug_phase_clocks(cgb_sampler *s, int32_t enable, double *out, uint64_t *nTasks)
{
    return guarded("cgb_sampler_debug_phase_clocks", [&]() { return cgb_sampler_debug_phase_clocks_body(s, enable, out, nTasks); });
}

// SURVEY 8(d), sparse model: the index bit-flags of the data column and of the factor column(s) — 2 x ceil(L/64)
// words of 8 bytes per scan, 3 for a same-row pair, twice 2 for a two-row pair; the per-element part (8 B + 4 k B per
// visited element) is counted on the device (StreamStats::visited) and added when the resident grid retires.
static double sparseFlagBytes(const DevProposal &p, uint32_t L)
{
    const double words = static_cast<double>((L + 63) / 64) * 8.0;
    if (p.type == 'B' || p.type == 'D') { return 2.0 * words; }
    return (p.r1 == p.r2) ? 3.0 * words : 4.0 * words;
}

// reference-formulation bytes of one evaluated proposal (fp32, L = scanned length)
static double algorithmicBytes(const DevProposal &p, const DevOutcome &o, uint32_t L)
{
    const double l = static_cast<double>(L);
    if (p.type == 'B') { return 16.0 * l + (o.accepted ? 4.0 * l : 0.0); }
    if (p.type == 'D') { return 16.0 * l + ((!o.accepted || o.mass1 != p.m1) ? 4.0 * l : 0.0); }
    const bool changed = o.accepted != 0;
    if (p.r1 == p.r2) { return 20.0 * l + (changed ? 4.0 * l : 0.0); }
    return (p.c1 == p.c2 ? 28.0 : 32.0) * l + (changed ? 8.0 * l : 0.0);
}

// ------------------------------------------------------------------------------------------------
// host <-> device proposal traffic
// ------------------------------------------------------------------------------------------------
static void fillProposal(const cgb_sampler *s, const HostProposal &hp, DevProposal &dp)
{
    dp.rng = hp.rng.state;
    dp.r1 = hp.r1; dp.c1 = hp.c1; dp.r2 = hp.r2; dp.c2 = hp.c2;
    dp.m1 = s->domain.atom(hp.atom1).mass;
    dp.m2 = (hp.atom2 != kNoAtom) ? s->domain.atom(hp.atom2).mass : 0.f;
    dp.type = static_cast<uint32_t>(hp.type);
    dp.variant = 0;
    dp.ch = 0.f;
    dp.pad = 0;
}

static inline bool isTwoRow(const DevProposal &p)
{
    return (p.type == 'M' || p.type == 'E') && p.r1 != p.r2;
}

// the host-visible half of AsynchronousGibbsSampler::birth/death/move/exchange (:126-219)
// rows whose AP line / factor element the device rewrites for an outcome: every CTA of the committing
// cluster bumps the row's version once and counts itself done once (kernels.cuh commit_task)
static inline void noteCommit(cgb_sampler *s, const DevProposal &dp, bool resident)
{
    s->rowVersion[dp.r1] += s->nSeg;
    if (isTwoRow(dp)) { s->rowVersion[dp.r2] += s->nSeg; }
    if (resident)
    {
        // the proposal travelled in the chunk whose tag is still current (outcomes are applied before the next
        // chunk is opened); every CTA of the committing cluster counts itself done once
        s->commitsExpected[s->mailSeq & 1ull] += s->nSeg;
        s->rowPending[dp.r1] = s->mailSeq;
        if (isTwoRow(dp)) { s->rowPending[dp.r2] = s->mailSeq; }
    }
}

// `resident`: the commit is made by the resident grid (no kernel boundary orders it before the next batch)
static int applyOutcome(cgb_sampler *s, const HostProposal &hp, const DevProposal &dp, bool resident, bool accepted, float mass1, float mass2)
{
    DevOutcome o;
    o.accepted = accepted ? 1u : 0u;
    o.mass1 = mass1;
    o.mass2 = mass2;
    s->counters.algorithmicBytes += s->sparse ? sparseFlagBytes(dp, s->L) : algorithmicBytes(dp, o, s->L);
    // rows whose AP line / factor element the device rewrites for this outcome: every CTA of the
    // committing cluster bumps the row's version once (kernels.cuh commit_task)
    if (hp.type != 'B' && hp.type != 'D' && hp.type != 'M' && hp.type != 'E') { return fail(CGB_EINTERNAL, "applyOutcome: corrupt proposal type"); }
    const bool commit = applyToDomain(s->domain, s->queue, hp, accepted, mass1, mass2, dp.m1);
    if (commit) { noteCommit(s, dp, resident); }
    return CGB_OK;
}

// true when the last commit to `row` is known to have landed, i.e. no rowVersion check is needed
static inline bool rowSettled(cgb_sampler *s, uint32_t row)
{
    const uint64_t pend = s->rowPending[row];
    if (pend == 0) { return true; }
    if (s->forceRowWait) { return false; } // tests (COGAPS_FORCE_ROW_WAIT): every such task checks rowVersion
    const uint64_t q = pend & 1ull;
    if (pend <= s->provenThrough[q]) { return true; }
    // Commits finish out of order, so only "all of them" proves anything, and only for the parity the chunk now
    // being posted does NOT use: its own commits are already landing in its own counter and could stand in for a
    // straggler of an earlier chunk.  (No outcome is applied while a chunk is posted, so the target stands still.)
    if (q != (s->mailSeq & 1ull) && s->hCommitsMirror[q] == s->commitsExpected[q])
    {
        s->provenThrough[q] = s->mailSeq - 1; // the newest chunk of that parity
        return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// resident mode: one grid per update(); every proposal is posted to its cluster's record ring the moment
// the generator queues it, outcomes come back as self-validating 16-byte records
// ------------------------------------------------------------------------------------------------
static size_t streamSmemBytes(const cgb_sampler *s) { return evalSmemBytes(s) + (s->tablesInSmem ? kStreamTableBytes : 0); }

static const int kResidentUnavailable = 1; // startPersistent / beginChunk: fall back to one launch per batch

static int startPersistent(cgb_sampler *s)
{
    if (g_serialisedLaunches.load() == 1) { s->usePersistent = false; return kResidentUnavailable; }
    cudaLaunchConfig_t cfg = cudaLaunchConfig_t();
    cfg.blockDim = dim3(s->sparse ? kSparseThreads : kThreads, 1, 1);
    cfg.dynamicSmemBytes = streamSmemBytes(s);
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = s->nSeg;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (s->persistentGrid == 0)
    {
        // every CTA polls for work until told to leave, so the whole grid must be resident at once;
        // the last cluster only mirrors the commit count, the others are workers
        cfg.gridDim = dim3(s->nSeg, 1, 1);
        int maxClusters = 0;
        if (s->sparse) { CGB_CUDA(cudaOccupancyMaxActiveClusters(&maxClusters, eval_stream_sparse_kernel, &cfg)); }
        else if (s->hasS) { CGB_CUDA(cudaOccupancyMaxActiveClusters(&maxClusters, eval_stream_kernel<true>, &cfg)); }
        else { CGB_CUDA(cudaOccupancyMaxActiveClusters(&maxClusters, eval_stream_kernel<false>, &cfg)); }
        const int cap = envInt("COGAPS_PERSISTENT_CLUSTERS", 0);
        if (cap > 1 && cap < maxClusters) { maxClusters = cap; }
        // several chains driven by several host threads share the device: every grid must fit beside the others
        const int share = s->residentShare > 0 ? s->residentShare : g_residentShare.load();
        if (share > 1) { maxClusters = std::max(2, maxClusters / share); }
        if (maxClusters < 2) { return fail(CGB_ECUDA, "resident kernel: fewer than two clusters fit on the device"); }
        s->persistentGrid = maxClusters * static_cast<int>(s->nSeg);
        s->nClusters = static_cast<uint32_t>(maxClusters - 1);
        s->nSlotRecords = static_cast<size_t>(s->nClusters) * kStreamRing * s->nSeg;
        CGB_CUDA(cudaHostAlloc(&s->hSlots, sizeof(StreamRecord) * (s->nSlotRecords + 1), cudaHostAllocMapped));
        std::memset(s->hSlots, 0, sizeof(StreamRecord) * (s->nSlotRecords + 1));
        s->slotOwner.assign(static_cast<size_t>(s->nClusters) * kStreamRing, ~0ull);
        void *alive = nullptr;
        CGB_CUDA(cudaHostAlloc(&alive, sizeof(uint32_t) * s->nClusters, cudaHostAllocMapped));
        std::memset(alive, 0, sizeof(uint32_t) * s->nClusters);
        s->hAlive = static_cast<volatile uint32_t*>(alive);
        s->clusterTicket.assign(s->nClusters, 0u);
        s->aliveSeen.assign(s->nClusters, 0);
    }
    cfg.gridDim = dim3(s->persistentGrid, 1, 1);
    ModelView mv;
    fillModelView(s, mv);
    mv.annealingTemp = s->annealingTemp; // constant for the whole update() this grid serves
    // every cluster's tickets start behind the highest one any cluster used under the previous grid
    std::fill(s->clusterTicket.begin(), s->clusterTicket.end(), s->ticketBase);
    std::fill(s->aliveSeen.begin(), s->aliveSeen.end(), 0);
    s->aliveList.clear();
    s->aliveNext = 0;
    s->aliveLooks = 0;
    if (++s->launchEpoch == 0u) { s->launchEpoch = 1u; }
    volatile unsigned long long *doorbell = reinterpret_cast<volatile unsigned long long*>(static_cast<StreamRecord*>(s->hSlots) + s->nSlotRecords);
    *doorbell = 0ull;
    __sync_synchronize();
    StreamParams sp;
    sp.slots = static_cast<const StreamRecord*>(s->hSlots);
    sp.doorbell = doorbell;
    sp.outcomes = static_cast<HostOutcome*>(s->hStreamOutcomes);
    sp.commitsMirror = s->hCommitsMirror;
    sp.stats = static_cast<StreamStats*>(s->dStreamStats);
    sp.ticket0 = s->ticketBase;
    sp.alive = const_cast<uint32_t*>(s->hAlive);
    sp.epoch = s->launchEpoch;
    sp.idleTimeoutNs = static_cast<unsigned long long>(envInt("COGAPS_PERSISTENT_IDLE_MS", 2000)) * 1000000ull;
    sp.nWorkers = s->nClusters;
    sp.pollSleepNs = static_cast<uint32_t>(envInt("COGAPS_POLL_SLEEP_NS", 0));
    CGB_CUDA(cudaMemsetAsync(sp.stats, 0, sizeof(StreamStats), s->stream));
    // the device counter restarts with the grid; nothing is pending across a kernel boundary
    s->hCommitsMirror[0] = s->hCommitsMirror[1] = 0ull;
    std::fill(s->rowPending.begin(), s->rowPending.end(), 0ull);
    s->commitsExpected[0] = s->commitsExpected[1] = 0;
    s->provenThrough[0] = s->provenThrough[1] = s->mailSeq;
    CGB_CUDA(cudaEventRecord(s->evStart, s->stream));
    const double tLaunch = nowSeconds();
    if (s->sparse) { CGB_CUDA(cudaLaunchKernelEx(&cfg, eval_stream_sparse_kernel, mv, sp)); }
    else if (s->hasS) { CGB_CUDA(cudaLaunchKernelEx(&cfg, eval_stream_kernel<true>, mv, sp)); }
    else { CGB_CUDA(cudaLaunchKernelEx(&cfg, eval_stream_kernel<false>, mv, sp)); }
    if (nowSeconds() - tLaunch > 0.5e-9 * static_cast<double>(sp.idleTimeoutNs))
    {
        // The launch call came back only after the workers' idle timeout: something unannounced serialises launches and
        // the grid has come and gone without us.  From here on this process uses one launch per batch.
        g_serialisedLaunches.store(1);
        s->usePersistent = false;
        ++g_kernelLaunches;
        CGB_CUDA(cudaStreamSynchronize(s->stream));
        return kResidentUnavailable;
    }
    // No CUDA call from here until the exit records are posted: the grid only ends when this thread says so, and a
    // runtime call can block behind another thread's cudaFree / allocation that is itself waiting for the device
    // to drain (several chains per process) — that would be a deadlock.  evStop is recorded in stopPersistent.
    ++g_kernelLaunches;
    s->persistentRunning = true;
    s->lastPostTime = nowSeconds();
    return CGB_OK;
}

// writes a record into its cluster's ring (one copy per CTA of the cluster); a torn read fails the
// checksum and is simply polled again
static inline void writeRecord(cgb_sampler *s, uint32_t cluster, uint32_t ticket, StreamRecord &rec)
{
    rec.ticket = ticket;
    rec.check = stream_check(reinterpret_cast<const uint32_t*>(&rec));
    StreamRecord *dst = static_cast<StreamRecord*>(s->hSlots)
        + (static_cast<size_t>(cluster) * kStreamRing + ((ticket - 1u) % kStreamRing)) * s->nSeg;
    for (uint32_t q = 0; q < s->nSeg; ++q) { std::memcpy(static_cast<void*>(dst + q), &rec, sizeof(rec)); }
}

static int stopPersistent(cgb_sampler *s)
{
    if (!s->persistentRunning) { return CGB_OK; }
    // an exit record in every worker cluster's next slot (all earlier records have been consumed: their
    // outcomes are in), and the exit bit for the mirror CTA
    StreamRecord rec;
    std::memset(&rec, 0, sizeof(rec));
    rec.type = kStreamExit;
    uint32_t top = s->ticketBase;
    for (uint32_t c = 0; c < s->nClusters; ++c)
    {
        writeRecord(s, c, s->clusterTicket[c] + 1u, rec);
        top = std::max(top, s->clusterTicket[c] + 1u);
    }
    s->ticketBase = top + 1u;
    volatile unsigned long long *doorbell = reinterpret_cast<volatile unsigned long long*>(static_cast<StreamRecord*>(s->hSlots) + s->nSlotRecords);
    *doorbell = kDoorbellExit;
    __sync_synchronize();
    CGB_CUDA(cudaEventRecord(s->evStop, s->stream)); // behind the grid on its stream: marks its end
    CGB_CUDA(cudaStreamSynchronize(s->stream));
    s->persistentRunning = false;
    s->provenThrough[0] = s->provenThrough[1] = s->mailSeq; // the grid has drained: every commit it made is complete
    float ms = 0.f;
    CGB_CUDA(cudaEventElapsedTime(&ms, s->evStart, s->evStop));
    s->counters.secondsKernel += static_cast<double>(ms) * 1e-3;
    if (s->sparse)
    {
        unsigned long long visited = 0;
        CGB_CUDA(cudaMemcpy(&visited, &static_cast<StreamStats*>(s->dStreamStats)->visited, sizeof(visited), cudaMemcpyDeviceToHost));
        s->counters.algorithmicBytes += static_cast<double>(visited) * (8.0 + 4.0 * s->k);
    }
    if (s->dPhaseClocks)
    {
        // phase profile: every task slot holds the stamps of the last task that used it
        std::vector<unsigned long long> h(static_cast<size_t>(2 * kMaxBatch) * kPhaseSlots);
        CGB_CUDA(cudaMemcpy(h.data(), s->dPhaseClocks, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        for (uint32_t t = 0; t < 2u * kMaxBatch; ++t)
        {
            const unsigned long long *p = h.data() + static_cast<size_t>(t) * kPhaseSlots;
            if (p[1] == 0 || p[8] <= p[1]) { continue; }
            for (int i = 2; i < 9; ++i) { s->phaseSum[i] += static_cast<double>(p[i] - p[1]); }
            s->phaseTasks += 1;
        }
        CGB_CUDA(cudaMemset(s->dPhaseClocks, 0, h.size() * sizeof(unsigned long long)));
    }
    if (envInt("COGAPS_PERSISTENT_DEBUG", 0))
    {
        StreamStats st;
        CGB_CUDA(cudaMemcpy(&st, s->dStreamStats, sizeof(st), cudaMemcpyDeviceToHost));
        const double nt = static_cast<double>(st.tasks ? st.tasks : 1), no = static_cast<double>(st.outcomes ? st.outcomes : 1);
        std::printf("[resident nSeg=%u grid=%d] kernel %.3f ms, tasks %llu | record seen -> commit done %.2f us (max %.2f) | "
                    "record seen -> outcome posted %.2f us | row-version wait %.3f us\n",
                    s->nSeg, s->persistentGrid, ms, st.tasks, st.taskNs / nt * 1e-3, st.maxTaskNs * 1e-3,
                    st.decideNs / no * 1e-3, st.verWaitNs / nt * 1e-3);
    }
    return CGB_OK;
}

// spins until the outcome of proposal `pi` of the current chunk is in host memory, then files it
static int waitOutcome(cgb_sampler *s, uint32_t pi)
{
    const size_t idx = s->chunkBase + pi;
    if (s->arrived[idx]) { return CGB_OK; }
    volatile HostOutcome *o = static_cast<HostOutcome*>(s->hStreamOutcomes) + pi;
    const uint32_t want = s->chunkTag;
    double t0 = 0.0; // the clock is only read when we actually have to wait
    uint64_t spins = 0;
    for (;;)
    {
        const uint32_t w2 = o->seqAndAccepted;
        if ((w2 >> 3) == want)
        {
            const uint32_t w0 = o->mass1Bits, w1 = o->mass2Bits, w3 = o->check;
            if (o->seqAndAccepted == w2 && w3 == outcome_check(w0, w1, w2))
            {
                DevOutcome &d = s->collected[idx];
                std::memcpy(&d.mass1, &w0, 4);
                std::memcpy(&d.mass2, &w1, 4);
                d.accepted = w2 & 1u;
                d.pad[0] = (w2 >> 1) & 3u;
                s->arrived[idx] = 1;
                break;
            }
        }
        if (spins == 0) { t0 = nowSeconds(); }
        __builtin_ia32_pause();
        if ((++spins & 0xfffff) == 0 && nowSeconds() - t0 > 20.0)
        {
            cudaError_t e = cudaStreamQuery(s->stream);
            s->persistentRunning = false;
            return fail(CGB_ECUDA, std::string("resident eval kernel did not answer within 20 s (stream state: ")
                + cudaGetErrorString(e) + "; waiting for outcome " + std::to_string(pi) + " of chunk " + std::to_string(s->mailSeq)
                + ", " + std::to_string(s->chunkPosted) + " posted; grid " + std::to_string(s->persistentGrid) + " nSeg "
                + std::to_string(s->nSeg) + ")");
        }
    }
    if (spins != 0) { s->counters.secondsDeviceWait += nowSeconds() - t0; }
    return CGB_OK;
}

// opens a chunk of at most kMaxPersistentBatch proposals whose outcomes share one tag
static int beginChunk(cgb_sampler *s, size_t chunkBase)
{
    const double t0 = nowSeconds();
    if (s->persistentRunning && t0 - s->lastPostTime > 1.0)
    {
        CGB_TRY(stopPersistent(s)); // the grid may be about to give up waiting (idle timeout): restart it
    }
    if (!s->persistentRunning)
    {
        const int rc = startPersistent(s);
        if (rc != CGB_OK) { return rc; } // including kResidentUnavailable
    }
    s->lastPostTime = t0;
    ++s->mailSeq;
    {
        // the parity this chunk will use: chunks of it up to mailSeq - 2 are provably complete if their counter
        // matches now, before this chunk adds to it; otherwise rows they touched keep their rowVersion check
        const uint64_t q = s->mailSeq & 1ull;
        if (s->mailSeq >= 2 && s->hCommitsMirror[q] == s->commitsExpected[q]) { s->provenThrough[q] = s->mailSeq - 2; }
    }
    s->chunkTag = static_cast<uint32_t>(s->mailSeq) & 0x1fffffffu;
    s->chunkPosted = 0;
    s->chunkBase = chunkBase;
    return CGB_OK;
}

static int postPrepared(cgb_sampler *s, size_t index);

// picks up clusters that have reported in since the last look
static void refreshAlive(cgb_sampler *s)
{
    for (uint32_t c = 0; c < s->nClusters; ++c)
    {
        if (!s->aliveSeen[c] && s->hAlive[c] == s->launchEpoch)
        {
            s->aliveSeen[c] = 1;
            s->aliveList.push_back(c);
        }
    }
}

static int waitForAnyCluster(cgb_sampler *s)
{
    const double t0 = nowSeconds();
    uint64_t spins = 0;
    while (s->aliveList.empty())
    {
        refreshAlive(s);
        __builtin_ia32_pause();
        if ((++spins & 0xffff) == 0 && nowSeconds() - t0 > 20.0)
        {
            return fail(CGB_ECUDA, "resident eval kernel: no cluster came up within 20 s (is the device fully occupied by other work?)");
        }
    }
    s->counters.secondsDeviceWait += nowSeconds() - t0;
    return CGB_OK;
}

// posts proposal `index` of the batch (ProposalQueue sink, or the chunk loop for very long batches)
static int postProposal(cgb_sampler *s, const HostProposal &hp, size_t index)
{
    if (index < s->chunkBase || index - s->chunkBase >= s->chunkCap) { return CGB_OK; }
    if (s->posted.size() <= index)
    {
        const size_t n = std::max<size_t>(index + 1, s->posted.size() * 2 + 256);
        s->posted.resize(n);
        s->collected.resize(n);
        s->arrived.resize(n, 0);
    }
    fillProposal(s, hp, s->posted[index]);
    return postPrepared(s, index);
}

// posts s->posted[index]
static int postPrepared(cgb_sampler *s, size_t index)
{
    DevProposal &dp = s->posted[index];
    s->arrived[index] = 0;
    const uint32_t pi = static_cast<uint32_t>(index - s->chunkBase);
    StreamRecord rec;
    rec.rng = dp.rng;
    rec.r1 = dp.r1; rec.c1 = dp.c1; rec.r2 = dp.r2; rec.c2 = dp.c2;
    rec.m1 = dp.m1; rec.m2 = dp.m2;
    rec.type = dp.type | ((dp.pad & 1u) ? kStreamSeq : 0u);
    rec.ver1 = s->rowVersion[dp.r1];
    rec.ver2 = 0u;
    if (!rowSettled(s, dp.r1)) { rec.type |= kStreamWait1; }
    if (dp.type == 'M' || dp.type == 'E')
    {
        rec.ver2 = s->rowVersion[dp.r2];
        if (!rowSettled(s, dp.r2)) { rec.type |= kStreamWait2; }
    }
    rec.pad = 0u;
    rec.batch = s->chunkTag;
    const uint32_t nParts = isTwoRow(dp) ? 2u : 1u;
    for (uint32_t part = 0; part < nParts; ++part)
    {
        // round robin over the clusters that have reported in
        if (s->aliveList.size() < s->nClusters && (++s->aliveLooks & 7u) == 0) { refreshAlive(s); } // until all are in
        if (s->aliveList.empty()) { CGB_TRY(waitForAnyCluster(s)); }
        if (s->aliveNext >= s->aliveList.size()) { s->aliveNext = 0; }
        const uint32_t cluster = s->aliveList[s->aliveNext++];
        const uint32_t ticket = ++s->clusterTicket[cluster];
        uint64_t &owner = s->slotOwner[static_cast<size_t>(cluster) * kStreamRing + ((ticket - 1u) % kStreamRing)];
        if ((owner >> 32) == (s->mailSeq & 0xffffffffull))
        {
            // the ring slot still holds a record of this chunk: its outcome proves it has been consumed
            CGB_TRY(waitOutcome(s, static_cast<uint32_t>(owner & 0xffffffffull)));
        }
        owner = ((s->mailSeq & 0xffffffffull) << 32) | pi;
        rec.piPart = pi | (part << 31);
        writeRecord(s, cluster, ticket, rec);
    }
    s->chunkPosted += 1;
    return CGB_OK;
}

struct SinkCtx
{
    cgb_sampler *s;
    int rc;
};

static void proposalSink(void *ctx, const HostProposal &hp, size_t index)
{
    SinkCtx *c = static_cast<SinkCtx*>(ctx);
    if (g_prof)
    {
        const unsigned long long t0 = tsc();
        if (c->rc == CGB_OK) { c->rc = postProposal(c->s, hp, index); }
        g_tscPost += tsc() - t0;
        ++g_nPosts;
        return;
    }
    if (c->rc == CGB_OK) { c->rc = postProposal(c->s, hp, index); }
}

static int evaluateQueue(cgb_sampler *s)
{
    std::vector<HostProposal> &q = s->queue.entries();
    size_t done = 0;
    if (s->usePersistent)
    {
        // the first chunk was streamed while the generator ran; longer batches continue in chunks
        while (done < q.size())
        {
            const size_t n = std::min<size_t>(s->chunkCap, q.size() - done);
            if (done > 0)
            {
                CGB_TRY(beginChunk(s, done));
                for (size_t i = 0; i < n; ++i) { CGB_TRY(postProposal(s, q[done + i], done + i)); }
            }
            // outcomes are applied in queue order as they arrive (generation is over, so the domain and
            // the atom-count window may change now); the early ones are long in, the tail hides the rest
            for (size_t i = 0; i < n; ++i)
            {
                const unsigned long long tw = g_prof ? tsc() : 0ull;
                CGB_TRY(waitOutcome(s, static_cast<uint32_t>(i)));
                if (g_prof) { g_tscWaitTail += tsc() - tw; }
                const DevOutcome &o = s->collected[done + i];
                const unsigned long long ta = g_prof ? tsc() : 0ull;
                CGB_TRY(applyOutcome(s, q[done + i], s->posted[done + i], true, o.accepted != 0u, o.mass1, o.mass2));
                if (g_prof) { g_tscApply += tsc() - ta; }
            }
            s->counters.nBatches += 1;
            done += n;
        }
        s->counters.nProposalsQueued += q.size();
        return CGB_OK;
    }
    static thread_local EvalParams params; // ~20 KB of kernel parameters, reused
    fillModelView(s, params.mv);
    while (done < q.size())
    {
        const uint32_t n = static_cast<uint32_t>(std::min<size_t>(kMaxBatch, q.size() - done));
        for (uint32_t i = 0; i < n; ++i) { fillProposal(s, q[done + i], params.props[i]); }
        params.nProps = n;
        CGB_TRY(launchEval(s, params));
        for (uint32_t i = 0; i < n; ++i)
        {
            const DevOutcome &o = s->hOutcomes[i];
            CGB_TRY(applyOutcome(s, q[done + i], params.props[i], false, o.accepted != 0u, o.mass1, o.mass2));
        }
        s->counters.nProposalsQueued += n;
        done += n;
    }
    return CGB_OK;
}

// ------------------------------------------------------------------------------------------------
// SingleThreadedGibbsSampler (gibbs_sampler/SingleThreadedGibbsSampler.h:94-257): one proposal at a time,
// one rng stream.  The host draws the type and the atoms, the device evaluates (same kernels, one task in
// flight) and reports how many draws it took from the stream; the host applies the outcome at once.
// ------------------------------------------------------------------------------------------------
static int evalOne(cgb_sampler *s, DevProposal &dp, DevOutcome &o)
{
    dp.pad = 1u;
    dp.variant = 0u;
    dp.ch = 0.f;
    dp.rng = s->seq.rng.state;
    if (s->usePersistent)
    {
        const int rcBegin = beginChunk(s, 0);
        if (rcBegin != CGB_OK && rcBegin != kResidentUnavailable) { return rcBegin; }
    }
    if (s->usePersistent)
    {
        if (s->posted.empty())
        {
            s->posted.resize(256);
            s->collected.resize(256);
            s->arrived.resize(256, 0);
        }
        s->posted[0] = dp;
        CGB_TRY(postPrepared(s, 0));
        CGB_TRY(waitOutcome(s, 0));
        o = s->collected[0];
    }
    else
    {
        static thread_local EvalParams params;
        fillModelView(s, params.mv);
        params.props[0] = dp;
        params.nProps = 1;
        CGB_TRY(launchEval(s, params));
        o = s->hOutcomes[0];
    }
    for (uint32_t d = 0; d < o.pad[0]; ++d) { s->seq.rng.advance(); }
    s->counters.nBatches += s->usePersistent ? 1 : 0;
    s->counters.nProposalsQueued += 1;
    s->counters.algorithmicBytes += s->sparse ? sparseFlagBytes(dp, s->L) : algorithmicBytes(dp, o, s->L);
    return CGB_OK;
}

static int sequentialUpdate(cgb_sampler *s, uint32_t nSteps)
{
    SequentialState &q = s->seq;
    AtomicDomain &dom = s->domain;
    for (uint32_t step = 0; step < nSteps; ++step)
    {
        // getUpdateType, :94-111
        char type = 'B';
        if (dom.size() >= 2)
        {
            const float u1 = q.rng.uniform();
            if (u1 < 0.5f)
            {
                const double nAtoms = static_cast<double>(dom.size());
                const double numer = nAtoms * q.domainLength;
                const float deathProb = static_cast<float>(numer / (numer + q.alpha * q.numBins * (q.domainLength - nAtoms)));
                type = (q.rng.uniform() < deathProb) ? 'D' : 'B';
            }
            else
            {
                type = (u1 < 0.75f) ? 'M' : 'E';
            }
        }
        DevProposal dp;
        std::memset(&dp, 0, sizeof(dp));
        dp.type = static_cast<uint32_t>(type);
        DevOutcome o;
        if (type == 'B')
        {
            // birth, :130-150
            uint64_t pos = q.rng.uniform64(1, dom.domainLength());
            while (dom.occupied(pos)) { pos = q.rng.uniform64(1, dom.domainLength()); }
            q.binOf(pos, dp.r1, dp.c1);
            CGB_TRY(evalOne(s, dp, o));
            if (o.accepted)
            {
                dom.insert(pos, o.mass1);
                noteCommit(s, dp, s->usePersistent);
            }
        }
        else if (type == 'D')
        {
            // death, :154-187
            const uint32_t id = dom.atIndex(q.rng.uniform32(0, static_cast<uint32_t>(dom.size() - 1)));
            q.binOf(dom.atom(id).pos, dp.r1, dp.c1);
            dp.m1 = dom.atom(id).mass;
            CGB_TRY(evalOne(s, dp, o));
            if (o.accepted)
            {
                if (o.mass1 != dp.m1)
                {
                    dom.atom(id).mass = o.mass1;
                    noteCommit(s, dp, s->usePersistent);
                }
            }
            else
            {
                dom.erase(id);
                noteCommit(s, dp, s->usePersistent);
            }
        }
        else if (type == 'M')
        {
            // move, :190-219
            const uint32_t id = dom.atIndex(q.rng.uniform32(0, static_cast<uint32_t>(dom.size() - 1)));
            const Atom &center = dom.atom(id);
            const uint64_t lbound = (center.left != kNoAtom) ? dom.atom(center.left).pos : 0;
            const uint64_t rbound = (center.right != kNoAtom) ? dom.atom(center.right).pos : referenceDoubleToU64(q.domainLength);
            const uint64_t pos = q.rng.uniform64(lbound + 1, rbound - 1);
            q.binOf(center.pos, dp.r1, dp.c1);
            q.binOf(pos, dp.r2, dp.c2);
            if (dp.r1 == dp.r2 && dp.c1 == dp.c2)
            {
                dom.move(id, pos);
                continue;
            }
            dp.m1 = center.mass;
            CGB_TRY(evalOne(s, dp, o));
            if (o.accepted)
            {
                dom.move(id, pos);
                noteCommit(s, dp, s->usePersistent);
            }
        }
        else
        {
            // exchange, :223-257 (same-bin exchanges are ignored; canUseGibbs(c1, c2) is tested on the device)
            const uint32_t id1 = dom.atIndex(q.rng.uniform32(0, static_cast<uint32_t>(dom.size() - 1)));
            const uint32_t right = dom.atom(id1).right;
            const uint32_t id2 = (right != kNoAtom) ? right : dom.front();
            q.binOf(dom.atom(id1).pos, dp.r1, dp.c1);
            q.binOf(dom.atom(id2).pos, dp.r2, dp.c2);
            if (dp.r1 == dp.r2 && dp.c1 == dp.c2) { continue; }
            dp.m1 = dom.atom(id1).mass;
            dp.m2 = dom.atom(id2).mass;
            CGB_TRY(evalOne(s, dp, o));
            if (o.accepted)
            {
                dom.atom(id1).mass = o.mass1;
                dom.atom(id2).mass = o.mass2;
                noteCommit(s, dp, s->usePersistent);
            }
        }
    }
    return CGB_OK;
}

// AsynchronousGibbsSampler::update, AsynchronousGibbsSampler.h:88-122

// ------------------------------------------------------------------------------------------------
// Row-parallel sweep (sweep.cuh; cgb_params.updateMode == CGB_UPDATE_SWEEP): the whole update() is ONE launch with one
// CTA per row of the factor matrix.  The atoms live on the device, one run sorted by position per row; the host only
// draws the Philox key from the seeder, forms the proposal weights, and reads the counters back.
// ------------------------------------------------------------------------------------------------
static const uint32_t kSweepInitialCap = 64;
static const size_t kSweepMaxSmem = 227u * 1024u;
static const uint32_t kSweepOrderFromRows = 148; // fewer rows than SMs: nothing to order
static const uint32_t kSweepThreadsLong = 512;    // threads per row beyond 10240 floats (measured at C3's 20000-long rows, see sweepStageFor)

static uint32_t sweepThreadsForLength(uint32_t L)
{
    const int forcedLong = envInt("COGAPS_SWEEP_THREADS_LONG", 0); // experiments: rows beyond 10240 floats only
    if (L > 10240u && (forcedLong == 256 || forcedLong == 512 || forcedLong == 1024)) { return static_cast<uint32_t>(forcedLong); }
    const int forced = envInt("COGAPS_SWEEP_THREADS", 0); // experiments: 128 / 256 / 512 / 1024 (changes the reduction order)
    if (forced == 128 || forced == 256 || forced == 512 || forced == 1024) { return static_cast<uint32_t>(forced); }
    return L <= 10240u ? 256u : kSweepThreadsLong;
}
static const int kSweepKeep = 10; // float4 per thread and column kept in registers between scan and commit (512-thread rows up to 20480 floats)

static int cgb_sweep_reduction_order_for_length_body(uint32_t rowLength, cgb_reduction_order *out)
{
    CGB_CHECK(out && rowLength, "cgb_sweep_reduction_order_for_length: bad argument");
    out->threadsPerSegment = sweepThreadsForLength(rowLength);
    out->vectorWidth = kVec;
    out->nSegments = 1;
    out->segmentLength = roundUp(rowLength, 4);
    return CGB_OK;
}

extern "C" int cgb_sweep_reduction_order_for_length(uint32_t rowLength, cgb_reduction_order *out)
{
    return guarded("cgb_sweep_reduction_order_for_length", [&]() { return cgb_sweep_reduction_order_for_length_body(rowLength, out); });
}

// The proposal weights of one update(), in f64, operation for operation what oracle/cogaps_oracle.c sweep_rates does
// (this file is compiled with -ffp-contract=off): a birth falls into a given row with probability (1 - pDeath)/2/nRows;
// a given atom is picked for a death / move / exchange with probability pDeath/2/n, 1/4/n, 1/4/n
// (ProposalQueue.cpp:123-160, with the atom count frozen at the start of the update).
struct SweepRates { double birthRow, deathAtom, moveAtom, exchAtom, perAtom; };

static SweepRates sweepRates(uint64_t nAtoms, uint32_t nRows, uint32_t k, uint64_t binLength, double alpha)
{
    SweepRates w;
    if (nAtoms < 2)
    {
        w.birthRow = 1.0 / static_cast<double>(nRows); // "always birth when 0 or 1 atoms exist", ProposalQueue.cpp:139-142
        w.deathAtom = w.moveAtom = w.exchAtom = 0.0;
    }
    else
    {
        const uint64_t nElements = static_cast<uint64_t>(nRows) * k;
        const double domainLength = static_cast<double>(binLength * nElements);
        const double numer = static_cast<double>(nAtoms) * domainLength;
        const float pDeath = static_cast<float>(numer / (numer + alpha * static_cast<double>(nElements) * (domainLength - static_cast<double>(nAtoms))));
        w.birthRow = 0.5 * (1.0 - static_cast<double>(pDeath)) / static_cast<double>(nRows);
        w.deathAtom = 0.5 * static_cast<double>(pDeath) / static_cast<double>(nAtoms);
        w.moveAtom = 0.25 / static_cast<double>(nAtoms);
        w.exchAtom = 0.25 / static_cast<double>(nAtoms);
    }
    w.perAtom = w.deathAtom + w.moveAtom + w.exchAtom;
    return w;
}

// (re)allocates the per-row atom store with `cap` slots per row, keeping what the rows hold
static int sweepEnsureStore(cgb_sampler *s, uint32_t cap)
{
    if (s->dSwCounters == nullptr)
    {
        CGB_CUDA(cudaMalloc(&s->dSwCounters, sizeof(SweepCounters)));
        CGB_CUDA(cudaHostAlloc(&s->hSwCounters, sizeof(SweepCounters), cudaHostAllocDefault));
    }
    if (s->dSwPos != nullptr && cap <= s->swCap) { return CGB_OK; }
    uint64_t *pos = nullptr;
    float *mass = nullptr;
    CGB_CUDA(cudaMalloc(&pos, static_cast<size_t>(s->nRows) * cap * sizeof(uint64_t)));
    if (cudaMalloc(&mass, static_cast<size_t>(s->nRows) * cap * sizeof(float)) != cudaSuccess)
    {
        cudaFree(pos);
        return fail(CGB_ENOMEM, "sweep: out of device memory for the atom store");
    }
    if (s->dSwPos != nullptr)
    {
        sweep_regrow_kernel<<<s->nRows, 64, 0, s->stream>>>(s->dSwPos, s->dSwMass, s->dSwCount, s->swCap, pos, mass, cap, s->nRows);
        g_kernelLaunches.fetch_add(1);
        CGB_CUDA(cudaGetLastError());
        CGB_CUDA(cudaStreamSynchronize(s->stream));
        cudaFree(s->dSwPos);
        cudaFree(s->dSwMass);
    }
    else
    {
        CGB_CUDA(cudaMalloc(&s->dSwCount, static_cast<size_t>(s->nRows) * sizeof(uint32_t)));
        CGB_CUDA(cudaMemsetAsync(s->dSwCount, 0, static_cast<size_t>(s->nRows) * sizeof(uint32_t), s->stream));
        CGB_CUDA(cudaStreamSynchronize(s->stream));
        s->swTotalAtoms = 0;
    }
    s->dSwPos = pos;
    s->dSwMass = mass;
    s->swCap = cap;
    return CGB_OK;
}

static uint32_t sweepCapFor(uint32_t cap, uint32_t maxCount)
{
    // the store grows between updates so that a row practically never fills up within one (oracle: sweep_update)
    return (2u * maxCount > cap) ? roundUp(4u * maxCount, 32u) : cap;
}

// host atomic domain -> per-row store (entering sweep mode)
static int sweepFromDomain(cgb_sampler *s)
{
    const uint64_t n = s->domain.size();
    const uint64_t binLength = s->domain.binLength();
    const uint64_t Lseg = binLength * s->k;
    std::vector<std::pair<uint64_t, float> > atoms(n);
    for (uint64_t i = 0; i < n; ++i)
    {
        const Atom &a = s->domain.atom(s->domain.atIndex(static_cast<uint32_t>(i)));
        atoms[i] = std::make_pair(a.pos, a.mass);
    }
    std::sort(atoms.begin(), atoms.end());
    std::vector<uint32_t> count(s->nRows, 0u);
    std::vector<uint32_t> rowOf(n);
    uint32_t maxCount = 0;
    for (uint64_t i = 0; i < n; ++i)
    {
        uint64_t row = atoms[i].first / Lseg;
        if (row >= s->nRows) { row = s->nRows - 1; } // pos == domainLength lies one past the last bin (ProposalQueue.cpp:214)
        rowOf[i] = static_cast<uint32_t>(row);
        maxCount = std::max(maxCount, ++count[row]);
    }
    uint32_t cap = std::max(kSweepInitialCap, s->swCap);
    cap = std::max(cap, sweepCapFor(cap, maxCount));
    if (s->dSwPos != nullptr && cap > s->swCap)
    {
        cudaFree(s->dSwPos); cudaFree(s->dSwMass); cudaFree(s->dSwCount);
        s->dSwPos = nullptr; s->dSwMass = nullptr; s->dSwCount = nullptr; s->swCap = 0;
    }
    CGB_TRY(sweepEnsureStore(s, cap));
    std::vector<uint64_t> pos(static_cast<size_t>(s->nRows) * s->swCap, 0ull);
    std::vector<float> mass(static_cast<size_t>(s->nRows) * s->swCap, 0.f);
    std::fill(count.begin(), count.end(), 0u);
    for (uint64_t i = 0; i < n; ++i)
    {
        const uint32_t row = rowOf[i];
        uint64_t local = atoms[i].first - static_cast<uint64_t>(row) * Lseg;
        if (local >= Lseg) { local = Lseg - 1; }
        const size_t at = static_cast<size_t>(row) * s->swCap + count[row]++;
        pos[at] = local;
        mass[at] = atoms[i].second;
    }
    CGB_CUDA(cudaMemcpy(s->dSwPos, pos.data(), pos.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
    CGB_CUDA(cudaMemcpy(s->dSwMass, mass.data(), mass.size() * sizeof(float), cudaMemcpyHostToDevice));
    CGB_CUDA(cudaMemcpy(s->dSwCount, count.data(), count.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CGB_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    s->swTotalAtoms = n;
    return CGB_OK;
}

// per-row store -> (position in the whole domain, mass), rows in order, ascending position within a row
static int sweepDownloadAtoms(const cgb_sampler *s, std::vector<uint64_t> &posOut, std::vector<float> &massOut)
{
    posOut.clear();
    massOut.clear();
    if (s->dSwPos == nullptr) { return CGB_OK; }
    CGB_CUDA(cudaStreamSynchronize(s->stream));
    std::vector<uint32_t> count(s->nRows);
    std::vector<uint64_t> pos(static_cast<size_t>(s->nRows) * s->swCap);
    std::vector<float> mass(static_cast<size_t>(s->nRows) * s->swCap);
    CGB_CUDA(cudaMemcpy(count.data(), s->dSwCount, count.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CGB_CUDA(cudaMemcpy(pos.data(), s->dSwPos, pos.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    CGB_CUDA(cudaMemcpy(mass.data(), s->dSwMass, mass.size() * sizeof(float), cudaMemcpyDeviceToHost));
    const uint64_t Lseg = (0xFFFFFFFFFFFFFFFFull / (static_cast<uint64_t>(s->nRows) * s->k)) * s->k;
    for (uint32_t r = 0; r < s->nRows; ++r)
    {
        CGB_CHECK(count[r] <= s->swCap, "sweep: a row's atom count exceeds its store");
        for (uint32_t i = 0; i < count[r]; ++i)
        {
            posOut.push_back(static_cast<uint64_t>(r) * Lseg + pos[static_cast<size_t>(r) * s->swCap + i]);
            massOut.push_back(mass[static_cast<size_t>(r) * s->swCap + i]);
        }
    }
    return CGB_OK;
}

static int setAtoms(cgb_sampler *s, const uint64_t *pos, const float *mass, uint64_t n);

static int cgb_sampler_set_update_mode_body(cgb_sampler *s, int32_t mode)
{
    CGB_CHECK(s != nullptr, "cgb_sampler_set_update_mode: NULL sampler");
    CGB_CHECK(mode == CGB_UPDATE_EXACT || mode == CGB_UPDATE_SWEEP, "cgb_sampler_set_update_mode: unknown mode");
    CGB_CHECK(!s->persistentRunning, "cgb_sampler_set_update_mode: called in the middle of an update");
    if (mode == s->updateMode) { return CGB_OK; }
    CGB_CUDA(cudaSetDevice(s->device));
    if (mode == CGB_UPDATE_SWEEP)
    {
        CGB_TRY(sweepFromDomain(s));
    }
    else
    {
        std::vector<uint64_t> pos;
        std::vector<float> mass;
        CGB_TRY(sweepDownloadAtoms(s, pos, mass));
        CGB_TRY(setAtoms(s, pos.data(), mass.data(), pos.size()));
    }
    s->updateMode = mode;
    return CGB_OK;
}

extern "C" int cgb_sampler_set_update_mode(cgb_sampler *s, int32_t mode)
{
    return guarded("cgb_sampler_set_update_mode", [&]() { return cgb_sampler_set_update_mode_body(s, mode); });
}

// The opt-in shared-memory limit is a property of the kernel instance, not of a sampler: it only ever grows, under a
// lock (several chains may share the process).
template <class Kernel>
static int sweepGrowSmemLimit(Kernel kernel, size_t smem)
{
    static std::mutex lock;
    static std::map<std::pair<int, const void*>, size_t> configured; // per device and instance (every instantiation has the same pointer type)
    int device = 0;
    CGB_CUDA(cudaGetDevice(&device));
    std::lock_guard<std::mutex> hold(lock);
    size_t &have = configured[std::make_pair(device, reinterpret_cast<const void*>(kernel))];
    if (have == 0u) { have = 48u * 1024u; }
    if (smem > have)
    {
        CGB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        have = smem;
    }
    return CGB_OK;
}

template <int T, int NV, bool HAS_S, int STAGE>
static int sweepLaunchInstance(cgb_sampler *s, const SweepArgs &args, size_t smem)
{
    CGB_TRY(sweepGrowSmemLimit(sweep_kernel<T, NV, HAS_S, STAGE>, smem));
    sweep_kernel<T, NV, HAS_S, STAGE><<<s->nRows, T, smem, s->stream>>>(args);
    g_kernelLaunches.fetch_add(1);
    return CGB_OK;
}

// stage: 0 rows walked where they live, 1 D and AP lines in shared memory, 2 the AP line only
template <int T, bool HAS_S>
static int sweepLaunchRows(cgb_sampler *s, const SweepArgs &args, size_t smem, int stage)
{
    constexpr int kKeep = (T == 512) ? kSweepKeep : 0; // at 1024 threads (64 registers) keeping the columns spills: measured slower
    const bool keep = stage == 1 && kKeep > 0 && (s->L + kVec - 1) / kVec <= static_cast<uint32_t>(kKeep) * T && envInt("COGAPS_SWEEP_KEEP", 1) != 0;
    if (stage == 0) { return sweepLaunchInstance<T, 0, HAS_S, 0>(s, args, smem); }
    if (stage == 2) { return sweepLaunchInstance<T, 0, HAS_S, 2>(s, args, smem); }
    if (keep) { return sweepLaunchInstance<T, kKeep, HAS_S, 1>(s, args, smem); }
    return sweepLaunchInstance<T, 0, HAS_S, 1>(s, args, smem);
}

static size_t sweepSparseScanBytes() { return static_cast<size_t>(4) * kSparseThreads * kSparseGroup * sizeof(float); }

template <int T, bool HAS_S, bool SPARSE = false>
static int sweepLaunchTransport(cgb_sampler *s, SweepArgs &args)
{
    // pairs (r, r+1) with even r on even-numbered updates of this sampler, odd r on odd-numbered ones
    const uint32_t colour = static_cast<uint32_t>(s->swUpdates & 1ull);
    const uint32_t pairs = (s->nRows - colour) / 2u;
    if (s->nRows >= 2u && pairs > 0u)
    {
        args.colour = colour;
        const size_t dyn = SPARSE ? ((static_cast<size_t>(s->ldR) * sizeof(float) + 127u) & ~static_cast<size_t>(127u)) + sweepSparseScanBytes() : 0u;
        sweep_transport_kernel<T, HAS_S, SPARSE><<<pairs, T, dyn, s->stream>>>(args);
        g_kernelLaunches.fetch_add(1);
        CGB_CUDA(cudaGetLastError());
    }
    return CGB_OK;
}

static int sweepUpdate(cgb_sampler *s, uint32_t nSteps)
{
    const double t0 = nowSeconds();
    CGB_TRY(sweepEnsureStore(s, std::max(kSweepInitialCap, s->swCap)));
    SweepArgs args;
    fillModelView(s, args.mv);
    args.pos = s->dSwPos;
    args.mass = s->dSwMass;
    args.count = s->dSwCount;
    args.qgamma = s->rs->dQgamma;
    args.counters = static_cast<SweepCounters*>(s->dSwCounters);
    args.order = nullptr;
    args.key = s->rs->seeder.next();                      // one seeder value per update()
    args.binLength = 0xFFFFFFFFFFFFFFFFull / (static_cast<uint64_t>(s->nRows) * s->k);
    args.binMagic = (args.binLength > 1) ? static_cast<uint64_t>((static_cast<unsigned __int128>(1) << 64) / args.binLength) : 0ull;
    args.colour = 0;
    args.profile = envInt("COGAPS_SWEEP_PROFILE", 0) != 0 ? 1u : 0u;
    const SweepRates w = sweepRates(s->swTotalAtoms, s->nRows, s->k, args.binLength, static_cast<double>(s->alpha));
    args.birthRow = w.birthRow;
    args.deathAtom = w.deathAtom;
    args.moveAtom = w.moveAtom;
    args.exchAtom = w.exchAtom;
    args.perAtom = w.perAtom;
    args.cap = s->swCap;
    args.nSteps = nSteps;
    const size_t base = s->sparse ? sweepSparseScanOffset(s->swCap, s->k, s->ldR) + sweepSparseScanBytes() : sweepRowOffset(s->swCap, s->k);
    const size_t lineBytes = static_cast<size_t>(s->ld) * sizeof(float);
    // What a row keeps in shared memory (COGAPS_SWEEP_STAGE overrides): its D and AP lines where four rows fit on an SM that
    // way (rows up to 10240 floats); beyond, the AP line only — two 20000-float rows per SM with D through L2 measured 15 %
    // faster than one row with both lines (C3's P side).  A row that does not fit falls back to the next smaller footprint.
    int stage = s->sparse ? 0 : envInt("COGAPS_SWEEP_STAGE", envInt("COGAPS_SWEEP_ROW_SMEM", 1) != 0 ? (s->L <= 10240u ? 1 : 2) : 0);
    if (stage == 1 && base + (s->hasS ? 3u : 2u) * lineBytes > kSweepMaxSmem) { stage = 2; }
    if (stage == 2 && base + lineBytes > kSweepMaxSmem) { stage = 0; }
    if (stage < 0 || stage > 2) { stage = 0; }
    const size_t smem = base + (stage == 1 ? (s->hasS ? 3u : 2u) * lineBytes : (stage == 2 ? lineBytes : 0u));
    CGB_CHECK(smem <= kSweepMaxSmem, "sweep: a row's atom store no longer fits in shared memory");
    CGB_CUDA(cudaMemsetAsync(s->dSwCounters, 0, sizeof(SweepCounters), s->stream));
    CGB_CUDA(cudaEventRecord(s->evStart, s->stream));
    if (s->nRows > kSweepOrderFromRows && envInt("COGAPS_SWEEP_ORDER", 1) != 0)
    {
        // longest chains first (sweep_order_kernel); with few rows every row has a CTA slot of its own from the start
        if (s->dSwOrder == nullptr) { CGB_CUDA(cudaMalloc(&s->dSwOrder, static_cast<size_t>(s->nRows) * sizeof(uint32_t))); }
        sweep_order_kernel<<<1, 1024, 0, s->stream>>>(s->dSwCount, s->nRows, s->dSwOrder);
        g_kernelLaunches.fetch_add(1);
        CGB_CUDA(cudaGetLastError());
        args.order = s->dSwOrder;
    }
    const uint32_t T = s->sparse ? static_cast<uint32_t>(kSparseThreads) : sweepThreadsForLength(s->L);
    if (s->sparse)
    {
        CGB_TRY(sweepGrowSmemLimit(sweep_sparse_kernel, smem));
        sweep_sparse_kernel<<<s->nRows, kSparseThreads, smem, s->stream>>>(args);
        g_kernelLaunches.fetch_add(1);
    }
    else if (T == 128u)
    {
        if (s->hasS) { CGB_TRY((sweepLaunchRows<128, true>(s, args, smem, stage))); } else { CGB_TRY((sweepLaunchRows<128, false>(s, args, smem, stage))); }
    }
    else if (T == 256u)
    {
        if (s->hasS) { CGB_TRY((sweepLaunchRows<256, true>(s, args, smem, stage))); } else { CGB_TRY((sweepLaunchRows<256, false>(s, args, smem, stage))); }
    }
    else if (T == 1024u)
    {
        if (s->hasS) { CGB_TRY((sweepLaunchRows<1024, true>(s, args, smem, stage))); } else { CGB_TRY((sweepLaunchRows<1024, false>(s, args, smem, stage))); }
    }
    else
    {
        if (s->hasS) { CGB_TRY((sweepLaunchRows<512, true>(s, args, smem, stage))); } else { CGB_TRY((sweepLaunchRows<512, false>(s, args, smem, stage))); }
    }
    CGB_CUDA(cudaGetLastError());
    if (envInt("COGAPS_SWEEP_TRANSPORT", 1) != 0)
    {
        // transport between adjacent rows (see sweep.cuh)
        if (s->sparse) { CGB_TRY((sweepLaunchTransport<kSparseThreads, false, true>(s, args))); }
        else if (T == 128u)
        {
            if (s->hasS) { CGB_TRY((sweepLaunchTransport<128, true>(s, args))); } else { CGB_TRY((sweepLaunchTransport<128, false>(s, args))); }
        }
        else if (T == 256u)
        {
            if (s->hasS) { CGB_TRY((sweepLaunchTransport<256, true>(s, args))); } else { CGB_TRY((sweepLaunchTransport<256, false>(s, args))); }
        }
        else if (T == 1024u)
        {
            if (s->hasS) { CGB_TRY((sweepLaunchTransport<1024, true>(s, args))); } else { CGB_TRY((sweepLaunchTransport<1024, false>(s, args))); }
        }
        else
        {
            if (s->hasS) { CGB_TRY((sweepLaunchTransport<512, true>(s, args))); } else { CGB_TRY((sweepLaunchTransport<512, false>(s, args))); }
        }
    }
    s->swUpdates += 1;
    CGB_CUDA(cudaEventRecord(s->evStop, s->stream));
    sweep_max_count_kernel<<<(s->nRows + 255u) / 256u, 256, 0, s->stream>>>(s->dSwCount, s->nRows, &static_cast<SweepCounters*>(s->dSwCounters)->maxCount);
    g_kernelLaunches.fetch_add(1);
    CGB_CUDA(cudaGetLastError());
    CGB_CUDA(cudaMemcpyAsync(s->hSwCounters, s->dSwCounters, sizeof(SweepCounters), cudaMemcpyDeviceToHost, s->stream));
    const double tw = nowSeconds();
    CGB_CUDA(cudaStreamSynchronize(s->stream));
    s->counters.secondsDeviceWait += nowSeconds() - tw;
    const SweepCounters &c = *static_cast<const SweepCounters*>(s->hSwCounters);
    float ms = 0.f;
    CGB_CUDA(cudaEventElapsedTime(&ms, s->evStart, s->evStop));
    s->counters.secondsKernel += static_cast<double>(ms) * 1e-3;
    s->counters.nBatches += 1;
    s->counters.nProposalsQueued += c.scans1 + c.scans2 + c.scansX;
    s->counters.nProposalsTotal += c.steps;
    if (s->sparse)
    {
        // SURVEY 8(d), sparse: 2 flag words streams per single-column scan (3 for a same-row pair, 4 for a two-row one) of
        // ceil(L / 64) * 8 bytes, plus 8 + 4k bytes per common non-zero
        const double flagBytes = static_cast<double>((s->L + 63u) / 64u) * 8.0;
        s->counters.algorithmicBytes += flagBytes * (2.0 * static_cast<double>(c.scans1) + 3.0 * static_cast<double>(c.scans2) + 4.0 * static_cast<double>(c.scansX))
            + static_cast<double>(c.visited) * (8.0 + 4.0 * s->k);
    }
    else
    {
        s->counters.algorithmicBytes += static_cast<double>(s->L) * (16.0 * static_cast<double>(c.scans1) + 20.0 * static_cast<double>(c.scans2) + 32.0 * static_cast<double>(c.scansX) + 4.0 * static_cast<double>(c.commits));
    }
    if (args.profile != 0u)
    {
        // debug: where thread 0 of a row's CTA spends its cycles, per evaluated proposal
        const double nEval = static_cast<double>(std::max<unsigned long long>(1ull, c.scans1 + c.scans2));
        static const char *const names[8] = {"wait:published", "scan+reduce", "decide", "wait:decided", "commit", "propose next", "no-eval proposals", "tail"};
        std::fprintf(stderr, "[sweep profile L=%u T=%u rows=%llu evals=%.0f steps=%llu kernel %.3f ms] cycles per evaluated proposal:", s->L, T,
                     c.rowsActive, nEval, c.steps, ms);
        double tot = 0.0;
        for (int i = 0; i < 8; ++i)
        {
            std::fprintf(stderr, " %s %.0f;", names[i], static_cast<double>(c.phase[i]) / nEval);
            tot += static_cast<double>(c.phase[i]);
        }
        std::fprintf(stderr, " total %.0f\n", tot / nEval);
    }
    s->swTotalAtoms = static_cast<uint64_t>(static_cast<long long>(s->swTotalAtoms) + c.atomDelta);
    s->swOverflow += c.overflow;
    const uint32_t cap = sweepCapFor(s->swCap, c.maxCount);
    if (cap > s->swCap) { CGB_TRY(sweepEnsureStore(s, cap)); }
    s->counters.secondsHostGenerate += tw - t0;
    return CGB_OK;
}

static int cgb_sampler_update_body(cgb_sampler *s, uint32_t nSteps, uint32_t nThreads)
{
    (void)nThreads;
    CGB_CHECK(s && s->other, "cgb_sampler_update: sync() has not been called");
    CGB_CUDA(cudaSetDevice(s->device));
    if (s->updateMode == CGB_UPDATE_SWEEP) { return sweepUpdate(s, nSteps); }
    if (s->sequential)
    {
        const double t0 = nowSeconds();
        const double waitBefore = s->counters.secondsDeviceWait;
        const int rcSeq = sequentialUpdate(s, nSteps);
        const int rcStop = stopPersistent(s);
        s->counters.secondsHostGenerate += (nowSeconds() - t0) - (s->counters.secondsDeviceWait - waitBefore);
        s->counters.nProposalsTotal += nSteps;
        return rcSeq != CGB_OK ? rcSeq : rcStop;
    }
    if (g_hostProfile < 0) { g_hostProfile = envInt("COGAPS_HOST_PROFILE", 0); }
    uint32_t n = 0;
    while (n < nSteps)
    {
        const double t0 = nowSeconds();
        const double waitBefore = s->counters.secondsDeviceWait;
        SinkCtx sink;
        sink.s = s;
        sink.rc = CGB_OK;
        int rcBegin = CGB_OK;
        if (s->usePersistent) { rcBegin = beginChunk(s, 0); }
        if (rcBegin != CGB_OK && rcBegin != kResidentUnavailable) { return rcBegin; }
        if (s->usePersistent)
        {
            g_prof = g_hostProfile > 0 && ((g_profCounter++ & 15ull) == 0ull);
            if (g_prof) { ++g_profBatches; }
            const unsigned long long tp = g_prof ? tsc() : 0ull;
            s->queue.populate(s->domain, nSteps - n, proposalSink, &sink);
            if (g_prof) { g_tscPopulate += tsc() - tp; }
        }
        else
        {
            s->queue.populate(s->domain, nSteps - n);
        }
        n += s->queue.nProcessed();
        if (n < nSteps)
        {
            s->numQueueSamples += 1.f;
            s->avgQueueLength *= (s->numQueueSamples - 1.f) / s->numQueueSamples;
            s->avgQueueLength += static_cast<float>(s->queue.entries().size()) / s->numQueueSamples;
        }
        int rcEval = sink.rc;
        if (rcEval == CGB_OK && !s->queue.entries().empty()) { rcEval = evaluateQueue(s); }
        if (rcEval != CGB_OK)
        {
            stopPersistent(s);
            return rcEval;
        }
        const unsigned long long tf = g_prof ? tsc() : 0ull;
        s->queue.clear();
        s->domain.flushEraseCache();
        if (g_prof) { g_tscFlush += tsc() - tf; }
        g_prof = false;
        s->counters.secondsHostGenerate += (nowSeconds() - t0) - (s->counters.secondsDeviceWait - waitBefore);
    }
    CGB_TRY(stopPersistent(s));
    if (g_hostProfile > 0)
    {
        const double nb = static_cast<double>(g_profBatches ? g_profBatches : 1);
        std::printf("[host profile nSeg=%u] kticks per sampled batch (%llu batches, %.1f posts each): populate incl. posting %.2f, posting %.2f, "
                    "outcome wait %.2f, apply %.2f, flush+clear %.2f\n",
                    s->nSeg, g_profBatches, g_nPosts / nb, g_tscPopulate / nb * 1e-3, g_tscPost / nb * 1e-3, g_tscWaitTail / nb * 1e-3,
                    g_tscApply / nb * 1e-3, g_tscFlush / nb * 1e-3);
        g_tscPost = g_tscApply = g_tscFlush = g_tscPopulate = g_nPosts = g_profBatches = g_tscWaitTail = 0;
    }
    s->counters.nProposalsTotal += nSteps;
    if (s->queue.minAtoms() != s->queue.maxAtoms() || s->queue.maxAtoms() != s->domain.size())
    {
        return fail(CGB_EINTERNAL, "cgb_sampler_update: atom bookkeeping out of step with the domain");
    }
    return CGB_OK;
}

extern "C" int cgb_sampler_update(cgb_sampler *s, uint32_t nSteps, uint32_t nThreads)
{
    return guarded("cgb_sampler_update", [&]() { return cgb_sampler_update_body(s, nSteps, nThreads); });
}

static int cgb_sampler_alpha_parameters_body(cgb_sampler *s, uint32_t n, const int32_t *variant,
                                            const uint32_t *r1, const uint32_t *c1, const uint32_t *r2,
                                            const uint32_t *c2, const float *ch, float *s_out, float *smu_out)
{
    CGB_CHECK(s && s->other && variant && r1 && c1 && r2 && c2 && ch && s_out && smu_out, "cgb_sampler_alpha_parameters: NULL argument or no sync()");
    CGB_CUDA(cudaSetDevice(s->device));
    static thread_local EvalParams params;
    fillModelView(s, params.mv);
    std::vector<DevProposal> all(n);
    std::vector<uint32_t> bulk, rest; // single-row queries go out in one launch; two-row ones need the pairing path
    for (uint32_t j = 0; j < n; ++j)
    {
        CGB_CHECK(variant[j] >= 0 && variant[j] <= 2, "cgb_sampler_alpha_parameters: bad variant");
        CGB_CHECK(r1[j] < s->nRows && r2[j] < s->nRows && c1[j] < s->k && c2[j] < s->k, "cgb_sampler_alpha_parameters: index out of range");
        DevProposal &dp = all[j];
        std::memset(&dp, 0, sizeof(dp));
        dp.type = kProbe;
        dp.variant = static_cast<uint32_t>(variant[j]);
        dp.r1 = r1[j]; dp.c1 = c1[j];
        dp.r2 = (variant[j] == 1) ? r2[j] : r1[j];
        dp.c2 = (variant[j] == 1) ? c2[j] : c1[j];
        dp.ch = ch[j];
        const bool twoRow = (variant[j] == 1) && (dp.r1 != dp.r2);
        ((s->sparse || twoRow || n <= static_cast<uint32_t>(kMaxBatch)) ? rest : bulk).push_back(j);
    }
    if (!bulk.empty())
    {
        const uint32_t m = static_cast<uint32_t>(bulk.size());
        std::vector<DevProposal> props(m);
        for (uint32_t i = 0; i < m; ++i) { props[i] = all[bulk[i]]; }
        std::vector<DevOutcome> outs(m);
        DevProposal *dProps = nullptr;
        DevOutcome *dOuts = nullptr;
        CGB_CUDA(cudaMalloc(&dProps, sizeof(DevProposal) * m));
        CGB_CUDA(cudaMalloc(&dOuts, sizeof(DevOutcome) * m));
        CGB_CUDA(cudaMemcpyAsync(dProps, props.data(), sizeof(DevProposal) * m, cudaMemcpyHostToDevice, s->stream));
        for (uint32_t first = 0; first < m; first += 65535u)
        {
            ProbeParams pp;
            pp.mv = params.mv;
            pp.props = dProps + first;
            pp.outs = dOuts + first;
            cudaLaunchConfig_t cfg = cudaLaunchConfig_t();
            cfg.gridDim = dim3(s->nSeg, std::min<uint32_t>(65535u, m - first), 1);
            cfg.blockDim = dim3(kThreads, 1, 1);
            cfg.dynamicSmemBytes = evalSmemBytes(s);
            cfg.stream = s->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = s->nSeg;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            if (s->timeKernels) { CGB_CUDA(cudaEventRecord(s->evStart, s->stream)); }
            if (s->hasS) { CGB_CUDA(cudaLaunchKernelEx(&cfg, probe_kernel<true>, pp)); }
            else { CGB_CUDA(cudaLaunchKernelEx(&cfg, probe_kernel<false>, pp)); }
            ++g_kernelLaunches;
            if (s->timeKernels)
            {
                CGB_CUDA(cudaEventRecord(s->evStop, s->stream));
                CGB_CUDA(cudaStreamSynchronize(s->stream));
                float ms = 0.f;
                CGB_CUDA(cudaEventElapsedTime(&ms, s->evStart, s->evStop));
                s->counters.secondsKernel += static_cast<double>(ms) * 1e-3;
            }
            s->counters.nBatches += 1;
        }
        CGB_CUDA(cudaMemcpyAsync(outs.data(), dOuts, sizeof(DevOutcome) * m, cudaMemcpyDeviceToHost, s->stream));
        CGB_CUDA(cudaStreamSynchronize(s->stream));
        cudaFree(dProps);
        cudaFree(dOuts);
        for (uint32_t i = 0; i < m; ++i)
        {
            s_out[bulk[i]] = outs[i].s;
            smu_out[bulk[i]] = outs[i].s_mu;
        }
    }
    size_t done = 0;
    while (done < rest.size())
    {
        const uint32_t m = static_cast<uint32_t>(std::min<size_t>(kMaxBatch, rest.size() - done));
        for (uint32_t i = 0; i < m; ++i) { params.props[i] = all[rest[done + i]]; }
        params.nProps = m;
        CGB_TRY(launchEval(s, params));
        for (uint32_t i = 0; i < m; ++i)
        {
            s_out[rest[done + i]] = s->hOutcomes[i].s;
            smu_out[rest[done + i]] = s->hOutcomes[i].s_mu;
        }
        done += m;
    }
    return CGB_OK;
}

extern "C" int cgb_sampler_alpha_parameters(cgb_sampler *s, uint32_t n, const int32_t *variant,
                                            const uint32_t *r1, const uint32_t *c1, const uint32_t *r2,
                                            const uint32_t *c2, const float *ch, float *s_out, float *smu_out)
{
    return guarded("cgb_sampler_alpha_parameters", [&]() { return cgb_sampler_alpha_parameters_body(s, n, variant, r1, c1, r2, c2, ch, s_out, smu_out); });
}

// ------------------------------------------------------------------------------------------------
// reductions and accessors
// ------------------------------------------------------------------------------------------------
static int sumPartials(cgb_sampler *s, int nBlocks, double *out)
{
    CGB_CUDA(cudaMemcpyAsync(s->hReducePartials, s->dReducePartials, sizeof(double) * nBlocks, cudaMemcpyDeviceToHost, s->stream));
    CGB_CUDA(cudaStreamSynchronize(s->stream));
    double t = 0.0;
    for (int i = 0; i < nBlocks; ++i) { t += s->hReducePartials[i]; }
    *out = t;
    return CGB_OK;
}

static int cgb_sampler_chisq_body(const cgb_sampler *cs, float *out)
{
    CGB_CHECK(cs && out, "cgb_sampler_chisq: NULL argument");
    cgb_sampler *s = const_cast<cgb_sampler*>(cs);
    CGB_CUDA(cudaSetDevice(s->device));
    const int blocks = static_cast<int>(std::min<uint32_t>(kReduceBlocks, s->nRows));
    if (s->sparse)
    {
        CGB_CHECK(s->other != nullptr, "cgb_sampler_chisq: sync() has not been called");
        CGB_CUDA(cudaStreamSynchronize(s->other->stream));
        sparse_chisq_kernel<<<blocks, kSparseThreads, s->ldR * sizeof(float), s->stream>>>(s->dD, s->ld, s->dMrows, s->other->dMrows,
            s->ldR, s->nRows, s->L, s->k, s->dReducePartials);
    }
    else
    {
        chisq_kernel<<<blocks, 256, 0, s->stream>>>(s->dD, s->hasS ? s->dS : nullptr, s->dAP, s->nRows, s->L, s->ld, s->dReducePartials);
    }
    ++g_kernelLaunches;
    CGB_CUDA(cudaGetLastError());
    double t = 0.0;
    CGB_TRY(sumPartials(s, blocks, &t));
    *out = s->sparse ? static_cast<float>(t) * 100.f : static_cast<float>(t);
    return CGB_OK;
}

extern "C" int cgb_sampler_chisq(const cgb_sampler *cs, float *out)
{
    return guarded("cgb_sampler_chisq", [&]() { return cgb_sampler_chisq_body(cs, out); });
}

static int cgb_sampler_n_atoms_body(const cgb_sampler *s, uint64_t *out)
{
    CGB_CHECK(s && out, "cgb_sampler_n_atoms: NULL argument");
    *out = (s->updateMode == CGB_UPDATE_SWEEP) ? s->swTotalAtoms : s->domain.size();
    return CGB_OK;
}

extern "C" int cgb_sampler_n_atoms(const cgb_sampler *s, uint64_t *out)
{
    return guarded("cgb_sampler_n_atoms", [&]() { return cgb_sampler_n_atoms_body(s, out); });
}

static int cgb_sampler_data_sparsity_body(const cgb_sampler *s, float *out)
{
    CGB_CHECK(s && out, "cgb_sampler_data_sparsity: NULL argument");
    *out = s->dataSparsity;
    return CGB_OK;
}

extern "C" int cgb_sampler_data_sparsity(const cgb_sampler *s, float *out)
{
    return guarded("cgb_sampler_data_sparsity", [&]() { return cgb_sampler_data_sparsity_body(s, out); });
}

static int cgb_sampler_average_queue_length_body(const cgb_sampler *s, float *out)
{
    CGB_CHECK(s && out, "cgb_sampler_average_queue_length: NULL argument");
    *out = s->avgQueueLength;
    return CGB_OK;
}

extern "C" int cgb_sampler_average_queue_length(const cgb_sampler *s, float *out)
{
    return guarded("cgb_sampler_average_queue_length", [&]() { return cgb_sampler_average_queue_length_body(s, out); });
}

static int cgb_sampler_get_matrix_body(const cgb_sampler *s, float *out)
{
    CGB_CHECK(s && out, "cgb_sampler_get_matrix: NULL argument");
    CGB_CUDA(cudaSetDevice(s->device));
    CGB_CUDA(cudaStreamSynchronize(s->stream));
    if (s->sparse)
    {
        // HybridMatrix -> Matrix conversions read the row copy (data_structures/Matrix.cpp)
        std::vector<float> rows(static_cast<size_t>(s->nRows) * s->ldR);
        CGB_CUDA(cudaMemcpy(rows.data(), s->dMrows, rows.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (uint32_t r = 0; r < s->nRows; ++r)
        {
            for (uint32_t c = 0; c < s->k; ++c) { out[static_cast<size_t>(r) * s->k + c] = rows[static_cast<size_t>(r) * s->ldR + c]; }
        }
        return CGB_OK;
    }
    std::vector<float> host(static_cast<size_t>(s->k) * s->ldM);
    CGB_CUDA(cudaMemcpy(host.data(), s->dM, host.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (uint32_t r = 0; r < s->nRows; ++r)
    {
        for (uint32_t c = 0; c < s->k; ++c) { out[static_cast<size_t>(r) * s->k + c] = host[static_cast<size_t>(c) * s->ldM + r]; }
    }
    return CGB_OK;
}

extern "C" int cgb_sampler_get_matrix(const cgb_sampler *s, float *out)
{
    return guarded("cgb_sampler_get_matrix", [&]() { return cgb_sampler_get_matrix_body(s, out); });
}

static int cgb_sampler_shape_body(const cgb_sampler *s, uint32_t *rows, uint32_t *nPatterns, uint32_t *rowLength)
{
    CGB_CHECK(s != nullptr, "cgb_sampler_shape: NULL sampler");
    if (rows) { *rows = s->nRows; }
    if (nPatterns) { *nPatterns = s->k; }
    if (rowLength) { *rowLength = s->L; }
    return CGB_OK;
}

extern "C" int cgb_sampler_shape(const cgb_sampler *s, uint32_t *rows, uint32_t *nPatterns, uint32_t *rowLength)
{
    return guarded("cgb_sampler_shape", [&]() { return cgb_sampler_shape_body(s, rows, nPatterns, rowLength); });
}

static int cgb_sampler_lambda_body(const cgb_sampler *s, float *lambda, float *maxGibbsMass)
{
    CGB_CHECK(s != nullptr, "cgb_sampler_lambda: NULL sampler");
    if (lambda) { *lambda = s->lambda; }
    if (maxGibbsMass) { *maxGibbsMass = s->maxGibbsMass; }
    return CGB_OK;
}

extern "C" int cgb_sampler_lambda(const cgb_sampler *s, float *lambda, float *maxGibbsMass)
{
    return guarded("cgb_sampler_lambda", [&]() { return cgb_sampler_lambda_body(s, lambda, maxGibbsMass); });
}

static int cgb_sampler_get_atoms_body(const cgb_sampler *s, uint64_t *pos, float *mass, uint64_t capacity, uint64_t *count)
{
    CGB_CHECK(s && count, "cgb_sampler_get_atoms: NULL argument");
    if (s->updateMode == CGB_UPDATE_SWEEP)
    {
        // position order (the sweep has no pick vector)
        std::vector<uint64_t> p;
        std::vector<float> m;
        CGB_CUDA(cudaSetDevice(s->device));
        CGB_TRY(sweepDownloadAtoms(s, p, m));
        *count = p.size();
        if (pos && mass)
        {
            const uint64_t n = std::min<uint64_t>(capacity, p.size());
            std::copy(p.begin(), p.begin() + n, pos);
            std::copy(m.begin(), m.begin() + n, mass);
        }
        return CGB_OK;
    }
    *count = s->domain.size();
    if (pos && mass)
    {
        const uint64_t n = std::min<uint64_t>(capacity, s->domain.size());
        for (uint64_t i = 0; i < n; ++i)
        {
            const Atom &a = s->domain.atom(s->domain.atIndex(static_cast<uint32_t>(i)));
            pos[i] = a.pos;
            mass[i] = a.mass;
        }
    }
    return CGB_OK;
}

extern "C" int cgb_sampler_get_atoms(const cgb_sampler *s, uint64_t *pos, float *mass, uint64_t capacity, uint64_t *count)
{
    return guarded("cgb_sampler_get_atoms", [&]() { return cgb_sampler_get_atoms_body(s, pos, mass, capacity, count); });
}

static int cgb_sampler_get_ap_row_body(const cgb_sampler *s, uint32_t row, float *out)
{
    CGB_CHECK(s && out && row < s->nRows, "cgb_sampler_get_ap_row: bad argument");
    if (s->sparse) { return fail(CGB_EUNSUPPORTED, "cgb_sampler_get_ap_row: the sparse model keeps no AP matrix"); }
    CGB_CUDA(cudaSetDevice(s->device));
    CGB_CUDA(cudaStreamSynchronize(s->stream));
    CGB_CUDA(cudaMemcpy(out, s->dAP + static_cast<size_t>(row) * s->ld, sizeof(float) * s->L, cudaMemcpyDeviceToHost));
    return CGB_OK;
}

extern "C" int cgb_sampler_get_ap_row(const cgb_sampler *s, uint32_t row, float *out)
{
    return guarded("cgb_sampler_get_ap_row", [&]() { return cgb_sampler_get_ap_row_body(s, row, out); });
}

static int cgb_sampler_get_counters_body(const cgb_sampler *s, cgb_sampler_counters *out)
{
    CGB_CHECK(s && out, "cgb_sampler_get_counters: NULL argument");
    *out = s->counters;
    return CGB_OK;
}

extern "C" int cgb_sampler_get_counters(const cgb_sampler *s, cgb_sampler_counters *out)
{
    return guarded("cgb_sampler_get_counters", [&]() { return cgb_sampler_get_counters_body(s, out); });
}

static int cgb_sampler_reset_counters_body(cgb_sampler *s)
{
    CGB_CHECK(s != nullptr, "cgb_sampler_reset_counters: NULL sampler");
    std::memset(static_cast<void*>(&s->counters), 0, sizeof(s->counters));
    return CGB_OK;
}

extern "C" int cgb_sampler_reset_counters(cgb_sampler *s)
{
    return guarded("cgb_sampler_reset_counters", [&]() { return cgb_sampler_reset_counters_body(s); });
}

static int cgb_sampler_set_kernel_timing_body(cgb_sampler *s, int32_t enabled)
{
    CGB_CHECK(s != nullptr, "cgb_sampler_set_kernel_timing: NULL sampler");
    s->timeKernels = enabled != 0;
    return CGB_OK;
}

extern "C" int cgb_sampler_set_kernel_timing(cgb_sampler *s, int32_t enabled)
{
    return guarded("cgb_sampler_set_kernel_timing", [&]() { return cgb_sampler_set_kernel_timing_body(s, enabled); });
}

static int cgb_sampler_set_persistent_body(cgb_sampler *s, int32_t enabled)
{
    CGB_CHECK(s != nullptr, "cgb_sampler_set_persistent: NULL sampler");
    CGB_CHECK(!s->persistentRunning, "cgb_sampler_set_persistent: called in the middle of an update");
    s->usePersistent = enabled != 0;
    return CGB_OK;
}

extern "C" int cgb_sampler_set_persistent(cgb_sampler *s, int32_t enabled)
{
    return guarded("cgb_sampler_set_persistent", [&]() { return cgb_sampler_set_persistent_body(s, enabled); });
}

static int cgb_sampler_set_resident_share_body(cgb_sampler *s, int32_t parts)
{
    CGB_CHECK(s != nullptr, "cgb_sampler_set_resident_share: NULL sampler");
    CGB_CHECK(parts >= 1 && parts <= 64, "cgb_sampler_set_resident_share: parts must be in 1..64");
    CGB_CHECK(s->persistentGrid == 0, "cgb_sampler_set_resident_share: the sampler's resident grid has already been sized (call before the first update)");
    s->residentShare = parts;
    return CGB_OK;
}

extern "C" int cgb_sampler_set_resident_share(cgb_sampler *s, int32_t parts)
{
    return guarded("cgb_sampler_set_resident_share", [&]() { return cgb_sampler_set_resident_share_body(s, parts); });
}

static int cgb_sampler_reduction_order_body(const cgb_sampler *s, cgb_reduction_order *out)
{
    CGB_CHECK(s && out, "cgb_sampler_reduction_order: NULL argument");
    out->threadsPerSegment = kThreads;
    out->vectorWidth = kVec;
    out->nSegments = s->nSeg;
    out->segmentLength = s->seg;
    return CGB_OK;
}

extern "C" int cgb_sampler_reduction_order(const cgb_sampler *s, cgb_reduction_order *out)
{
    return guarded("cgb_sampler_reduction_order", [&]() { return cgb_sampler_reduction_order_body(s, out); });
}

static int cgb_sampler_device_matrix_body(const cgb_sampler *s, void **dev, uint64_t *ld)
{
    CGB_CHECK(s && dev && ld, "cgb_sampler_device_matrix: NULL argument");
    *dev = s->dM;
    *ld = s->ldM;
    return CGB_OK;
}

extern "C" int cgb_sampler_device_matrix(const cgb_sampler *s, void **dev, uint64_t *ld)
{
    return guarded("cgb_sampler_device_matrix", [&]() { return cgb_sampler_device_matrix_body(s, dev, ld); });
}

// ------------------------------------------------------------------------------------------------
// GapsStatistics on the device
// ------------------------------------------------------------------------------------------------
struct cgb_stats
{
    uint32_t nGenes, nSamples, k, ldA, ldP;
    float *dAmean, *dAsq, *dPmean, *dPsq, *dPump, *dNorms, *dScratch;
    unsigned statUpdates, pumpUpdates;
    int device;
};

extern "C" void cgb_stats_destroy(cgb_stats *st)
{
    if (!st) { return; }
    cudaSetDevice(st->device);
    cudaFree(st->dAmean); cudaFree(st->dAsq); cudaFree(st->dPmean); cudaFree(st->dPsq);
    cudaFree(st->dPump); cudaFree(st->dNorms); cudaFree(st->dScratch);
    delete st;
}

static int cgb_stats_create_body(uint32_t nGenes, uint32_t nSamples, uint32_t nPatterns, cgb_stats **out)
{
    CGB_CHECK(out && nGenes && nSamples && nPatterns, "cgb_stats_create: bad argument");
    CGB_TRY(ensureDevice());
    cgb_stats *st = new (std::nothrow) cgb_stats();
    if (!st) { return fail(CGB_ENOMEM, "cgb_stats_create: out of memory"); }
    std::memset(static_cast<void*>(st), 0, sizeof(*st));
    st->nGenes = nGenes; st->nSamples = nSamples; st->k = nPatterns;
    st->ldA = roundUp(nGenes, 32);
    st->ldP = roundUp(nSamples, 32);
    st->device = g_device;
    const size_t aBytes = static_cast<size_t>(nPatterns) * st->ldA * sizeof(float);
    const size_t pBytes = static_cast<size_t>(nPatterns) * st->ldP * sizeof(float);
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) { e = cudaMalloc(&st->dAmean, aBytes); }
    if (e == cudaSuccess) { e = cudaMalloc(&st->dAsq, aBytes); }
    if (e == cudaSuccess) { e = cudaMalloc(&st->dPump, aBytes); }
    if (e == cudaSuccess) { e = cudaMalloc(&st->dScratch, aBytes); }
    if (e == cudaSuccess) { e = cudaMalloc(&st->dPmean, pBytes); }
    if (e == cudaSuccess) { e = cudaMalloc(&st->dPsq, pBytes); }
    if (e == cudaSuccess) { e = cudaMalloc(&st->dNorms, sizeof(float) * nPatterns); }
    if (e == cudaSuccess) { e = cudaMemset(st->dAmean, 0, aBytes); }
    if (e == cudaSuccess) { e = cudaMemset(st->dAsq, 0, aBytes); }
    if (e == cudaSuccess) { e = cudaMemset(st->dPump, 0, aBytes); }
    if (e == cudaSuccess) { e = cudaMemset(st->dPmean, 0, pBytes); }
    if (e == cudaSuccess) { e = cudaMemset(st->dPsq, 0, pBytes); }
    if (e == cudaSuccess) { e = cudaStreamSynchronize(cudaStreamLegacy); } // the kernels run on the samplers' non-blocking streams
    if (e != cudaSuccess)
    {
        cgb_stats_destroy(st);
        return fail(e == cudaErrorMemoryAllocation ? CGB_ENOMEM : CGB_ECUDA, std::string("cgb_stats_create: ") + cudaGetErrorString(e));
    }
    *out = st;
    return CGB_OK;
}

extern "C" int cgb_stats_create(uint32_t nGenes, uint32_t nSamples, uint32_t nPatterns, cgb_stats **out)
{
    return guarded("cgb_stats_create", [&]() { return cgb_stats_create_body(nGenes, nSamples, nPatterns, out); });
}

// mode 0: update (both, P normalised by its column max); 1: updateA; 2: updateP (norm forced to 1)
static int statsUpdate(cgb_stats *st, const cgb_sampler *A, const cgb_sampler *P, int mode)
{
    CGB_CHECK(st && A && P, "cgb_stats_update: NULL argument");
    CGB_CHECK(A->nRows == st->nGenes && P->nRows == st->nSamples && A->k == st->k && P->k == st->k, "cgb_stats_update: shape mismatch");
    CGB_CUDA(cudaSetDevice(st->device));
    cudaStream_t stream = P->stream;
    CGB_CUDA(cudaStreamSynchronize(A->stream));
    ++st->statUpdates;
    col_max_kernel<<<st->k, 256, 0, stream>>>(P->dM, P->nRows, P->ldM, st->dNorms, mode != 0 ? 1 : 0);
    ++g_kernelLaunches;
    if (mode == 0 || mode == 2)
    {
        dim3 grid((P->nRows + 255) / 256, st->k);
        stats_accumulate_kernel<<<grid, 256, 0, stream>>>(P->dM, P->nRows, P->ldM, st->dNorms, 1, st->dPmean, st->dPsq);
        ++g_kernelLaunches;
    }
    if (mode == 0 || mode == 1)
    {
        dim3 grid((A->nRows + 255) / 256, st->k);
        stats_accumulate_kernel<<<grid, 256, 0, stream>>>(A->dM, A->nRows, A->ldM, st->dNorms, 0, st->dAmean, st->dAsq);
        ++g_kernelLaunches;
    }
    CGB_CUDA(cudaGetLastError());
    CGB_CUDA(cudaStreamSynchronize(stream));
    return CGB_OK;
}

extern "C" int cgb_stats_update(cgb_stats *st, const cgb_sampler *A, const cgb_sampler *P) { return guarded("cgb_stats_update", [&]() { return statsUpdate(st, A, P, 0); }); }
extern "C" int cgb_stats_update_a(cgb_stats *st, const cgb_sampler *A, const cgb_sampler *P) { return guarded("cgb_stats_update_a", [&]() { return statsUpdate(st, A, P, 1); }); }
extern "C" int cgb_stats_update_p(cgb_stats *st, const cgb_sampler *A, const cgb_sampler *P) { return guarded("cgb_stats_update_p", [&]() { return statsUpdate(st, A, P, 2); }); }

static int cgb_stats_update_pump_body(cgb_stats *st, const cgb_sampler *A)
{
    CGB_CHECK(st && A && A->nRows == st->nGenes, "cgb_stats_update_pump: bad argument");
    CGB_CUDA(cudaSetDevice(st->device));
    ++st->pumpUpdates;
    pump_kernel<<<(A->nRows + 255) / 256, 256, 0, A->stream>>>(A->dM, A->nRows, A->ldM, st->k, 1.f, st->dPump);
    ++g_kernelLaunches;
    CGB_CUDA(cudaGetLastError());
    CGB_CUDA(cudaStreamSynchronize(A->stream));
    return CGB_OK;
}

extern "C" int cgb_stats_update_pump(cgb_stats *st, const cgb_sampler *A)
{
    return guarded("cgb_stats_update_pump", [&]() { return cgb_stats_update_pump_body(st, A); });
}

static int downloadFactor(const float *dev, uint32_t rows, uint32_t k, uint32_t ld, std::vector<float> &host)
{
    host.resize(static_cast<size_t>(k) * ld);
    CGB_CUDA(cudaMemcpy(host.data(), dev, host.size() * sizeof(float), cudaMemcpyDeviceToHost));
    (void)rows;
    return CGB_OK;
}

// Amean / Pmean: sums / nUpdates (GapsStatistics.cpp:13-21,38-46)
static int statsMean(const cgb_stats *st, const float *dev, uint32_t rows, uint32_t ld, float div, float *out)
{
    CGB_CHECK(st && out, "cgb_stats: NULL argument");
    CGB_CUDA(cudaSetDevice(st->device));
    std::vector<float> host;
    CGB_TRY(downloadFactor(dev, rows, st->k, ld, host));
    for (uint32_t i = 0; i < rows; ++i)
    {
        for (uint32_t c = 0; c < st->k; ++c) { out[static_cast<size_t>(i) * st->k + c] = host[static_cast<size_t>(c) * ld + i] / div; }
    }
    return CGB_OK;
}

// Asd / Psd (GapsStatistics.cpp:23-36,48-61)
static int statsSd(const cgb_stats *st, const float *devMean, const float *devSq, uint32_t rows, uint32_t ld, float *out)
{
    CGB_CHECK(st && out, "cgb_stats: NULL argument");
    CGB_CUDA(cudaSetDevice(st->device));
    std::vector<float> mean, sq;
    CGB_TRY(downloadFactor(devMean, rows, st->k, ld, mean));
    CGB_TRY(downloadFactor(devSq, rows, st->k, ld, sq));
    const float n = static_cast<float>(st->statUpdates);
    for (uint32_t i = 0; i < rows; ++i)
    {
        for (uint32_t c = 0; c < st->k; ++c)
        {
            const size_t idx = static_cast<size_t>(c) * ld + i;
            const float meanTerm = (mean[idx] * mean[idx]) / n;
            const float numer = gmax(0.f, sq[idx] - meanTerm);
            out[static_cast<size_t>(i) * st->k + c] = std::sqrt(numer / (n - 1.f));
        }
    }
    return CGB_OK;
}

extern "C" int cgb_stats_amean(const cgb_stats *st, float *out) { return guarded("cgb_stats_amean", [&]() { return statsMean(st, st ? st->dAmean : nullptr, st ? st->nGenes : 0, st ? st->ldA : 0, st ? static_cast<float>(st->statUpdates) : 1.f, out); }); }
extern "C" int cgb_stats_pmean(const cgb_stats *st, float *out) { return guarded("cgb_stats_pmean", [&]() { return statsMean(st, st ? st->dPmean : nullptr, st ? st->nSamples : 0, st ? st->ldP : 0, st ? static_cast<float>(st->statUpdates) : 1.f, out); }); }
extern "C" int cgb_stats_asd(const cgb_stats *st, float *out) { return guarded("cgb_stats_asd", [&]() { return statsSd(st, st ? st->dAmean : nullptr, st ? st->dAsq : nullptr, st ? st->nGenes : 0, st ? st->ldA : 0, out); }); }
extern "C" int cgb_stats_psd(const cgb_stats *st, float *out) { return guarded("cgb_stats_psd", [&]() { return statsSd(st, st ? st->dPmean : nullptr, st ? st->dPsq : nullptr, st ? st->nSamples : 0, st ? st->ldP : 0, out); }); }

static int cgb_stats_pump_matrix_body(const cgb_stats *st, float *out)
{
    const float denom = (st && st->pumpUpdates != 0) ? static_cast<float>(st->pumpUpdates) : 1.f;
    return statsMean(st, st ? st->dPump : nullptr, st ? st->nGenes : 0, st ? st->ldA : 0, denom, out);
}

extern "C" int cgb_stats_pump_matrix(const cgb_stats *st, float *out)
{
    return guarded("cgb_stats_pump_matrix", [&]() { return cgb_stats_pump_matrix_body(st, out); });
}

static int cgb_stats_mean_pattern_body(const cgb_stats *cst, float *out)
{
    CGB_CHECK(cst && out, "cgb_stats_mean_pattern: NULL argument");
    cgb_stats *st = const_cast<cgb_stats*>(cst);
    CGB_CUDA(cudaSetDevice(st->device));
    const size_t aBytes = static_cast<size_t>(st->k) * st->ldA * sizeof(float);
    CGB_CUDA(cudaMemset(st->dScratch, 0, aBytes));
    pump_kernel<<<(st->nGenes + 255) / 256, 256>>>(st->dAmean, st->nGenes, st->ldA, st->k, static_cast<float>(st->statUpdates), st->dScratch);
    ++g_kernelLaunches;
    CGB_CUDA(cudaGetLastError());
    CGB_CUDA(cudaDeviceSynchronize());
    return statsMean(st, st->dScratch, st->nGenes, st->ldA, 1.f, out);
}

extern "C" int cgb_stats_mean_pattern(const cgb_stats *cst, float *out)
{
    return guarded("cgb_stats_mean_pattern", [&]() { return cgb_stats_mean_pattern_body(cst, out); });
}

static int cgb_stats_mean_chisq_body(const cgb_stats *st, const cgb_sampler *cP, float *out)
{
    CGB_CHECK(st && cP && out, "cgb_stats_mean_chisq: NULL argument");
    cgb_sampler *P = const_cast<cgb_sampler*>(cP);
    CGB_CHECK(P->nRows == st->nSamples && P->L == st->nGenes, "cgb_stats_mean_chisq: expects the P sampler");
    CGB_CUDA(cudaSetDevice(st->device));
    const float n = static_cast<float>(st->statUpdates);
    const int blocks = static_cast<int>(std::min<uint32_t>(kReduceBlocks, P->nRows));
    mean_chisq_kernel<<<blocks, 256, 0, P->stream>>>(P->dD, P->hasS ? P->dS : nullptr, P->nRows, P->L, P->ld,
        st->dAmean, st->ldA, st->dPmean, st->ldP, st->k, n * n, P->dReducePartials);
    ++g_kernelLaunches;
    CGB_CUDA(cudaGetLastError());
    double t = 0.0;
    CGB_TRY(sumPartials(P, blocks, &t));
    *out = static_cast<float>(t);
    return CGB_OK;
}

extern "C" int cgb_stats_mean_chisq(const cgb_stats *st, const cgb_sampler *cP, float *out)
{
    return guarded("cgb_stats_mean_chisq", [&]() { return cgb_stats_mean_chisq_body(st, cP, out); });
}

static int cgb_stats_device_sums_body(const cgb_stats *st, void **AmeanSum, void **AsqSum, void **PmeanSum,
                                     void **PsqSum, uint64_t *ldA, uint64_t *ldP, uint32_t *nUpdates)
{
    CGB_CHECK(st != nullptr, "cgb_stats_device_sums: NULL stats");
    if (AmeanSum) { *AmeanSum = st->dAmean; }
    if (AsqSum) { *AsqSum = st->dAsq; }
    if (PmeanSum) { *PmeanSum = st->dPmean; }
    if (PsqSum) { *PsqSum = st->dPsq; }
    if (ldA) { *ldA = st->ldA; }
    if (ldP) { *ldP = st->ldP; }
    if (nUpdates) { *nUpdates = st->statUpdates; }
    return CGB_OK;
}

extern "C" int cgb_stats_device_sums(const cgb_stats *st, void **AmeanSum, void **AsqSum, void **PmeanSum,
                                     void **PsqSum, uint64_t *ldA, uint64_t *ldP, uint32_t *nUpdates)
{
    return guarded("cgb_stats_device_sums", [&]() { return cgb_stats_device_sums_body(st, AmeanSum, AsqSum, PmeanSum, PsqSum, ldA, ldP, nUpdates); });
}

// device portable-log probe (tests)
static int cgb_debug_logf_body(const float *in, float *out, uint32_t n)
{
    CGB_CHECK(in && out, "cgb_debug_logf: NULL argument");
    CGB_TRY(ensureDevice());
    float *dIn = nullptr, *dOut = nullptr;
    CGB_CUDA(cudaMalloc(&dIn, sizeof(float) * n));
    CGB_CUDA(cudaMalloc(&dOut, sizeof(float) * n));
    CGB_CUDA(cudaMemcpy(dIn, in, sizeof(float) * n, cudaMemcpyHostToDevice));
    logf_probe_kernel<<<(n + 255) / 256, 256>>>(dIn, dOut, n);
    ++g_kernelLaunches;
    CGB_CUDA(cudaGetLastError());
    CGB_CUDA(cudaMemcpy(out, dOut, sizeof(float) * n, cudaMemcpyDeviceToHost));
    cudaFree(dIn);
    cudaFree(dOut);
    return CGB_OK;
}

extern "C" int cgb_debug_logf(const float *in, float *out, uint32_t n)
{
    return guarded("cgb_debug_logf", [&]() { return cgb_debug_logf_body(in, out, n); });
}

// host-logic probe: the generator's multiply-high division (atomic_domain.h FastDivU64) against the hardware divide
extern "C" uint64_t cgb_debug_fastdiv(uint64_t divisor, uint64_t x)
{
    FastDivU64 f;
    f.init(divisor);
    return f.div(x);
}

// host portable-log probe: the same header compiled for the host (tests compare both with the oracle)
extern "C" float cgb_debug_host_logf(float x) { return portable_logf(x); }

// the running sum / positive count behind lambda exactly as the samplers take it: over the rows of a row-major
// nrow x ncol matrix (byColumns == 0) or down its columns (the blocked walk)
extern "C" int cgb_debug_running_sum(const float *data, uint32_t nrow, uint32_t ncol, int32_t byColumns, float *sum, uint32_t *nnz)
{
    if (!data || !sum || !nnz) { return fail(CGB_EINVAL, "cgb_debug_running_sum: NULL argument"); }
    return guarded("cgb_debug_running_sum", [&]()
    {
        unsigned n = 0;
        if (byColumns) { runningSum(data, ncol, nrow, 1, ncol, *sum, n); }
        else { runningSum(data, nrow, ncol, ncol, 1, *sum, n); }
        *nnz = n;
        return CGB_OK;
    });
}

// ------------------------------------------------------------------------------------------------
// Checkpoints: device-resident state <-> the images of checkpoint.h (Archive << / >> of the Sampler concept,
// AsynchronousGibbsSampler.h:221-233; GapsStatistics.cpp:164-176)
// ------------------------------------------------------------------------------------------------
static const char *kSequentialCheckpointMsg =
    "checkpoints need the asynchronous sampler: the reference's SingleThreadedGibbsSampler does not archive its rng "
    "and cannot read its own archive (SingleThreadedGibbsSampler.h:260-273)";

// [k][ld] on the device -> [k][rows] on the host
static int downloadPatternMajor(const float *dev, uint32_t rows, uint32_t k, uint32_t ld, std::vector<float> &out)
{
    std::vector<float> host(static_cast<size_t>(k) * ld);
    CGB_CUDA(cudaMemcpy(host.data(), dev, host.size() * sizeof(float), cudaMemcpyDeviceToHost));
    out.resize(static_cast<size_t>(k) * rows);
    for (uint32_t c = 0; c < k; ++c) { std::memcpy(out.data() + static_cast<size_t>(c) * rows, host.data() + static_cast<size_t>(c) * ld, sizeof(float) * rows); }
    return CGB_OK;
}

// [k][rows] on the host -> [k][ld] on the device, padding zeroed
static int uploadPatternMajor(float *dev, uint32_t rows, uint32_t k, uint32_t ld, const std::vector<float> &in)
{
    std::vector<float> host(static_cast<size_t>(k) * ld, 0.f);
    for (uint32_t c = 0; c < k; ++c) { std::memcpy(host.data() + static_cast<size_t>(c) * ld, in.data() + static_cast<size_t>(c) * rows, sizeof(float) * rows); }
    CGB_CUDA(cudaMemcpy(dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    return CGB_OK;
}

static int samplerToImage(const cgb_sampler *s, SamplerImage &img)
{
    if (s->sequential) { return fail(CGB_EUNSUPPORTED, kSequentialCheckpointMsg); }
    CGB_CHECK(!s->persistentRunning, "checkpoint: called in the middle of an update");
    CGB_CUDA(cudaSetDevice(s->device));
    CGB_CUDA(cudaStreamSynchronize(s->stream));
    img.sparse = s->sparse;
    img.nRows = s->nRows;
    img.k = s->k;
    CGB_TRY(downloadPatternMajor(s->dM, s->nRows, s->k, s->ldM, img.cols));
    img.rows.clear();
    img.beta = 0.f;
    if (s->sparse)
    {
        std::vector<float> rows(static_cast<size_t>(s->nRows) * s->ldR);
        CGB_CUDA(cudaMemcpy(rows.data(), s->dMrows, rows.size() * sizeof(float), cudaMemcpyDeviceToHost));
        img.rows.resize(static_cast<size_t>(s->nRows) * s->k);
        for (uint32_t r = 0; r < s->nRows; ++r) { std::memcpy(img.rows.data() + static_cast<size_t>(r) * s->k, rows.data() + static_cast<size_t>(r) * s->ldR, sizeof(float) * s->k); }
        img.beta = 100.f; // SparseNormalModel.h:77
    }
    img.domainLength = s->domain.domainLength();
    s->queue.save(img.queue);
    if (s->updateMode == CGB_UPDATE_SWEEP)
    {
        // the atoms live in the device's per-row store; archived in position order, the window collapsed onto the count
        CGB_TRY(sweepDownloadAtoms(s, img.pos, img.mass));
        img.queue.minAtoms = img.queue.maxAtoms = img.pos.size();
        return CGB_OK;
    }
    const size_t n = static_cast<size_t>(s->domain.size());
    img.pos.resize(n);
    img.mass.resize(n);
    for (size_t i = 0; i < n; ++i)
    {
        const Atom &a = s->domain.atom(s->domain.atIndex(static_cast<uint32_t>(i)));
        img.pos[i] = a.pos;
        img.mass[i] = a.mass;
    }
    return CGB_OK;
}

static int setAtoms(cgb_sampler *s, const uint64_t *pos, const float *mass, uint64_t n)
{
    const uint64_t nBins = static_cast<uint64_t>(s->nRows) * s->k;
    CGB_CHECK(n <= nBins, "checkpoint: more atoms than the domain can ever hold is not a state this sampler produced");
    s->domain.init(nBins);
    for (uint64_t i = 0; i < n; ++i)
    {
        if (pos[i] > s->domain.domainLength() || s->domain.occupied(pos[i]))
        {
            s->domain.init(nBins);
            return fail(CGB_EINVAL, "checkpoint: atom position outside the domain or used twice");
        }
        if (!(mass[i] >= 0.f) || !(mass[i] <= 3.0e38f))
        {
            s->domain.init(nBins);
            return fail(CGB_EINVAL, "checkpoint: atom mass negative or not finite");
        }
        s->domain.insert(pos[i], mass[i]);
    }
    // the generator's atom-count window (ProposalQueue.cpp:59-60 asserts min == max == domain.size() before it draws)
    if (!s->sequential) { s->queue.setAtomCount(n); }
    if (s->updateMode == CGB_UPDATE_SWEEP) { CGB_TRY(sweepFromDomain(s)); }
    return CGB_OK;
}

static int imageToSampler(cgb_sampler *s, const SamplerImage &img)
{
    if (s->sequential) { return fail(CGB_EUNSUPPORTED, kSequentialCheckpointMsg); }
    CGB_CHECK(!s->persistentRunning, "checkpoint: called in the middle of an update");
    CGB_CHECK(img.sparse == s->sparse, "checkpoint: made with the other data model (useSparseOptimization differs)");
    CGB_CHECK(img.nRows == s->nRows && img.k == s->k, "checkpoint: factor matrix shape differs from this sampler's");
    CGB_CHECK(img.domainLength == s->domain.domainLength(), "checkpoint: atomic domain length differs from this sampler's");
    CGB_CHECK(img.pos.size() == img.mass.size(), "checkpoint: atom arrays differ in length");
    // a checkpoint is taken between updates, where the window has collapsed (ProposalQueue.cpp:59-60); anything else would
    // let the first update pick from atoms that do not exist
    CGB_CHECK(img.queue.minAtoms == img.queue.maxAtoms && img.queue.maxAtoms == img.pos.size(),
              "checkpoint: archived atom counts of the proposal queue disagree with the archived atoms");
    QueueState probe;
    s->queue.save(probe);
    CGB_CHECK(probe.binLength == img.queue.binLength && probe.numCols == img.queue.numCols && probe.numBins == img.queue.numBins
              && probe.domainLength == img.queue.domainLength, "checkpoint: proposal queue geometry differs from this sampler's");
    CGB_CUDA(cudaSetDevice(s->device));
    CGB_CUDA(cudaStreamSynchronize(s->stream));
    CGB_TRY(uploadPatternMajor(s->dM, s->nRows, s->k, s->ldM, img.cols));
    if (s->sparse)
    {
        CGB_CHECK(img.rows.size() == static_cast<size_t>(s->nRows) * s->k, "checkpoint: row copy of the factor matrix missing");
        std::vector<float> rows(static_cast<size_t>(s->nRows) * s->ldR, 0.f);
        for (uint32_t r = 0; r < s->nRows; ++r) { std::memcpy(rows.data() + static_cast<size_t>(r) * s->ldR, img.rows.data() + static_cast<size_t>(r) * s->k, sizeof(float) * s->k); }
        CGB_CUDA(cudaMemcpy(s->dMrows, rows.data(), rows.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    CGB_CUDA(cudaStreamSynchronize(cudaStreamLegacy)); // the kernels run on non-blocking streams (see samplerCreateImpl)
    CGB_TRY(refreshColNonzero(s));
    CGB_TRY(setAtoms(s, img.pos.data(), img.mass.data(), img.pos.size()));
    if (!s->queue.restore(img.queue)) { return fail(CGB_EINTERNAL, "checkpoint: queue state rejected after it was checked"); }
    return CGB_OK;
}

static int statsToImage(const cgb_stats *st, StatsImage &img)
{
    // every statistics update ends with a synchronise of the stream it ran on (statsUpdate), nothing is in flight
    CGB_CUDA(cudaSetDevice(st->device));
    img.nGenes = st->nGenes;
    img.nSamples = st->nSamples;
    img.k = st->k;
    CGB_TRY(downloadPatternMajor(st->dAmean, st->nGenes, st->k, st->ldA, img.aMean));
    CGB_TRY(downloadPatternMajor(st->dAsq, st->nGenes, st->k, st->ldA, img.aSq));
    CGB_TRY(downloadPatternMajor(st->dPmean, st->nSamples, st->k, st->ldP, img.pMean));
    CGB_TRY(downloadPatternMajor(st->dPsq, st->nSamples, st->k, st->ldP, img.pSq));
    img.statUpdates = st->statUpdates;
    img.numPatterns = st->k;
    return CGB_OK;
}

static int imageToStats(cgb_stats *st, const StatsImage &img)
{
    CGB_CHECK(img.nGenes == st->nGenes && img.nSamples == st->nSamples && img.k == st->k && img.numPatterns == st->k,
              "checkpoint: statistics shape differs from this run's");
    CGB_CUDA(cudaSetDevice(st->device));
    CGB_TRY(uploadPatternMajor(st->dAmean, st->nGenes, st->k, st->ldA, img.aMean));
    CGB_TRY(uploadPatternMajor(st->dAsq, st->nGenes, st->k, st->ldA, img.aSq));
    CGB_TRY(uploadPatternMajor(st->dPmean, st->nSamples, st->k, st->ldP, img.pMean));
    CGB_TRY(uploadPatternMajor(st->dPsq, st->nSamples, st->k, st->ldP, img.pSq));
    CGB_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    st->statUpdates = img.statUpdates;
    return CGB_OK;
}

static int copyOut(const std::vector<uint8_t> &bytes, void *buf, uint64_t capacity, uint64_t *size, const char *who)
{
    *size = bytes.size();
    if (buf != nullptr)
    {
        if (capacity < bytes.size()) { return fail(CGB_EINVAL, std::string(who) + ": buffer too small"); }
        std::memcpy(buf, bytes.data(), bytes.size());
    }
    return CGB_OK;
}

static int cgb_sampler_serialize_body(const cgb_sampler *s, void *buf, uint64_t capacity, uint64_t *size)
{
    CGB_CHECK(s && size, "cgb_sampler_serialize: NULL argument");
    SamplerImage img;
    CGB_TRY(samplerToImage(s, img));
    ByteWriter w;
    putSampler(w, img);
    return copyOut(w.bytes(), buf, capacity, size, "cgb_sampler_serialize");
}

extern "C" int cgb_sampler_serialize(const cgb_sampler *s, void *buf, uint64_t capacity, uint64_t *size)
{
    return guarded("cgb_sampler_serialize", [&]() { return cgb_sampler_serialize_body(s, buf, capacity, size); });
}

static int cgb_sampler_deserialize_body(cgb_sampler *s, const void *buf, uint64_t size)
{
    CGB_CHECK(s && buf, "cgb_sampler_deserialize: NULL argument");
    SamplerImage img;
    std::string err;
    ByteReader r(static_cast<const uint8_t*>(buf), static_cast<size_t>(size));
    if (!getSampler(r, s->sparse, img, err)) { return fail(CGB_EINVAL, "cgb_sampler_deserialize: " + err); }
    CGB_CHECK(r.remaining() == 0, "cgb_sampler_deserialize: trailing bytes");
    return imageToSampler(s, img);
}

extern "C" int cgb_sampler_deserialize(cgb_sampler *s, const void *buf, uint64_t size)
{
    return guarded("cgb_sampler_deserialize", [&]() { return cgb_sampler_deserialize_body(s, buf, size); });
}

static int cgb_sampler_set_atoms_body(cgb_sampler *s, const uint64_t *pos, const float *mass, uint64_t n)
{
    CGB_CHECK(s && (n == 0 || (pos && mass)), "cgb_sampler_set_atoms: NULL argument");
    CGB_CHECK(!s->persistentRunning, "cgb_sampler_set_atoms: called in the middle of an update");
    return setAtoms(s, pos, mass, n);
}

extern "C" int cgb_sampler_set_atoms(cgb_sampler *s, const uint64_t *pos, const float *mass, uint64_t n)
{
    return guarded("cgb_sampler_set_atoms", [&]() { return cgb_sampler_set_atoms_body(s, pos, mass, n); });
}

static int cgb_stats_serialize_body(const cgb_stats *st, void *buf, uint64_t capacity, uint64_t *size)
{
    CGB_CHECK(st && size, "cgb_stats_serialize: NULL argument");
    StatsImage img;
    CGB_TRY(statsToImage(st, img));
    ByteWriter w;
    putStats(w, img);
    return copyOut(w.bytes(), buf, capacity, size, "cgb_stats_serialize");
}

extern "C" int cgb_stats_serialize(const cgb_stats *st, void *buf, uint64_t capacity, uint64_t *size)
{
    return guarded("cgb_stats_serialize", [&]() { return cgb_stats_serialize_body(st, buf, capacity, size); });
}

static int cgb_stats_deserialize_body(cgb_stats *st, const void *buf, uint64_t size)
{
    CGB_CHECK(st && buf, "cgb_stats_deserialize: NULL argument");
    StatsImage img;
    std::string err;
    ByteReader r(static_cast<const uint8_t*>(buf), static_cast<size_t>(size));
    if (!getStats(r, img, err)) { return fail(CGB_EINVAL, "cgb_stats_deserialize: " + err); }
    CGB_CHECK(r.remaining() == 0, "cgb_stats_deserialize: trailing bytes");
    return imageToStats(st, img);
}

extern "C" int cgb_stats_deserialize(cgb_stats *st, const void *buf, uint64_t size)
{
    return guarded("cgb_stats_deserialize", [&]() { return cgb_stats_deserialize_body(st, buf, size); });
}

static int cgb_randstate_get_state_body(const cgb_randstate *rs, uint64_t state[2])
{
    CGB_CHECK(rs && state, "cgb_randstate_get_state: NULL argument");
    rs->seeder.getState(state);
    return CGB_OK;
}

extern "C" int cgb_randstate_get_state(const cgb_randstate *rs, uint64_t state[2])
{
    return guarded("cgb_randstate_get_state", [&]() { return cgb_randstate_get_state_body(rs, state); });
}

static int cgb_randstate_set_state_body(cgb_randstate *rs, const uint64_t state[2])
{
    CGB_CHECK(rs && state, "cgb_randstate_set_state: NULL argument");
    rs->seeder.setState(state);
    return CGB_OK;
}

extern "C" int cgb_randstate_set_state(cgb_randstate *rs, const uint64_t state[2])
{
    return guarded("cgb_randstate_set_state", [&]() { return cgb_randstate_set_state_body(rs, state); });
}

static int cgb_rng_get_state_body(const cgb_rng *r, uint64_t *state)
{
    CGB_CHECK(r && state, "cgb_rng_get_state: NULL argument");
    *state = r->rng.state;
    return CGB_OK;
}

extern "C" int cgb_rng_get_state(const cgb_rng *r, uint64_t *state)
{
    return guarded("cgb_rng_get_state", [&]() { return cgb_rng_get_state_body(r, state); });
}

static int cgb_rng_set_state_body(cgb_rng *r, uint64_t state)
{
    CGB_CHECK(r != nullptr, "cgb_rng_set_state: NULL argument");
    r->rng.state = state;
    return CGB_OK;
}

extern "C" int cgb_rng_set_state(cgb_rng *r, uint64_t state)
{
    return guarded("cgb_rng_set_state", [&]() { return cgb_rng_set_state_body(r, state); });
}

static int cgb_checkpoint_info_read_body(const char *path, cgb_checkpoint_info *out)
{
    CGB_CHECK(path && out, "cgb_checkpoint_info_read: NULL argument");
    CGB_CHECK(out->struct_size == sizeof(cgb_checkpoint_info), "cgb_checkpoint_info_read: cgb_checkpoint_info ABI mismatch");
    std::vector<uint8_t> raw;
    std::string err;
    if (!readWholeFile(path, raw, err)) { return fail(CGB_EINVAL, "cgb_checkpoint_info_read: " + err); }
    CheckpointImage c;
    ByteReader r(raw.data(), raw.size());
    if (!getCheckpoint(r, c, err)) { return fail(CGB_EINVAL, std::string("cgb_checkpoint_info_read: ") + path + ": " + err); }
    out->seed = c.params.seed;
    out->nGenes = c.params.nGenes;
    out->nSamples = c.params.nSamples;
    out->nPatterns = c.params.nPatterns;
    out->nIterations = c.params.nIterations;
    out->alphaA = c.params.alphaA;
    out->alphaP = c.params.alphaP;
    out->maxGibbsMassA = c.params.maxGibbsMassA;
    out->maxGibbsMassP = c.params.maxGibbsMassP;
    out->useSparseOptimization = c.params.useSparseOptimization ? 1 : 0;
    out->checkpointInterval = c.params.checkpointInterval;
    out->phase = c.phase;
    out->iter = c.iter;
    out->nAtomsA = c.A.pos.size();
    out->nAtomsP = c.P.pos.size();
    out->statUpdates = c.stats.statUpdates;
    out->fileBytes = raw.size();
    return CGB_OK;
}

extern "C" int cgb_checkpoint_info_read(const char *path, cgb_checkpoint_info *out)
{
    return guarded("cgb_checkpoint_info_read", [&]() { return cgb_checkpoint_info_read_body(path, out); });
}

static int cgb_checkpoint_rewrite_body(const char *inPath, const char *outPath)
{
    CGB_CHECK(inPath && outPath, "cgb_checkpoint_rewrite: NULL argument");
    CheckpointImage c;
    std::string err;
    if (!readCheckpointFile(inPath, c, err)) { return fail(CGB_EINVAL, std::string("cgb_checkpoint_rewrite: ") + inPath + ": " + err); }
    if (!writeCheckpointFile(outPath, c, err)) { return fail(CGB_EINVAL, "cgb_checkpoint_rewrite: " + err); }
    return CGB_OK;
}

extern "C" int cgb_checkpoint_rewrite(const char *inPath, const char *outPath)
{
    return guarded("cgb_checkpoint_rewrite", [&]() { return cgb_checkpoint_rewrite_body(inPath, outPath); });
}

// ------------------------------------------------------------------------------------------------
// gaps::run — runCoGAPSAlgorithm / runOnePhase / updateSampler / displayStatus
// (src/GapsRunner.cpp:161-222, 272-327, 381-503)
// ------------------------------------------------------------------------------------------------
struct RunGuard
{
    cgb_randstate *rs;
    cgb_sampler *A, *P;
    cgb_stats *st;
    cgb_rng *rng;
    RunGuard() : rs(nullptr), A(nullptr), P(nullptr), st(nullptr), rng(nullptr) {}
    ~RunGuard()
    {
        const double t0 = nowSeconds();
        cgb_rng_destroy(rng);
        cgb_stats_destroy(st);
        cgb_sampler_destroy(A);
        cgb_sampler_destroy(P);
        cgb_randstate_destroy(rs);
        if (envInt("COGAPS_HOST_PROFILE", 0)) { std::printf("[cgb_run] teardown %.3f s\n", nowSeconds() - t0); }
    }
};

static const float *g_tableOverride[3] = {nullptr, nullptr, nullptr};
// tables used by cgb_run for the GapsRandomState it creates internally (NULLs restore the built-ins)
static int cgb_run_set_tables_body(const float *erf, const float *erfinv, const float *qgamma)
{
    g_tableOverride[0] = erf;
    g_tableOverride[1] = erfinv;
    g_tableOverride[2] = qgamma;
    return CGB_OK;
}

extern "C" int cgb_run_set_tables(const float *erf, const float *erfinv, const float *qgamma)
{
    return guarded("cgb_run_set_tables", [&]() { return cgb_run_set_tables_body(erf, erfinv, qgamma); });
}

// createCheckpoint (GapsRunner.cpp:226-256): archive the run, then rebuild AP from the factors so that the chain
// that carries on is the one a resumed run will follow
static int createCheckpoint(const cgb_params *p, uint32_t nGenes, uint32_t nSamples, uint32_t interval, const char *path,
                            RunGuard &g, int phase, uint32_t iter)
{
    CheckpointImage c;
    c.params.seed = p->seed;
    c.params.nGenes = nGenes;
    c.params.nSamples = nSamples;
    c.params.nPatterns = p->nPatterns;
    c.params.nIterations = p->nIterations;
    c.params.alphaA = p->alphaA;
    c.params.alphaP = p->alphaP;
    c.params.maxGibbsMassA = p->maxGibbsMassA;
    c.params.maxGibbsMassP = p->maxGibbsMassP;
    c.params.useSparseOptimization = p->useSparseOptimization != 0;
    c.params.checkpointInterval = interval;
    g.rs->seeder.getState(c.seeder);
    CGB_TRY(samplerToImage(g.A, c.A));
    CGB_TRY(samplerToImage(g.P, c.P));
    CGB_TRY(statsToImage(g.st, c.stats));
    c.phase = phase;
    c.iter = iter;
    c.rng = g.rng->rng.state;
    std::string err;
    if (!writeCheckpointFile(path, c, err)) { return fail(CGB_EINVAL, "checkpoint: " + err); }
    CGB_TRY(cgb_sampler_extra_initialization(g.A));
    CGB_TRY(cgb_sampler_extra_initialization(g.P));
    return CGB_OK;
}

static int cgb_run_body(const float *data, uint32_t nrow, uint32_t ncol, int32_t colmajor,
                       const float *uncertainty, const cgb_params *p, cgb_result *r)
{
    return cgb_run_ex(data, nrow, ncol, colmajor, uncertainty, p, nullptr, r);
}

extern "C" int cgb_run(const float *data, uint32_t nrow, uint32_t ncol, int32_t colmajor,
                       const float *uncertainty, const cgb_params *p, cgb_result *r)
{
    return guarded("cgb_run", [&]() { return cgb_run_body(data, nrow, ncol, colmajor, uncertainty, p, r); });
}

// both orientations of a Matrix-Market file in compressed rows (see cgb_run_file)
struct CsrPair
{
    std::vector<uint32_t> rowPtr, rowIdx, colPtr, colIdx; // by file row / by file column
    std::vector<float> rowVal, colVal;
};

static int runCore(const float *data, const CsrPair *csr, uint32_t nrow, uint32_t ncol, int32_t colmajor, const float *uncertainty,
                   const cgb_params *p0, const cgb_run_options *opt, cgb_result *r);

static int cgb_run_ex_body(const float *data, uint32_t nrow, uint32_t ncol, int32_t colmajor, const float *uncertainty,
                          const cgb_params *p0, const cgb_run_options *opt, cgb_result *r)
{
    CGB_CHECK(data != nullptr, "cgb_run: NULL argument");
    return runCore(data, nullptr, nrow, ncol, colmajor, uncertainty, p0, opt, r);
}

extern "C" int cgb_run_ex(const float *data, uint32_t nrow, uint32_t ncol, int32_t colmajor, const float *uncertainty,
                          const cgb_params *p0, const cgb_run_options *opt, cgb_result *r)
{
    return guarded("cgb_run_ex", [&]() { return cgb_run_ex_body(data, nrow, ncol, colmajor, uncertainty, p0, opt, r); });
}

static int runCore(const float *data, const CsrPair *csr, uint32_t nrow, uint32_t ncol, int32_t colmajor, const float *uncertainty,
                   const cgb_params *p0, const cgb_run_options *opt, cgb_result *r)
{
    CGB_CHECK((data || csr) && p0 && r, "cgb_run: NULL argument");
    CGB_CHECK(p0->struct_size == sizeof(cgb_params), "cgb_run: cgb_params ABI mismatch");
    CGB_CHECK(r->struct_size == sizeof(cgb_result), "cgb_run: cgb_result ABI mismatch");
    CGB_CHECK(opt == nullptr || opt->struct_size == sizeof(cgb_run_options), "cgb_run_ex: cgb_run_options ABI mismatch");
    cgb_params pv = *p0; // a checkpoint overwrites some of the caller's parameters (run_helper, GapsRunner.cpp:99-105)
    const cgb_params *p = &pv;
    uint32_t ckInterval = opt ? opt->checkpointInterval : 0;
    const char *ckIn = (opt && opt->checkpointInFile && opt->checkpointInFile[0]) ? opt->checkpointInFile : nullptr;
    const char *ckOut = (opt && opt->checkpointOutFile && opt->checkpointOutFile[0]) ? opt->checkpointOutFile : "gaps_checkpoint.out";
    uint64_t ckSeeder[2] = {0, 0};
    if (ckIn)
    {
        ParamsImage pi;
        std::string err;
        if (!readCheckpointHeader(ckIn, pi, ckSeeder, err)) { return fail(CGB_EINVAL, std::string("cgb_run_ex: ") + ckIn + ": " + err); }
        // the caller sized its result arrays from its own nPatterns; the reference would silently switch to the file's
        CGB_CHECK(pi.nPatterns == p0->nPatterns, "cgb_run_ex: nPatterns differs from the checkpoint's (cgb_checkpoint_info_read tells what it holds)");
        pv.seed = pi.seed;
        pv.nPatterns = pi.nPatterns;
        pv.nIterations = pi.nIterations;
        pv.alphaA = pi.alphaA;
        pv.alphaP = pi.alphaP;
        pv.maxGibbsMassA = pi.maxGibbsMassA;
        pv.maxGibbsMassP = pi.maxGibbsMassP;
        pv.useSparseOptimization = pi.useSparseOptimization ? 1 : 0;
        ckInterval = pi.checkpointInterval;
    }
    if ((ckInterval > 0 || ckIn) && !p->asynchronousUpdates) { return fail(CGB_EUNSUPPORTED, kSequentialCheckpointMsg); }
    const int fixed = p->whichMatrixFixed ? p->whichMatrixFixed : 'N';
    CGB_CHECK(fixed == 'N' || fixed == 'A' || fixed == 'P', "cgb_run: whichMatrixFixed must be 'N', 'A' or 'P'");
    const bool useFixed = p->fixedPatterns != nullptr && fixed != 'N';

    uint32_t nGenes = p->transposeData ? ncol : nrow;
    uint32_t nSamples = p->transposeData ? nrow : ncol;
    if (p->nSubsetIndices && p->subsetGenes) { nGenes = p->nSubsetIndices; }
    if (p->nSubsetIndices && !p->subsetGenes) { nSamples = p->nSubsetIndices; }

    RunGuard g;
    const double tEnter = nowSeconds();
    CGB_TRY(cgb_randstate_create(p0->seed, &g.rs)); // the caller builds GapsRandomState from ITS seed (Cogaps.cpp:141-142)
    if (ckIn) { g.rs->seeder.setState(ckSeeder); }
    if (g_tableOverride[0]) { CGB_TRY(cgb_randstate_set_tables(g.rs, g_tableOverride[0], g_tableOverride[1], g_tableOverride[2])); }
    // GapsRunner.cpp:402-406.  The two orientations are prepared concurrently (each is a pass over the whole
    // matrix on the host: orientation, the fp32 running sum behind lambda, upload); the generators are then
    // built A first, P second — that fixes which seeds the two queue rngs get.
    CGB_TRY(ensureDevice());
    CGB_TRY(uploadTables(g.rs));
    {
        // which of the two sees its rows contiguous in the caller's matrix (exactly one does, subsets aside)
        const bool dense = !p->useSparseOptimization && p->nSubsetIndices == 0;
        const bool straightA = ((!p->transposeData) != (colmajor != 0));
        TwinLink link;
        TwinLink *pubA = (dense && straightA) ? &link : nullptr, *twinA = (dense && !straightA) ? &link : nullptr;
        TwinLink *pubP = (dense && !straightA) ? &link : nullptr, *twinP = (dense && straightA) ? &link : nullptr;
        // compressed-row input: a sampler built with transpose == true has the file's rows as its rows
        // (orientData: nRows = transpose ? nrow : ncol), the other one the file's columns
        CsrView byRow = {nrow, ncol, nullptr, nullptr, nullptr}, byCol = {ncol, nrow, nullptr, nullptr, nullptr};
        if (csr)
        {
            CGB_CHECK(p->useSparseOptimization && p->nSubsetIndices == 0 && !uncertainty, "cgb_run: compressed-row input needs the sparse model, the whole matrix and default uncertainty");
            byRow.ptr = &csr->rowPtr; byRow.idx = &csr->rowIdx; byRow.val = &csr->rowVal;
            byCol.ptr = &csr->colPtr; byCol.idx = &csr->colIdx; byCol.val = &csr->colVal;
        }
        const CsrView *csrP = csr ? (p->transposeData ? &byRow : &byCol) : nullptr;
        const CsrView *csrA = csr ? (p->transposeData ? &byCol : &byRow) : nullptr;
        int rcP = CGB_OK;
        std::string errP;
        std::thread prepP([&]()
        {
            rcP = guarded("cgb_run (P sampler)", [&]() { return samplerCreateImpl(data, nrow, ncol, colmajor, p->transposeData, p->subsetGenes, p->alphaP, p->maxGibbsMassP, p, g.rs, false, &g.P, pubP, twinP, csrP); });
            if (rcP != CGB_OK)
            {
                errP = g_lastError;
                if (pubP && pubP->state.load() == 0) { pubP->state.store(-1); } // never leave the twin waiting
            }
        });
        struct JoinOnExit { std::thread &t; ~JoinOnExit() { if (t.joinable()) { t.join(); } } } joinP = {prepP};
        const int rcA = guarded("cgb_run (A sampler)", [&]() { return samplerCreateImpl(data, nrow, ncol, colmajor, !p->transposeData, !p->subsetGenes, p->alphaA, p->maxGibbsMassA, p, g.rs, false, &g.A, pubA, twinA, csrA); });
        if (rcA != CGB_OK && pubA && pubA->state.load() == 0) { pubA->state.store(-1); }
        prepP.join();
        if (rcA != CGB_OK) { return rcA; }
        if (rcP != CGB_OK) { return fail(rcP, errP); }
        initGenerator(g.A, p, g.rs);
        initGenerator(g.P, p, g.rs);
    }
    if (uncertainty)
    {
        CGB_TRY(cgb_sampler_set_uncertainty(g.A, uncertainty, nrow, ncol, colmajor, !p->transposeData, !p->subsetGenes, p));
        CGB_TRY(cgb_sampler_set_uncertainty(g.P, uncertainty, nrow, ncol, colmajor, p->transposeData, p->subsetGenes, p));
    }
    if (useFixed)
    {
        if (fixed == 'A') { CGB_TRY(cgb_sampler_set_matrix(g.A, p->fixedPatterns)); }
        if (fixed == 'P') { CGB_TRY(cgb_sampler_set_matrix(g.P, p->fixedPatterns)); }
    }
    CGB_CHECK(p->updateMode == CGB_UPDATE_EXACT || p->updateMode == CGB_UPDATE_SWEEP, "cgb_run: unknown updateMode");
    CGB_TRY(cgb_stats_create(nGenes, nSamples, p->nPatterns, &g.st));
    CGB_TRY(cgb_rng_create(g.rs, &g.rng)); // GapsRunner.cpp:437
    int startPhase = CGB_PHASE_EQUILIBRATION;
    uint32_t startIter = 0;
    if (ckIn)
    {
        // processCheckpoint, GapsRunner.cpp:258-270
        CheckpointImage c;
        std::string err;
        if (!readCheckpointFile(ckIn, c, err)) { return fail(CGB_EINVAL, std::string("cgb_run_ex: ") + ckIn + ": " + err); }
        CGB_CHECK(c.params.nGenes == nGenes && c.params.nSamples == nSamples, "cgb_run_ex: the checkpoint was made from data of another shape");
        CGB_CHECK(c.phase == CGB_PHASE_EQUILIBRATION || c.phase == CGB_PHASE_SAMPLING, "cgb_run_ex: checkpoint holds an unknown phase");
        g.rs->seeder.setState(c.seeder);
        CGB_TRY(imageToSampler(g.A, c.A));
        CGB_TRY(imageToSampler(g.P, c.P));
        CGB_TRY(imageToStats(g.st, c.stats));
        startPhase = c.phase;
        startIter = c.iter;
        g.rng->rng.state = c.rng;
    }

    CGB_TRY(cgb_sampler_sync(g.A, g.P));
    CGB_TRY(cgb_sampler_sync(g.P, g.A));
    CGB_TRY(cgb_sampler_extra_initialization(g.A));
    CGB_TRY(cgb_sampler_extra_initialization(g.P));
    CGB_TRY(cgb_sampler_set_update_mode(g.A, p->updateMode));
    CGB_TRY(cgb_sampler_set_update_mode(g.P, p->updateMode));

    const double tStart = nowSeconds();
    if (envInt("COGAPS_HOST_PROFILE", 0)) { std::printf("[cgb_run] setup (tables, both orientations, upload, sync, AP rebuild) %.3f s\n", tStart - tEnter); }
    uint64_t totalUpdates = 0;
    uint32_t nHist = 0, nSnapEq = 0, nSnapSamp = 0;
    double secondsA = 0.0, secondsP = 0.0;
    for (int phase = startPhase; phase <= CGB_PHASE_SAMPLING; ++phase)
    {
        if (p->printMessages) { std::printf(phase == CGB_PHASE_EQUILIBRATION ? "-- Equilibration Phase --\n" : "-- Sampling Phase --\n"); }
        for (uint32_t iter = (phase == startPhase) ? startIter : 0; iter < p->nIterations; ++iter)
        {
            // gaps_check_interrupt + createCheckpoint, GapsRunner.cpp:280-282
            if (opt && opt->interrupt && opt->interrupt(opt->interruptUser) != 0) { return fail(CGB_EINTERRUPTED, "cgb_run_ex: interrupted by the caller"); }
            if (ckInterval > 0 && ((iter + 1) % ckInterval) == 0 && p->nSubsetIndices == 0)
            {
                CGB_TRY(createCheckpoint(p, nGenes, nSamples, ckInterval, ckOut, g, phase, iter));
            }
            if (phase == CGB_PHASE_EQUILIBRATION)
            {
                const float temp = static_cast<float>(2 * iter) / static_cast<float>(p->nIterations);
                g.A->annealingTemp = gmin(1.f, temp);
                g.P->annealingTemp = gmin(1.f, temp);
            }
            uint64_t nAtomsA = 0, nAtomsP = 0;
            CGB_TRY(cgb_sampler_n_atoms(g.A, &nAtomsA));
            CGB_TRY(cgb_sampler_n_atoms(g.P, &nAtomsP));
            const unsigned atomsA = static_cast<unsigned>(nAtomsA);
            const unsigned atomsP = static_cast<unsigned>(nAtomsP);
            const unsigned nA = static_cast<unsigned>(g.rng->rng.poisson(static_cast<double>(atomsA < 10u ? 10u : atomsA)));
            const unsigned nP = static_cast<unsigned>(g.rng->rng.poisson(static_cast<double>(atomsP < 10u ? 10u : atomsP)));
            // updateSampler, GapsRunner.cpp:201-222
            if (fixed != 'A')
            {
                const double t0 = nowSeconds();
                CGB_TRY(cgb_sampler_update(g.A, nA, p->maxThreads));
                secondsA += nowSeconds() - t0;
                if (fixed != 'P') { CGB_TRY(cgb_sampler_sync(g.P, g.A)); }
            }
            if (fixed != 'P')
            {
                const double t0 = nowSeconds();
                CGB_TRY(cgb_sampler_update(g.P, nP, p->maxThreads));
                secondsP += nowSeconds() - t0;
                if (fixed != 'A') { CGB_TRY(cgb_sampler_sync(g.A, g.P)); }
            }
            totalUpdates += nA + nP;
            if (phase == CGB_PHASE_SAMPLING)
            {
                if (useFixed)
                {
                    if (fixed == 'A') { CGB_TRY(cgb_stats_update_p(g.st, g.A, g.P)); }
                    else { CGB_TRY(cgb_stats_update_a(g.st, g.A, g.P)); }
                }
                else
                {
                    CGB_TRY(cgb_stats_update(g.st, g.A, g.P));
                    if (p->takePumpSamples) { CGB_TRY(cgb_stats_update_pump(g.st, g.A)); }
                }
            }
            if (static_cast<int>(p->snapshotPhase) == phase || p->snapshotPhase == CGB_PHASE_ALL)
            {
                if (p->snapshotFrequency > 0 && ((iter + 1) % p->snapshotFrequency) == 0)
                {
                    const uint32_t slot = nSnapEq + nSnapSamp;
                    if (slot < r->snapshotCapacity)
                    {
                        if (r->snapshotsA) { CGB_TRY(cgb_sampler_get_matrix(g.A, r->snapshotsA + static_cast<size_t>(slot) * nGenes * p->nPatterns)); }
                        if (r->snapshotsP) { CGB_TRY(cgb_sampler_get_matrix(g.P, r->snapshotsP + static_cast<size_t>(slot) * nSamples * p->nPatterns)); }
                    }
                    // only snapshots that were stored are counted: the caller splits its arrays by these numbers
                    if (slot < r->snapshotCapacity) { if (phase == CGB_PHASE_EQUILIBRATION) { ++nSnapEq; } else { ++nSnapSamp; } }
                }
            }
            // displayStatus, GapsRunner.cpp:161-199
            if (p->outputFrequency > 0 && ((iter + 1) % p->outputFrequency) == 0)
            {
                float cs = 0.f;
                CGB_TRY(cgb_sampler_chisq(fixed == 'P' ? g.A : g.P, &cs));
                uint64_t na = 0, nb = 0;
                CGB_TRY(cgb_sampler_n_atoms(g.A, &na));
                CGB_TRY(cgb_sampler_n_atoms(g.P, &nb));
                const unsigned a = static_cast<unsigned>(na), b = static_cast<unsigned>(nb);
                if (nHist < r->historyCapacity)
                {
                    if (r->chisqHistory) { r->chisqHistory[nHist] = cs; }
                    if (r->atomHistoryA) { r->atomHistoryA[nHist] = a; }
                    if (r->atomHistoryP) { r->atomHistoryP[nHist] = b; }
                    ++nHist;
                }
                if (p->printMessages)
                {
                    std::printf("%d of %d, Atoms: %d(A), %d(P), ChiSq: %.0f, elapsed %.1f s\n", iter + 1, p->nIterations, a, b, cs, nowSeconds() - tStart);
                    std::fflush(stdout);
                }
            }
        }
    }
    r->totalRunningTime = nowSeconds() - tStart;

    if (r->Amean) { CGB_TRY(cgb_stats_amean(g.st, r->Amean)); }
    if (r->Asd) { CGB_TRY(cgb_stats_asd(g.st, r->Asd)); }
    if (r->Pmean) { CGB_TRY(cgb_stats_pmean(g.st, r->Pmean)); }
    if (r->Psd) { CGB_TRY(cgb_stats_psd(g.st, r->Psd)); }
    r->nHistory = nHist;
    r->nSnapshotsEquilibration = nSnapEq;
    r->nSnapshotsSampling = nSnapSamp;
    r->seed = p->seed;
    r->totalUpdates = totalUpdates;
    if (p->updateMode == CGB_UPDATE_SWEEP)
    {
        // the sweep rounds each row's share of nSteps stochastically: report the proposals actually made
        r->totalUpdates = (fixed != 'A' ? g.A->counters.nProposalsTotal : 0) + (fixed != 'P' ? g.P->counters.nProposalsTotal : 0);
    }
    r->averageQueueLengthA = g.A->avgQueueLength;
    r->averageQueueLengthP = g.P->avgQueueLength;
    r->meanChiSq = 0.f; // GapsRunner.cpp:478-485: zero whenever a matrix is fixed
    if (fixed == 'N') { CGB_TRY(cgb_stats_mean_chisq(g.st, g.P, &r->meanChiSq)); }
    if (p->takePumpSamples)
    {
        if (r->pumpMatrix) { CGB_TRY(cgb_stats_pump_matrix(g.st, r->pumpMatrix)); }
        if (r->meanPatternAssignment) { CGB_TRY(cgb_stats_mean_pattern(g.st, r->meanPatternAssignment)); }
    }
    r->nBatchesA = g.A->counters.nBatches;
    r->nBatchesP = g.P->counters.nBatches;
    r->secondsUpdateA = secondsA;
    r->secondsUpdateP = secondsP;
    r->secondsDevice = g.A->counters.secondsKernel + g.P->counters.secondsKernel;
    r->algorithmicBytes = g.A->counters.algorithmicBytes + g.P->counters.algorithmicBytes;
    if (envInt("COGAPS_HOST_PROFILE", 0))
    {
        std::printf("[cgb_run] loop %.3f s, results %.3f s (teardown follows)\n", r->totalRunningTime, nowSeconds() - tStart - r->totalRunningTime);
    }
    return CGB_OK;
}


// ------------------------------------------------------------------------------------------------
// Multi-GPU: one process per GPU, each running an independent chain on its own shard (what distributed
// CoGAPS does per set, R/DistributedCogaps.R:48-119).  The only exchange on the path is the concatenation of per-shard
// factor rows (stitchTogether, R/DistributedCogaps.R:226-278): an NCCL all-gather straight from device memory.  NCCL is
// bound at run time (dlopen) so that a process which already carries one (torch) shares it and a single-GPU host needs none.
// ------------------------------------------------------------------------------------------------
#include <dlfcn.h>

namespace {
struct NcclUniqueId { char internal[CGB_COMM_UNIQUE_ID_BYTES]; };
typedef int (*NcclGetUniqueIdFn)(NcclUniqueId*);
typedef int (*NcclCommInitRankFn)(void**, int, NcclUniqueId, int);
typedef int (*NcclCommDestroyFn)(void*);
typedef int (*NcclAllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char *(*NcclGetErrorStringFn)(int);
struct NcclApi
{
    void *handle;
    NcclGetUniqueIdFn getUniqueId;
    NcclCommInitRankFn commInitRank;
    NcclCommDestroyFn commDestroy;
    NcclAllGatherFn allGather;
    NcclGetErrorStringFn errorString;
};
NcclApi g_nccl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
std::mutex g_ncclLock;
const int kNcclFloat32 = 7; // ncclFloat32 (nccl.h)
}

static int loadNccl()
{
    std::lock_guard<std::mutex> hold(g_ncclLock);
    if (g_nccl.handle) { return CGB_OK; }
    const char *names[] = {std::getenv("COGAPS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (size_t i = 0; i < sizeof(names) / sizeof(names[0]) && !h; ++i)
    {
        if (names[i] && names[i][0]) { h = dlopen(names[i], RTLD_NOW | RTLD_LOCAL); }
    }
    if (!h) { return fail(CGB_EUNSUPPORTED, "multi-GPU: libnccl.so.2 not found (set COGAPS_NCCL_LIB to its path)"); }
    NcclApi api;
    api.handle = h;
    api.getUniqueId = reinterpret_cast<NcclGetUniqueIdFn>(dlsym(h, "ncclGetUniqueId"));
    api.commInitRank = reinterpret_cast<NcclCommInitRankFn>(dlsym(h, "ncclCommInitRank"));
    api.commDestroy = reinterpret_cast<NcclCommDestroyFn>(dlsym(h, "ncclCommDestroy"));
    api.allGather = reinterpret_cast<NcclAllGatherFn>(dlsym(h, "ncclAllGather"));
    api.errorString = reinterpret_cast<NcclGetErrorStringFn>(dlsym(h, "ncclGetErrorString"));
    if (!api.getUniqueId || !api.commInitRank || !api.commDestroy || !api.allGather)
    {
        dlclose(h);
        return fail(CGB_EUNSUPPORTED, "multi-GPU: the NCCL library lacks ncclGetUniqueId / ncclCommInitRank / ncclAllGather");
    }
    g_nccl = api;
    return CGB_OK;
}

static int ncclFail(const char *what, int rc)
{
    const char *msg = g_nccl.errorString ? g_nccl.errorString(rc) : "";
    return fail(CGB_ECUDA, std::string(what) + ": NCCL error " + std::to_string(rc) + " " + (msg ? msg : ""));
}

struct cgb_comm
{
    void *comm;
    int rank, nRanks, device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    float *dSend, *dRecv;   // staging: [k][ldMax] and [nRanks][k][ldMax]
    size_t sendFloats, recvFloats;
};

static int cgb_comm_get_unique_id_body(uint8_t *id)
{
    CGB_CHECK(id != nullptr, "cgb_comm_get_unique_id: NULL argument");
    CGB_TRY(loadNccl());
    NcclUniqueId u;
    const int rc = g_nccl.getUniqueId(&u);
    if (rc != 0) { return ncclFail("cgb_comm_get_unique_id", rc); }
    std::memcpy(id, u.internal, CGB_COMM_UNIQUE_ID_BYTES);
    return CGB_OK;
}

extern "C" int cgb_comm_get_unique_id(uint8_t *id)
{
    return guarded("cgb_comm_get_unique_id", [&]() { return cgb_comm_get_unique_id_body(id); });
}

extern "C" void cgb_comm_destroy(cgb_comm *c)
{
    if (!c) { return; }
    cudaSetDevice(c->device);
    if (c->comm && g_nccl.commDestroy) { g_nccl.commDestroy(c->comm); }
    cudaFree(c->dSend);
    cudaFree(c->dRecv);
    if (c->ev0) { cudaEventDestroy(c->ev0); }
    if (c->ev1) { cudaEventDestroy(c->ev1); }
    if (c->stream) { cudaStreamDestroy(c->stream); }
    delete c;
}

static int cgb_comm_init_body(const uint8_t *id, int32_t rank, int32_t nRanks, cgb_comm **out)
{
    CGB_CHECK(id && out, "cgb_comm_init: NULL argument");
    CGB_CHECK(nRanks >= 1 && rank >= 0 && rank < nRanks, "cgb_comm_init: rank must be in [0, nRanks)");
    CGB_TRY(ensureDevice());
    CGB_TRY(loadNccl());
    cgb_comm *c = new (std::nothrow) cgb_comm();
    if (!c) { return fail(CGB_ENOMEM, "cgb_comm_init: out of memory"); }
    c->rank = rank;
    c->nRanks = nRanks;
    c->device = g_device;
    NcclUniqueId u;
    std::memcpy(u.internal, id, CGB_COMM_UNIQUE_ID_BYTES);
    int rc = CGB_OK;
    do
    {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&c->ev0) != cudaSuccess
            || cudaEventCreate(&c->ev1) != cudaSuccess) { rc = fail(CGB_ECUDA, "cgb_comm_init: stream / event creation failed"); break; }
        const int nrc = g_nccl.commInitRank(&c->comm, nRanks, u, rank);
        if (nrc != 0) { c->comm = nullptr; rc = ncclFail("cgb_comm_init (ncclCommInitRank)", nrc); break; }
    } while (false);
    if (rc != CGB_OK) { cgb_comm_destroy(c); return rc; }
    *out = c;
    return CGB_OK;
}

extern "C" int cgb_comm_init(const uint8_t *id, int32_t rank, int32_t nRanks, cgb_comm **out)
{
    return guarded("cgb_comm_init", [&]() { return cgb_comm_init_body(id, rank, nRanks, out); });
}

// gathers pattern-major device blocks ([k][ld], element (row r, pattern p) at dev[p*ld + r]; rowsPerRank[rank] rows here)
// of every rank into one row-major host matrix (sum of rows) x k, ranks in order
// dev: this rank's factor block — pattern-major [k][ld] (the dense model's matrix), or, rowMajor, [rows][ld] with k values per
// row (the sparse model's row copy, the one HybridMatrix -> Matrix conversions and cgb_sampler_get_matrix read)
static int allgatherRows(cgb_comm *c, const float *dev, uint64_t ld, uint32_t k, const uint32_t *rowsPerRank, float *out, double *deviceMs,
                         bool rowMajor = false)
{
    CGB_CHECK(c && dev && rowsPerRank && out, "cgb_allgather_rows: NULL argument");
    CGB_CUDA(cudaSetDevice(c->device));
    uint32_t maxRows = 0;
    uint64_t totalRows = 0;
    for (int r = 0; r < c->nRanks; ++r) { maxRows = std::max(maxRows, rowsPerRank[r]); totalRows += rowsPerRank[r]; }
    const uint32_t mine = rowsPerRank[c->rank];
    CGB_CHECK(rowMajor ? k <= ld : mine <= ld, "cgb_allgather_rows: rowsPerRank[rank] exceeds the block's stride");
    const size_t ldMax = rowMajor ? static_cast<size_t>(ld) : roundUp(std::max(maxRows, 1u), 32);
    const size_t sendFloats = rowMajor ? static_cast<size_t>(std::max(maxRows, 1u)) * ldMax : static_cast<size_t>(k) * ldMax;
    const size_t recvFloats = sendFloats * c->nRanks;
    if (sendFloats > c->sendFloats)
    {
        cudaFree(c->dSend); c->dSend = nullptr; c->sendFloats = 0;
        CGB_CUDA(cudaMalloc(&c->dSend, sendFloats * sizeof(float)));
        c->sendFloats = sendFloats;
    }
    if (recvFloats > c->recvFloats)
    {
        cudaFree(c->dRecv); c->dRecv = nullptr; c->recvFloats = 0;
        CGB_CUDA(cudaMalloc(&c->dRecv, recvFloats * sizeof(float)));
        c->recvFloats = recvFloats;
    }
    // equal-size blocks for the collective: this rank's k columns, each padded to the widest shard
    CGB_CUDA(cudaMemsetAsync(c->dSend, 0, sendFloats * sizeof(float), c->stream));
    if (rowMajor) { CGB_CUDA(cudaMemcpyAsync(c->dSend, dev, static_cast<size_t>(mine) * ld * sizeof(float), cudaMemcpyDeviceToDevice, c->stream)); }
    else
    {
        CGB_CUDA(cudaMemcpy2DAsync(c->dSend, ldMax * sizeof(float), dev, ld * sizeof(float), static_cast<size_t>(mine) * sizeof(float), k,
                                   cudaMemcpyDeviceToDevice, c->stream));
    }
    CGB_CUDA(cudaEventRecord(c->ev0, c->stream));
    const int nrc = g_nccl.allGather(c->dSend, c->dRecv, sendFloats, kNcclFloat32, c->comm, c->stream);
    if (nrc != 0) { return ncclFail("cgb_allgather_rows (ncclAllGather)", nrc); }
    CGB_CUDA(cudaEventRecord(c->ev1, c->stream));
    std::vector<float> host(recvFloats);
    CGB_CUDA(cudaMemcpyAsync(host.data(), c->dRecv, recvFloats * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CGB_CUDA(cudaStreamSynchronize(c->stream));
    if (deviceMs)
    {
        float ms = 0.f;
        CGB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        *deviceMs = ms;
    }
    uint64_t row0 = 0;
    for (int r = 0; r < c->nRanks; ++r)
    {
        const float *blk = host.data() + static_cast<size_t>(r) * sendFloats;
        for (uint32_t i = 0; i < rowsPerRank[r]; ++i)
        {
            for (uint32_t p = 0; p < k; ++p) { out[(row0 + i) * k + p] = rowMajor ? blk[static_cast<size_t>(i) * ldMax + p] : blk[static_cast<size_t>(p) * ldMax + i]; }
        }
        row0 += rowsPerRank[r];
    }
    (void)totalRows;
    return CGB_OK;
}

static int cgb_allgather_rows_body(cgb_comm *c, const cgb_sampler *s, const uint32_t *rowsPerRank, float *out, double *deviceMs)
{
    CGB_CHECK(c && s && rowsPerRank, "cgb_allgather_rows: NULL argument");
    CGB_CHECK(rowsPerRank[c->rank] == s->nRows, "cgb_allgather_rows: rowsPerRank[rank] is not this sampler's row count");
    CGB_CUDA(cudaSetDevice(s->device));
    CGB_CUDA(cudaStreamSynchronize(s->stream));
    // the values cgb_sampler_get_matrix returns: the dense model's matrix, the sparse model's row copy (its column copy
    // stores values below epsilon as 0, data_structures/HybridMatrix.cpp:25-39)
    if (s->sparse) { return allgatherRows(c, s->dMrows, s->ldR, s->k, rowsPerRank, out, deviceMs, true); }
    return allgatherRows(c, s->dM, s->ldM, s->k, rowsPerRank, out, deviceMs);
}

extern "C" int cgb_allgather_rows(cgb_comm *c, const cgb_sampler *s, const uint32_t *rowsPerRank, float *out, double *deviceMs)
{
    return guarded("cgb_allgather_rows", [&]() { return cgb_allgather_rows_body(c, s, rowsPerRank, out, deviceMs); });
}

static int cgb_allgather_device_rows_body(cgb_comm *c, const void *dev, uint64_t ld, uint32_t k, const uint32_t *rowsPerRank, float *out, double *deviceMs)
{
    return allgatherRows(c, static_cast<const float*>(dev), ld, k, rowsPerRank, out, deviceMs);
}

extern "C" int cgb_allgather_device_rows(cgb_comm *c, const void *dev, uint64_t ld, uint32_t k, const uint32_t *rowsPerRank, float *out, double *deviceMs)
{
    return guarded("cgb_allgather_device_rows", [&]() { return cgb_allgather_device_rows_body(c, dev, ld, k, rowsPerRank, out, deviceMs); });
}

// ------------------------------------------------------------------------------------------------
// gaps::run(const std::string &data, ...) — the path overload (src/GapsRunner.h:19-24, GapsRunner.cpp:119-159)
// ------------------------------------------------------------------------------------------------
namespace cgb {
bool loadMatrixFile(const char *path, std::vector<float> &out, uint32_t &nrow, uint32_t &ncol, std::string &err);
bool loadMtxTriplets(const char *path, std::vector<uint32_t> &rows, std::vector<uint32_t> &cols, std::vector<float> &vals,
                     uint32_t &nrow, uint32_t &ncol, std::string &err);
void compressTriplets(const std::vector<uint32_t> &major, const std::vector<uint32_t> &minor, const std::vector<float> &vals,
                      uint32_t nMajor, std::vector<uint32_t> &ptr, std::vector<uint32_t> &idx, std::vector<float> &out,
                      bool &anyNegative);
}

namespace cgb { bool loadColumnNames(const char *path, std::vector<std::string> &names, std::string &err); }

// getFileInfo_cpp's colNames (src/Cogaps.cpp:245-256): the names, each followed by a NUL, packed into buf
extern "C" int cgb_file_col_names(const char *path, char *buf, uint64_t capacity, uint64_t *needed, uint32_t *count)
{
    if (!path || !needed || !count) { return fail(CGB_EINVAL, "cgb_file_col_names: NULL argument"); }
    return guarded("cgb_file_col_names", [&]()
    {
        std::vector<std::string> names;
        std::string err;
        if (!loadColumnNames(path, names, err)) { return fail(CGB_EINVAL, "cgb_file_col_names: " + err); }
        uint64_t total = 0;
        for (size_t i = 0; i < names.size(); ++i) { total += names[i].size() + 1; }
        *needed = total;
        *count = static_cast<uint32_t>(names.size());
        if (buf != nullptr)
        {
            if (capacity < total) { return fail(CGB_EINVAL, "cgb_file_col_names: buffer too small"); }
            char *p = buf;
            for (size_t i = 0; i < names.size(); ++i)
            {
                std::memcpy(p, names[i].c_str(), names[i].size() + 1);
                p += names[i].size() + 1;
            }
        }
        return CGB_OK;
    });
}

namespace cgb { bool writeMatrixCsv(const char *path, const float *mat, uint32_t nrow, uint32_t ncol, std::string &err); }

extern "C" int cgb_write_matrix_csv(const char *path, const float *mat, uint32_t nrow, uint32_t ncol)
{
    if (!path || !mat) { return fail(CGB_EINVAL, "cgb_write_matrix_csv: NULL argument"); }
    return guarded("cgb_write_matrix_csv", [&]()
    {
        std::string err;
        if (!writeMatrixCsv(path, mat, nrow, ncol, err)) { return fail(CGB_EINVAL, "cgb_write_matrix_csv: " + err); }
        return CGB_OK;
    });
}

// GapsResult::writeToFile (GapsResult.cpp:27-35)
extern "C" int cgb_result_write_files(const char *pathPrefix, uint32_t nGenes, uint32_t nSamples, uint32_t nPatterns, const cgb_result *r)
{
    if (!pathPrefix || !r || !r->Amean || !r->Pmean || !r->Asd || !r->Psd) { return fail(CGB_EINVAL, "cgb_result_write_files: NULL argument"); }
    return guarded("cgb_result_write_files", [&]()
    {
        const std::string label = std::string(pathPrefix) + "_" + std::to_string(nPatterns) + "_";
        CGB_TRY(cgb_write_matrix_csv((label + "Amean.csv").c_str(), r->Amean, nGenes, nPatterns));
        CGB_TRY(cgb_write_matrix_csv((label + "Pmean.csv").c_str(), r->Pmean, nSamples, nPatterns));
        CGB_TRY(cgb_write_matrix_csv((label + "Asd.csv").c_str(), r->Asd, nGenes, nPatterns));
        CGB_TRY(cgb_write_matrix_csv((label + "Psd.csv").c_str(), r->Psd, nSamples, nPatterns));
        return CGB_OK;
    });
}

static bool endsWith(const char *s, const char *suffix)
{
    const size_t n = std::strlen(s), m = std::strlen(suffix);
    return n >= m && std::strcmp(s + n - m, suffix) == 0;
}

// Matrix-Market -> both orientations in compressed rows, no dense matrix in between.  False (with err empty) when the
// file holds something the compressed form cannot carry bit for bit (a negative value: lambda sums those too).
static bool loadMtxCsrPair(const char *path, CsrPair &pair, uint32_t &nrow, uint32_t &ncol, std::string &err)
{
    std::vector<uint32_t> rows, cols;
    std::vector<float> vals;
    if (!loadMtxTriplets(path, rows, cols, vals, nrow, ncol, err)) { return false; }
    if (vals.size() > 0xFFFFFFF0ull) { err = "more than 2^32 entries"; return false; }
    bool negRow = false, negCol = false;
    compressTriplets(rows, cols, vals, nrow, pair.rowPtr, pair.rowIdx, pair.rowVal, negRow);
    compressTriplets(cols, rows, vals, ncol, pair.colPtr, pair.colIdx, pair.colVal, negCol);
    return !(negRow || negCol);
}

/* Host only: the compressed rows (byRows != 0) or compressed columns of a Matrix-Market file as the sparse model's
 * loader builds them.  ptr has nMajor + 1 entries; idx / val are written when non-NULL and large enough. */
static int cgb_read_matrix_csr_body(const char *path, int32_t byRows, uint32_t *nrow, uint32_t *ncol, uint32_t *ptr,
                                   uint64_t ptrCapacity, uint32_t *idx, float *val, uint64_t capacity, uint64_t *nnz)
{
    CGB_CHECK(path && nrow && ncol && nnz, "cgb_read_matrix_csr: NULL argument");
    CsrPair pair;
    std::string err;
    if (!loadMtxCsrPair(path, pair, *nrow, *ncol, err))
    {
        return fail(err.empty() ? CGB_EUNSUPPORTED : CGB_EINVAL, std::string("cgb_read_matrix_csr: ") + (err.empty() ? "negative entries: use the dense loader" : err));
    }
    const std::vector<uint32_t> &p = byRows ? pair.rowPtr : pair.colPtr, &i = byRows ? pair.rowIdx : pair.colIdx;
    const std::vector<float> &v = byRows ? pair.rowVal : pair.colVal;
    *nnz = v.size();
    if (ptr)
    {
        CGB_CHECK(ptrCapacity >= p.size(), "cgb_read_matrix_csr: ptr buffer too small");
        std::memcpy(ptr, p.data(), p.size() * sizeof(uint32_t));
    }
    if (idx && val)
    {
        CGB_CHECK(capacity >= v.size(), "cgb_read_matrix_csr: idx / val buffers too small");
        if (!v.empty())
        {
            std::memcpy(idx, i.data(), i.size() * sizeof(uint32_t));
            std::memcpy(val, v.data(), v.size() * sizeof(float));
        }
    }
    return CGB_OK;
}

extern "C" int cgb_read_matrix_csr(const char *path, int32_t byRows, uint32_t *nrow, uint32_t *ncol, uint32_t *ptr,
                                   uint64_t ptrCapacity, uint32_t *idx, float *val, uint64_t capacity, uint64_t *nnz)
{
    return guarded("cgb_read_matrix_csr", [&]() { return cgb_read_matrix_csr_body(path, byRows, nrow, ncol, ptr, ptrCapacity, idx, val, capacity, nnz); });
}

static int cgb_read_matrix_file_body(const char *path, float *out, uint64_t capacity, uint32_t *nrow, uint32_t *ncol)
{
    CGB_CHECK(path && nrow && ncol, "cgb_read_matrix_file: NULL argument");
    std::vector<float> m;
    std::string err;
    if (!loadMatrixFile(path, m, *nrow, *ncol, err)) { return fail(CGB_EINVAL, std::string("cgb_read_matrix_file: ") + err); }
    if (out != nullptr)
    {
        CGB_CHECK(capacity >= m.size(), "cgb_read_matrix_file: output buffer too small");
        std::memcpy(out, m.data(), m.size() * sizeof(float));
    }
    return CGB_OK;
}

extern "C" int cgb_read_matrix_file(const char *path, float *out, uint64_t capacity, uint32_t *nrow, uint32_t *ncol)
{
    return guarded("cgb_read_matrix_file", [&]() { return cgb_read_matrix_file_body(path, out, capacity, nrow, ncol); });
}

static int cgb_run_file_ex_body(const char *dataPath, const char *uncertaintyPath, const cgb_params *p, const cgb_run_options *opt, cgb_result *r)
{
    CGB_CHECK(dataPath && p && r, "cgb_run_file: NULL argument");
    CGB_CHECK(p->struct_size == sizeof(cgb_params), "cgb_run_file: cgb_params ABI mismatch");
    std::vector<float> data, unc;
    uint32_t nrow = 0, ncol = 0, urow = 0, ucol = 0;
    std::string err;
    // The sparse model on a whole Matrix-Market file: triplets go straight to the compressed rows both samplers keep
    // (SparseMatrix(path, ...), data_structures/SparseMatrix.cpp:52-107, never forms a dense matrix either); the dense
    // copy the chi-square kernels read is rebuilt on the device.  COGAPS_MTX_DENSE=1 forces the dense route (tests).
    const bool uncGiven = uncertaintyPath != nullptr && uncertaintyPath[0] != 0; // then the dense route validates it as before
    // a resumed run takes useSparseOptimization from the checkpoint (run_helper, GapsRunner.cpp:99-105): that flag, not
    // the caller's, picks the route
    bool sparseModel = p->useSparseOptimization != 0;
    if (opt && opt->checkpointInFile && opt->checkpointInFile[0])
    {
        ParamsImage pi;
        uint64_t seederState[2];
        std::string herr;
        if (readCheckpointHeader(opt->checkpointInFile, pi, seederState, herr)) { sparseModel = pi.useSparseOptimization != 0; }
    }
    if (sparseModel && p->nSubsetIndices == 0 && !uncGiven && endsWith(dataPath, ".mtx") && envInt("COGAPS_MTX_DENSE", 0) == 0)
    {
        CsrPair pair;
        if (loadMtxCsrPair(dataPath, pair, nrow, ncol, err))
        {
            return runCore(nullptr, &pair, nrow, ncol, 0, nullptr, p, opt, r); // uncertainty is ignored by the sparse model
        }
        if (!err.empty()) { return fail(CGB_EINVAL, std::string("cgb_run_file: ") + err); }
        // negative entries: fall through to the dense route, which carries them into lambda like the reference
    }
    if (!loadMatrixFile(dataPath, data, nrow, ncol, err)) { return fail(CGB_EINVAL, std::string("cgb_run_file: ") + err); }
    const bool haveUnc = uncertaintyPath != nullptr && uncertaintyPath[0] != 0;
    if (haveUnc)
    {
        if (!loadMatrixFile(uncertaintyPath, unc, urow, ucol, err)) { return fail(CGB_EINVAL, std::string("cgb_run_file: ") + err); }
        CGB_CHECK(urow == nrow && ucol == ncol, "cgb_run_file: uncertainty and data differ in shape");
    }
    // a subset read from a file keeps the rows in increasing index order whatever the order of the indices
    // (Matrix.cpp:113-131 sorts them; the in-memory constructor does not, :30-69)
    cgb_params q = *p;
    std::vector<uint32_t> sorted;
    if (p->nSubsetIndices > 0 && p->subsetIndices != nullptr)
    {
        sorted.assign(p->subsetIndices, p->subsetIndices + p->nSubsetIndices);
        std::sort(sorted.begin(), sorted.end());
        q.subsetIndices = sorted.data();
    }
    return cgb_run_ex(data.data(), nrow, ncol, 0, haveUnc ? unc.data() : nullptr, &q, opt, r);
}

extern "C" int cgb_run_file_ex(const char *dataPath, const char *uncertaintyPath, const cgb_params *p, const cgb_run_options *opt, cgb_result *r)
{
    return guarded("cgb_run_file_ex", [&]() { return cgb_run_file_ex_body(dataPath, uncertaintyPath, p, opt, r); });
}

extern "C" int cgb_run_file(const char *dataPath, const char *uncertaintyPath, const cgb_params *p, cgb_result *r)
{
    return guarded("cgb_run_file", [&]() { return cgb_run_file_ex_body(dataPath, uncertaintyPath, p, nullptr, r); });
}
