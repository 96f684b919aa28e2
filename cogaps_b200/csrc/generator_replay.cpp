// generator_replay.cpp — test hook, host only: the library's own proposal generator driven through a whole run with
// the outcomes of every proposal taken from a recorded trace.
//
// The sequential half of the hot path — ProposalQueue::populate and the atomic domain (atomic/ProposalQueue.cpp:53-283,
// atomic/ConcurrentAtomicDomain.cpp:14-132) — decides which proposals exist at all; the device only evaluates them.
// With the evaluation replaced by a trace (the oracle's, which is pinned to the reference), the generator can be held
// to the reference WITHOUT a GPU: same run loop as runCore (seed consumption order GapsRunner.cpp:402-437, Poisson
// step counts :294-295, updateSampler :201-222), same batch loop as cgb_sampler_update, and after every populate()
// each queued proposal is compared field by field with the next trace record before its recorded outcome is applied
// through the very function the product applies outcomes with (applyToDomain, sampler.h).
#include "sampler.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace cgb;

namespace {

struct Side
{
    AtomicDomain domain;
    ProposalQueue queue;
    uint32_t nRows;
    char name;
};

// lambda as the model's constructor derives it (DenseNormalModel.h:75-77) from gaps::nonZeroMean over the sampler's
// orientation — through the same runningSum the samplers use (sampler.h), so the replay also holds that to the oracle:
// a same-bin exchange draws its new mass with scale 1 / lambda (ProposalQueue.cpp:266-276)
float lambdaOf(const float *data, uint32_t nrow, uint32_t ncol, bool rowsAreDataRows, float alpha, uint32_t k)
{
    float sum = 0.f;
    unsigned nnz = 0;
    if (rowsAreDataRows) { runningSum(data, nrow, ncol, ncol, 1, sum, nnz); }
    else { runningSum(data, ncol, nrow, 1, ncol, sum, nnz); }
    const float meanD = sum / static_cast<float>(nnz);
    return alpha * std::sqrt(static_cast<float>(static_cast<uint64_t>(k)) / meanD);
}

struct Replay
{
    const cgb_trace_record *trace;
    uint64_t n, next;
    std::string what;

    bool fail(const char *field, unsigned long long got, unsigned long long want)
    {
        char buf[256];
        std::snprintf(buf, sizeof(buf), "record %llu: %s is %llu, the trace has %llu", static_cast<unsigned long long>(next), field, got, want);
        what = buf;
        return false;
    }

    // one update() of one sampler: cgb_sampler_update's batch loop with the evaluation replaced by the trace
    bool update(Side &s, uint32_t nSteps, uint32_t phase, uint32_t iter)
    {
        uint32_t done = 0, batch = 0;
        while (done < nSteps)
        {
            s.queue.populate(s.domain, nSteps - done);
            done += s.queue.nProcessed();
            std::vector<HostProposal> &q = s.queue.entries();
            for (size_t i = 0; i < q.size(); ++i)
            {
                const HostProposal &hp = q[i];
                if (next >= n) { what = "the generator queued more proposals than the trace holds"; return false; }
                const cgb_trace_record &t = trace[next];
                if (t.phase != phase) { return fail("phase", phase, t.phase); }
                if (t.iter != iter) { return fail("iteration", iter, t.iter); }
                if (t.side != static_cast<uint32_t>(s.name)) { return fail("side", s.name, t.side); }
                if (t.batch != batch) { return fail("batch", batch, t.batch); }
                if (t.type != static_cast<uint32_t>(hp.type)) { return fail("type", hp.type, t.type); }
                if (t.r1 != hp.r1) { return fail("r1", hp.r1, t.r1); }
                if (t.c1 != hp.c1) { return fail("c1", hp.c1, t.c1); }
                const bool two = hp.type == 'M' || hp.type == 'E';
                if (two && t.r2 != hp.r2) { return fail("r2", hp.r2, t.r2); }
                if (two && t.c2 != hp.c2) { return fail("c2", hp.c2, t.c2); }
                if ((hp.type == 'B' || hp.type == 'M') && t.pos != hp.pos) { return fail("pos", hp.pos, t.pos); }
                if (t.rngState != hp.rng.state) { return fail("rng state", hp.rng.state, t.rngState); }
                const Atom &a1 = s.domain.atom(hp.atom1);
                if (t.atom1Pos != a1.pos) { return fail("atom1 position", a1.pos, t.atom1Pos); }
                uint32_t got, want;
                std::memcpy(&got, &a1.mass, 4);
                std::memcpy(&want, &t.mass1, 4);
                if (got != want) { return fail("atom1 mass bits", got, want); }
                if (hp.type == 'E')
                {
                    const Atom &a2 = s.domain.atom(hp.atom2);
                    if (t.atom2Pos != a2.pos) { return fail("atom2 position", a2.pos, t.atom2Pos); }
                    std::memcpy(&got, &a2.mass, 4);
                    std::memcpy(&want, &t.mass2, 4);
                    if (got != want) { return fail("atom2 mass bits", got, want); }
                }
                applyToDomain(s.domain, s.queue, hp, t.accepted != 0, t.newMass1, t.newMass2, a1.mass);
                ++next;
            }
            s.queue.clear();
            s.domain.flushEraseCache();
            ++batch;
        }
        if (s.queue.minAtoms() != s.queue.maxAtoms() || s.queue.maxAtoms() != s.domain.size())
        {
            what = "atom bookkeeping out of step with the domain at the end of an update";
            return false;
        }
        if (!s.domain.checkInvariants()) { what = "atomic domain invariants broken"; return false; }
        return true;
    }
};

} // namespace

static thread_local std::string g_replayMessage;

extern "C" const char *cgb_debug_replay_message(void) { return g_replayMessage.c_str(); }

extern "C" int cgb_debug_replay_generator(const float *data, uint32_t nrow, uint32_t ncol, const cgb_params *p,
                                          const cgb_trace_record *trace, uint64_t n, uint64_t *checked)
{
    g_replayMessage.clear();
    if (!data || !p || !trace || !checked || p->struct_size != sizeof(cgb_params)) { g_replayMessage = "bad argument"; return CGB_EINVAL; }
    if (p->transposeData || p->nSubsetIndices || !p->asynchronousUpdates)
    {
        g_replayMessage = "the replay covers the asynchronous sampler on the whole, untransposed matrix";
        return CGB_EUNSUPPORTED;
    }
    try
    {
        const int fixed = p->whichMatrixFixed ? p->whichMatrixFixed : 'N';
        cgb_randstate rs(p->seed);
        Side A, P;
        A.nRows = nrow; A.name = 'A';
        P.nRows = ncol; P.name = 'P';
        // A first, P second, then the run's own rng: the order the seeder is consumed in (GapsRunner.cpp:402-437)
        A.domain.init(static_cast<uint64_t>(A.nRows) * p->nPatterns);
        A.queue.init(static_cast<uint64_t>(A.nRows) * p->nPatterns, p->nPatterns, &rs, p->alphaA, lambdaOf(data, nrow, ncol, true, p->alphaA, p->nPatterns));
        P.domain.init(static_cast<uint64_t>(P.nRows) * p->nPatterns);
        P.queue.init(static_cast<uint64_t>(P.nRows) * p->nPatterns, p->nPatterns, &rs, p->alphaP, lambdaOf(data, nrow, ncol, false, p->alphaP, p->nPatterns));
        HostRng rng(rs.seeder);
        Replay r;
        r.trace = trace; r.n = n; r.next = 0;
        bool ok = true;
        for (uint32_t phase = CGB_PHASE_EQUILIBRATION; ok && phase <= CGB_PHASE_SAMPLING; ++phase)
        {
            for (uint32_t iter = 0; ok && iter < p->nIterations; ++iter)
            {
                const unsigned atomsA = static_cast<unsigned>(A.domain.size()), atomsP = static_cast<unsigned>(P.domain.size());
                const unsigned nA = static_cast<unsigned>(rng.poisson(static_cast<double>(atomsA < 10u ? 10u : atomsA)));
                const unsigned nP = static_cast<unsigned>(rng.poisson(static_cast<double>(atomsP < 10u ? 10u : atomsP)));
                if (fixed != 'A') { ok = r.update(A, nA, phase, iter); }
                if (ok && fixed != 'P') { ok = r.update(P, nP, phase, iter); }
            }
        }
        *checked = r.next;
        if (!ok)
        {
            g_replayMessage = r.what;
            return CGB_EINTERNAL;
        }
        if (r.next != n)
        {
            g_replayMessage = "the trace holds proposals the generator never queued";
            return CGB_EINTERNAL;
        }
        return CGB_OK;
    }
    catch (const std::exception &e)
    {
        g_replayMessage = e.what();
        return CGB_ENOMEM;
    }
}
