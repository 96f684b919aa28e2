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

#include <algorithm>
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

// ------------------------------------------------------------------------------------------------
// Test hook, host only: the bin-indexed atomic domain (atomic_domain.h) against the reference's own data structures
// restated naively — std::map<position, atom> for order and neighbours, std::vector for the pick order with
// swap-with-last erases (ConcurrentAtomicDomain.cpp:14-132) — under a random workload of inserts, batched erases
// (cacheErase + flushEraseCache) and in-gap moves, including positions in the last bin and at domainLength itself.
// ------------------------------------------------------------------------------------------------
#include <map>

extern "C" int cgb_debug_domain_fuzz(uint64_t seed, uint64_t nBins, uint32_t nOps, uint64_t *opsDone)
{
    g_replayMessage.clear();
    if (!opsDone || nBins == 0) { g_replayMessage = "bad argument"; return CGB_EINVAL; }
    try
    {
        AtomicDomain dom;
        dom.init(nBins);
        Xoroshiro128plus seeder(seed);
        HostRng rng(seeder);
        std::map<uint64_t, uint32_t> byPos;      // position -> atom id in `dom`
        std::vector<uint64_t> pick;              // the reference's mAtoms, as positions
        std::map<uint64_t, size_t> pickIndex;    // position -> index in `pick` (the model's ConcurrentAtom::mIndex)
        const uint32_t growUntil = nOps / 8;     // inserts dominate until the domain holds this many atoms
        const uint64_t len = dom.domainLength();
        for (uint32_t op = 0; op < nOps; ++op)
        {
            *opsDone = op;
            const uint32_t kind = rng.uniform32(0, 9);
            if (kind < (pick.size() < growUntil ? 7u : 4u) || pick.size() < 4)
            {
                // insert at a free position; now and then in the last bin or at the very end of the domain
                uint64_t pos = rng.uniform64(1, len);
                if (kind == 0) { pos = len - rng.uniform64(0, 3); }
                if (byPos.count(pos)) { if (!dom.occupied(pos)) { g_replayMessage = "occupied() misses an atom"; return CGB_EINTERNAL; } continue; }
                if (dom.occupied(pos)) { g_replayMessage = "occupied() reports a free position"; return CGB_EINTERNAL; }
                const uint32_t id = dom.insert(pos, static_cast<float>(op));
                byPos[pos] = id;
                pickIndex[pos] = pick.size();
                pick.push_back(pos);
            }
            else if (kind < 8)
            {
                // a batch of erases: cached in any order, flushed in increasing position order, swap-with-last each
                const uint32_t n = rng.uniform32(1, 3);
                std::vector<uint64_t> victims;
                for (uint32_t i = 0; i < n && victims.size() < pick.size(); ++i)
                {
                    const uint64_t pos = pick[rng.uniform32(0, static_cast<uint32_t>(pick.size() - 1))];
                    bool dup = false;
                    for (size_t j = 0; j < victims.size(); ++j) { dup = dup || victims[j] == pos; }
                    if (!dup) { victims.push_back(pos); dom.cacheErase(byPos[pos]); }
                }
                dom.flushEraseCache();
                std::sort(victims.begin(), victims.end());
                for (size_t j = 0; j < victims.size(); ++j)
                {
                    const size_t at = pickIndex[victims[j]];
                    pick[at] = pick.back();
                    pickIndex[pick[at]] = at;
                    pick.pop_back();
                    pickIndex.erase(victims[j]);
                    byPos.erase(victims[j]);
                }
            }
            else
            {
                // move inside the gap between the neighbours (ProposalQueue.cpp:209-248)
                const uint32_t vi = rng.uniform32(0, static_cast<uint32_t>(pick.size() - 1));
                const uint64_t pos = pick[vi];
                std::map<uint64_t, uint32_t>::iterator it = byPos.find(pos), nx = it, pv = it;
                ++nx;
                const uint64_t hi = (nx != byPos.end()) ? nx->first : len + 1;
                const uint64_t lo = (it != byPos.begin()) ? (--pv)->first : 0;
                if (hi - lo < 2) { continue; }
                const uint64_t to = rng.uniform64(lo + 1, hi - 1);
                if (to == pos) { continue; }
                const uint32_t id = it->second;
                dom.move(id, to);
                byPos.erase(it);
                byPos[to] = id;
                pickIndex.erase(pos);
                pickIndex[to] = vi;
                pick[vi] = to;
            }
            // compare: size, pick order, sorted order with neighbours, first atom
            if (dom.size() != pick.size() || byPos.size() != pick.size()) { g_replayMessage = "size differs from the model"; return CGB_EINTERNAL; }
            if ((op & 1023u) == 0 || op + 1 == nOps || (pick.size() < 64 && (op & 15u) == 0))
            {
                if (!dom.checkInvariants()) { g_replayMessage = "checkInvariants failed"; return CGB_EINTERNAL; }
                for (size_t i = 0; i < pick.size(); ++i)
                {
                    if (dom.atom(dom.atIndex(static_cast<uint32_t>(i))).pos != pick[i]) { g_replayMessage = "pick order differs from the model"; return CGB_EINTERNAL; }
                }
                uint32_t a = dom.front();
                for (std::map<uint64_t, uint32_t>::const_iterator it = byPos.begin(); it != byPos.end(); ++it)
                {
                    if (a == kNoAtom || dom.atom(a).pos != it->first || a != it->second) { g_replayMessage = "sorted order differs from the model"; return CGB_EINTERNAL; }
                    a = dom.atom(a).right;
                }
                if (a != kNoAtom) { g_replayMessage = "the domain holds atoms the model does not"; return CGB_EINTERNAL; }
            }
        }
        *opsDone = nOps;
        return CGB_OK;
    }
    catch (const std::exception &e)
    {
        g_replayMessage = e.what();
        return CGB_ENOMEM;
    }
}
