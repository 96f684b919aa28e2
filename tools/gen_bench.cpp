// gen_bench.cpp — host-only microbenchmark of the proposal generator + atomic domain at the bench
// workload's shape (no GPU needed): outcomes are drawn at random with the acceptance rates the real
// chain shows, so the domain reaches and keeps a steady size.
//   nvcc -O2 -std=c++17 -Xcompiler -O2,-ffp-contract=off -o /tmp/gen_bench tools/gen_bench.cpp cogaps_b200/csrc/proposal_queue.cpp
#include "../cogaps_b200/csrc/sampler.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>

using namespace cgb;

int main(int argc, char **argv)
{
    const uint32_t nRows = argc > 1 ? std::atoi(argv[1]) : 20000, k = argc > 2 ? std::atoi(argv[2]) : 20;
    const uint64_t target = argc > 3 ? std::atoll(argv[3]) : 82000;
    cgb_randstate rs(42);
    AtomicDomain domain;
    ProposalQueue queue;
    domain.init(static_cast<uint64_t>(nRows) * k);
    const float alpha = argc > 4 ? std::atof(argv[4]) : 0.0228f; // alpha * nBins / (1 - accept imbalance) sets the steady size
    queue.init(static_cast<uint64_t>(nRows) * k, k, &rs, alpha, 0.05f);
    HostRng orng(rs.seeder);
    uint64_t total = 0, batches = 0, queued = 0;
    double tGen = 0, tApply = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    for (int phase = 0; phase < 2; ++phase)
    {
        // phase 0 grows the domain (births mostly accepted), phase 1 measures at steady state
        const uint64_t steps = phase == 0 ? 40 * target : 4000000;
        uint64_t done = 0;
        total = batches = queued = 0;
        tGen = tApply = 0;
        while (done < steps)
        {
            const double t0 = now();
            queue.populate(domain, 1000000);
            done += queue.nProcessed();
            total += queue.nProcessed();
            const double t1 = now();
            std::vector<HostProposal> &q = queue.entries();
            for (size_t i = 0; i < q.size(); ++i)
            {
                const HostProposal &hp = q[i];
                const float u = orng.uniform();
                const bool grow = domain.size() < target;
                switch (hp.type)
                {
                    case 'B':
                        if (u < (grow ? 0.9f : 0.5f)) { queue.acceptBirth(); domain.atom(hp.atom1).mass = 0.5f + u; }
                        else { queue.rejectBirth(); domain.cacheErase(hp.atom1); }
                        break;
                    case 'D':
                        if (u < (grow ? 0.9f : 0.5f)) { queue.rejectDeath(); domain.atom(hp.atom1).mass = 0.3f + u; }
                        else { queue.acceptDeath(); domain.cacheErase(hp.atom1); }
                        break;
                    case 'M':
                        if (u < 0.3f) { domain.move(hp.atom1, hp.pos); }
                        break;
                    default:
                        if (u < 0.5f) { domain.atom(hp.atom1).mass += 0.01f; }
                        break;
                }
            }
            queued += q.size();
            ++batches;
            queue.clear();
            domain.flushEraseCache();
            tGen += t1 - t0;
            tApply += now() - t1;
        }
        std::printf("phase %d: atoms %llu, %llu proposals (%llu queued) in %llu batches (%.1f per batch); generate %.1f ns/proposal, "
                    "apply+flush %.1f ns/proposal\n", phase, (unsigned long long)domain.size(), (unsigned long long)total,
                    (unsigned long long)queued, (unsigned long long)batches, double(queued) / batches, tGen / total * 1e9, tApply / total * 1e9);
    }
    if (!domain.checkInvariants()) { std::printf("INVARIANTS BROKEN\n"); return 1; }
    return 0;
}
