// proposal_queue.cpp — sequential generation of the longest conflict-free prefix of the proposal
// stream.  Semantics follow atomic/ProposalQueue.cpp:19-283 exactly (draw order, seed roll-back on a
// failed proposal, cached (u1,u2) across a failure, min/max atom-count window, row / atom / move-
// interval conflicts); the bookkeeping underneath is ours (epoch stamps instead of linear-scan sets,
// the bin-indexed AtomicDomain instead of std::map).
#include "sampler.h"


namespace cgb {

void ProposalQueue::init(uint64_t nElements, uint64_t nPatterns, cgb_randstate *rs, float alpha, float lambda)
{
    mQueue.clear();
    mUsedRows.assign(static_cast<size_t>(nElements / nPatterns), 0u);
    mMoveLo.clear();
    mMoveHi.clear();
    mEpoch = 1;
    mRandState = rs;
    mRng = HostRng(rs->seeder);                       // ProposalQueue.cpp:24 — one seed at construction
    mMinAtoms = mMaxAtoms = 0;
    mBinLength = 0xFFFFFFFFFFFFFFFFull / nElements;
    mNumCols = nPatterns;
    mBinDiv.init(mBinLength);
    mColDiv.init(mNumCols);
    mBirthIPart = 0xFFFFFFFFFFFFFFFFull / (mBinLength * nElements);
    mAlpha = static_cast<double>(alpha);              // setAlpha, :39-42
    mDomainLength = static_cast<double>(mBinLength * nElements);
    mNumBins = static_cast<double>(nElements);
    mLambda = lambda;                                 // setLambda, :44-47
    mU1 = mU2 = 0.f;
    mNumProcessed = 0;
    mUseCachedRng = false;
}

void ProposalQueue::save(QueueState &out) const
{
    out.rng = mRng.state;
    out.minAtoms = mMinAtoms;
    out.maxAtoms = mMaxAtoms;
    out.binLength = mBinLength;
    out.numCols = mNumCols;
    out.alpha = mAlpha;
    out.domainLength = mDomainLength;
    out.numBins = mNumBins;
    out.lambda = mLambda;
    out.useCachedRng = mUseCachedRng;
    out.u1 = mU1;
    out.u2 = mU2;
}

bool ProposalQueue::restore(const QueueState &in)
{
    if (in.binLength != mBinLength || in.numCols != mNumCols || in.numBins != mNumBins || in.domainLength != mDomainLength)
    {
        return false;
    }
    mRng.state = in.rng;
    mMinAtoms = in.minAtoms;
    mMaxAtoms = in.maxAtoms;
    mAlpha = in.alpha;
    mLambda = in.lambda;
    mUseCachedRng = in.useCachedRng;
    mU1 = in.u1;
    mU2 = in.u2;
    return true;
}

void ProposalQueue::populate(AtomicDomain &domain, unsigned limit, SinkFn sink, void *sinkCtx)
{
    bool success = true;
    mNumProcessed = 0;
    while (mNumProcessed < limit && success)
    {
        const size_t before = mQueue.size();
        if (!makeProposal(domain))
        {
            success = false;
            mUseCachedRng = true;
        }
        else
        {
            ++mNumProcessed;
            // same-bin moves / exchanges succeed without queueing anything
            if (sink != nullptr && mQueue.size() > before) { sink(sinkCtx, mQueue.back(), before); }
        }
    }
}

void ProposalQueue::clear()
{
    mQueue.clear();
    mMoveLo.clear();
    mMoveHi.clear();
    if (++mEpoch == 0) // stamp wrapped: start over with clean tables
    {
        std::fill(mUsedRows.begin(), mUsedRows.end(), 0u);
        mEpoch = 1;
    }
}

float ProposalQueue::deathProb(double nAtoms) const
{
    double numer = nAtoms * mDomainLength;
    return static_cast<float>(numer / (numer + mAlpha * mNumBins * (mDomainLength - nAtoms)));
}

bool ProposalQueue::moveOverlap(uint64_t pos) const
{
    const size_t n = mMoveLo.size();
    for (size_t i = 0; i < n; ++i)
    {
        if (mMoveLo[i] < pos && pos < mMoveHi[i]) { return true; }
    }
    return false;
}

bool ProposalQueue::makeProposal(AtomicDomain &domain)
{
    mU1 = mUseCachedRng ? mU1 : mRng.uniform();
    mU2 = mUseCachedRng ? mU2 : mRng.uniform();
    mUseCachedRng = false;

    if (mMinAtoms < 2 && mMaxAtoms >= 2) { return false; } // indeterminate
    if (mMaxAtoms < 2) { return birth(domain); }

    if (mU1 < 0.5f)
    {
        const float lowerBound = deathProb(static_cast<double>(mMinAtoms));
        if (mU2 < lowerBound) { return death(domain); }
        const float upperBound = deathProb(static_cast<double>(mMaxAtoms));
        if (mU2 >= upperBound) { return birth(domain); }
        return false; // birth/death undecidable until the batch resolves
    }
    return (mU1 < 0.75f) ? move(domain) : exchange(domain);
}

bool ProposalQueue::birth(AtomicDomain &domain)
{
    HostProposal prop;
    prop.rng = HostRng(mRandState->seeder);
    prop.type = 'B';
    prop.atom2 = kNoAtom;
    prop.r2 = prop.c2 = 0;
    // randomFreePosition, ConcurrentAtomicDomain.cpp:53-60
    uint64_t pos = prop.rng.uniform64Fixed(1, domain.domainLength(), mBirthIPart);
    while (domain.occupied(pos)) { pos = prop.rng.uniform64Fixed(1, domain.domainLength(), mBirthIPart); }
    prop.pos = 0;

    if (moveOverlap(pos))
    {
        mRandState->seeder.rollBackOnce();
        return false;
    }
    binOf(pos, prop.r1, prop.c1);
    if (rowUsed(prop.r1))
    {
        mRandState->seeder.rollBackOnce();
        return false;
    }
    prop.atom1 = domain.insert(pos, 0.f);
    useRow(prop.r1);
    domain.atom(prop.atom1).usedEpoch = mEpoch;
    mQueue.push_back(prop);
    ++mMaxAtoms;
    return true;
}

bool ProposalQueue::death(AtomicDomain &domain)
{
    HostProposal prop;
    prop.rng = HostRng(mRandState->seeder);
    prop.type = 'D';
    prop.atom2 = kNoAtom;
    prop.r2 = prop.c2 = 0;
    prop.pos = 0;
    prop.atom1 = domain.atIndex(prop.rng.uniform32(0, static_cast<uint32_t>(domain.size() - 1)));
    const uint64_t p1 = domain.atom(prop.atom1).pos;
    binOf(p1, prop.r1, prop.c1);
    if (rowUsed(prop.r1))
    {
        mRandState->seeder.rollBackOnce();
        return false;
    }
    useRow(prop.r1);
    domain.atom(prop.atom1).usedEpoch = mEpoch;
    mQueue.push_back(prop);
    --mMinAtoms;
    return true;
}

bool ProposalQueue::move(AtomicDomain &domain)
{
    HostProposal prop;
    prop.rng = HostRng(mRandState->seeder);
    prop.type = 'M';
    prop.atom2 = kNoAtom;
    prop.atom1 = domain.atIndex(prop.rng.uniform32(0, static_cast<uint32_t>(domain.size() - 1)));
    const Atom &center = domain.atom(prop.atom1);
    const uint32_t left = center.left, right = center.right;
    const uint64_t lbound = (left != kNoAtom) ? domain.atom(left).pos : 0;
    const uint64_t rbound = (right != kNoAtom) ? domain.atom(right).pos : referenceDoubleToU64(mDomainLength);

    // mUsedAtoms.contains(lbound) || contains(rbound): positions of atoms held by this batch are
    // unique and frozen while it is generated, so membership by position == membership by atom
    if ((left != kNoAtom && domain.atom(left).usedEpoch == mEpoch)
    || (right != kNoAtom && domain.atom(right).usedEpoch == mEpoch))
    {
        mRandState->seeder.rollBackOnce();
        return false;
    }

    prop.pos = prop.rng.uniform64(lbound + 1, rbound - 1);
    const uint64_t p1 = center.pos;
    binOf(p1, prop.r1, prop.c1);
    binOf(prop.pos, prop.r2, prop.c2);

    if (rowUsed(prop.r1) || rowUsed(prop.r2))
    {
        mRandState->seeder.rollBackOnce();
        return false;
    }
    if (prop.r1 == prop.r2 && prop.c1 == prop.c2)
    {
        domain.move(prop.atom1, prop.pos); // same bin: accepted on the spot, never queued
        return true;
    }
    mQueue.push_back(prop);
    useRow(prop.r1);
    useRow(prop.r2);
    domain.atom(prop.atom1).usedEpoch = mEpoch;
    mMoveLo.push_back(p1 < prop.pos ? p1 : prop.pos);
    mMoveHi.push_back(p1 < prop.pos ? prop.pos : p1);
    return true;
}

bool ProposalQueue::exchange(AtomicDomain &domain)
{
    HostProposal prop;
    prop.rng = HostRng(mRandState->seeder);
    prop.type = 'E';
    prop.pos = 0;
    prop.atom1 = domain.atIndex(prop.rng.uniform32(0, static_cast<uint32_t>(domain.size() - 1)));
    const uint32_t right = domain.atom(prop.atom1).right;
    prop.atom2 = (right != kNoAtom) ? right : domain.front();
    const uint64_t p1 = domain.atom(prop.atom1).pos, p2 = domain.atom(prop.atom2).pos;
    binOf(p1, prop.r1, prop.c1);
    binOf(p2, prop.r2, prop.c2);

    if (rowUsed(prop.r1) || rowUsed(prop.r2))
    {
        mRandState->seeder.rollBackOnce();
        return false;
    }
    if (prop.r1 == prop.r2 && prop.c1 == prop.c2)
    {
        // same bin: resample the split of the combined mass right here (ProposalQueue.cpp:266-276)
        Atom &a1 = domain.atom(prop.atom1);
        Atom &a2 = domain.atom(prop.atom2);
        const float newMass = prop.rng.truncGammaUpper(mRandState->tables.qgamma, a1.mass + a2.mass, 1.f / mLambda);
        const float delta = (a1.mass > a2.mass) ? newMass - a1.mass : a2.mass - newMass;
        if (a1.mass + delta > kEpsilon && a2.mass - delta > kEpsilon)
        {
            const float m1 = a1.mass + delta, m2 = a2.mass - delta;
            a1.mass = m1;
            a2.mass = m2;
        }
        return true;
    }
    mQueue.push_back(prop);
    useRow(prop.r1);
    useRow(prop.r2);
    return true;
}

} // namespace cgb
