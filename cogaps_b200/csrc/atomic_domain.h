// atomic_domain.h — the atomic domain of one factor matrix, host side.
//
// Same observable behaviour as the reference's ConcurrentAtomicDomain
// (atomic/ConcurrentAtomicDomain.cpp:14-132): atoms ordered by u64 position, an insertion-ordered
// vector with swap-with-last erase for uniform picks (randomAtom must return the same atom for the
// same draw), left/right neighbours, deferred erases flushed in position order.
//
// Different structure: the reference walks a red-black tree (std::map) for every insert, erase and
// occupancy test; on the GPU path proposal generation is the serial term, so here the ordered view
// is a bin-indexed table.  A position's matrix bin is pos / binLength (ProposalQueue.cpp:174-175), atoms
// are uniform over bins and there is ~one atom per two bins, so "first atom of each bin" plus a
// hierarchical bitmap of non-empty bins finds a predecessor in O(1) expected with two or three cache
// lines touched, and neighbours are a doubly linked list threaded through a flat atom pool.
#ifndef CGB_ATOMIC_DOMAIN_H
#define CGB_ATOMIC_DOMAIN_H

#include <stdint.h>
#include <algorithm>
#include <vector>

namespace cgb {

static const uint32_t kNoAtom = 0xFFFFFFFFu;

// floor(x / d) for a divisor fixed per sampler (bin length, pattern count): one 64x64->128 multiply and a
// fix-up instead of a hardware divide; exact for every x
struct FastDivU64
{
    uint64_t d, m;
    void init(uint64_t divisor)
    {
        d = divisor;
        m = (divisor > 1) ? static_cast<uint64_t>((static_cast<unsigned __int128>(1) << 64) / divisor) : 0;
    }
    uint64_t div(uint64_t x) const
    {
        if (d == 1) { return x; }
        uint64_t q = static_cast<uint64_t>((static_cast<unsigned __int128>(x) * m) >> 64); // floor(x/d) - {0,1}
        uint64_t r = x - q * d;
        while (r >= d)
        {
            ++q;
            r -= d;
        }
        return q;
    }
};

struct Atom
{
    uint64_t pos;
    float mass;
    uint32_t left;      // pool index of the neighbour with the next smaller position
    uint32_t right;
    uint32_t vecIndex;  // ConcurrentAtom::mIndex — slot in the pick vector
    uint32_t usedEpoch; // stamp of the proposal batch that holds this atom (SmallHashSetU64 of the reference)
    uint32_t pad;
};

// set bits = non-empty bins; levels of 64-way summaries so "next set bit after b" is a few ctz's
class BinBitmap
{
public:
    void init(uint64_t nBits)
    {
        mLevels.clear();
        uint64_t n = nBits;
        do
        {
            n = (n + 63) / 64;
            mLevels.push_back(std::vector<uint64_t>(n, 0));
        } while (n > 1);
    }
    void set(uint64_t b)
    {
        for (size_t l = 0; l < mLevels.size(); ++l)
        {
            uint64_t &w = mLevels[l][b >> 6];
            const bool wasZero = (w == 0);
            w |= (1ull << (b & 63));
            if (!wasZero) { break; }
            b >>= 6;
        }
    }
    void clear(uint64_t b)
    {
        for (size_t l = 0; l < mLevels.size(); ++l)
        {
            uint64_t &w = mLevels[l][b >> 6];
            w &= ~(1ull << (b & 63));
            if (w != 0) { break; }
            b >>= 6;
        }
    }
    // smallest set bit strictly greater than b, or UINT64_MAX
    uint64_t nextAfter(uint64_t b) const
    {
        size_t l = 0;
        uint64_t idx = b;
        // climb until a word has a set bit above idx
        for (;;)
        {
            const uint64_t word = mLevels[l][idx >> 6];
            const unsigned bit = static_cast<unsigned>(idx & 63);
            const uint64_t above = (bit == 63) ? 0 : (word & (~0ull << (bit + 1)));
            if (above != 0)
            {
                idx = (idx & ~63ull) | static_cast<uint64_t>(__builtin_ctzll(above));
                break;
            }
            idx >>= 6;
            ++l;
            if (l == mLevels.size()) { return ~0ull; }
        }
        // descend to level 0 taking the lowest set bit each time
        while (l > 0)
        {
            --l;
            const uint64_t word = mLevels[l][idx];
            idx = (idx << 6) | static_cast<uint64_t>(__builtin_ctzll(word));
        }
        return idx;
    }
private:
    std::vector<std::vector<uint64_t> > mLevels;
};

class AtomicDomain
{
public:
    void init(uint64_t nBins)
    {
        mNumBins = nBins;
        mBinLength = 0xFFFFFFFFFFFFFFFFull / nBins;
        mBinDiv.init(mBinLength);
        mDomainLength = mBinLength * nBins; // ConcurrentAtomicDomain.cpp:14-18
        mBinFirst.assign(nBins, kNoAtom);
        mBitmap.init(nBins);
        mPool.clear();
        mFree.clear();
        mVec.clear();
        mEraseCache.clear();
        mHead = mTail = kNoAtom;
    }

    uint64_t size() const { return mVec.size(); }
    uint64_t domainLength() const { return mDomainLength; }
    uint64_t binLength() const { return mBinLength; }
    Atom &atom(uint32_t id) { return mPool[id]; }
    const Atom &atom(uint32_t id) const { return mPool[id]; }
    uint32_t front() const { return mHead; }               // ConcurrentAtomicDomain.cpp:20-30
    uint32_t atIndex(uint32_t i) const { return mVec[i]; } // mAtoms[index], :34-47
    uint64_t binOf(uint64_t pos) const
    {
        const uint64_t b = mBinDiv.div(pos);
        return b < mNumBins ? b : mNumBins - 1; // pos == domainLength would index one past the end
    }

    bool occupied(uint64_t pos) const // mAtomMap.count(pos), :53-60
    {
        const uint64_t b = binOf(pos);
        uint32_t a = mBinFirst[b];
        while (a != kNoAtom && mPool[a].pos <= pos)
        {
            if (mPool[a].pos == pos) { return true; }
            a = mPool[a].right;
        }
        return false;
    }

    // ConcurrentAtomicDomain::insert, :82-105
    uint32_t insert(uint64_t pos, float mass)
    {
        uint32_t id;
        if (!mFree.empty())
        {
            id = mFree.back();
            mFree.pop_back();
        }
        else
        {
            id = static_cast<uint32_t>(mPool.size());
            mPool.push_back(Atom());
        }
        const uint64_t b = binOf(pos);
        uint32_t pred, succ;
        const uint32_t first = mBinFirst[b];
        if (first == kNoAtom)
        {
            const uint64_t nb = mBitmap.nextAfter(b);
            succ = (nb == ~0ull) ? kNoAtom : mBinFirst[nb];
            pred = (succ == kNoAtom) ? mTail : mPool[succ].left;
            mBinFirst[b] = id;
            mBitmap.set(b);
        }
        else if (pos < mPool[first].pos)
        {
            succ = first;
            pred = mPool[first].left;
            mBinFirst[b] = id;
        }
        else
        {
            pred = first;
            succ = mPool[first].right;
            while (succ != kNoAtom && mPool[succ].pos < pos)
            {
                pred = succ;
                succ = mPool[succ].right;
            }
        }
        Atom &a = mPool[id];
        a.pos = pos;
        a.mass = mass;
        a.left = pred;
        a.right = succ;
        a.vecIndex = static_cast<uint32_t>(mVec.size());
        a.usedEpoch = 0;
        mVec.push_back(id);
        if (pred != kNoAtom) { mPool[pred].right = id; } else { mHead = id; }
        if (succ != kNoAtom) { mPool[succ].left = id; } else { mTail = id; }
        return id;
    }

    // ConcurrentAtomicDomain::erase, :108-123
    void erase(uint32_t id)
    {
        Atom &a = mPool[id];
        unlinkFromBin(id, binOf(a.pos));
        if (a.left != kNoAtom) { mPool[a.left].right = a.right; } else { mHead = a.right; }
        if (a.right != kNoAtom) { mPool[a.right].left = a.left; } else { mTail = a.left; }
        const uint32_t vi = a.vecIndex;
        mVec[vi] = mVec.back();
        mPool[mVec[vi]].vecIndex = vi;
        mVec.pop_back();
        mFree.push_back(id);
    }

    // ConcurrentAtomicDomain::move, :126-132 — the caller guarantees no neighbour is crossed
    void move(uint32_t id, uint64_t newPos)
    {
        Atom &a = mPool[id];
        const uint64_t ob = binOf(a.pos), nb = binOf(newPos);
        if (ob != nb)
        {
            unlinkFromBin(id, ob);
            const uint32_t first = mBinFirst[nb];
            if (first == kNoAtom)
            {
                mBinFirst[nb] = id;
                mBitmap.set(nb);
            }
            else if (newPos < mPool[first].pos)
            {
                mBinFirst[nb] = id;
            }
        }
        a.pos = newPos;
    }

    void cacheErase(uint32_t id) { mEraseCache.push_back(id); } // :62-69

    // flushEraseCache, :71-79: erase in increasing position order (fixes the pick-vector permutation)
    void flushEraseCache()
    {
        if (mEraseCache.empty()) { return; }
        if (mEraseCache.size() > 1)
        {
            const std::vector<Atom> &pool = mPool;
            std::sort(mEraseCache.begin(), mEraseCache.end(),
                [&pool](uint32_t x, uint32_t y) { return pool[x].pos < pool[y].pos; });
        }
        for (size_t i = 0; i < mEraseCache.size(); ++i) { erase(mEraseCache[i]); }
        mEraseCache.clear();
    }

    // debug invariant of the reference (isSorted, :152-167) extended to the bin table
    bool checkInvariants() const
    {
        uint64_t n = 0;
        uint32_t a = mHead, prev = kNoAtom;
        while (a != kNoAtom)
        {
            if (mPool[a].left != prev) { return false; }
            if (prev != kNoAtom && !(mPool[prev].pos < mPool[a].pos)) { return false; }
            const uint64_t b = binOf(mPool[a].pos);
            const bool firstInBin = (prev == kNoAtom) || binOf(mPool[prev].pos) != b;
            if (firstInBin && mBinFirst[b] != a) { return false; }
            if (mVec[mPool[a].vecIndex] != a) { return false; }
            prev = a;
            a = mPool[a].right;
            ++n;
        }
        return prev == mTail && n == mVec.size();
    }

private:
    void unlinkFromBin(uint32_t id, uint64_t b)
    {
        if (mBinFirst[b] != id) { return; }
        const uint32_t r = mPool[id].right;
        if (r != kNoAtom && binOf(mPool[r].pos) == b)
        {
            mBinFirst[b] = r;
        }
        else
        {
            mBinFirst[b] = kNoAtom;
            mBitmap.clear(b);
        }
    }

    uint64_t mNumBins, mBinLength, mDomainLength;
    FastDivU64 mBinDiv;
    std::vector<uint32_t> mBinFirst;
    BinBitmap mBitmap;
    std::vector<Atom> mPool;
    std::vector<uint32_t> mFree;
    std::vector<uint32_t> mVec;
    std::vector<uint32_t> mEraseCache;
    uint32_t mHead, mTail;
};

} // namespace cgb

#endif // CGB_ATOMIC_DOMAIN_H
