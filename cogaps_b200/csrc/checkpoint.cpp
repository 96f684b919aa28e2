// checkpoint.cpp — serialisation of checkpoint images in the reference's Archive wire format (see checkpoint.h).
#include "checkpoint.h"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>

namespace cgb {

void ByteReader::take(void *out, size_t n)
{
    if (!mOk || mSize - mOff < n)
    {
        mOk = false;
        std::memset(out, 0, n);
        return;
    }
    std::memcpy(out, mData + mOff, n);
    mOff += n;
}

// ---- GapsParameters (GapsParameters.cpp:82-96) ----
void putParams(ByteWriter &w, const ParamsImage &p)
{
    w.put<uint32_t>(p.seed);
    w.put<uint32_t>(p.nGenes);
    w.put<uint32_t>(p.nSamples);
    w.put<uint32_t>(p.nPatterns);
    w.put<uint32_t>(p.nIterations);
    w.put<float>(p.alphaA);
    w.put<float>(p.alphaP);
    w.put<float>(p.maxGibbsMassA);
    w.put<float>(p.maxGibbsMassP);
    w.put<uint8_t>(p.useSparseOptimization ? 1 : 0); // bool, one byte
    w.put<uint32_t>(p.checkpointInterval);
}

bool getParams(ByteReader &r, ParamsImage &p)
{
    p.seed = r.get<uint32_t>();
    p.nGenes = r.get<uint32_t>();
    p.nSamples = r.get<uint32_t>();
    p.nPatterns = r.get<uint32_t>();
    p.nIterations = r.get<uint32_t>();
    p.alphaA = r.get<float>();
    p.alphaP = r.get<float>();
    p.maxGibbsMassA = r.get<float>();
    p.maxGibbsMassP = r.get<float>();
    p.useSparseOptimization = r.get<uint8_t>() != 0;
    p.checkpointInterval = r.get<uint32_t>();
    return r.ok();
}

// ---- Matrix (Matrix.cpp:182-204): nRows nCols, every column as a Vector (Vector.cpp:90-111: size, floats) ----
static void putMatrix(ByteWriter &w, const std::vector<float> &colMajor, uint32_t rows, uint32_t cols)
{
    w.put<uint32_t>(rows);
    w.put<uint32_t>(cols);
    for (uint32_t c = 0; c < cols; ++c)
    {
        w.put<uint32_t>(rows);
        w.putFloats(colMajor.data() + static_cast<size_t>(c) * rows, rows);
    }
}

// a matrix as large as the file could possibly hold, so that a corrupt header cannot ask for terabytes
static bool plausible(const ByteReader &r, uint64_t rows, uint64_t cols)
{
    return rows > 0 && cols > 0 && cols <= (r.remaining() / sizeof(float)) / rows; // no product: it could wrap
}

static bool getMatrix(ByteReader &r, std::vector<float> &colMajor, uint32_t &rows, uint32_t &cols, std::string &err)
{
    rows = r.get<uint32_t>();
    cols = r.get<uint32_t>();
    if (!r.ok() || !plausible(r, rows, cols)) { err = "matrix header does not fit the file"; r.fail(); return false; }
    colMajor.assign(static_cast<size_t>(rows) * cols, 0.f);
    for (uint32_t c = 0; c < cols; ++c)
    {
        if (r.get<uint32_t>() != rows) { err = "column length differs from the matrix header"; r.fail(); return false; }
        r.getFloats(colMajor.data() + static_cast<size_t>(c) * rows, rows);
    }
    if (!r.ok()) { err = "file ends inside a matrix"; }
    return r.ok();
}

// ---- AsynchronousGibbsSampler (AsynchronousGibbsSampler.h:221-233) ----
void putSampler(ByteWriter &w, const SamplerImage &s)
{
    if (!s.sparse)
    {
        putMatrix(w, s.cols, s.nRows, s.k); // DenseNormalModel.cpp:260-264
    }
    else
    {
        // SparseNormalModel.cpp:313-317 -> HybridMatrix.cpp:85-97
        w.put<uint32_t>(s.nRows);
        w.put<uint32_t>(s.k);
        for (uint32_t r = 0; r < s.nRows; ++r)
        {
            w.put<uint32_t>(s.k);
            w.putFloats(s.rows.data() + static_cast<size_t>(r) * s.k, s.k);
        }
        // HybridVector.cpp:103-115: size, size/64+1 flag words (bit i set iff element i is not 0), the floats
        std::vector<uint64_t> flags(s.nRows / 64 + 1);
        for (uint32_t c = 0; c < s.k; ++c)
        {
            const float *col = s.cols.data() + static_cast<size_t>(c) * s.nRows;
            std::fill(flags.begin(), flags.end(), 0ull);
            for (uint32_t r = 0; r < s.nRows; ++r) { if (col[r] != 0.f) { flags[r / 64] |= 1ull << (r % 64); } }
            w.put<uint32_t>(s.nRows);
            for (size_t i = 0; i < flags.size(); ++i) { w.put<uint64_t>(flags[i]); }
            w.putFloats(col, s.nRows);
        }
        w.put<float>(s.beta);
    }
    // ConcurrentAtomicDomain.cpp:134-142 (size_t count = 8 bytes), ConcurrentAtom.cpp:98-102
    w.put<uint64_t>(s.domainLength);
    w.put<uint64_t>(static_cast<uint64_t>(s.pos.size()));
    for (size_t i = 0; i < s.pos.size(); ++i)
    {
        w.put<uint64_t>(s.pos[i]);
        w.put<float>(s.mass[i]);
    }
    // ProposalQueue.cpp:285-291
    const QueueState &q = s.queue;
    w.put<uint64_t>(q.rng);
    w.put<uint64_t>(q.minAtoms);
    w.put<uint64_t>(q.maxAtoms);
    w.put<uint64_t>(q.binLength);
    w.put<uint64_t>(q.numCols);
    w.put<double>(q.alpha);
    w.put<double>(q.domainLength);
    w.put<double>(q.numBins);
    w.put<float>(q.lambda);
    w.put<uint8_t>(q.useCachedRng ? 1 : 0);
    w.put<float>(q.u1);
    w.put<float>(q.u2);
}

bool getSampler(ByteReader &r, bool sparse, SamplerImage &s, std::string &err)
{
    s.sparse = sparse;
    if (!sparse)
    {
        if (!getMatrix(r, s.cols, s.nRows, s.k, err)) { return false; }
        s.rows.clear();
        s.beta = 0.f;
    }
    else
    {
        s.nRows = r.get<uint32_t>();
        s.k = r.get<uint32_t>();
        if (!r.ok() || !plausible(r, s.nRows, 2ull * s.k)) { err = "hybrid matrix header does not fit the file"; r.fail(); return false; }
        s.rows.assign(static_cast<size_t>(s.nRows) * s.k, 0.f);
        s.cols.assign(static_cast<size_t>(s.nRows) * s.k, 0.f);
        for (uint32_t i = 0; i < s.nRows; ++i)
        {
            if (r.get<uint32_t>() != s.k) { err = "row length differs from the hybrid matrix header"; r.fail(); return false; }
            r.getFloats(s.rows.data() + static_cast<size_t>(i) * s.k, s.k);
        }
        for (uint32_t c = 0; c < s.k; ++c)
        {
            if (r.get<uint32_t>() != s.nRows) { err = "column length differs from the hybrid matrix header"; r.fail(); return false; }
            r.skip((static_cast<size_t>(s.nRows) / 64 + 1) * sizeof(uint64_t)); // implied by the values
            r.getFloats(s.cols.data() + static_cast<size_t>(c) * s.nRows, s.nRows);
        }
        s.beta = r.get<float>();
    }
    s.domainLength = r.get<uint64_t>();
    const uint64_t n = r.get<uint64_t>();
    if (!r.ok() || n > r.remaining() / 12) { err = "atom count does not fit the file"; r.fail(); return false; }
    s.pos.resize(static_cast<size_t>(n));
    s.mass.resize(static_cast<size_t>(n));
    for (size_t i = 0; i < s.pos.size(); ++i)
    {
        s.pos[i] = r.get<uint64_t>();
        s.mass[i] = r.get<float>();
    }
    QueueState &q = s.queue;
    q.rng = r.get<uint64_t>();
    q.minAtoms = r.get<uint64_t>();
    q.maxAtoms = r.get<uint64_t>();
    q.binLength = r.get<uint64_t>();
    q.numCols = r.get<uint64_t>();
    q.alpha = r.get<double>();
    q.domainLength = r.get<double>();
    q.numBins = r.get<double>();
    q.lambda = r.get<float>();
    q.useCachedRng = r.get<uint8_t>() != 0;
    q.u1 = r.get<float>();
    q.u2 = r.get<float>();
    if (!r.ok()) { err = "file ends inside a sampler"; }
    return r.ok();
}

// ---- GapsStatistics (GapsStatistics.cpp:164-176) ----
void putStats(ByteWriter &w, const StatsImage &st)
{
    putMatrix(w, st.aMean, st.nGenes, st.k);
    putMatrix(w, st.aSq, st.nGenes, st.k);
    putMatrix(w, st.pMean, st.nSamples, st.k);
    putMatrix(w, st.pSq, st.nSamples, st.k);
    w.put<uint32_t>(st.statUpdates);
    w.put<uint32_t>(st.numPatterns);
}

bool getStats(ByteReader &r, StatsImage &st, std::string &err)
{
    uint32_t rows = 0, cols = 0;
    if (!getMatrix(r, st.aMean, st.nGenes, st.k, err)) { return false; }
    if (!getMatrix(r, st.aSq, rows, cols, err)) { return false; }
    if (rows != st.nGenes || cols != st.k) { err = "statistics matrices differ in shape"; r.fail(); return false; }
    if (!getMatrix(r, st.pMean, st.nSamples, cols, err)) { return false; }
    if (cols != st.k) { err = "statistics matrices differ in shape"; r.fail(); return false; }
    if (!getMatrix(r, st.pSq, rows, cols, err)) { return false; }
    if (rows != st.nSamples || cols != st.k) { err = "statistics matrices differ in shape"; r.fail(); return false; }
    st.statUpdates = r.get<uint32_t>();
    st.numPatterns = r.get<uint32_t>();
    if (!r.ok()) { err = "file ends inside the statistics"; }
    return r.ok();
}

// ---- the whole file (GapsRunner.cpp:237-240) ----
void putCheckpoint(ByteWriter &w, const CheckpointImage &c)
{
    w.put<uint32_t>(kArchiveMagic);
    putParams(w, c.params);
    w.put<uint64_t>(c.seeder[0]);
    w.put<uint64_t>(c.seeder[1]);
    putSampler(w, c.A);
    putSampler(w, c.P);
    putStats(w, c.stats);
    w.put<int32_t>(c.phase);
    w.put<uint32_t>(c.iter);
    w.put<uint64_t>(c.rng);
}

static bool getHeader(ByteReader &r, ParamsImage &p, uint64_t seeder[2], std::string &err)
{
    if (r.get<uint32_t>() != kArchiveMagic || !r.ok()) { err = "incompatible checkpoint file"; r.fail(); return false; } // Archive.h:33-36
    if (!getParams(r, p)) { err = "file ends inside the parameters"; return false; }
    seeder[0] = r.get<uint64_t>();
    seeder[1] = r.get<uint64_t>();
    if (!r.ok()) { err = "file ends inside the random state"; }
    return r.ok();
}

bool getCheckpoint(ByteReader &r, CheckpointImage &c, std::string &err)
{
    if (!getHeader(r, c.params, c.seeder, err)) { return false; }
    if (!getSampler(r, c.params.useSparseOptimization, c.A, err)) { return false; }
    if (!getSampler(r, c.params.useSparseOptimization, c.P, err)) { return false; }
    if (!getStats(r, c.stats, err)) { return false; }
    c.phase = r.get<int32_t>();
    c.iter = r.get<uint32_t>();
    c.rng = r.get<uint64_t>();
    if (!r.ok()) { err = "file ends before the run state"; return false; }
    if (r.remaining() != 0) { err = "trailing bytes after the run state"; return false; }
    if (c.A.nRows != c.params.nGenes || c.P.nRows != c.params.nSamples || c.A.k != c.params.nPatterns || c.P.k != c.params.nPatterns
        || c.stats.nGenes != c.params.nGenes || c.stats.nSamples != c.params.nSamples || c.stats.k != c.params.nPatterns)
    {
        err = "sampler / statistics shapes disagree with the archived parameters";
        return false;
    }
    return true;
}

bool readWholeFile(const char *path, std::vector<uint8_t> &out, std::string &err)
{
    std::FILE *f = std::fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path + ": " + std::strerror(errno); return false; }
    out.clear();
    uint8_t buf[1 << 16];
    size_t n = 0;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) { out.insert(out.end(), buf, buf + n); }
    const bool bad = std::ferror(f) != 0;
    std::fclose(f);
    if (bad) { err = std::string("read error on ") + path; }
    return !bad;
}

bool writeWholeFile(const char *path, const std::vector<uint8_t> &bytes, std::string &err)
{
    std::FILE *f = std::fopen(path, "wb");
    if (!f) { err = std::string("cannot create ") + path + ": " + std::strerror(errno); return false; }
    const bool wrote = bytes.empty() || std::fwrite(bytes.data(), 1, bytes.size(), f) == bytes.size();
    const bool closed = std::fclose(f) == 0;
    if (!wrote || !closed) { err = std::string("write error on ") + path; }
    return wrote && closed;
}

bool readCheckpointHeader(const char *path, ParamsImage &p, uint64_t seeder[2], std::string &err)
{
    std::vector<uint8_t> raw;
    if (!readWholeFile(path, raw, err)) { return false; }
    ByteReader r(raw.data(), raw.size());
    return getHeader(r, p, seeder, err);
}

bool readCheckpointFile(const char *path, CheckpointImage &c, std::string &err)
{
    std::vector<uint8_t> raw;
    if (!readWholeFile(path, raw, err)) { return false; }
    ByteReader r(raw.data(), raw.size());
    return getCheckpoint(r, c, err);
}

bool writeCheckpointFile(const char *path, const CheckpointImage &c, std::string &err)
{
    ByteWriter w;
    putCheckpoint(w, c);
    const std::string backup = std::string(path) + ".backup";
    std::rename(path, backup.c_str()); // fails harmlessly when there is no previous checkpoint
    if (!writeWholeFile(path, w.bytes(), err)) { return false; } // the backup stays behind, as in the reference
    std::remove(backup.c_str());
    return true;
}

} // namespace cgb
