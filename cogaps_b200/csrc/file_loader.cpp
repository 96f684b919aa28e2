// file_loader.cpp — the path overload of gaps::run (src/GapsRunner.h:19-24): reads a data / uncertainty matrix
// from a Matrix-Market (.mtx), comma- or tab-separated (.csv / .tsv) or .gct file into a dense fp32 row-major array,
// with the reference's parsing rules restated:
//   file type by extension .................. file_parser/FileParser.cpp:69-86
//   .mtx: '%' comment lines, "nrow ncol [nnz]", then 1-based "row col value" triplets, absent entries are 0
//                                             file_parser/MtxParser.cpp:8-62
//   .csv/.tsv: first line = column names; the first cell is empty when row names are present (then the first
//   cell of every later line is skipped); cells are trimmed of spaces, quotes, CR/LF ... CharacterDelimitedParser.cpp:8-150
//   .gct: line 2 holds "nrow ncol", line 3 the column names, two leading cells (name, description) per row
//   values: digits . - only -> decimal parse; otherwise base "e" exponent -> base * powf(10, exponent) in fp32
//                                             file_parser/MatrixElement.cpp:10-46
#include "../../include/cogaps_b200.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <system_error>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <limits>
#include <string>
#include <utility>
#include <vector>

namespace cgb {

static const char *kTrimChars = " \r\n\"";

static std::string trim(const std::string &s)
{
    const std::size_t a = s.find_first_not_of(kTrimChars);
    if (a == std::string::npos) { return std::string(); }
    const std::size_t b = s.find_last_not_of(kTrimChars);
    return s.substr(a, b - a + 1);
}

static bool isNumber(const std::string &s)
{
    return !s.empty() && s.find_first_not_of("0123456789.-") == std::string::npos;
}

// what `std::stringstream ss(s); float v; ss >> v;` yields for a token made of digits, '.' and '-': the longest
// valid prefix, 0 when there is none
static float streamFloat(const std::string &s)
{
    // from_chars is correctly rounded like glibc's strtof and takes the longest valid prefix too; it is several times
    // faster.  It reports overflow / underflow instead of saturating, so those rare tokens go through strtof.
    float fast = 0.f;
    const std::from_chars_result res = std::from_chars(s.data(), s.data() + s.size(), fast, std::chars_format::fixed);
    if (res.ec == std::errc()) { return fast; }
    if (res.ec == std::errc::invalid_argument) { return 0.f; }
    char *end = nullptr;
    const float v = std::strtof(s.c_str(), &end);
    if (end == s.c_str()) { return 0.f; }
    // out of range: the stream extractor stores +/-FLT_MAX for an overflow (num_get, LWG 23) where strtof returns
    // +/-inf; an underflow keeps strtof's (denormal or zero) value
    if (v > std::numeric_limits<float>::max()) { return std::numeric_limits<float>::max(); }
    if (v < -std::numeric_limits<float>::max()) { return -std::numeric_limits<float>::max(); }
    return v;
}

static bool parseValue(const std::string &s, float &out, std::string &err)
{
    if (isNumber(s))
    {
        out = streamFloat(s);
        return true;
    }
    const std::size_t pos = s.find('e');
    if (pos != std::string::npos)
    {
        const std::string base = s.substr(0, pos), expo = s.substr(pos + 1);
        if (isNumber(base) && isNumber(expo))
        {
            out = streamFloat(base) * std::pow(10.f, streamFloat(expo));
            return true;
        }
    }
    err = "Invalid entry found in input data: " + s;
    return false;
}

enum FileType { kInvalid, kMtx, kCsv, kTsv, kGct };

static FileType fileType(const std::string &path)
{
    const std::size_t pos = path.find_last_of('.');
    if (pos == std::string::npos) { return kInvalid; }
    const std::string ext = path.substr(pos);
    if (ext.find('/') != std::string::npos) { return kInvalid; }
    if (ext == ".mtx") { return kMtx; }
    if (ext == ".csv") { return kCsv; }
    if (ext == ".tsv") { return kTsv; }
    if (ext == ".gct") { return kGct; }
    return kInvalid;
}

static std::vector<std::string> split(const std::string &line, char delimiter)
{
    std::vector<std::string> tokens;
    std::string cell;
    std::stringstream ss(line);
    while (std::getline(ss, cell, delimiter)) { tokens.push_back(trim(cell)); }
    return tokens;
}

static bool readMtx(std::ifstream &f, std::vector<float> &out, uint32_t &nrow, uint32_t &ncol, std::string &err)
{
    std::string line = "%";
    while (line.find('%') != std::string::npos)
    {
        if (!std::getline(f, line))
        {
            err = "Invalid MTX file";
            return false;
        }
    }
    std::stringstream dims(line);
    unsigned long r = 0, c = 0;
    dims >> r >> c;
    if (r == 0 || c == 0)
    {
        err = "Invalid MTX file";
        return false;
    }
    if (r > 0xFFFFFFFFul || c > 0xFFFFFFFFul)
    {
        err = "MTX dimensions exceed 32 bits";
        return false;
    }
    nrow = static_cast<uint32_t>(r);
    ncol = static_cast<uint32_t>(c);
    out.assign(static_cast<size_t>(nrow) * ncol, 0.f);
    unsigned long row = 0, col = 0;
    std::string val;
    while (f >> row >> col >> val)
    {
        float v;
        if (!parseValue(val, v, err)) { return false; }
        if (row < 1 || row > nrow || col < 1 || col > ncol)
        {
            err = "MTX entry outside the declared dimensions";
            return false;
        }
        out[static_cast<size_t>(row - 1) * ncol + (col - 1)] = v;
    }
    return true;
}

// ---- Matrix-Market straight to compressed rows ("loaders straight to device layouts") ----
// Same parsing as readMtx, but the triplets are kept as triplets.  Semantics of the dense reader that must survive:
// a later entry for the same cell overwrites an earlier one; explicit zeros are cells like any other (they simply are
// not positive); everything not listed is 0.
// what `stream >> unsignedLong` accepts: optional whitespace, digits.  False (cursor unchanged) when no digit follows —
// the reference's read loop simply ends there (MtxParser.cpp:35-47 reads while the stream is good)
static bool scanUnsigned(const char *&p, const char *end, unsigned long &out)
{
    const char *q = p;
    while (q < end && (*q == ' ' || *q == '\t' || *q == '\n' || *q == '\r' || *q == '\v' || *q == '\f')) { ++q; }
    if (q < end && *q == '+') { ++q; }
    if (q >= end || *q < '0' || *q > '9') { return false; }
    unsigned long v = 0;
    while (q < end && *q >= '0' && *q <= '9')
    {
        v = v * 10 + static_cast<unsigned long>(*q - '0');
        ++q;
    }
    out = v;
    p = q;
    return true;
}

// what `stream >> std::string` yields: the next run of non-whitespace characters
static bool scanToken(const char *&p, const char *end, std::string &out)
{
    const char *q = p;
    while (q < end && (*q == ' ' || *q == '\t' || *q == '\n' || *q == '\r' || *q == '\v' || *q == '\f')) { ++q; }
    if (q >= end) { return false; }
    const char *b = q;
    while (q < end && !(*q == ' ' || *q == '\t' || *q == '\n' || *q == '\r' || *q == '\v' || *q == '\f')) { ++q; }
    out.assign(b, q);
    p = q;
    return true;
}

bool loadMtxTriplets(const char *path, std::vector<uint32_t> &rows, std::vector<uint32_t> &cols, std::vector<float> &vals,
                     uint32_t &nrow, uint32_t &ncol, std::string &err)
{
    if (fileType(path) != kMtx)
    {
        err = "not a Matrix-Market file";
        return false;
    }
    // the whole file in one read, then a hand-rolled scan: the stream extractors of the dense reader cost about a
    // microsecond per entry, which is minutes at the 10^8 entries of a single-cell matrix
    std::string text;
    {
        std::FILE *f = std::fopen(path, "rb");
        if (!f)
        {
            err = std::string("cannot open ") + path;
            return false;
        }
        char buf[1 << 16];
        size_t n = 0;
        while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) { text.append(buf, n); }
        std::fclose(f);
    }
    const char *p = text.data(), *end = text.data() + text.size();
    // header: lines are skipped while they contain a '%' anywhere (MtxParser.cpp:14-20), the first one that does not
    // holds "nrow ncol [nnz]"
    std::string line;
    for (;;)
    {
        if (p >= end)
        {
            err = "Invalid MTX file";
            return false;
        }
        const char *nl = static_cast<const char*>(std::memchr(p, '\n', static_cast<size_t>(end - p)));
        const char *lineEnd = nl ? nl : end;
        line.assign(p, lineEnd);
        p = nl ? nl + 1 : end;
        if (line.find('%') == std::string::npos) { break; }
    }
    unsigned long r = 0, c = 0, declared = 0;
    {
        const char *q = line.data(), *qe = line.data() + line.size();
        if (scanUnsigned(q, qe, r) && scanUnsigned(q, qe, c)) { scanUnsigned(q, qe, declared); }
    }
    if (r == 0 || c == 0)
    {
        err = "Invalid MTX file";
        return false;
    }
    if (r > 0xFFFFFFFFul || c > 0xFFFFFFFFul)
    {
        err = "MTX dimensions exceed 32 bits";
        return false;
    }
    nrow = static_cast<uint32_t>(r);
    ncol = static_cast<uint32_t>(c);
    rows.clear(); cols.clear(); vals.clear();
    if (declared > 0 && declared <= static_cast<unsigned long>(nrow) * ncol)
    {
        rows.reserve(declared); cols.reserve(declared); vals.reserve(declared);
    }
    unsigned long row = 0, col = 0;
    std::string val;
    while (scanUnsigned(p, end, row) && scanUnsigned(p, end, col) && scanToken(p, end, val))
    {
        float v;
        if (!parseValue(val, v, err)) { return false; }
        if (row < 1 || row > nrow || col < 1 || col > ncol)
        {
            err = "MTX entry outside the declared dimensions";
            return false;
        }
        if (vals.size() >= 0xFFFFFFFFul)
        {
            err = "MTX file holds more than 2^32 - 1 entries";   // the compressed rows index them with 32 bits
            return false;
        }
        rows.push_back(static_cast<uint32_t>(row - 1));
        cols.push_back(static_cast<uint32_t>(col - 1));
        vals.push_back(v);
    }
    return true;
}

// Compressed rows of an nMajor x nMinor matrix from triplets (major[i], minor[i], vals[i]): within a row ascending
// minor index, of several entries for one cell the LAST in file order, entries that are not positive dropped
// (SparseVector keeps v > 0 only, SparseVector.cpp:20-35).  anyNegative reports a value < 0 anywhere (the dense path's
// lambda sums those too, so the caller falls back to it).
void compressTriplets(const std::vector<uint32_t> &major, const std::vector<uint32_t> &minor, const std::vector<float> &vals,
                      uint32_t nMajor, std::vector<uint32_t> &ptr, std::vector<uint32_t> &idx, std::vector<float> &out,
                      bool &anyNegative)
{
    const size_t n = vals.size();
    // two stable counting sorts (by minor, then by major): entries end up grouped by row, ascending minor inside a
    // row, and entries for the same cell in file order — O(n), no comparison sort
    uint32_t nMinor = 0;
    for (size_t i = 0; i < n; ++i) { if (minor[i] >= nMinor) { nMinor = minor[i] + 1; } }
    std::vector<uint32_t> byMinor(n), order(n);
    {
        std::vector<size_t> fill(static_cast<size_t>(nMinor) + 1, 0);
        for (size_t i = 0; i < n; ++i) { ++fill[minor[i] + 1]; }
        for (uint32_t m = 0; m < nMinor; ++m) { fill[m + 1] += fill[m]; }
        for (size_t i = 0; i < n; ++i) { byMinor[fill[minor[i]]++] = static_cast<uint32_t>(i); }
    }
    {
        std::vector<size_t> fill(static_cast<size_t>(nMajor) + 1, 0);
        for (size_t i = 0; i < n; ++i) { ++fill[major[i] + 1]; }
        for (uint32_t r = 0; r < nMajor; ++r) { fill[r + 1] += fill[r]; }
        for (size_t j = 0; j < n; ++j) { order[fill[major[byMinor[j]]]++] = byMinor[j]; }
    }
    ptr.assign(static_cast<size_t>(nMajor) + 1, 0u);
    idx.clear(); out.clear();
    idx.reserve(n); out.reserve(n);
    anyNegative = false;
    size_t j = 0;
    for (uint32_t r = 0; r < nMajor; ++r)
    {
        while (j < n && major[order[j]] == r)
        {
            const uint32_t e = order[j];
            const bool overwrittenLater = j + 1 < n && major[order[j + 1]] == r && minor[order[j + 1]] == minor[e];
            if (!overwrittenLater)
            {
                const float v = vals[e];
                if (v < 0.f) { anyNegative = true; }
                if (v > 0.f)
                {
                    idx.push_back(minor[e]);
                    out.push_back(v);
                }
            }
            ++j;
        }
        ptr[r + 1] = static_cast<uint32_t>(idx.size());
    }
}

static bool readDelimited(std::ifstream &f, char delimiter, bool gct, std::vector<float> &out, uint32_t &nrow, uint32_t &ncol,
                          std::string &err)
{
    std::vector<std::string> lines;
    std::string line;
    while (std::getline(f, line)) { lines.push_back(line); }
    size_t firstData;
    bool rowNames = false;
    size_t skipCells = 0;
    if (gct)
    {
        if (lines.size() < 3)
        {
            err = "Invalid character delimited file";
            return false;
        }
        std::stringstream dims(lines[1]);
        unsigned long r = 0, c = 0;
        dims >> r >> c;
        nrow = static_cast<uint32_t>(r);
        ncol = static_cast<uint32_t>(c);
        firstData = 3;
        skipCells = 2;
    }
    else
    {
        if (lines.empty())
        {
            err = "Invalid character delimited file";
            return false;
        }
        // header: the cells of the first line; an empty first cell means row names are present
        std::string header = lines[0];
        std::vector<std::string> cells;
        {
            std::string cell;
            std::stringstream ss(header);
            while (std::getline(ss, cell, delimiter)) { cells.push_back(cell); }
            if (!header.empty() && header[header.size() - 1] == delimiter) { cells.push_back(std::string()); }
        }
        if (cells.empty())
        {
            err = "Invalid character delimited file";
            return false;
        }
        rowNames = trim(cells[0]).empty();
        ncol = static_cast<uint32_t>(cells.size() - (rowNames ? 1 : 0));
        nrow = static_cast<uint32_t>(lines.size() - 1); // the reference counts every remaining line
        firstData = 1;
        skipCells = rowNames ? 1 : 0;
    }
    if (nrow == 0 || ncol == 0)
    {
        err = "Invalid character delimited file";
        return false;
    }
    out.assign(static_cast<size_t>(nrow) * ncol, 0.f);
    uint32_t r = 0;
    for (size_t i = firstData; i < lines.size() && r < nrow; ++i)
    {
        // the reference stops at trailing whitespace-only content (hasNext skips whitespace, then EOF)
        if (lines[i].find_first_not_of(" \t\r\n") == std::string::npos)
        {
            bool onlyBlankAfter = true;
            for (size_t j = i + 1; j < lines.size(); ++j)
            {
                if (lines[j].find_first_not_of(" \t\r\n") != std::string::npos) { onlyBlankAfter = false; }
            }
            if (onlyBlankAfter) { break; }
        }
        std::vector<std::string> cells = split(lines[i], delimiter);
        for (size_t c = skipCells; c < cells.size(); ++c)
        {
            const size_t col = c - skipCells;
            if (col >= ncol)
            {
                err = "row with more cells than the header";
                return false;
            }
            float v;
            if (!parseValue(cells[c], v, err)) { return false; }
            out[static_cast<size_t>(r) * ncol + col] = v;
        }
        ++r;
    }
    return true;
}

// FileParser::writeToCsv (file_parser/FileParser.h:59-89): "" then "Col<j>" headers, "Row<i>" row names, values
// through operator<<(float) — the default stream format, i.e. %g with six significant digits
bool writeMatrixCsv(const char *path, const float *mat, uint32_t nrow, uint32_t ncol, std::string &err)
{
    if (fileType(path) != kCsv)
    {
        err = "output file must be a csv";
        return false;
    }
    std::FILE *f = std::fopen(path, "w");
    if (!f)
    {
        err = std::string("cannot create ") + path;
        return false;
    }
    std::string text = "\"\"";
    char buf[48];
    for (uint32_t j = 0; j < ncol; ++j)
    {
        std::snprintf(buf, sizeof(buf), ",\"Col%u\"", j);
        text += buf;
    }
    text += "\n";
    for (uint32_t i = 0; i < nrow; ++i)
    {
        std::snprintf(buf, sizeof(buf), "\"Row%u\"", i);
        text += buf;
        for (uint32_t j = 0; j < ncol; ++j)
        {
            std::snprintf(buf, sizeof(buf), ",%g", static_cast<double>(mat[static_cast<size_t>(i) * ncol + j]));
            text += buf;
        }
        text += "\n";
        if (text.size() > (1u << 20))
        {
            if (std::fwrite(text.data(), 1, text.size(), f) != text.size()) { std::fclose(f); err = std::string("write error on ") + path; return false; }
            text.clear();
        }
    }
    const bool wrote = text.empty() || std::fwrite(text.data(), 1, text.size(), f) == text.size();
    const bool closed = std::fclose(f) == 0;
    if (!wrote || !closed) { err = std::string("write error on ") + path; }
    return wrote && closed;
}

// FileParser::colNames (what getFileInfo_cpp hands to R, src/Cogaps.cpp:245-256): the trimmed cells of the header line
// of a .csv / .tsv file, without the empty corner cell when row names are present
// (CharacterDelimitedParser.cpp:78-98).  .gct and .mtx files have none — and no parser of the reference ever fills
// rowNames, so those are always empty.
bool loadColumnNames(const char *path, std::vector<std::string> &names, std::string &err)
{
    names.clear();
    const FileType t = fileType(path);
    if (t == kInvalid)
    {
        err = "Invalid file type";
        return false;
    }
    if (t == kMtx || t == kGct) { return true; }
    std::ifstream f(path);
    if (!f.is_open())
    {
        err = std::string("cannot open ") + path;
        return false;
    }
    std::string header;
    if (!std::getline(f, header))
    {
        err = "Invalid character delimited file";
        return false;
    }
    const char delimiter = (t == kCsv) ? ',' : '\t';
    std::vector<std::string> cells;
    {
        std::string cell;
        std::stringstream ss(header);
        while (std::getline(ss, cell, delimiter)) { cells.push_back(cell); }
        if (!header.empty() && header[header.size() - 1] == delimiter) { cells.push_back(std::string()); }
    }
    for (size_t i = 0; i < cells.size(); ++i)
    {
        const std::string name = trim(cells[i]);
        if (i == 0 && name.empty()) { continue; } // the corner cell above the row names
        names.push_back(name);
    }
    return true;
}

bool loadMatrixFile(const char *path, std::vector<float> &out, uint32_t &nrow, uint32_t &ncol, std::string &err)
{
    const FileType t = fileType(path);
    if (t == kInvalid)
    {
        err = "Invalid file type";
        return false;
    }
    std::ifstream f(path);
    if (!f.is_open())
    {
        err = std::string("cannot open ") + path;
        return false;
    }
    switch (t)
    {
        case kMtx: return readMtx(f, out, nrow, ncol, err);
        case kCsv: return readDelimited(f, ',', false, out, nrow, ncol, err);
        case kTsv: return readDelimited(f, '\t', false, out, nrow, ncol, err);
        default: return readDelimited(f, '\t', true, out, nrow, ncol, err);
    }
}

} // namespace cgb
