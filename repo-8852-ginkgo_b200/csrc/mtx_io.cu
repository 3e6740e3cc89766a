// mtx_io.cu — host code: the callers' wire / disk formats.
// A native reader / writer for what gko::read_generic_raw / write_raw / write_binary_raw accept
// and produce (reference core/base/mtx_io.cpp):
//   * MatrixMarket text: "%%MatrixMarket matrix (coordinate|array) (real|integer|complex|pattern)
//     (general|symmetric|skew-symmetric|hermitian)" (:690-722), comment lines, 1-based
//     coordinates, column-major array layout, symmetric / skew-symmetric / hermitian expansion
//     (:293-463: the mirrored entry is inserted right after the stored one; skew negates it),
//     pattern entries = 1;
//   * the "GINKGO" binary format (:776-925): 32-byte header {magic, rows, cols, entries}, then
//     (row, col, value) records in the file's index / value type (I|L x S|D; complex files are
//     rejected for real value types like the reference does).
// Like the reference, the result is in row-major order (matrix_data::ensure_row_major_order;
// the sort is stable here, the reference's std::sort leaves duplicates' order unspecified).
// No CUDA in this file: it feeds assembly (setup_kernels.cu) through host arrays.
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <stdexcept>
#include <numeric>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#include <parallel/algorithm>
#endif

#include "../../include/gko_b200.h"

namespace {

thread_local std::string g_error;

struct MtxData {
    int64_t n_rows = 0, n_cols = 0;
    std::vector<int64_t> rows, cols;
    std::vector<double> vals;
};

int fail(const std::string& msg)
{
    g_error = msg;
    return GKOB200_EINVAL;
}

bool read_file(const char* path, std::string& out)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? static_cast<size_t>(n) : 0);
    const size_t got = n > 0 ? std::fread(&out[0], 1, static_cast<size_t>(n), f) : 0;
    std::fclose(f);
    return got == out.size();
}

// whitespace-separated tokens of a buffer, like successive `stream >> x`
struct Tokens {
    const char* p;
    const char* end;
    bool next_i64(int64_t& v)
    {
        while (p < end && std::isspace(static_cast<unsigned char>(*p))) ++p;
        if (p >= end) return false;
        char* q = nullptr;
        v = std::strtoll(p, &q, 10);
        if (q == p) return false;
        p = q;
        return true;
    }
    bool next_f64(double& v)
    {
        while (p < end && std::isspace(static_cast<unsigned char>(*p))) ++p;
        if (p >= end) return false;
        char* q = nullptr;
        v = std::strtod(p, &q);
        if (q == p) return false;
        p = q;
        return true;
    }
};

// ---- parallel token parsing ------------------------------------------------------------------
// The body of a MatrixMarket file is a stream of whitespace-separated tokens (the reference reads
// it with successive `stream >> x`, line structure does not matter).  The buffer is cut into
// chunks at token boundaries, tokens are counted per chunk (pass 1), and every chunk then knows
// the global index of its first token and parses its own tokens (pass 2): field = index % fields.
inline bool is_space(char ch) { return std::isspace(static_cast<unsigned char>(ch)) != 0; }

struct TokenTable {
    std::vector<const char*> chunk_begin;   // first token start of every chunk (+ end sentinel)
    std::vector<int64_t> first_index;       // global index of that token
    int64_t total = 0;
};

TokenTable index_tokens(const char* begin, const char* end)
{
    int n_chunks = 1;
#ifdef _OPENMP
    n_chunks = omp_get_max_threads() * 4;
#endif
    const size_t len = static_cast<size_t>(end - begin);
    if (len < (1u << 20)) n_chunks = 1;
    TokenTable t;
    t.chunk_begin.resize(n_chunks + 1);
    t.first_index.assign(n_chunks + 1, 0);
    for (int c = 0; c <= n_chunks; ++c) {
        const char* p = begin + len * static_cast<size_t>(c) / n_chunks;
        if (c == n_chunks) {
            p = end;
        } else if (p > begin) {
            while (p < end && !is_space(p[-1])) ++p;   // finish the token the cut landed in
        }
        while (p < end && is_space(*p)) ++p;           // next token start
        t.chunk_begin[c] = p;
    }
    std::vector<int64_t> count(n_chunks, 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < n_chunks; ++c) {
        int64_t k = 0;
        const char* p = t.chunk_begin[c];
        const char* e = t.chunk_begin[c + 1];
        while (p < e) {
            ++k;
            while (p < e && !is_space(*p)) ++p;
            while (p < e && is_space(*p)) ++p;
        }
        count[c] = k;
    }
    for (int c = 0; c < n_chunks; ++c) t.first_index[c + 1] = t.first_index[c] + count[c];
    t.total = t.first_index[n_chunks];
    return t;
}

// Parses the first n_entries * fields tokens: field f < int_fields as int64 into ints[f][entry],
// the others as double into vals[entry].  Returns the index of the first entry that failed to
// parse, or -1.
int64_t parse_tokens(const TokenTable& t, int64_t n_entries, int fields, int int_fields, std::vector<int64_t>* ints,
                     std::vector<double>& vals)
{
    int64_t bad = std::numeric_limits<int64_t>::max();
    const int n_chunks = static_cast<int>(t.chunk_begin.size()) - 1;
#pragma omp parallel for schedule(dynamic, 1) reduction(min : bad)
    for (int c = 0; c < n_chunks; ++c) {
        const char* p = t.chunk_begin[c];
        const char* e = t.chunk_begin[c + 1];
        int64_t g = t.first_index[c];
        while (p < e && g < n_entries * fields) {
            const int64_t entry = g / fields;
            const int f = static_cast<int>(g % fields);
            char* q = nullptr;
            if (f < int_fields) {
                ints[f][entry] = std::strtoll(p, &q, 10);
            } else {
                vals[entry] = std::strtod(p, &q);
            }
            // the token has to be consumed completely (`stream >> x` would leave the rest for the
            // next extraction, which then fails)
            if (q == p || (q < e && !is_space(*q))) bad = std::min(bad, entry);
            while (p < e && !is_space(*p)) ++p;
            while (p < e && is_space(*p)) ++p;
            ++g;
        }
    }
    return bad == std::numeric_limits<int64_t>::max() ? -1 : bad;
}

enum Layout { COORDINATE, ARRAY };
enum Entry { REAL, INTEGER, COMPLEX, PATTERN };
enum Modifier { GENERAL, SYMMETRIC, SKEW, HERMITIAN };

void insert(MtxData& d, Modifier m, int64_t r, int64_t c, double v)
{
    d.rows.push_back(r);
    d.cols.push_back(c);
    d.vals.push_back(v);
    if (m == GENERAL) return;
    if (m == SKEW) {   // always mirrored, negated (mtx_io.cpp:395-401)
        d.rows.push_back(c);
        d.cols.push_back(r);
        d.vals.push_back(-v);
    } else if (r != c) {   // symmetric / hermitian (conj of a real value is the value)
        d.rows.push_back(c);
        d.cols.push_back(r);
        d.vals.push_back(v);
    }
}

void sort_row_major(MtxData& d)
{
    const size_t n = d.vals.size();
    bool sorted = true;
    for (size_t i = 1; i < n && sorted; ++i)
        sorted = d.rows[i - 1] < d.rows[i] || (d.rows[i - 1] == d.rows[i] && d.cols[i - 1] <= d.cols[i]);
    if (sorted) return;   // the usual case for files written by a library
    std::vector<size_t> perm(n);
    std::iota(perm.begin(), perm.end(), size_t(0));
    auto less = [&](size_t a, size_t b) {
        return d.rows[a] != d.rows[b] ? d.rows[a] < d.rows[b] : d.cols[a] < d.cols[b];
    };
#ifdef _OPENMP
    __gnu_parallel::stable_sort(perm.begin(), perm.end(), less);
#else
    std::stable_sort(perm.begin(), perm.end(), less);
#endif
    MtxData s;
    s.rows.resize(n);
    s.cols.resize(n);
    s.vals.resize(n);
    for (size_t i = 0; i < n; ++i) {
        s.rows[i] = d.rows[perm[i]];
        s.cols[i] = d.cols[perm[i]];
        s.vals[i] = d.vals[perm[i]];
    }
    d.rows.swap(s.rows);
    d.cols.swap(s.cols);
    d.vals.swap(s.vals);
}

int read_text(const std::string& buf, MtxData& d)
{
    size_t pos = 0;
    auto getline = [&](std::string& line) {
        if (pos >= buf.size()) return false;
        const size_t e = buf.find('\n', pos);
        line = buf.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        pos = e == std::string::npos ? buf.size() : e + 1;
        return true;
    };
    std::string line;
    do {
        if (!getline(line)) return fail("error when reading the header line");
    } while (line.empty());
    for (auto& ch : line) ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
    const std::string prefix = "%%matrixmarket matrix ";
    if (line.compare(0, prefix.size(), prefix) != 0) return fail("error parsing the header line");
    // exactly three single-space separated words
    std::vector<std::string> w;
    size_t a = prefix.size();
    while (a <= line.size()) {
        const size_t b = line.find(' ', a);
        w.push_back(line.substr(a, b == std::string::npos ? std::string::npos : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    if (w.size() != 3) return fail("error parsing the header line");
    Layout layout;
    Entry entry;
    Modifier mod;
    if (w[0] == "coordinate") layout = COORDINATE; else if (w[0] == "array") layout = ARRAY; else return fail("error parsing the header line: layout");
    if (w[1] == "real") entry = REAL; else if (w[1] == "integer") entry = INTEGER; else if (w[1] == "complex") entry = COMPLEX;
    else if (w[1] == "pattern") entry = PATTERN; else return fail("error parsing the header line: value type");
    if (w[2] == "general") mod = GENERAL; else if (w[2] == "symmetric") mod = SYMMETRIC; else if (w[2] == "skew-symmetric") mod = SKEW;
    else if (w[2] == "hermitian") mod = HERMITIAN; else return fail("error parsing the header line: modifier");
    if (entry == COMPLEX) return fail("trying to read a complex matrix into a real storage type");
    do {
        if (!getline(line)) return fail("error when reading the dimensions line");
    } while (!line.empty() && line[0] == '%');
    Tokens dims{line.data(), line.data() + line.size()};
    int64_t nnz = 0;
    if (!dims.next_i64(d.n_rows) || !dims.next_i64(d.n_cols) || d.n_rows < 0 || d.n_cols < 0)
        return fail("error when determining matrix size, expected: rows cols nnz");
    const TokenTable tokens = index_tokens(buf.data() + pos, buf.data() + buf.size());
    if (layout == COORDINATE) {
        if (!dims.next_i64(nnz) || nnz < 0) return fail("error when determining matrix size, expected: rows cols nnz");
        const int fields = entry == PATTERN ? 2 : 3;
        std::vector<int64_t> ints[2];
        // entries the file does not have tokens for fail like a stream at its end; the header's
        // nnz is not trusted for allocation: at most `have` entries (bounded by the file size)
        const int64_t have = tokens.total / fields;
        const int64_t n_parse = std::min(nnz, have);
        std::vector<double> vals(static_cast<size_t>(n_parse), 1.0);
        ints[0].resize(static_cast<size_t>(n_parse));
        ints[1].resize(static_cast<size_t>(n_parse));
        const int64_t bad = parse_tokens(tokens, n_parse, fields, 2, ints, vals);
        const int64_t first_bad = bad >= 0 ? bad : (have < nnz ? have : -1);
        if (first_bad >= 0) {
            // which extraction failed: the coordinates or the value
            const bool coord_missing = bad < 0 && tokens.total - have * fields < 2;
            return fail((bad >= 0 || coord_missing ? "error when reading coordinates of matrix entry "
                                                   : "error when reading matrix entry ") +
                        std::to_string(first_bad));
        }
        const size_t reserve = static_cast<size_t>(mod == GENERAL ? nnz : 2 * nnz);
        d.rows.reserve(reserve);
        d.cols.reserve(reserve);
        d.vals.reserve(reserve);
        for (int64_t i = 0; i < nnz; ++i) {
            // coordinates feed device assembly: they must address the declared matrix
            if (ints[0][i] < 1 || ints[0][i] > d.n_rows || ints[1][i] < 1 || ints[1][i] > d.n_cols)
                return fail("matrix entry " + std::to_string(i) + " lies outside of the " + std::to_string(d.n_rows) +
                            " x " + std::to_string(d.n_cols) + " matrix");
            insert(d, mod, ints[0][i] - 1, ints[1][i] - 1, vals[i]);
        }
    } else {
        int64_t count = 0;
        for (int64_t c = 0; c < d.n_cols; ++c) {
            const int64_t start = mod == GENERAL ? 0 : mod == SKEW ? c + 1 : c;
            count += std::max<int64_t>(0, d.n_rows - start);
        }
        if (entry != PATTERN && tokens.total < count) {
            // fewer values than the dimensions ask for: fail at the first missing (or malformed)
            // one without allocating what the header claims
            std::vector<double> part(static_cast<size_t>(tokens.total), 1.0);
            const int64_t bad = parse_tokens(tokens, tokens.total, 1, 0, nullptr, part);
            return fail("error when reading matrix entry " + std::to_string(bad >= 0 ? bad : tokens.total));
        }
        std::vector<double> vals(static_cast<size_t>(count), 1.0);   // pattern entries are ones and read nothing
        if (entry != PATTERN) {
            const int64_t n_parse = std::min(count, tokens.total);
            const int64_t bad = parse_tokens(tokens, n_parse, 1, 0, nullptr, vals);
            const int64_t first_bad = bad >= 0 ? bad : (tokens.total < count ? tokens.total : -1);
            if (first_bad >= 0) return fail("error when reading matrix entry " + std::to_string(first_bad));
        }
        int64_t k = 0;
        for (int64_t c = 0; c < d.n_cols; ++c) {
            const int64_t start = mod == GENERAL ? 0 : mod == SKEW ? c + 1 : c;
            for (int64_t r = start; r < d.n_rows; ++r) insert(d, mod, r, c, vals[k++]);
        }
    }
    sort_row_major(d);
    return 0;
}

// 'G','I','N','K','G','O', value char, index char — little endian (mtx_io.cpp:776-803)
uint64_t magic(char value_bit, char index_bit)
{
    const unsigned char b[8] = {'G', 'I', 'N', 'K', 'G', 'O', static_cast<unsigned char>(value_bit),
                                static_cast<unsigned char>(index_bit)};
    uint64_t m = 0;
    for (int i = 7; i >= 0; --i) m = m * 256 + b[i];
    return m;
}

int read_binary(const std::string& buf, MtxData& d)
{
    if (buf.size() < 32) return fail("failed reading header");
    uint64_t hdr[4];
    std::memcpy(hdr, buf.data(), 32);
    char vb = 0, ib = 0;
    for (char v : {'D', 'S', 'Z', 'C'})
        for (char i : {'I', 'L'})
            if (hdr[0] == magic(v, i)) {
                vb = v;
                ib = i;
            }
    if (!vb) return fail("invalid header magic number '" + buf.substr(0, 8) + "'");
    if (vb == 'Z' || vb == 'C') return fail("cannot read into this format, would assign complex to real");
    const size_t isz = ib == 'I' ? 4 : 8, vsz = vb == 'D' ? 8 : 4, rec = 2 * isz + vsz;
    const uint64_t n = hdr[3];
    // (checked: 32 + n * rec can wrap for a hostile n)
    if (n > (buf.size() - 32) / rec) return fail("failed reading entry " + std::to_string((buf.size() - 32) / rec));
    if (hdr[1] > static_cast<uint64_t>(INT64_MAX) || hdr[2] > static_cast<uint64_t>(INT64_MAX))
        return fail("invalid matrix dimensions in the header");
    d.n_rows = static_cast<int64_t>(hdr[1]);
    d.n_cols = static_cast<int64_t>(hdr[2]);
    d.rows.resize(n);
    d.cols.resize(n);
    d.vals.resize(n);
    const char* p = buf.data() + 32;
    for (uint64_t i = 0; i < n; ++i, p += rec) {
        if (isz == 4) {
            int32_t r, c;
            std::memcpy(&r, p, 4);
            std::memcpy(&c, p + 4, 4);
            d.rows[i] = r;
            d.cols[i] = c;
        } else {
            std::memcpy(&d.rows[i], p, 8);
            std::memcpy(&d.cols[i], p + 8, 8);
        }
        if (vsz == 8) {
            std::memcpy(&d.vals[i], p + 2 * isz, 8);
        } else {
            float v;
            std::memcpy(&v, p + 2 * isz, 4);
            d.vals[i] = v;
        }
    }
    for (uint64_t i = 0; i < n; ++i)
        if (d.rows[i] < 0 || d.rows[i] >= d.n_rows || d.cols[i] < 0 || d.cols[i] >= d.n_cols)
            return fail("matrix entry " + std::to_string(i) + " lies outside of the " + std::to_string(d.n_rows) +
                        " x " + std::to_string(d.n_cols) + " matrix");
    sort_row_major(d);
    return 0;
}

template <typename V, typename I>
int copy_out(const MtxData* d, I* rows, I* cols, V* vals)
{
    if (!d) return GKOB200_EINVAL;
    const int64_t lim = static_cast<int64_t>(std::numeric_limits<I>::max());
    if (d->n_rows > lim || d->n_cols > lim) return fail("cannot read into this format, its index type would overflow");
    const size_t n = d->vals.size();
    if (n && (!rows || !cols || !vals)) return GKOB200_EINVAL;
    for (size_t i = 0; i < n; ++i) {
        rows[i] = static_cast<I>(d->rows[i]);
        cols[i] = static_cast<I>(d->cols[i]);
        vals[i] = static_cast<V>(d->vals[i]);
    }
    return 0;
}

template <typename V, typename I>
int write_file(const char* path, int format, int precision, int64_t n_rows, int64_t n_cols, int64_t nnz, const I* rows,
               const I* cols, const V* vals)
{
    if (!path || nnz < 0 || (nnz > 0 && (!rows || !cols || !vals))) return GKOB200_EINVAL;
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(std::string("cannot open ") + path);
    if (format == 1) {   // GINKGO binary (mtx_io.cpp:927-958)
        const uint64_t hdr[4] = {magic(sizeof(V) == 8 ? 'D' : 'S', sizeof(I) == 4 ? 'I' : 'L'),
                                 static_cast<uint64_t>(n_rows), static_cast<uint64_t>(n_cols), static_cast<uint64_t>(nnz)};
        std::fwrite(hdr, 8, 4, f);
        for (int64_t i = 0; i < nnz; ++i) {
            std::fwrite(&rows[i], sizeof(I), 1, f);
            std::fwrite(&cols[i], sizeof(I), 1, f);
            std::fwrite(&vals[i], sizeof(V), 1, f);
        }
    } else if (format == 0) {   // coordinate real general (mtx_io.cpp:552-573)
        if (precision <= 0) precision = 6;   // the reference streams doubles at the default precision
        std::fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n%lld %lld %lld\n", static_cast<long long>(n_rows),
                     static_cast<long long>(n_cols), static_cast<long long>(nnz));
        for (int64_t i = 0; i < nnz; ++i)
            std::fprintf(f, "%lld %lld %.*g\n", static_cast<long long>(rows[i]) + 1, static_cast<long long>(cols[i]) + 1,
                         precision, static_cast<double>(vals[i]));
    } else if (format == 2) {   // array real general: column-major, zeros filled in (mtx_io.cpp:627-655)
        if (precision <= 0) precision = 6;
        std::vector<int64_t> perm(static_cast<size_t>(nnz));
        std::iota(perm.begin(), perm.end(), int64_t(0));
        std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) {
            return cols[a] != cols[b] ? cols[a] < cols[b] : rows[a] < rows[b];
        });
        std::fprintf(f, "%%%%MatrixMarket matrix array real general\n%lld %lld\n", static_cast<long long>(n_rows),
                     static_cast<long long>(n_cols));
        size_t pos = 0;
        for (int64_t j = 0; j < n_cols; ++j)
            for (int64_t i = 0; i < n_rows; ++i) {
                double v = 0.0;
                if (pos < perm.size() && rows[perm[pos]] == i && cols[perm[pos]] == j) v = static_cast<double>(vals[perm[pos++]]);
                std::fprintf(f, "%.*g\n", precision, v);
            }
    } else {
        std::fclose(f);
        return GKOB200_EUNSUPPORTED;
    }
    const bool ok = std::fclose(f) == 0;
    return ok ? 0 : fail("error when writing matrix data");
}

}  // namespace

extern "C" {

const char* gkob200_mtx_last_error(void) { return g_error.c_str(); }

int gkob200_mtx_read_open(const char* path, void** handle, int64_t* n_rows, int64_t* n_cols, int64_t* nnz)
{
    if (!path || !handle || !n_rows || !n_cols || !nnz) return GKOB200_EINVAL;
    *handle = nullptr;
    std::string buf;
    try {
        if (!read_file(path, buf)) return fail(std::string("failed reading from stream: ") + path);
    } catch (const std::exception& e) {
        return fail(std::string("failed reading from stream: ") + e.what());
    }
    if (buf.empty()) return fail("failed reading from stream");
    auto* d = new (std::nothrow) MtxData();
    if (!d) return fail("out of memory");
    // read_generic_raw: a '%' first byte selects the text reader (mtx_io.cpp:911-925)
    int rc;
    try {   // nothing may propagate through the C boundary (bad_alloc / length_error on hostile sizes)
        rc = buf[0] == '%' ? read_text(buf, *d) : read_binary(buf, *d);
    } catch (const std::exception& e) {
        rc = fail(std::string("cannot hold the matrix described by the file: ") + e.what());
    }
    if (rc) {
        delete d;
        return rc;
    }
    *handle = d;
    *n_rows = d->n_rows;
    *n_cols = d->n_cols;
    *nnz = static_cast<int64_t>(d->vals.size());
    return 0;
}

int gkob200_mtx_read_close(void* handle)
{
    delete static_cast<MtxData*>(handle);
    return 0;
}

#define GKOB200_DEF_MTX(V, VT, I, IT)                                                                               \
    int gkob200_mtx_read_copy_##V##_##I(void* handle, IT* rows, IT* cols, VT* vals)                                  \
    { return copy_out<VT, IT>(static_cast<const MtxData*>(handle), rows, cols, vals); }                              \
    int gkob200_mtx_write_##V##_##I(const char* path, int format, int precision, int64_t n_rows, int64_t n_cols,     \
                                    int64_t nnz, const IT* rows, const IT* cols, const VT* vals)                     \
    {                                                                                                                \
        try {                                                                                                        \
            return write_file<VT, IT>(path, format, precision, n_rows, n_cols, nnz, rows, cols, vals);              \
        } catch (const std::exception& e) {                                                                          \
            return fail(std::string("error when writing matrix data: ") + e.what());                                \
        }                                                                                                            \
    }
GKOB200_DEF_MTX(f64, double, i32, int32_t)
GKOB200_DEF_MTX(f32, float, i32, int32_t)
GKOB200_DEF_MTX(f64, double, i64, int64_t)
GKOB200_DEF_MTX(f32, float, i64, int64_t)
#undef GKOB200_DEF_MTX

}  // extern "C"
