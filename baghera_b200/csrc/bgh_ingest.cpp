// bgh_ingest.cpp — native reader of the annotated SNP table, host only.
//
// Replaces the `pd.read_csv` of ref: baghera_tool/snps.py:32-64 for the file written by
// `baghera-tool generate-snp-file` (schema ref: docs/createdata.rst:45-47:
// chr, rs_id, position, cm, maf, l, gene, sample_size, z): the file is memory-mapped, cut into one byte range
// per thread at line boundaries, and every thread parses its lines into typed columns.  The two label
// columns come back dictionary-encoded with the dictionary SORTED, so the gene code of a SNP is already the
// `cat` index of ref: baghera_tool/gene_regression.py:61-71 (rank of the label in sorted order).  A missing
// gene (empty field or one of pandas' default NA spellings) gets code -1 (the reference renames it to
// "NonCoding" afterwards, ref: snps.py:98-102).  Numeric fields are parsed with std::from_chars (correctly
// rounded).  Everything after the parse (filters, stats = z^2, l transform) stays where the reference has it.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "baghera_b200.h"

namespace {

thread_local std::string g_ingest_error;

struct Dict {
    std::vector<std::string> labels;       // sorted
};

}  // namespace

struct bgh_table {
    int64_t n_rows = 0;
    std::vector<std::string> names;                    // header, file order
    std::vector<int> kind;                             // per column: 0 = f64, 1 = label, 2 = skipped
    std::vector<std::vector<double>> f64;              // per column (empty unless kind 0)
    std::vector<std::vector<int32_t>> codes;           // per column (empty unless kind 1)
    std::vector<Dict> dict;                            // per column
};

namespace {

bool is_na(std::string_view s)
{
    // pandas' default na_values (io/parsers: STR_NA_VALUES)
    static const char* const na[] = {"", "#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN",
                                     "<NA>", "N/A", "NA", "NULL", "NaN", "None", "n/a", "nan", "null"};
    for (const char* x : na)
        if (s == x) return true;
    return false;
}

struct Field {
    const char* b;
    const char* e;
};

// splits one line [p, end) into fields; handles "quoted" fields (the quotes are dropped, "" inside is not unescaped)
int split_line(const char* p, const char* end, char sep, Field* out, int max_fields)
{
    int n = 0;
    while (true) {
        const char* b = p;
        const char* e;
        if (p < end && *p == '"') {
            ++b;
            const char* q = b;
            while (q < end && *q != '"') ++q;
            e = q;
            p = q < end ? q + 1 : q;
            while (p < end && *p != sep) ++p;
        } else {
            while (p < end && *p != sep) ++p;
            e = p;
        }
        if (n < max_fields) out[n] = Field{b, e};
        ++n;
        if (p >= end) break;
        ++p;   // separator
        if (p == end) {   // trailing separator: one more empty field
            if (n < max_fields) out[n] = Field{p, p};
            ++n;
            break;
        }
    }
    return n;
}

bool parse_f64(std::string_view s, double* v)
{
    while (!s.empty() && (s.front() == ' ' || s.front() == '\t')) s.remove_prefix(1);
    while (!s.empty() && (s.back() == ' ' || s.back() == '\t')) s.remove_suffix(1);
    if (is_na(s)) {
        *v = std::numeric_limits<double>::quiet_NaN();
        return true;
    }
    if (!s.empty() && s.front() == '+') s.remove_prefix(1);
    const auto r = std::from_chars(s.data(), s.data() + s.size(), *v);
    if (r.ec == std::errc() && r.ptr == s.data() + s.size()) return true;
    if (s == "inf" || s == "Inf" || s == "INF" || s == "infinity") { *v = INFINITY; return true; }
    if (s == "-inf" || s == "-Inf" || s == "-INF" || s == "-infinity") { *v = -INFINITY; return true; }
    return false;
}

struct Chunk {
    std::vector<std::vector<double>> f64;
    std::vector<std::vector<int32_t>> codes;                                  // local codes
    std::vector<std::unordered_map<std::string, int32_t>> local;              // label -> local code
    std::vector<std::vector<std::string>> local_labels;
    int64_t rows = 0;
    std::string err;
    const char* err_at = nullptr;      // start of the offending line (for the file line number of the message)
};

}  // namespace

extern "C" {

const char* bgh_ingest_last_error(void) { return g_ingest_error.c_str(); }

int bgh_csv_read(const char* path, int32_t sep_char, int32_t n_threads, bgh_table** out)
{
    if (!path || !out) { g_ingest_error = "bgh_csv_read: null argument"; return BGH_EINVAL; }
    *out = nullptr;
    const char sep = (char)sep_char;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { g_ingest_error = std::string("bgh_csv_read: cannot open ") + path; return BGH_EINVAL; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); g_ingest_error = "bgh_csv_read: empty or unreadable file"; return BGH_EINVAL; }
    const size_t size = (size_t)st.st_size;
    const char* data = (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (data == MAP_FAILED) { g_ingest_error = "bgh_csv_read: mmap failed"; return BGH_ENOMEM; }
    madvise((void*)data, size, MADV_SEQUENTIAL);
    const char* const end = data + size;

    // ---- header ---------------------------------------------------------------------------------
    const char* nl = (const char*)memchr(data, '\n', size);
    const char* hend = nl ? nl : end;
    const char* body = nl ? nl + 1 : end;
    if (hend > data && hend[-1] == '\r') --hend;
    Field hf[256];
    const int ncol = split_line(data, hend, sep, hf, 256);
    if (ncol < 1 || ncol > 256) { munmap((void*)data, size); g_ingest_error = "bgh_csv_read: bad header"; return BGH_EINVAL; }
    auto* T = new (std::nothrow) bgh_table;
    if (!T) { munmap((void*)data, size); return BGH_ENOMEM; }
    T->names.resize(ncol); T->kind.assign(ncol, 0); T->f64.resize(ncol); T->codes.resize(ncol); T->dict.resize(ncol);
    // first data line: columns outside the schema are typed by what their first value looks like
    std::vector<Field> first(ncol);
    int n_first = 0;
    {
        const char* p = body;
        while (p < end && (*p == '\n' || *p == '\r')) ++p;
        const char* q = p < end ? (const char*)memchr(p, '\n', (size_t)(end - p)) : nullptr;
        const char* le = q ? q : end;
        if (le > p && le[-1] == '\r') --le;
        if (le > p) n_first = split_line(p, le, sep, first.data(), ncol);
    }
    for (int c = 0; c < ncol; ++c) {
        T->names[c].assign(hf[c].b, hf[c].e);
        const std::string& nm = T->names[c];
        // label columns of the schema; identifiers nobody reads downstream are skipped
        if (nm == "gene" || nm == "chr") T->kind[c] = 1;
        else if (nm == "rs_id") T->kind[c] = 2;
        else if (nm == "position" || nm == "cm" || nm == "maf" || nm == "l" || nm == "sample_size" || nm == "z") T->kind[c] = 0;
        else if (n_first == ncol) {
            double v;
            T->kind[c] = parse_f64(std::string_view(first[c].b, (size_t)(first[c].e - first[c].b)), &v) ? 0 : 1;
        }
    }

    // ---- cut the body into line-aligned byte ranges ------------------------------------------------
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    const size_t body_size = (size_t)(end - body);
    if ((size_t)nt > body_size / 65536 + 1) nt = (int)(body_size / 65536 + 1);
    std::vector<const char*> cut(nt + 1);
    cut[0] = body; cut[nt] = end;
    for (int t = 1; t < nt; ++t) {
        const char* p = body + body_size * (size_t)t / (size_t)nt;
        if (p < cut[t - 1]) p = cut[t - 1];
        const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
        cut[t] = q ? q + 1 : end;
    }
    std::vector<Chunk> chunks(nt);
    auto work = [&](int t) {
        Chunk& ch = chunks[t];
        ch.f64.resize(ncol); ch.codes.resize(ncol); ch.local.resize(ncol); ch.local_labels.resize(ncol);
        const size_t guess = (size_t)(cut[t + 1] - cut[t]) / 40 + 16;
        for (int c = 0; c < ncol; ++c) {
            if (T->kind[c] == 0) ch.f64[c].reserve(guess);
            if (T->kind[c] == 1) ch.codes[c].reserve(guess);
        }
        std::vector<Field> f(ncol);
        std::string key;
        const char* p = cut[t];
        while (p < cut[t + 1]) {
            const char* q = (const char*)memchr(p, '\n', (size_t)(cut[t + 1] - p));
            const char* le = q ? q : cut[t + 1];
            const char* next = q ? q + 1 : cut[t + 1];
            if (le > p && le[-1] == '\r') --le;
            if (le == p) { p = next; continue; }   // blank line (pandas skips them)
            const int n = split_line(p, le, sep, f.data(), ncol);
            if (n != ncol) {
                ch.err = "has " + std::to_string(n) + " fields, the header has " + std::to_string(ncol);
                ch.err_at = p;
                return;
            }
            for (int c = 0; c < ncol; ++c) {
                const std::string_view sv(f[c].b, (size_t)(f[c].e - f[c].b));
                if (T->kind[c] == 0) {
                    double v;
                    if (!parse_f64(sv, &v)) {
                        ch.err = "column '" + T->names[c] + "': cannot parse '" + std::string(sv) + "' as a number";
                        ch.err_at = p;
                        return;
                    }
                    ch.f64[c].push_back(v);
                } else if (T->kind[c] == 1) {
                    int32_t code = -1;
                    if (!is_na(sv)) {
                        key.assign(sv);
                        auto it = ch.local.at(c).find(key);
                        if (it == ch.local[c].end()) {
                            code = (int32_t)ch.local_labels[c].size();
                            ch.local[c].emplace(key, code);
                            ch.local_labels[c].push_back(key);
                        } else {
                            code = it->second;
                        }
                    }
                    ch.codes[c].push_back(code);
                }
            }
            ++ch.rows;
            p = next;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < nt; ++t)
        if (!chunks[t].err.empty()) {
            // the FILE line of the offending row (1-based, header = line 1), whichever thread range it fell into
            int64_t line = 1;
            for (const char* c = data; chunks[t].err_at && c < chunks[t].err_at;) {
                const char* q = (const char*)memchr(c, '\n', (size_t)(chunks[t].err_at - c));
                if (!q) break;
                ++line;
                c = q + 1;
            }
            g_ingest_error = "bgh_csv_read: line " + std::to_string(line) + ": " + chunks[t].err;
            munmap((void*)data, size);
            delete T;
            return BGH_EINVAL;
        }
    munmap((void*)data, size);

    // ---- merge: global sorted dictionaries, concatenated columns -------------------------------------
    int64_t rows = 0;
    std::vector<int64_t> row0(nt);
    for (int t = 0; t < nt; ++t) { row0[t] = rows; rows += chunks[t].rows; }
    T->n_rows = rows;
    for (int c = 0; c < ncol; ++c) {
        if (T->kind[c] == 0) {
            T->f64[c].resize((size_t)rows);
        } else if (T->kind[c] == 1) {
            std::vector<std::string> all;
            for (int t = 0; t < nt; ++t) all.insert(all.end(), chunks[t].local_labels[c].begin(), chunks[t].local_labels[c].end());
            std::sort(all.begin(), all.end());            // byte order == Python's str order for ASCII labels
            all.erase(std::unique(all.begin(), all.end()), all.end());
            T->dict[c].labels = std::move(all);
            T->codes[c].resize((size_t)rows);
        }
    }
    auto merge = [&](int t) {
        Chunk& ch = chunks[t];
        for (int c = 0; c < ncol; ++c) {
            if (T->kind[c] == 0) {
                if (ch.rows) memcpy(T->f64[c].data() + row0[t], ch.f64[c].data(), sizeof(double) * (size_t)ch.rows);
            } else if (T->kind[c] == 1) {
                const auto& L = T->dict[c].labels;
                std::vector<int32_t> remap(ch.local_labels[c].size());
                for (size_t i = 0; i < remap.size(); ++i)
                    remap[i] = (int32_t)(std::lower_bound(L.begin(), L.end(), ch.local_labels[c][i]) - L.begin());
                int32_t* dst = T->codes[c].data() + row0[t];
                for (int64_t i = 0; i < ch.rows; ++i) dst[i] = ch.codes[c][(size_t)i] < 0 ? -1 : remap[(size_t)ch.codes[c][(size_t)i]];
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(merge, t);
        merge(0);
        for (auto& x : th) x.join();
    }
    *out = T;
    return BGH_OK;
}

int64_t bgh_table_rows(const bgh_table* t) { return t ? t->n_rows : 0; }
int32_t bgh_table_n_columns(const bgh_table* t) { return t ? (int32_t)t->names.size() : 0; }
const char* bgh_table_column_name(const bgh_table* t, int32_t c)
{
    return (t && c >= 0 && c < (int32_t)t->names.size()) ? t->names[c].c_str() : nullptr;
}
/* 0 = float64 column, 1 = dictionary-encoded label column, 2 = skipped (rs_id), -1 = bad index */
int32_t bgh_table_column_kind(const bgh_table* t, int32_t c)
{
    return (t && c >= 0 && c < (int32_t)t->kind.size()) ? t->kind[c] : -1;
}
const double* bgh_table_f64(const bgh_table* t, int32_t c)
{
    return (t && c >= 0 && c < (int32_t)t->kind.size() && t->kind[c] == 0) ? t->f64[c].data() : nullptr;
}
const int32_t* bgh_table_codes(const bgh_table* t, int32_t c)
{
    return (t && c >= 0 && c < (int32_t)t->kind.size() && t->kind[c] == 1) ? t->codes[c].data() : nullptr;
}
int32_t bgh_table_n_labels(const bgh_table* t, int32_t c)
{
    return (t && c >= 0 && c < (int32_t)t->kind.size() && t->kind[c] == 1) ? (int32_t)t->dict[c].labels.size() : 0;
}
const char* bgh_table_label(const bgh_table* t, int32_t c, int32_t i)
{
    if (!t || c < 0 || c >= (int32_t)t->kind.size() || t->kind[c] != 1) return nullptr;
    return (i >= 0 && i < (int32_t)t->dict[c].labels.size()) ? t->dict[c].labels[(size_t)i].c_str() : nullptr;
}
void bgh_table_free(bgh_table* t) { delete t; }

}  // extern "C"
