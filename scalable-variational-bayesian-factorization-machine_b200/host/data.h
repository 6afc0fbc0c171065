// host/data.h -- libFM data sets on the host: text parser, binary .x/.xt/.y (.data/.datat/.target) files and the
// in-memory counting transpose. Formats and indexing follow the reference bit for bit
// (src/libfm/src/Data.h:106-283, 457-509; src/util/fmatrix.h:46-86, 157-172; src/util/matrix.h:280-328);
// the containers are flat arrays (ptr / id / val) instead of the reference's row-pointer objects.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace svbfm_host {

struct SparseMatrix {          // rows x cols, row-compressed
    uint32_t num_rows = 0, num_cols = 0;
    std::vector<uint64_t> ptr{0};
    std::vector<uint32_t> id;
    std::vector<float> val;
    uint64_t nnz() const { return id.size(); }
};

inline bool file_exists(const std::string& f) { std::ifstream in(f.c_str()); return in.is_open(); }

// Host threads of the loaders (text parser, binary reader, transpose): SVBFM_HOST_THREADS, else the hardware's count.
// The results do not depend on it: every thread owns a contiguous range of lines / cases and the pieces are joined in order.
inline unsigned host_threads() {
    long t = 0;
    if (const char* e = getenv("SVBFM_HOST_THREADS")) t = atol(e);
    if (t <= 0) t = (long)std::thread::hardware_concurrency();
    return (unsigned)std::min<long>(std::max<long>(t, 1), 64);
}
template <class F>
inline void parallel_chunks(unsigned nthreads, F&& body) {        // body(t) for t in [0, nthreads), one thread each
    if (nthreads <= 1) { body(0u); return; }
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nthreads; t++) th.emplace_back([&body, t] { body(t); });
    body(0u);
    for (auto& x : th) x.join();
}

// A whole file mapped read-only (the loaders work on the bytes in place).
struct MappedFile {
    const char* data = nullptr;
    size_t size = 0;
    int fd = -1;
    bool open(const std::string& path) {
        fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); fd = -1; return false; }
        size = (size_t)st.st_size;
        if (size) {
            void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p == MAP_FAILED) { ::close(fd); fd = -1; return false; }
            madvise(p, size, MADV_SEQUENTIAL);
            data = (const char*)p;
        }
        return true;
    }
    ~MappedFile() {
        if (data) munmap((void*)data, size);
        if (fd >= 0) ::close(fd);
    }
};

// 24-byte header of .x / .xt (fmatrix.h:46-52)
#pragma pack(push, 1)
struct XFileHeader { uint32_t id, float_size; uint64_t num_values; uint32_t num_rows, num_cols; };
#pragma pack(pop)

inline void write_x_file(const std::string& path, const SparseMatrix& m) {
    std::ofstream out(path.c_str(), std::ios::binary);
    if (!out.is_open()) throw "could not open " + path;
    XFileHeader h{2, 4, m.nnz(), m.num_rows, m.num_cols};
    out.write(reinterpret_cast<const char*>(&h), sizeof(h));
    std::vector<char> buf;                       // rows are staged in blocks of ~4 MB: {uint size; {uint id; float value}[size]} each
    buf.reserve((4u << 20) + 64);
    for (uint32_t r = 0; r < m.num_rows; r++) {
        uint32_t size = (uint32_t)(m.ptr[r + 1] - m.ptr[r]);
        size_t at = buf.size();
        buf.resize(at + 4 + (size_t)size * 8);
        char* b = buf.data() + at;
        memcpy(b, &size, 4);
        for (uint32_t k = 0; k < size; k++) {
            memcpy(b + 4 + (size_t)k * 8, &m.id[m.ptr[r] + k], 4);
            memcpy(b + 8 + (size_t)k * 8, &m.val[m.ptr[r] + k], 4);
        }
        if (buf.size() >= (4u << 20)) { out.write(buf.data(), (std::streamsize)buf.size()); buf.clear(); }
    }
    if (!buf.empty()) out.write(buf.data(), (std::streamsize)buf.size());
}

inline void read_x_file(const std::string& path, SparseMatrix& m) {
    MappedFile f;
    if (!f.open(path)) throw "could not open " + path;
    XFileHeader h;
    if (f.size < sizeof(h)) throw "bad header in " + path;
    memcpy(&h, f.data, sizeof(h));
    if (h.id != 2 || h.float_size != 4) throw "bad header in " + path;
    m.num_rows = h.num_rows; m.num_cols = h.num_cols;
    m.ptr.assign((size_t)h.num_rows + 1, 0);
    // pass 1 (sequential, reads one size field per row): row pointers + the byte offset where every thread's block of rows starts
    const unsigned T = (unsigned)std::min<uint64_t>(host_threads(), std::max<uint32_t>(h.num_rows / 65536u, 1u));
    std::vector<size_t> start_at(T + 1, 0);
    std::vector<uint32_t> start_row(T + 1, h.num_rows);
    for (unsigned t = 0; t < T; t++) start_row[t] = (uint32_t)((uint64_t)h.num_rows * t / T);
    size_t at = sizeof(h);
    uint64_t w = 0;
    unsigned next = 0;
    for (uint32_t r = 0; r < h.num_rows; r++) {
        while (next < T && start_row[next] == r) start_at[next++] = at;
        uint32_t size;
        if (at + 4 > f.size) throw "truncated file " + path;
        memcpy(&size, f.data + at, 4);
        if (w + size > h.num_values || at + 4 + (size_t)size * 8 > f.size) throw "truncated file " + path;
        at += 4 + (size_t)size * 8;
        w += size;
        m.ptr[r + 1] = w;
    }
    while (next <= T) start_at[next++] = at;
    // pass 2 (parallel): de-interleave {id, value} pairs
    m.id.resize(h.num_values); m.val.resize(h.num_values);
    parallel_chunks(T, [&](unsigned t) {
        size_t a = start_at[t];
        for (uint32_t r = start_row[t]; r < start_row[t + 1]; r++) {
            const uint32_t size = (uint32_t)(m.ptr[r + 1] - m.ptr[r]);
            const char* b = f.data + a + 4;
            uint32_t* id = m.id.data() + m.ptr[r];
            float* val = m.val.data() + m.ptr[r];
            for (uint32_t k = 0; k < size; k++) { memcpy(id + k, b + (size_t)k * 8, 4); memcpy(val + k, b + (size_t)k * 8 + 4, 4); }
            a += 4 + (size_t)size * 8;
        }
    });
}

// .y: 12-byte header {version=1, data_size=4, num_rows} + floats (matrix.h:280-294)
inline void write_y_file(const std::string& path, const std::vector<float>& y) {
    std::ofstream out(path.c_str(), std::ios::binary);
    if (!out.is_open()) throw "unable to open " + path;
    uint32_t hdr[3] = {1, 4, (uint32_t)y.size()};
    out.write(reinterpret_cast<const char*>(hdr), 12);
    out.write(reinterpret_cast<const char*>(y.data()), (std::streamsize)y.size() * 4);
}
inline void read_y_file(const std::string& path, std::vector<float>& y) {
    std::ifstream in(path.c_str(), std::ios::binary);
    if (!in.is_open()) throw "unable to open " + path;
    uint32_t hdr[3];
    in.read(reinterpret_cast<char*>(hdr), 12);
    if (!in || hdr[0] != 1 || hdr[1] != 4) throw "bad header in " + path;
    y.resize(hdr[2]);
    in.read(reinterpret_cast<char*>(y.data()), (std::streamsize)y.size() * 4);
}

// One libFM text line -> target + (id, value) pairs with the reference's acceptance rules
// (Data.h:192-216): leading blanks skipped; empty and '#' lines ignored (returns false); `%f` target; `%d:%f`
// pairs; only blanks or a '#' comment may follow.
struct LineParser {
    static bool skip_line(const char*& p) {
        while (*p == ' ' || *p == '\t') p++;
        return *p == 0 || *p == '#';
    }
    static bool scan_float(const char*& p, float& v) {       // sscanf("%f"): leading whitespace skipped
        const char* q = p;
        while (*q == ' ' || *q == '\t' || *q == '\r' || *q == '\v' || *q == '\f') q++;
        char* end = nullptr;
        float f = strtof(q, &end);
        if (end == q) return false;
        v = f; p = end;
        return true;
    }
    static bool scan_pair(const char*& p, long& id, float& v) {   // sscanf("%d:%f")
        const char* q = p;
        while (*q == ' ' || *q == '\t' || *q == '\r' || *q == '\v' || *q == '\f') q++;
        char* end = nullptr;
        long i = strtol(q, &end, 10);
        if (end == q || *end != ':') return false;
        const char* r = end + 1;
        float f;
        if (!scan_float(r, f)) return false;
        id = i; v = f; p = r;
        return true;
    }
    static void finish(const char* p, const std::string& line) {
        while (*p == ' ' || *p == '\t') p++;
        if (*p != 0 && *p != '#') throw "cannot parse line \"" + line + "\" at character " + p[0];
    }
};

// Parallel text parser. The file is mapped, cut into one contiguous range of lines per thread, and every range is parsed into its
// own arrays, which are joined in file order: the result is what the sequential getline loop of Data::load (Data.h:185-278) builds.
// A line made of plain tokens (`[-]digits` target, `digits:digits` pairs with at most 7 digits of value, blanks or tabs between
// them) is read by a hand-written scanner whose results are exact; every other line (decimal points, exponents, '\r', comments after
// the pairs, anything malformed) goes through LineParser, i.e. strtof / strtol like the reference's sscanf, and throws like it.
struct ParsedText {
    std::vector<float> target;
    std::vector<uint64_t> ptr{0};
    std::vector<uint32_t> id;
    std::vector<float> val;
    long max_id = 0;
    bool has_feature = false;
};

struct TextChunk {
    std::vector<float> target;
    std::vector<uint32_t> row_nnz;
    std::vector<uint32_t> id;
    std::vector<float> val;
    long max_id = 0;
    bool has_feature = false;
    bool failed = false;
    std::string error;

    static inline bool plain_uint(const char*& p, const char* e, int max_digits, long& out) {
        const char* q = p;
        long v = 0;
        while (q < e && (unsigned)(*q - '0') <= 9u && q - p < max_digits) v = v * 10 + (*q++ - '0');
        if (q == p || (q < e && (unsigned)(*q - '0') <= 9u)) return false;
        out = v; p = q;
        return true;
    }
    // the plain form of one line; false = not plain (nothing was committed)
    inline bool fast_line(const char* p, const char* e, bool keep) {
        bool neg = false;
        if (*p == '-') { neg = true; p++; }
        long t;
        if (!plain_uint(p, e, 7, t) || (p < e && *p != ' ' && *p != '\t')) return false;
        const size_t id0 = id.size();
        long mx = max_id; bool any = false; uint32_t n = 0;
        for (;;) {
            while (p < e && (*p == ' ' || *p == '\t')) p++;
            if (p == e) break;
            long i, v;
            if (!plain_uint(p, e, 9, i) || p == e || *p != ':') goto not_plain;
            p++;
            if (!plain_uint(p, e, 7, v) || (p < e && *p != ' ' && *p != '\t')) goto not_plain;
            if (keep) { id.push_back((uint32_t)i); val.push_back((float)v); }
            if (i > mx) mx = i;
            any = true; n++;
        }
        target.push_back(neg ? -(float)t : (float)t);
        if (keep) row_nnz.push_back(n);
        max_id = mx; has_feature = has_feature || any;
        return true;
    not_plain:
        id.resize(id0); val.resize(id0);
        return false;
    }
    inline void slow_line(const char* b, const char* e, bool keep) {
        std::string line(b, e);
        const char* p = line.c_str();
        if (LineParser::skip_line(p)) return;
        float t;
        if (!LineParser::scan_float(p, t)) throw "cannot parse line \"" + line + "\" at character " + p[0];
        target.push_back(t);
        long i; float v; uint32_t n = 0;
        while (LineParser::scan_pair(p, i, v)) {
            if (keep) { id.push_back((uint32_t)i); val.push_back(v); }
            if (i > max_id) max_id = i;
            has_feature = true; n++;
        }
        if (keep) row_nnz.push_back(n);
        LineParser::finish(p, line);
    }
    void parse(const char* b, const char* e, bool keep) {
        try {
            while (b < e) {
                const char* nl = (const char*)memchr(b, '\n', (size_t)(e - b));
                const char* le = nl ? nl : e;
                const char* p = b;
                while (p < le && (*p == ' ' || *p == '\t')) p++;
                if (p < le && *p != '#' && !fast_line(p, le, keep)) slow_line(b, le, keep);
                b = le + 1;
            }
        } catch (std::string& err) { failed = true; error = err; }
    }
};

inline void parse_text_file(const std::string& filename, ParsedText& out, bool keep_entries) {
    MappedFile f;
    if (!f.open(filename)) throw "unable to open " + filename;
    const unsigned T = (unsigned)std::min<size_t>(host_threads(), std::max<size_t>(f.size >> 20, 1));
    std::vector<size_t> cut(T + 1, f.size);
    cut[0] = 0;
    for (unsigned t = 1; t < T; t++) {                       // a range starts right after a newline
        size_t a = std::max(f.size * t / T, cut[t - 1]);
        const char* nl = a < f.size ? (const char*)memchr(f.data + a, '\n', f.size - a) : nullptr;
        cut[t] = nl ? (size_t)(nl - f.data) + 1 : f.size;
    }
    std::vector<TextChunk> ch(T);
    parallel_chunks(T, [&](unsigned t) {
        size_t bytes = cut[t + 1] - cut[t];
        ch[t].target.reserve(bytes / 12 + 16);
        if (keep_entries) { ch[t].row_nnz.reserve(bytes / 12 + 16); ch[t].id.reserve(bytes / 6 + 16); ch[t].val.reserve(bytes / 6 + 16); }
        ch[t].parse(f.data + cut[t], f.data + cut[t + 1], keep_entries);
    });
    for (auto& c : ch) if (c.failed) throw c.error;           // the first bad line in file order, like the sequential loop
    std::vector<uint64_t> row0(T + 1, 0), ent0(T + 1, 0);
    for (unsigned t = 0; t < T; t++) { row0[t + 1] = row0[t] + ch[t].target.size(); ent0[t + 1] = ent0[t] + ch[t].id.size(); }
    out.target.resize(row0[T]);
    out.max_id = 0; out.has_feature = false;
    for (auto& c : ch) { if (c.max_id > out.max_id) out.max_id = c.max_id; out.has_feature = out.has_feature || c.has_feature; }
    if (keep_entries) { out.ptr.resize(row0[T] + 1); out.ptr[0] = 0; out.id.resize(ent0[T]); out.val.resize(ent0[T]); }
    parallel_chunks(T, [&](unsigned t) {
        TextChunk& c = ch[t];
        if (!c.target.empty()) memcpy(out.target.data() + row0[t], c.target.data(), c.target.size() * 4);
        if (!keep_entries) return;
        if (!c.id.empty()) { memcpy(out.id.data() + ent0[t], c.id.data(), c.id.size() * 4); memcpy(out.val.data() + ent0[t], c.val.data(), c.val.size() * 4); }
        uint64_t w = ent0[t];
        for (size_t r = 0; r < c.row_nnz.size(); r++) { w += c.row_nnz[r]; out.ptr[row0[t] + r + 1] = w; }
        std::vector<float>().swap(c.target); std::vector<uint32_t>().swap(c.id); std::vector<float>().swap(c.val); std::vector<uint32_t>().swap(c.row_nnz);
    });
}

struct DataSet {
    bool has_x = true, has_xt = true;
    // device_transpose: the learner takes the cases row-wise (svbfm_set_csr) and the DEVICE builds the transposed matrix (SURVEY
    // section 8f rank 1): a text file is parsed into `x` only, binary input needs <name>.x + <name>.y only (<name>.xt is used when
    // it is there: nothing to transpose then). csr_only says which of the two the load ended with.
    bool device_transpose = false, csr_only = false;
    SparseMatrix x;       // cases x features (CSR of X)
    SparseMatrix xt;      // features x cases (CSC of X), case ids ascending inside a feature
    std::vector<float> target;
    int num_feature = 0;
    uint32_t num_cases = 0;
    float min_target = +FLT_MAX, max_target = -FLT_MAX;

    DataSet(bool has_x_, bool has_xt_) : has_x(has_x_), has_xt(has_xt_) {}

    // Data::create_data_t (Data.h:457-509): counting transpose, case order kept inside each feature
    static void transpose(const SparseMatrix& in, uint32_t out_rows, SparseMatrix& out) {
        out.num_rows = out_rows; out.num_cols = in.num_rows;
        out.ptr.assign((size_t)out_rows + 1, 0);
        // every thread owns a contiguous block of cases: per-(thread, feature) counts, one scan over (feature, thread), then every
        // thread scatters its block into its own slots -- cases stay ascending inside a feature, as in the sequential transpose
        unsigned T = (unsigned)std::min<uint64_t>(host_threads(), std::max<uint64_t>(in.nnz() >> 18, 1));
        while (T > 1 && (uint64_t)T * out_rows * 8 > (1ull << 31)) T--;          // bound the counters (2 GB)
        std::vector<uint32_t> row_cut(T + 1, in.num_rows);
        for (unsigned t = 0; t < T; t++) row_cut[t] = (uint32_t)((uint64_t)in.num_rows * t / T);
        std::vector<std::vector<uint64_t>> cnt(T);
        std::vector<char> bad(T, 0);
        parallel_chunks(T, [&](unsigned t) {
            cnt[t].assign(out_rows, 0);
            for (uint64_t p = in.ptr[row_cut[t]]; p < in.ptr[row_cut[t + 1]]; p++) {
                uint32_t c = in.id[p];
                if (c >= out_rows) { bad[t] = 1; return; }
                cnt[t][c]++;
            }
        });
        for (char b : bad) if (b) throw std::string("feature id out of range in transpose");
        uint64_t w = 0;
        for (uint32_t j = 0; j < out_rows; j++) {
            for (unsigned t = 0; t < T; t++) { uint64_t c = cnt[t][j]; cnt[t][j] = w; w += c; }
            out.ptr[j + 1] = w;
        }
        out.id.resize(in.nnz()); out.val.resize(in.nnz());
        parallel_chunks(T, [&](unsigned t) {
            std::vector<uint64_t>& fill = cnt[t];
            for (uint32_t r = row_cut[t]; r < row_cut[t + 1]; r++)
                for (uint64_t p = in.ptr[r]; p < in.ptr[r + 1]; p++) {
                    uint64_t d = fill[in.id[p]]++;
                    out.id[d] = r; out.val[d] = in.val[p];
                }
        });
    }

    void scan_targets() {
        min_target = +FLT_MAX; max_target = -FLT_MAX;
        for (float t : target) { if (t < min_target) min_target = t; if (t > max_target) max_target = t; }
        num_cases = (uint32_t)target.size();
    }

    // text branch of Data::load (Data.h:173-283); forced_num_feature > 0 = the (file, num_attribute) overload (:287-454)
    void load_text(const std::string& filename, uint32_t forced_num_feature = 0) {
        ParsedText t;
        parse_text_file(filename, t, true);
        x = SparseMatrix();
        x.ptr.swap(t.ptr); x.id.swap(t.id); x.val.swap(t.val);
        target.swap(t.target);
        const long max_id = t.max_id; const bool has_feature = t.has_feature;
        num_feature = forced_num_feature ? (int)forced_num_feature : (int)(has_feature ? max_id + 1 : 0);
        x.num_rows = (uint32_t)target.size(); x.num_cols = (uint32_t)num_feature;
        scan_targets();
        if (!forced_num_feature)
            std::cout << "num_rows=" << x.num_rows << "\tnum_values=" << x.nnz() << "\tnum_features=" << num_feature << "\tmin_target=" << min_target
                      << "\tmax_target=" << max_target << std::endl;
        csr_only = device_transpose;
        if (has_xt && !csr_only) transpose(x, (uint32_t)num_feature, xt);
    }

    // Data::load (Data.h:106-171): binary files next to `filename` win over the text file
    void load(const std::string& filename) {
        int from = 0;
        if ((!has_x || file_exists(filename + ".data")) && (!has_xt || file_exists(filename + ".datat")) && file_exists(filename + ".target")) from = 1;
        else if ((!has_x || file_exists(filename + ".x")) && (!has_xt || file_exists(filename + ".xt")) && file_exists(filename + ".y")) from = 2;
        csr_only = false;
        if (from == 0 && device_transpose) {      // the rows alone will do: <name>.data / <name>.x without the transposed file
            const bool legacy = file_exists(filename + ".data") && file_exists(filename + ".target");
            if (legacy || (file_exists(filename + ".x") && file_exists(filename + ".y"))) {
                read_y_file(filename + (legacy ? ".target" : ".y"), target);
                read_x_file(filename + (legacy ? ".data" : ".x"), x);
                if (target.size() != x.num_rows) throw std::string("target and data disagree on the number of cases");
                num_feature = (int)x.num_cols;
                csr_only = true;
                scan_targets();
                std::cout << "num_cases=" << num_cases << "\tnum_values=" << x.nnz() << "\tnum_features=" << num_feature << "\tmin_target=" << min_target
                          << "\tmax_target=" << max_target << std::endl;
                return;
            }
        }
        if (from == 0) { load_text(filename); return; }
        read_y_file(filename + (from == 1 ? ".target" : ".y"), target);
        uint64_t num_values = 0;
        if (has_x) {
            read_x_file(filename + (from == 1 ? ".data" : ".x"), x);
            if (target.size() != x.num_rows) throw std::string("target and data disagree on the number of cases");
            num_feature = (int)x.num_cols; num_values = x.nnz();
        }
        if (has_xt) {
            read_x_file(filename + (from == 1 ? ".datat" : ".xt"), xt);
            num_feature = (int)xt.num_rows; num_values = xt.nnz();
            if (has_x && (x.num_cols != xt.num_rows || x.num_rows != xt.num_cols || x.nnz() != xt.nnz()))
                throw std::string("data and its transpose disagree");
        }
        scan_targets();
        std::cout << "num_cases=" << num_cases << "\tnum_values=" << num_values << "\tnum_features=" << num_feature << "\tmin_target=" << min_target
                  << "\tmax_target=" << max_target << std::endl;
    }

    void debug() const {
        if (!has_x) return;
        for (uint32_t r = 0; r < x.num_rows && r < 4; r++) {
            std::cout << target[r];
            for (uint64_t p = x.ptr[r]; p < x.ptr[r + 1]; p++) std::cout << " " << x.id[p] << ":" << x.val[p];
            std::cout << std::endl;
        }
    }
};

// DataMetaInfo (Data.h:35-69): attribute -> group, group sizes over ALL attributes
struct MetaInfo {
    std::vector<uint32_t> attr_group;
    uint32_t num_attr_groups = 1;
    std::vector<uint32_t> num_attr_per_group;
    explicit MetaInfo(uint32_t num_attributes) : attr_group(num_attributes, 0), num_attr_per_group(1, num_attributes) {}
    void load_groups(const std::string& filename) {
        std::ifstream in(filename.c_str());
        if (!in.is_open()) throw "Unable to open file " + filename;
        for (auto& g : attr_group) { uint32_t v = 0; in >> v; g = v; }
        num_attr_groups = 0;
        for (uint32_t g : attr_group) if (g + 1 > num_attr_groups) num_attr_groups = g + 1;
        num_attr_per_group.assign(num_attr_groups, 0);
        for (uint32_t g : attr_group) num_attr_per_group[g]++;
    }
    void debug() const {
        std::cout << "#attr=" << attr_group.size() << "\t#groups=" << num_attr_groups << std::endl;
        for (uint32_t g = 0; g < num_attr_groups; g++) std::cout << "#attr_in_group[" << g << "]=" << num_attr_per_group[g] << std::endl;
    }
};

}  // namespace svbfm_host
