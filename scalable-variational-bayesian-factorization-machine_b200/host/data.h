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
#include <vector>

namespace svbfm_host {

struct SparseMatrix {          // rows x cols, row-compressed
    uint32_t num_rows = 0, num_cols = 0;
    std::vector<uint64_t> ptr{0};
    std::vector<uint32_t> id;
    std::vector<float> val;
    uint64_t nnz() const { return id.size(); }
};

inline bool file_exists(const std::string& f) { std::ifstream in(f.c_str()); return in.is_open(); }

// 24-byte header of .x / .xt (fmatrix.h:46-52)
#pragma pack(push, 1)
struct XFileHeader { uint32_t id, float_size; uint64_t num_values; uint32_t num_rows, num_cols; };
#pragma pack(pop)

inline void write_x_file(const std::string& path, const SparseMatrix& m) {
    std::ofstream out(path.c_str(), std::ios::binary);
    if (!out.is_open()) throw "could not open " + path;
    XFileHeader h{2, 4, m.nnz(), m.num_rows, m.num_cols};
    out.write(reinterpret_cast<const char*>(&h), sizeof(h));
    std::vector<char> buf;
    for (uint32_t r = 0; r < m.num_rows; r++) {
        uint32_t size = (uint32_t)(m.ptr[r + 1] - m.ptr[r]);
        buf.resize(4 + (size_t)size * 8);
        memcpy(buf.data(), &size, 4);
        for (uint32_t k = 0; k < size; k++) {
            memcpy(buf.data() + 4 + (size_t)k * 8, &m.id[m.ptr[r] + k], 4);
            memcpy(buf.data() + 8 + (size_t)k * 8, &m.val[m.ptr[r] + k], 4);
        }
        out.write(buf.data(), (std::streamsize)buf.size());
    }
}

inline void read_x_file(const std::string& path, SparseMatrix& m) {
    std::ifstream in(path.c_str(), std::ios::binary);
    if (!in.is_open()) throw "could not open " + path;
    XFileHeader h;
    in.read(reinterpret_cast<char*>(&h), sizeof(h));
    if (!in || h.id != 2 || h.float_size != 4) throw "bad header in " + path;
    m.num_rows = h.num_rows; m.num_cols = h.num_cols;
    m.ptr.assign((size_t)h.num_rows + 1, 0);
    m.id.resize(h.num_values); m.val.resize(h.num_values);
    uint64_t w = 0;
    std::vector<char> buf;
    for (uint32_t r = 0; r < h.num_rows; r++) {
        uint32_t size = 0;
        in.read(reinterpret_cast<char*>(&size), 4);
        if (!in || w + size > h.num_values) throw "truncated file " + path;
        buf.resize((size_t)size * 8);
        in.read(buf.data(), (std::streamsize)buf.size());
        for (uint32_t k = 0; k < size; k++) {
            memcpy(&m.id[w], buf.data() + (size_t)k * 8, 4);
            memcpy(&m.val[w], buf.data() + (size_t)k * 8 + 4, 4);
            w++;
        }
        m.ptr[r + 1] = w;
    }
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

struct DataSet {
    bool has_x = true, has_xt = true;
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
        for (uint32_t c : in.id) {
            if (c >= out_rows) throw std::string("feature id out of range in transpose");
            out.ptr[c + 1]++;
        }
        for (uint32_t j = 0; j < out_rows; j++) out.ptr[j + 1] += out.ptr[j];
        out.id.resize(in.nnz()); out.val.resize(in.nnz());
        std::vector<uint64_t> fill(out.ptr.begin(), out.ptr.end() - 1);
        for (uint32_t r = 0; r < in.num_rows; r++)
            for (uint64_t p = in.ptr[r]; p < in.ptr[r + 1]; p++) {
                uint64_t d = fill[in.id[p]]++;
                out.id[d] = r; out.val[d] = in.val[p];
            }
    }

    void scan_targets() {
        min_target = +FLT_MAX; max_target = -FLT_MAX;
        for (float t : target) { if (t < min_target) min_target = t; if (t > max_target) max_target = t; }
        num_cases = (uint32_t)target.size();
    }

    // text branch of Data::load (Data.h:173-283); forced_num_feature > 0 = the (file, num_attribute) overload (:287-454)
    void load_text(const std::string& filename, uint32_t forced_num_feature = 0) {
        std::ifstream f(filename.c_str());
        if (!f.is_open()) throw "unable to open " + filename;
        x = SparseMatrix();
        target.clear();
        long max_id = 0; bool has_feature = false;
        std::string line;
        while (std::getline(f, line)) {
            const char* p = line.c_str();
            if (LineParser::skip_line(p)) continue;
            float t;
            if (!LineParser::scan_float(p, t)) throw "cannot parse line \"" + line + "\" at character " + p[0];
            target.push_back(t);
            long id; float v;
            while (LineParser::scan_pair(p, id, v)) {
                x.id.push_back((uint32_t)id); x.val.push_back(v);
                if (id > max_id) max_id = id;
                has_feature = true;
            }
            x.ptr.push_back(x.id.size());
            LineParser::finish(p, line);
        }
        num_feature = forced_num_feature ? (int)forced_num_feature : (int)(has_feature ? max_id + 1 : 0);
        x.num_rows = (uint32_t)target.size(); x.num_cols = (uint32_t)num_feature;
        scan_targets();
        if (!forced_num_feature)
            std::cout << "num_rows=" << x.num_rows << "\tnum_values=" << x.nnz() << "\tnum_features=" << num_feature << "\tmin_target=" << min_target
                      << "\tmax_target=" << max_target << std::endl;
        if (has_xt) transpose(x, (uint32_t)num_feature, xt);
    }

    // Data::load (Data.h:106-171): binary files next to `filename` win over the text file
    void load(const std::string& filename) {
        int from = 0;
        if ((!has_x || file_exists(filename + ".data")) && (!has_xt || file_exists(filename + ".datat")) && file_exists(filename + ".target")) from = 1;
        else if ((!has_x || file_exists(filename + ".x")) && (!has_xt || file_exists(filename + ".xt")) && file_exists(filename + ".y")) from = 2;
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
