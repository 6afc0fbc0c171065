// host/tools/transpose.cpp -- binary .x -> .xt (features x cases), byte-identical to the reference tool
// (src/libfm/tools/transpose.cpp:54-172): `transpose --ifile <in.x> --ofile <out.xt> [--cache_size n]`.
// The reference fills as many output rows as fit in its cache per pass over the input; the result does not depend
// on the cache size, so this tool does one in-memory counting transpose (-cache_size is accepted and ignored).
#include <iostream>
#include "../cmdline.h"
#include "../data.h"

#include <dlfcn.h>
using namespace svbfm_host;

static std::string exe_dir(const char* argv0) {
    char buf[4096];
    ssize_t n = readlink("/proc/self/exe", buf, sizeof(buf) - 1);
    std::string p = n > 0 ? std::string(buf, (size_t)n) : std::string(argv0);
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string(".") : p.substr(0, k);
}

int main(int argc, char** argv) {
    try {
        CmdLine cmd(argc, argv);
        const std::string p_in = cmd.reg("ifile", "input file name, file has to be in binary sparse format [MANDATORY]");
        const std::string p_out = cmd.reg("ofile", "output file name [MANDATORY]");
        cmd.reg("cache_size", "accepted for compatibility; the transpose is done in memory");
        const std::string p_dev = cmd.reg("device", "CUDA device ordinal: transpose on the GPU (libsvbfm.so is loaded at run time); default: host threads");
        const std::string p_help = cmd.reg("help", "this screen");
        if (cmd.has(p_help) || argc == 1) { cmd.print_help(); return 0; }
        cmd.check();
        SparseMatrix in, out;
        read_x_file(cmd.get(p_in), in);
        std::cout << "num_rows=" << in.num_rows << "\tnum_values=" << in.nnz() << "\tnum_features=" << in.num_cols << std::endl;
        if (cmd.has(p_dev)) {                         // the engine's device transpose (svbfm_transpose_csr): same bytes
            typedef int (*fn_t)(int32_t, uint32_t, uint32_t, const uint64_t*, const uint32_t*, const float*, uint64_t*, uint32_t*, float*);
            typedef const char* (*err_t)(const void*);
            const char* lib = getenv("SVBFM_LIB");
            std::string path = lib ? lib : (exe_dir(argv[0]) + "/../libsvbfm.so");
            void* so = dlopen(path.c_str(), RTLD_NOW);
            if (!so) throw std::string("cannot load ") + path + ": " + dlerror();
            fn_t fn = (fn_t)dlsym(so, "svbfm_transpose_csr");
            err_t last = (err_t)dlsym(so, "svbfm_last_error");
            if (!fn || !last) throw std::string("svbfm_transpose_csr is missing from ") + path;
            out.num_rows = in.num_cols; out.num_cols = in.num_rows;
            out.ptr.assign((size_t)in.num_cols + 1, 0); out.id.resize(in.nnz()); out.val.resize(in.nnz());
            if (fn((int32_t)cmd.get_int(p_dev, 0), in.num_rows, in.num_cols, in.ptr.data(), in.id.data(), in.val.data(), out.ptr.data(), out.id.data(), out.val.data()) != 0)
                throw std::string(last(nullptr));
        } else
        DataSet::transpose(in, in.num_cols, out);     // rows = features, ids = case ids ascending (transpose.cpp:129-162)
        std::cout << "output to " << cmd.get(p_out) << std::endl;
        write_x_file(cmd.get(p_out), out);            // header with num_rows / num_cols swapped (transpose.cpp:104-110)
    } catch (std::string& e) {
        std::cerr << e << std::endl;
    }
    return 0;
}
