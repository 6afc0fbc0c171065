// host/tools/transpose.cpp -- binary .x -> .xt (features x cases), byte-identical to the reference tool
// (src/libfm/tools/transpose.cpp:54-172): `transpose --ifile <in.x> --ofile <out.xt> [--cache_size n]`.
// The reference fills as many output rows as fit in its cache per pass over the input; the result does not depend
// on the cache size, so this tool does one in-memory counting transpose (-cache_size is accepted and ignored).
#include <iostream>
#include "../cmdline.h"
#include "../data.h"

using namespace svbfm_host;

int main(int argc, char** argv) {
    try {
        CmdLine cmd(argc, argv);
        const std::string p_in = cmd.reg("ifile", "input file name, file has to be in binary sparse format [MANDATORY]");
        const std::string p_out = cmd.reg("ofile", "output file name [MANDATORY]");
        cmd.reg("cache_size", "accepted for compatibility; the transpose is done in memory");
        const std::string p_help = cmd.reg("help", "this screen");
        if (cmd.has(p_help) || argc == 1) { cmd.print_help(); return 0; }
        cmd.check();
        SparseMatrix in, out;
        read_x_file(cmd.get(p_in), in);
        std::cout << "num_rows=" << in.num_rows << "\tnum_values=" << in.nnz() << "\tnum_features=" << in.num_cols << std::endl;
        DataSet::transpose(in, in.num_cols, out);     // rows = features, ids = case ids ascending (transpose.cpp:129-162)
        std::cout << "output to " << cmd.get(p_out) << std::endl;
        write_x_file(cmd.get(p_out), out);            // header with num_rows / num_cols swapped (transpose.cpp:104-110)
    } catch (std::string& e) {
        std::cerr << e << std::endl;
    }
    return 0;
}
