// host/tools/convert.cpp -- libFM text -> binary .x (cases x features) + .y, byte-identical to the reference tool
// (src/libfm/tools/convert.cpp:55-205): `convert --ifile <text> --ofilex <out.x> --ofiley <out.y>`.
#include <iostream>
#include "../cmdline.h"
#include "../data.h"

using namespace svbfm_host;

int main(int argc, char** argv) {
    try {
        CmdLine cmd(argc, argv);
        const std::string p_in = cmd.reg("ifile", "input file name (libFM text format) [MANDATORY]");
        const std::string p_x = cmd.reg("ofilex", "output file name for x [MANDATORY]");
        const std::string p_y = cmd.reg("ofiley", "output file name for y [MANDATORY]");
        const std::string p_help = cmd.reg("help", "this screen");
        if (cmd.has(p_help) || argc == 1) { cmd.print_help(); return 0; }
        cmd.check();
        DataSet d(true, false);
        d.load_text(cmd.get(p_in));                   // prints num_rows / num_values / num_features like the reference (convert.cpp:136)
        write_x_file(cmd.get(p_x), d.x);              // header {2, 4, num_values, num_rows, num_cols = max id + 1} (convert.cpp:147-153)
        write_y_file(cmd.get(p_y), d.target);         // convert.cpp:159-163, 177
    } catch (std::string& e) {
        std::cerr << e << std::endl;
    }
    return 0;
}
