// host/libfm.cpp -- libFM-compatible driver for the CUDA learners. Same command line as the reference
// (src/libfm/libfm.cpp:84-109): -task -train -test -dim -iter -method vb|vb_online|mcmc|als -meta -out -rlog
// -regular -init_stdev -stdev -verbosity -seed -batch -cache_size -validation -relation -learn_rate -help,
// same data formats, same per-iteration outputs. Differences, all deliberate:
//   * only the vb / vb_online / mcmc (als) learners exist here; sgd*, relations and classification are out of scope
//   * `-seed` is honoured when given (the reference ignores it and always uses time(NULL), libfm.cpp:123)
//   * v_file.txt (fm_model.h:98) is only written when SVBFM_WRITE_V_FILE=1
//   * multi-GPU: one process per GPU, RANK / WORLD_SIZE / LOCAL_RANK from the environment, id exchange through
//     the file named by SVBFM_COMM_FILE
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <ctime>
#include <iostream>
#include "cmdline.h"
#include "fm_learn_cuda.h"

using namespace svbfm_host;

// pre-scan for the online method (libfm.cpp:528-599): number of cases, largest feature id (NOT +1), target range
static void find_max_feature(DataSet& d, const std::string& file) {
    ParsedText t;
    parse_text_file(file, t, false);              // targets + largest id only (data.h: the parallel text parser)
    d.num_cases = (uint32_t)t.target.size();
    d.num_feature = (int)t.max_id;
    d.min_target = +FLT_MAX; d.max_target = -FLT_MAX;
    for (float v : t.target) { d.min_target = std::min(v, d.min_target); d.max_target = std::max(v, d.max_target); }
}

// The NCCL unique id travels from rank 0 to the others through a file every rank can read. A file left behind by an earlier launch
// must not be taken for this one's: the record carries a nonce every rank of ONE launch derives alike (SVBFM_COMM_NONCE, else
// MASTER_PORT / TORCHELASTIC_RUN_ID of the launcher and the parent's pid), rank 0 removes any old file before it asks for the id,
// records older than two minutes at the start of this process are ignored, and rank 0 removes the file once the communicator
// stands (svbfm_comm_init is collective: every rank has read the record by then; fm_learn_cuda.h open_engine).
static uint64_t launch_nonce() {
    std::string key;
    for (const char* name : {"SVBFM_COMM_NONCE", "TORCHELASTIC_RUN_ID", "MASTER_PORT", "SLURM_JOB_ID"})
        if (const char* v = getenv(name)) { key += name; key += '='; key += v; key += ';'; }
    key += "ppid=" + std::to_string((long)getppid());
    uint64_t h = 1469598103934665603ull;              // FNV-1a
    for (unsigned char c : key) { h ^= c; h *= 1099511628211ull; }
    return h;
}
struct CommRecord { char magic[8]; uint64_t nonce; uint8_t id[SVBFM_COMM_ID_BYTES]; };
static void exchange_comm_id(ShardInfo& sh) {
    const char* path = getenv("SVBFM_COMM_FILE");
    if (!path) throw std::string("WORLD_SIZE > 1 needs SVBFM_COMM_FILE (a path every rank can read)");
    const time_t started = time(nullptr);
    const uint64_t nonce = launch_nonce();
    std::string tmp = std::string(path) + ".tmp";
    if (sh.rank == 0) {
        unlink(path);
        if (svbfm_comm_get_unique_id(sh.comm_id) != 0) throw std::string("svbfm_comm_get_unique_id: ") + svbfm_last_error(nullptr);
        CommRecord rec;
        memcpy(rec.magic, "SVBFMID1", 8); rec.nonce = nonce; memcpy(rec.id, sh.comm_id, SVBFM_COMM_ID_BYTES);
        { std::ofstream f(tmp.c_str(), std::ios::binary); f.write(reinterpret_cast<const char*>(&rec), sizeof(rec)); }
        rename(tmp.c_str(), path);
    } else {
        for (int tries = 0; tries < 12000; tries++) {
            struct stat st;
            if (stat(path, &st) == 0 && st.st_mtime + 120 >= started) {
                std::ifstream f(path, std::ios::binary);
                CommRecord rec;
                if (f.is_open()) {
                    f.read(reinterpret_cast<char*>(&rec), sizeof(rec));
                    if (f.gcount() == (std::streamsize)sizeof(rec) && memcmp(rec.magic, "SVBFMID1", 8) == 0 && rec.nonce == nonce) {
                        memcpy(sh.comm_id, rec.id, SVBFM_COMM_ID_BYTES);
                        return;
                    }
                }
            }
            usleep(10000);
        }
        throw std::string("timed out waiting for this launch's record in ") + path;
    }
}
int main(int argc, char** argv) {
    try {
        CmdLine cmd(argc, argv);
        std::cout << "----------------------------------------------------------------------------" << std::endl;
        std::cout << "libFM (B200 build of the vb / vb_online / mcmc learners)" << std::endl;
        std::cout << "  command line, data formats and outputs of libFM 1.4.2 + VBFM/OVBFM fork" << std::endl;
        std::cout << "----------------------------------------------------------------------------" << std::endl;
        const std::string p_task = cmd.reg("task", "r=regression, c=binary classification [MANDATORY]");
        const std::string p_meta = cmd.reg("meta", "filename for meta information about data set");
        const std::string p_train = cmd.reg("train", "filename for training data [MANDATORY]");
        const std::string p_test = cmd.reg("test", "filename for test data [MANDATORY]");
        const std::string p_val = cmd.reg("validation", "filename for validation data (only for SGDA)");
        const std::string p_out = cmd.reg("out", "filename for output");
        const std::string p_dim = cmd.reg("dim", "'k0,k1,k2': k0=use bias, k1=use 1-way interactions, k2=dim of 2-way interactions; default=1,1,8");
        const std::string p_reg = cmd.reg("regular", "'r0,r1,r2' for ALS/MCMC: r0=bias regularization, r1=1-way regularization, r2=2-way regularization");
        const std::string p_init = cmd.reg("init_stdev", "stdev for initialization of 2-way factors; default=0.1");
        const std::string p_stdev = cmd.reg("stdev", "standard deviation for the model; default=1");
        const std::string p_iter = cmd.reg("iter", "number of iterations; default=100");
        const std::string p_lr = cmd.reg("learn_rate", "learn_rate for SGD; default=0.1");
        const std::string p_method = cmd.reg("method", "learning method (vb, vb_online, mcmc, als); default=MCMC");
        const std::string p_verb = cmd.reg("verbosity", "how much infos to print; default=0");
        const std::string p_rlog = cmd.reg("rlog", "write measurements within iterations to a file; default=''");
        const std::string p_seed = cmd.reg("seed", "integer value, default=time(NULL)");
        const std::string p_help = cmd.reg("help", "this screen");
        const std::string p_rel = cmd.reg("relation", "BS: filenames for the relations, default=''");
        const std::string p_cache = cmd.reg("cache_size", "cache size for data storage (accepted; data is held in memory)");
        const std::string p_batch = cmd.reg("batch", "How many batches for online algorithm");
        const std::string p_save = cmd.reg("save_model", "filename for writing the model (parameters, hyper-parameters, residuals) after learning");
        const std::string p_load = cmd.reg("load_model", "filename of a saved model to continue from instead of the random initial state");
        const std::string p_sampling = "do_sampling", p_multilevel = "do_multilevel", p_evalcases = "num_eval_cases";
        cmd.reg(p_sampling, "hidden"); cmd.reg(p_multilevel, "hidden"); cmd.reg(p_evalcases, "hidden");
        cmd.reg("device", "CUDA device ordinal (default: LOCAL_RANK or 0)");
        if (cmd.has(p_help) || argc == 1) { cmd.print_help(); return 0; }
        cmd.check();

        long seed = cmd.has(p_seed) ? cmd.get_int(p_seed, 1) : (long)time(NULL);      // libfm.cpp:123
        std::cout << "In libfm" << std::endl;
        if (!cmd.has(p_method)) cmd.set(p_method, "mcmc");
        if (!cmd.has(p_init)) cmd.set(p_init, "0.1");
        if (!cmd.has(p_stdev)) cmd.set(p_stdev, "1");
        if (!cmd.has(p_dim)) cmd.set(p_dim, "1,1,8");
        if (cmd.get(p_method) == "als") {                       // als = mcmc without sampling and hyper-priors (libfm.cpp:131-135)
            cmd.set(p_method, "mcmc");
            if (!cmd.has(p_sampling)) cmd.set(p_sampling, "0");
            if (!cmd.has(p_multilevel)) cmd.set(p_multilevel, "0");
        }
        const std::string method = cmd.get(p_method);
        if (method != "vb" && method != "vb_online" && method != "mcmc") throw std::string("unknown method in libfm (this build has vb, vb_online, mcmc, als)");
        if (!cmd.get_list(p_rel).empty()) throw std::string("relations (-relation) are out of scope of the CUDA path");

        ShardInfo sh;
        if (getenv("WORLD_SIZE")) sh.world = atoi(getenv("WORLD_SIZE"));
        if (getenv("RANK")) sh.rank = atoi(getenv("RANK"));
        int device = (int)cmd.get_int("device", getenv("LOCAL_RANK") ? atoi(getenv("LOCAL_RANK")) : 0);
        if (sh.world > 1) exchange_comm_id(sh);

        bool is_mcmc = method == "mcmc";
        DataSet train(!is_mcmc, true), test(!is_mcmc, true);     // libfm.cpp:137-146
        // the transposed matrix is built on the device unless a <name>.xt is there already (SVBFM_HOST_TRANSPOSE=1: on the host, as before)
        train.device_transpose = test.device_transpose = (method != "vb_online") && !getenv("SVBFM_HOST_TRANSPOSE");
        if (method != "vb_online") {
            std::cout << "Loading train...\t" << std::endl;
            train.load(cmd.get(p_train));
            if (cmd.get_int(p_verb, 0) > 0) train.debug();
            std::cout << "Loading test... \t" << std::endl;
            test.load(cmd.get(p_test));
            if (cmd.get_int(p_verb, 0) > 0) test.debug();
        } else {
            std::cout << "Loading train data members...\t" << std::endl;
            std::cout << "Loading test data... \t" << std::endl;
            test.load(cmd.get(p_test));
            DataSet tr_scan(true, true), te_scan(true, true);
            find_max_feature(tr_scan, cmd.get(p_train));
            find_max_feature(te_scan, cmd.get(p_test));
            // the online learner parses the train file itself with num_attribute columns (vbos.h:108-109)
            uint32_t nattr = (uint32_t)std::max(tr_scan.num_feature, te_scan.num_feature) + 1;
            train.load_text(cmd.get(p_train), nattr);
            train.num_feature = tr_scan.num_feature;             // libfm.cpp:552: largest id, not +1
            test.num_feature = te_scan.num_feature;
        }
        if (cmd.has(p_val)) std::cout << "WARNING: Validation data is only used for SGDA. The data is ignored." << std::endl;
        std::cout << "#relations: 0" << std::endl;
        std::cout << "Loading meta data...\t" << std::endl;
        uint32_t num_all_attribute = (uint32_t)std::max(train.num_feature, test.num_feature) + 1;   // libfm.cpp:215 (fork-specific +1)
        MetaInfo meta(num_all_attribute);
        if (cmd.has(p_meta)) meta.load_groups(cmd.get(p_meta));
        if (cmd.get_int(p_verb, 0) > 0) meta.debug();

        fm_model fm;
        fm.num_attribute = num_all_attribute;
        fm.init_stdev = cmd.get_double(p_init, 0.1);
        fm.stdev = cmd.get_double(p_stdev, 1.0);
        {
            std::vector<int> dim = cmd.get_ints(p_dim);
            if (dim.size() != 3) throw std::string("-dim needs 'k0,k1,k2'");
            fm.k0 = dim[0] != 0; fm.k1 = dim[1] != 0; fm.num_factor = dim[2];
        }
        std::cout << "stdev " << fm.stdev << std::endl;

        fm_learn_cuda* fml = nullptr;
        if (method == "mcmc") {
            auto* l = new fm_learn_mcmc_cuda();
            l->do_sample = cmd.get_int(p_sampling, 1) != 0;
            l->do_multilevel = cmd.get_int(p_multilevel, 1) != 0;
            fml = l;
        } else if (method == "vb") {
            fml = new fm_learn_vb_cuda();
        } else {
            auto* l = new fm_learn_vb_online_cuda();
            l->training_file = cmd.get(p_train); l->testing_file = cmd.get(p_test);
            l->num_batch = (unsigned)cmd.get_int(p_batch, 50);
            fml = l;
        }
        fml->num_iter = (unsigned)cmd.get_int(p_iter, 100);
        fml->num_eval_cases = (unsigned)cmd.get_int(p_evalcases, test.num_cases);
        fml->fm = &fm; fml->meta = &meta;
        fml->max_target = train.max_target; fml->min_target = train.min_target;     // libfm.cpp:332-333
        fml->seed = seed; fml->device = device; fml->shard = sh;
        if (cmd.has(p_save)) fml->save_model = cmd.get(p_save);
        if (cmd.has(p_load)) fml->load_model = cmd.get(p_load);
        const std::string task = cmd.get(p_task);
        if (task == "r") fml->task = 0;
        else if (task == "c") {                                   // libfm.cpp:337-343: every target <= 0 becomes -1, the others +1
            if (method != "mcmc") throw std::string("classification (-task c) is on the CUDA path for -method mcmc / als only");
            fml->task = 1;
            for (float& t : train.target) t = (t <= 0.0f) ? -1.0f : 1.0f;
            for (float& t : test.target) t = (t <= 0.0f) ? -1.0f : 1.0f;
        }
        else throw std::string("unknown task");
        {                                                         // regularisation (libfm.cpp:367-427)
            std::vector<double> reg = cmd.get_doubles(p_reg);
            if (reg.size() == 1) { fm.reg0 = fm.regw = fm.regv = reg[0]; }
            else if (reg.size() == 3) { fm.reg0 = reg[0]; fm.regw = reg[1]; fm.regv = reg[2]; }
            else if (!reg.empty()) throw std::string("-regular needs 1 or 3 values on the CUDA path");
        }
        RLog* rlog = nullptr;
        std::ofstream* rlog_stream = nullptr;
        if (cmd.has(p_rlog) && sh.rank == 0) {
            rlog_stream = new std::ofstream(cmd.get(p_rlog).c_str());
            if (!rlog_stream->is_open()) throw "Unable to open file " + cmd.get(p_rlog);
            std::cout << "logging to " << cmd.get(p_rlog) << std::endl;
            rlog = new RLog(rlog_stream);
        }
        fml->log = rlog;
        if (getenv("SVBFM_WRITE_V_FILE") && sh.rank == 0) {      // fm_model::init side effect (fm_model.h:98)
            InitialState s;
            init_state(seed, fm.num_attribute, fm.num_factor, fm.init_stdev, SVBFM_MCMC, s);
            std::ofstream vf("v_file.txt");
            for (int f = 0; f < fm.num_factor; f++) {
                for (uint32_t j = 0; j < fm.num_attribute; j++) vf << (j ? "\t" : "") << s.v_mean[(size_t)f * fm.num_attribute + j];
                vf << std::endl;
            }
        }
        fml->init();                                             // srand(seed) + the reference's draw order (host/init_state.h)
        if (rlog) rlog->init();
        if (cmd.get_int(p_verb, 0) > 0) { fm.debug(); fml->debug(); }

        fml->learn(train, test);
        if (sh.rank == 0) {
            std::cout << "after learn\n";
            if (method != "mcmc")                                 // libfm.cpp:509-511
                std::cout << "Final\t" << "Train=" << fml->evaluate(train) << "\tTest=" << fml->evaluate(test) << std::endl;
            if (cmd.has(p_out)) {                                 // libfm.cpp:514-519
                if (sh.world > 1) throw std::string("-out with WORLD_SIZE > 1 is not supported");
                std::vector<double> pred;
                fml->predict(test, pred);
                std::ofstream of(cmd.get(p_out).c_str());
                for (double v : pred) of << v << std::endl;
            }
        }
        delete fml;
    } catch (std::string& e) {
        std::cerr << std::endl << "ERROR: " << e << std::endl;   // and exit 0, like the reference (libfm.cpp:521-527)
    } catch (char const*& e) {
        std::cerr << std::endl << "ERROR: " << e << std::endl;
    }
    return 0;
}
