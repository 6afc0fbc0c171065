// host/fm_learn_cuda.h -- the learner shells: same interface as the reference's `class fm_learn`
// (src/libfm/src/fm_learn.h:38-265) and its vb / vb_online / mcmc subclasses, with the sweep delegated to the
// CUDA engine through the C-ABI (include/svbfm.h). What stays on the host is what the reference also does on
// the host around the sweep: the libc-rand() initial state, the per-iteration files in the CWD
// (test_rmse_<k0k1K>_<method>, free_energy_<k0k1K>_vb), the "#Iter=" lines and the -rlog fields.
#pragma once
#include <unistd.h>
#include <sys/resource.h>
#include <cmath>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <limits>
#include <sstream>
#include "../../include/svbfm.h"
#include "data.h"
#include "init_state.h"
#include "rlog.h"

namespace svbfm_host {

struct fm_model {                     // the fields of fm_model the drivers set (src/fm_core/fm_model.h:35-64; libfm.cpp:259-274)
    uint32_t num_attribute = 0;
    bool k0 = true, k1 = true;
    int num_factor = 0;
    double init_stdev = 0.01, init_mean = 0.0, stdev = 1.0;
    double reg0 = 0.0, regw = 0.0, regv = 0.0;
    void debug() const {
        std::cout << "num_attributes=" << num_attribute << std::endl << "use w0=" << k0 << std::endl << "use w1=" << k1 << std::endl
                  << "dim v =" << num_factor << std::endl << "reg_w0=" << reg0 << std::endl << "reg_w=" << regw << std::endl
                  << "reg_v=" << regv << std::endl << "init ~ N(" << init_mean << "," << init_stdev << ")" << std::endl;
    }
};

inline double user_time() {          // getusertime (src/util/util.h:70-80)
    struct rusage ru;
    getrusage(RUSAGE_SELF, &ru);
    return (double)ru.ru_utime.tv_sec + (double)ru.ru_utime.tv_usec / 1e6;
}

struct ShardInfo { int rank = 0, world = 1; uint8_t comm_id[SVBFM_COMM_ID_BYTES] = {0}; };

class fm_learn {
public:
    MetaInfo* meta = nullptr;
    fm_model* fm = nullptr;
    double min_target = 0, max_target = 0;
    int task = 0;                    // 0 = regression, 1 = binary classification (mcmc / als only on this path; targets -1 / +1)
    DataSet* validation = nullptr;
    RLog* log = nullptr;
    unsigned num_iter = 100, num_eval_cases = 0;
    long seed = 0;
    int device = 0;
    ShardInfo shard;
    virtual ~fm_learn() {}
    virtual void init() {            // fm_learn::init (fm_learn.h:80-100)
        if (log) {
            if (task == 0) { log->addField("rmse", nan_()); log->addField("mae", nan_()); }
            else if (task == 1) log->addField("accuracy", nan_());
            else throw std::string("unknown task");
            log->addField("time_pred", nan_()); log->addField("time_learn", nan_());
            log->addField("time_learn2", nan_()); log->addField("time_learn4", nan_());
        }
    }
    virtual void learn(DataSet& train, DataSet& test) = 0;
    virtual void predict(DataSet& data, std::vector<double>& out) = 0;
    virtual double evaluate(DataSet&) { return nan_(); }          // vb.h:27 / mcmc.h:69
    virtual void debug() {
        std::cout << "task=" << task << std::endl << "min_target=" << min_target << std::endl << "max_target=" << max_target << std::endl;
    }
protected:
    static double nan_() { return std::numeric_limits<double>::quiet_NaN(); }
};

// common CUDA plumbing of the three learners
class fm_learn_cuda : public fm_learn {
public:
    int method = SVBFM_VB;
    bool do_sample = true, do_multilevel = true;
    // -save_model / -load_model (the fork itself has no model files): parameters (means and variances
    // of w0, w, v), hyper-parameters and -- one GPU -- the cached residuals + sum_i T_i, raw doubles. A vb run that loads what
    // another vb run saved continues bit for bit; mcmc / vb_online take the parameters and hyper-parameters as a warm start.
    std::string save_model, load_model;
    ~fm_learn_cuda() override { if (h_) svbfm_destroy(h_); }

    void init() override {
        fm_learn::init();
        if (task != 0 && !(task == 1 && method == SVBFM_MCMC))
            throw std::string("task not supported on the CUDA path (regression; binary classification with mcmc / als)");
        init_state(seed, fm->num_attribute, fm->num_factor, fm->init_stdev, method, state_);   // host/init_state.h
        if (log) {                   // fm_learn_vb::init / fm_learn_mcmc::init log fields (vb.h:714-742, mcmc.h:1120-1150)
            log->addField("alpha", nan_());
            log->addField("rmse_mcmc_this", nan_());
            log->addField("rmse_mcmc_all", nan_());
            if (method == SVBFM_MCMC) log->addField("rmse_mcmc_all_but5", nan_());
            for (uint32_t g = 0; g < meta->num_attr_groups; g++) {
                std::ostringstream a, b;
                a << "wmu[" << g << "]"; b << "wlambda[" << g << "]";
                log->addField(a.str(), nan_()); log->addField(b.str(), nan_());
                for (int f = 0; f < fm->num_factor; f++) {
                    std::ostringstream c, d;
                    c << "vmu[" << g << "," << f << "]"; d << "vlambda[" << g << "," << f << "]";
                    log->addField(c.str(), nan_()); log->addField(d.str(), nan_());
                }
            }
        }
    }

    void predict(DataSet& data, std::vector<double>& out) override {
        out.assign(data.num_cases, 0.0);
        ck(svbfm_predict(h_, SVBFM_TEST, out.data()), "svbfm_predict");
    }
    void debug() override { fm_learn::debug(); std::cout << "num_eval_cases=" << num_eval_cases << std::endl; }

protected:
    svbfm_t* h_ = nullptr;
    InitialState state_;
    std::string tag_;                // "<k0><k1><K>"

    void ck(int rc, const char* what) {
        if (rc != 0) throw std::string(what) + ": " + svbfm_last_error(h_);     // -> "ERROR: ..." in main (libfm.cpp:521-525)
    }
    bool root() const { return shard.rank == 0; }

    void open_engine(DataSet& train, DataSet& test) {
        svbfm_config c;
        memset(&c, 0, sizeof(c));
        c.struct_size = sizeof(c); c.method = method; c.num_attribute = fm->num_attribute; c.num_factor = fm->num_factor;
        c.k0 = fm->k0; c.k1 = fm->k1; c.task = task; c.min_target = min_target; c.max_target = max_target; c.device = device;
        c.do_sample = do_sample; c.do_multilevel = do_multilevel; c.seed = (uint64_t)seed; c.reg0 = fm->reg0; c.regw = fm->regw; c.regv = fm->regv;
        int rc = svbfm_create(&h_, &c);
        if (rc != 0) throw std::string("svbfm_create: ") + svbfm_last_error(nullptr);
        if (shard.world > 1) {
            ck(svbfm_comm_init(h_, shard.comm_id, shard.rank, shard.world), "svbfm_comm_init");
            // the communicator stands on every rank: the id record must not outlive the launch (libfm.cpp exchange_comm_id)
            if (shard.rank == 0) if (const char* path = getenv("SVBFM_COMM_FILE")) unlink(path);
        }
        ck(svbfm_set_groups(h_, meta->attr_group.data(), meta->num_attr_groups), "svbfm_set_groups");
        push(SVBFM_TRAIN, train);
        push(SVBFM_TEST, test);
        if (!load_model.empty()) read_model();          // replaces the random initial state
        ck(svbfm_set_state(h_, state_.w0_mean, state_.w0_var, state_.w_mean.data(), state_.w_var.data(), state_.v_mean.data(), state_.v_var.data()),
           "svbfm_set_state");
        if (!load_model.empty())
            ck(svbfm_set_hyper(h_, model_.alpha, model_.sigma_0, model_.sigma_w.data(), model_.sigma_v.empty() ? nullptr : model_.sigma_v.data()), "svbfm_set_hyper");
        std::ostringstream t;
        t << fm->k0 << fm->k1 << fm->num_factor;
        tag_ = t.str();
    }

    struct ModelHeader { char magic[8]; uint32_t method, D; int32_t K; uint32_t G, has_resid, n_resid, reserved[2]; };
    struct ModelExtra { double alpha = 1.0, sigma_0 = 1.0, sum_t = 0.0; std::vector<double> sigma_w, sigma_v, resid; bool has_resid = false; } model_;

    void read_model() {
        std::ifstream f(load_model.c_str(), std::ios::binary);
        if (!f.is_open()) throw "Unable to open file " + load_model;
        ModelHeader hd;
        f.read(reinterpret_cast<char*>(&hd), sizeof(hd));
        if (f.gcount() != (std::streamsize)sizeof(hd) || memcmp(hd.magic, "SVBFMMD1", 8) != 0) throw load_model + " is not a model file of this program";
        const uint32_t D = fm->num_attribute, G = meta->num_attr_groups;
        const int K = fm->num_factor;
        if (hd.D != D || hd.K != K || hd.G != G) throw "model file " + load_model + " was saved with other dimensions (attributes, factors or groups)";
        auto rd = [&](double* p, size_t n) {
            f.read(reinterpret_cast<char*>(p), (std::streamsize)(n * 8));
            if ((size_t)f.gcount() != n * 8) throw "model file " + load_model + " is truncated";
        };
        double sc[5];
        rd(sc, 5);
        state_.w0_mean = sc[0]; state_.w0_var = sc[1]; model_.alpha = sc[2]; model_.sigma_0 = sc[3]; model_.sum_t = sc[4];
        state_.w_mean.resize(D); state_.w_var.resize(D); state_.v_mean.resize((size_t)K * D); state_.v_var.resize((size_t)K * D);
        rd(state_.w_mean.data(), D); rd(state_.w_var.data(), D); rd(state_.v_mean.data(), (size_t)K * D); rd(state_.v_var.data(), (size_t)K * D);
        model_.sigma_w.resize(G); model_.sigma_v.resize((size_t)G * K);
        rd(model_.sigma_w.data(), G); rd(model_.sigma_v.data(), (size_t)G * K);
        model_.has_resid = hd.has_resid != 0 && (int)hd.method == method;
        if (hd.has_resid) { model_.resid.resize(hd.n_resid); rd(model_.resid.data(), hd.n_resid); }
    }
    // after svbfm_begin: the saved residuals replace the freshly predicted ones (same run, same train split, one GPU)
    void restore_residuals(uint32_t num_train_cases) {
        if (method != SVBFM_VB || load_model.empty() || !model_.has_resid || shard.world > 1 || model_.resid.size() != num_train_cases) return;
        ck(svbfm_set_residuals(h_, model_.resid.data(), model_.sum_t), "svbfm_set_residuals");
    }
    void write_model(uint32_t num_train_cases) {
        if (save_model.empty()) return;
        const uint32_t D = fm->num_attribute, G = meta->num_attr_groups;
        const int K = fm->num_factor;
        std::vector<double> wm(D), wv(D), vm((size_t)K * D), vv((size_t)K * D), sw(G), sv((size_t)G * std::max(K, 1)), resid;
        double sc[5] = {0, 0, 0, 0, 0};
        ck(svbfm_get_state(h_, &sc[0], &sc[1], wm.data(), wv.data(), vm.data(), vv.data()), "svbfm_get_state");
        ck(svbfm_get_hyper(h_, &sc[2], &sc[3], sw.data(), sv.data()), "svbfm_get_hyper");
        ck(svbfm_get_sum_t(h_, &sc[4]), "svbfm_get_sum_t");
        const bool with_resid = shard.world <= 1;
        if (with_resid) { resid.resize(num_train_cases); ck(svbfm_get_residuals(h_, resid.data()), "svbfm_get_residuals"); }
        if (!root()) return;
        ModelHeader hd;
        memset(&hd, 0, sizeof(hd));
        memcpy(hd.magic, "SVBFMMD1", 8);
        hd.method = (uint32_t)method; hd.D = D; hd.K = K; hd.G = G; hd.has_resid = with_resid ? 1u : 0u; hd.n_resid = with_resid ? num_train_cases : 0u;
        std::ofstream f(save_model.c_str(), std::ios::binary);
        if (!f.is_open()) throw "Unable to open file " + save_model;
        auto wr = [&](const double* p, size_t n) { f.write(reinterpret_cast<const char*>(p), (std::streamsize)(n * 8)); };
        f.write(reinterpret_cast<const char*>(&hd), sizeof(hd));
        wr(sc, 5); wr(wm.data(), D); wr(wv.data(), D); wr(vm.data(), (size_t)K * D); wr(vv.data(), (size_t)K * D); wr(sw.data(), G); wr(sv.data(), (size_t)G * K);
        if (with_resid) wr(resid.data(), resid.size());
        if (!f.good()) throw "could not write " + save_model;
    }

    // hands data_t + target of this rank's shard of cases to the engine
    static bool all_ones(const std::vector<float>& v) {
        for (float f : v) if (f != 1.0f) return false;
        return true;
    }
    // Several GPUs, vb / mcmc on two complete one-hot fields with x = 1 (`u:1 i:1` rows, every first id below every second id): cross
    // shards: this rank's block of first-field columns goes in as SVBFM_TRAIN, its block of second-field columns
    // as SVBFM_TRAIN_SECOND; both block sets are cut where the running cost `entries + 8 per column` passes rank / world of the total.
    // SVBFM_SHARD=range keeps contiguous case ranges (both fields allreduced). Returns false when the data does not qualify.
    bool push_cross(DataSet& d) {
        const char* mode = getenv("SVBFM_SHARD");
        if (shard.world <= 1 || method == SVBFM_VB_ONLINE || task != 0 || !d.csr_only || (mode && std::string(mode) == "range")) return false;
        const uint32_t n = d.num_cases;
        const SparseMatrix& x = d.x;
        if (n == 0 || x.nnz() != 2ull * n || !all_ones(x.val)) return false;
        uint32_t max0 = 0, min1 = 0xffffffffu;
        for (uint32_t i = 0; i < n; i++) {
            if (x.ptr[i + 1] - x.ptr[i] != 2) return false;
            const uint32_t a = x.id[2ull * i], b = x.id[2ull * i + 1];
            if (a >= b) return false;
            max0 = std::max(max0, a); min1 = std::min(min1, b);
        }
        if (max0 >= min1) return false;
        const uint32_t nf = (uint32_t)d.num_feature, c1 = min1;               // fields: [0, c1) and [c1, nf)
        std::vector<uint64_t> cnt(nf, 0);
        for (uint64_t p = 0; p < x.nnz(); p++) cnt[x.id[p]]++;
        auto cuts = [&](uint32_t lo, uint32_t hi) {
            double total = 0;
            for (uint32_t j = lo; j < hi; j++) total += (double)cnt[j] + (cnt[j] ? 8.0 : 0.0);
            std::vector<uint32_t> b(shard.world + 1, hi);
            b[0] = lo;
            double run = 0; int r = 1;
            for (uint32_t j = lo; j < hi && r < shard.world; j++) {
                run += (double)cnt[j] + (cnt[j] ? 8.0 : 0.0);
                while (r < shard.world && run >= total * r / shard.world) b[r++] = j + 1;
            }
            return b;
        };
        const std::vector<uint32_t> b0 = cuts(0, c1), b1 = cuts(c1, nf);
        for (int which = 0; which < 2; which++) {
            const uint32_t lo = which ? b1[shard.rank] : b0[shard.rank], hi = which ? b1[shard.rank + 1] : b0[shard.rank + 1];
            std::vector<uint64_t> rp(1, 0);
            std::vector<uint32_t> ids;
            std::vector<float> y;
            for (uint32_t i = 0; i < n; i++) {
                const uint32_t key = x.id[2ull * i + which];
                if (key < lo || key >= hi) continue;
                ids.push_back(x.id[2ull * i]); ids.push_back(x.id[2ull * i + 1]);
                rp.push_back(ids.size());
                y.push_back(d.target[i]);
            }
            int rc = svbfm_set_csr(h_, which ? SVBFM_TRAIN_SECOND : SVBFM_TRAIN, (uint32_t)y.size(), nf, rp.data(), ids.data(), nullptr, y.data());
            if (rc != 0 && which == 1) {      // refused on every rank alike: the user-block shard stands by itself (item sums allreduced)
                if (root()) std::cout << "cross shards refused (" << svbfm_last_error(h_) << "): user-block shards" << std::endl;
                return true;
            }
            ck(rc, "svbfm_set_csr");
        }
        if (root()) std::cout << "multi-GPU: cross shards (first field by blocks of columns for the first residual copy, second field for the second)" << std::endl;
        return true;
    }
    void push(int split, DataSet& d) {
        if (split == SVBFM_TRAIN && push_cross(d)) return;
        // one-hot indicator data: the values are not shipped at all (x = NULL: include/svbfm.h)
        const bool ones = all_ones(d.csr_only ? d.x.val : d.xt.val);
        if (d.csr_only) {       // rows as loaded: a rank's shard is a slice of the row pointer; the device transposes (svbfm_set_csr)
            const uint32_t lo = (uint32_t)((uint64_t)d.num_cases * shard.rank / shard.world), hi = (uint32_t)((uint64_t)d.num_cases * (shard.rank + 1) / shard.world);
            const uint64_t e0 = d.x.ptr[lo];
            std::vector<uint64_t> rp(d.x.ptr.begin() + lo, d.x.ptr.begin() + hi + 1);
            for (uint64_t& v : rp) v -= e0;
            ck(svbfm_set_csr(h_, split, hi - lo, (uint32_t)d.num_feature, rp.data(), d.x.id.data() + e0, ones ? nullptr : d.x.val.data() + e0, d.target.data() + lo), "svbfm_set_csr");
            return;
        }
        if (shard.world <= 1) {
            ck(svbfm_set_csc(h_, split, d.num_cases, d.xt.num_rows, d.xt.ptr.data(), d.xt.id.data(), ones ? nullptr : d.xt.val.data(), d.target.data()), "svbfm_set_csc");
            return;
        }
        uint32_t lo = (uint32_t)((uint64_t)d.num_cases * shard.rank / shard.world), hi = (uint32_t)((uint64_t)d.num_cases * (shard.rank + 1) / shard.world);
        SparseMatrix s;
        s.num_rows = d.xt.num_rows; s.ptr.assign(1, 0);
        for (uint32_t j = 0; j < d.xt.num_rows; j++) {
            for (uint64_t p = d.xt.ptr[j]; p < d.xt.ptr[j + 1]; p++)
                if (d.xt.id[p] >= lo && d.xt.id[p] < hi) { s.id.push_back(d.xt.id[p] - lo); s.val.push_back(d.xt.val[p]); }
            s.ptr.push_back(s.id.size());
        }
        ck(svbfm_set_csc(h_, split, hi - lo, s.num_rows, s.ptr.data(), s.id.data(), s.val.data(), d.target.data() + lo), "svbfm_set_csc");
    }

    void truncate_file(const std::string& name) { if (root()) { std::ofstream f(name.c_str()); } }
    void append_value(const std::string& name, double v) { if (root()) { std::ofstream f(name.c_str(), std::ios_base::app); f << v << "\n"; } }
};

// fm_learn_vb_simultaneous (src/libfm/src/fm_learn_vb_simultaneous.h:15-259)
class fm_learn_vb_cuda : public fm_learn_cuda {
public:
    fm_learn_vb_cuda() { method = SVBFM_VB; }
    void learn(DataSet& train, DataSet& test) override {
        open_engine(train, test);
        ck(svbfm_begin(h_), "svbfm_begin");
        restore_residuals(train.num_cases);
        truncate_file("test_rmse_" + tag_ + "_vb");                 // vbs.h:66-73
        truncate_file("free_energy_" + tag_ + "_vb");
        for (unsigned i = 0; i < num_iter; i++) {
            double t0 = user_time(); clock_t c0 = clock(); time_t w0 = time(nullptr);
            svbfm_iter_stats s;
            ck(svbfm_vb_sweep(h_, &s), "svbfm_vb_sweep");
            if (s.nan_inf_count > 0 && root()) std::cout << "#nans/infs reverted:\t" << s.nan_inf_count << std::endl;
            if (s.has_free_energy) {
                append_value("free_energy_" + tag_ + "_vb", -s.free_energy);                          // vb.h:678 (file holds -F)
                if (root()) std::cout << "free energy " << s.free_energy << std::endl;                 // vb.h:680
            }
            append_value("test_rmse_" + tag_ + "_vb", s.test_rmse);                                   // vbs.h:221
            if (root())
                std::cout << "#Iter=" << std::setw(3) << i << "\tTrain=" << s.train_stat << "\tTest=" << s.test_rmse << std::endl;   // vbs.h:222
            if (log && root()) {
                log->log("time_learn", user_time() - t0);
                log->log("time_learn2", (double)(clock() - c0) / CLOCKS_PER_SEC);
                log->log("time_learn4", (double)(time(nullptr) - w0));
                log->log("alpha", s.alpha);
                log->log("rmse_mcmc_this", s.test_rmse);
                log->newLine();
            }
        }
        write_model(train.num_cases);
    }
};

// fm_learn_mcmc_simultaneous (src/libfm/src/fm_learn_mcmc_simultaneous.h:47-305)
class fm_learn_mcmc_cuda : public fm_learn_cuda {
public:
    fm_learn_mcmc_cuda() { method = SVBFM_MCMC; }
    void learn(DataSet& train, DataSet& test) override {
        open_engine(train, test);
        ck(svbfm_begin(h_), "svbfm_begin");
        truncate_file("test_rmse_" + tag_ + "_mcmc");               // mcmcs.h:60-62
        for (unsigned i = 0; i < num_iter; i++) {
            double t0 = user_time(); clock_t c0 = clock(); time_t w0 = time(nullptr);
            svbfm_iter_stats s;
            ck(svbfm_mcmc_sweep(h_, &s), "svbfm_mcmc_sweep");
            if (s.nan_inf_count > 0 && root()) std::cout << "#nans/infs reverted:\t" << s.nan_inf_count << std::endl;
            if (task == 1) {
                // classification: accuracies on stdout only, nothing goes to test_rmse_* (mcmcs.h:262-275). MAP@5 comes from a side
                // file with a hard-coded path (fm_learn.h:124); without that file the reference prints 0, and so does this line.
                if (root())
                    std::cout << "#Iter=" << std::setw(3) << i << "\tTrain=" << s.train_stat << "\tTest=" << s.test_rmse << "\tMAP@5= " << 0 << std::endl;
                if (log && root()) {
                    log->log("time_learn", user_time() - t0);
                    log->log("time_learn2", (double)(clock() - c0) / CLOCKS_PER_SEC);
                    log->log("time_learn4", (double)(time(nullptr) - w0));
                    log->log("alpha", s.alpha);
                    log->log("accuracy", s.test_rmse);
                    log->newLine();
                }
                continue;
            }
            if (root())
                std::cout << "#Iter=" << std::setw(3) << i << "\tTrain=" << s.train_stat << "\tTest=" << s.test_rmse << std::endl;   // mcmcs.h:244
            append_value("test_rmse_" + tag_ + "_mcmc", s.test_rmse);                                  // mcmcs.h:245
            if (log && root()) {
                log->log("time_learn", user_time() - t0);
                log->log("time_learn2", (double)(clock() - c0) / CLOCKS_PER_SEC);
                log->log("time_learn4", (double)(time(nullptr) - w0));
                log->log("alpha", s.alpha);
                log->log("rmse", s.test_rmse);
                log->log("rmse_mcmc_this", s.rmse_this);
                log->log("rmse_mcmc_all", s.test_rmse);
                log->newLine();
            }
        }
        write_model(train.num_cases);
    }
};

// fm_learn_vb_online_simultaneous (src/libfm/src/fm_learn_vb_online_simultaneous.h:18-290). The reference
// re-reads the training file and writes one text file per batch every epoch; here the file is parsed once
// and only the case -> batch rule (vbos.h:74-95, libc shuffle stream) is replayed.
class fm_learn_vb_online_cuda : public fm_learn_cuda {
public:
    unsigned num_batch = 50;
    std::string training_file, testing_file;
    fm_learn_vb_online_cuda() { method = SVBFM_VB_ONLINE; }
    void learn(DataSet& train, DataSet& test) override {
        open_engine(train, test);
        ck(svbfm_begin(h_), "svbfm_begin");
        truncate_file("test_rmse_" + tag_ + "_vb_online");           // vbos.h:45-52
        truncate_file("free_energy_" + tag_ + "_vb_online");         // stays empty: free_energy() appends to the _vb file (vbo.h:637)
        uint32_t n = train.num_cases;
        uint32_t size_except_last = (uint32_t)std::ceil((double)n / num_batch);
        std::vector<uint32_t> shuffle(n), batch(n);
        for (uint32_t i = 0; i < n; i++) shuffle[i] = i + 1;
        uint32_t lo = (uint32_t)((uint64_t)n * shard.rank / shard.world), hi = (uint32_t)((uint64_t)n * (shard.rank + 1) / shard.world);
        for (unsigned k = 0; k < num_iter; k++) {
            double t0 = user_time(); clock_t c0 = clock(); time_t w0 = time(nullptr);
            libc_random_shuffle(shuffle.data(), n);                   // vbos.h:74
            for (uint32_t r = 0; r < n; r++) batch[r] = (uint32_t)std::ceil((double)shuffle[r] / size_except_last) - 1;   // vbos.h:93
            svbfm_iter_stats s;
            ck(svbfm_vb_online_epoch(h_, batch.data() + lo, num_batch, &s), "svbfm_vb_online_epoch");
            (void)hi;
            if (s.has_free_energy) {                                  // batch 1 and batch B of the epoch (vbos.h:143-146)
                if (num_batch > 1) { append_value("free_energy_" + tag_ + "_vb", -s.free_energy_first); if (root()) std::cout << "free energy " << s.free_energy_first << std::endl; }
                append_value("free_energy_" + tag_ + "_vb", -s.free_energy);
                if (root()) std::cout << "free energy " << s.free_energy << std::endl;
            }
            if (root()) std::cout << "#Iter=" << std::setw(3) << k << "\tTest=" << s.test_rmse << std::endl;     // vbos.h:244
            append_value("test_rmse_" + tag_ + "_vb_online", s.test_rmse);                                  // vbos.h:243
            if (log && root()) {
                log->log("time_learn", user_time() - t0);
                log->log("time_learn2", (double)(clock() - c0) / CLOCKS_PER_SEC);
                log->log("time_learn4", (double)(time(nullptr) - w0));
                log->log("rmse_mcmc_this", s.test_rmse);
                log->newLine();
            }
        }
        write_model(train.num_cases);
    }
};

}  // namespace svbfm_host
