/* oracle/svbfm_oracle.h -- TEST INFRASTRUCTURE ONLY. Never linked into, imported by, or shipped with
 * the product. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may use it, and only as the checker.
 *
 * CPU restatement (plain C, sequential, one thread) of the reference's VB / vb_online / MCMC sweeps,
 * its libc-rand() based initial state, and its data formats. Every function cites the reference
 * file:line it follows (paths relative to the reference root, src/...).
 *
 * PARITY PIN: the reference ships no golden vectors or tests. This oracle is pinned
 * against the UNMODIFIED reference binary compiled from /root/reference into oracle/_ref/ and run with
 * a fixed seed (oracle/fixtime.c); the outputs of those runs are committed under tests/golden/ together
 * with the generating script, and tests/test_oracle_golden.py checks this file against them.
 */
#ifndef SVBFM_ORACLE_H_
#define SVBFM_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc orc_t;

enum { ORC_VB = 0, ORC_VB_ONLINE = 1, ORC_MCMC = 2 };

typedef struct orc_stats {
    double test_rmse;     /* value the reference appends to test_rmse_<k0k1K>_<method> */
    double train_stat;    /* value printed as "Train=" on stdout */
    double free_energy;   /* +F as printed on stdout (the file holds -F) */
    double alpha;
    double rmse_this;     /* mcmc: rmse of this draw's prediction (rlog rmse_mcmc_this) */
    int    has_free_energy; /* 0 when update_all returned early (non-finite alpha) */
    uint32_t nan_inf_count;
} orc_stats;

/* CSR in, as parsed from libFM text or a .x file. n_feat_* = (max feature id + 1) of that split as the
 * reference's loaders compute it (Data.h:220-222). D = fm.num_attribute (libfm.cpp:215). */
orc_t* orc_create(int method, uint32_t D, int K, int k0, int k1);
void   orc_destroy(orc_t* h);
int    orc_set_split(orc_t* h, int split /*0=train,1=test*/, uint32_t n_rows, uint32_t n_feat,
                     const uint64_t* rowptr, const uint32_t* col, const float* val, const float* y);
int    orc_set_groups(orc_t* h, const uint32_t* attr_group /*[D]*/, uint32_t n_groups);
/* replay of srand(seed) + fm_model::init + fm.w.init_normal + learner::init RNG consumption
 * (libfm.cpp:123-124,273,298/307/313,366; fm_model.h:92-101; fm_learn_vb.h:685-712; matrix.h:358-393) */
int    orc_init(orc_t* h, long seed, double init_stdev);
int    orc_set_mcmc_options(orc_t* h, int do_sample, int do_multilevel);

/* one-time work before the iteration loop (fm_learn_vb_simultaneous.h:37-44; fm_learn_mcmc_simultaneous.h:75-80) */
int    orc_begin(orc_t* h);
/* one outer iteration: vb -> update_all + test predict + eval; mcmc -> draw_all + predict + eval;
 * vb_online -> one epoch over num_batch batches */
int    orc_iterate(orc_t* h, orc_stats* out);
int    orc_set_num_batch(orc_t* h, uint32_t num_batch);
int    orc_set_task(orc_t* h, int task);   /* 1 = binary classification, targets already mapped to -1 / +1 (libfm.cpp:337-343); mcmc only */
int    orc_set_regular(orc_t* h, double r0, double rw, double rv);   /* mcmc/als -regular (libfm.cpp:367-405); after orc_init */

/* state access (row-major [K][D] for the matrices, like DMatrix::value[f][j]) */
int    orc_get_state(orc_t* h, double* w0_mean, double* w0_var, double* w_mean /*[D]*/, double* w_var /*[D]*/,
                     double* v_mean /*[K*D]*/, double* v_var /*[K*D]*/);
int    orc_set_state(orc_t* h, double w0_mean, double w0_var, const double* w_mean, const double* w_var,
                     const double* v_mean, const double* v_var);
int    orc_get_hyper(orc_t* h, double* alpha, double* sigma_0, double* sigma_w /*[G]*/, double* sigma_v /*[G*K]*/);
int    orc_get_train_cache(orc_t* h, double* e /*[N]*/, double* t /*[N] or NULL*/);
int    orc_get_test_pred(orc_t* h, double* p /*[Nt]*/);

/* ---- data formats (Data.h:106-283,457-509; fmatrix.h:46-86; matrix.h:280-294; convert.cpp; transpose.cpp) ---- */
typedef struct orc_csr {
    uint32_t n_rows, n_feat; uint64_t nnz;
    uint64_t* rowptr; uint32_t* col; float* val; float* y;
    float min_target, max_target;
} orc_csr;
int  orc_parse_text(const char* path, orc_csr* out);            /* Data::load text branch */
void orc_csr_free(orc_csr* m);
int  orc_transpose(const orc_csr* in, uint32_t n_out_rows, orc_csr* out); /* Data::create_data_t */
int  orc_write_x(const char* path, const orc_csr* m, uint32_t num_cols); /* convert.cpp:147-187 / transpose.cpp:104-160 */
int  orc_write_y(const char* path, const float* y, uint32_t n);          /* convert.cpp:159-177 */
int  orc_read_x(const char* path, orc_csr* out);                         /* fmatrix.h:157-172 */
int  orc_read_y(const char* path, float** y, uint32_t* n);               /* matrix.h:311-328 */

/* RNG restatement (util/random.h:118-176) exposed for tests */
double orc_ran_uniform(void);
double orc_ran_gaussian(void);
double orc_ran_gamma(double alpha);
double orc_erf(double x);                  /* random.h:47-61 */
double orc_cdf_gaussian(double x);         /* random.h:67-69 */
double orc_ran_left_tgaussian(double left, double mean, double stdev);    /* random.h:72-106 */
double orc_ran_right_tgaussian(double right, double mean, double stdev);  /* random.h:108-114 */

#ifdef __cplusplus
}
#endif
#endif
