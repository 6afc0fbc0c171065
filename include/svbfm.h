/* include/svbfm.h -- C-ABI of the B200 VB / vb_online / MCMC coordinate-sweep engine.
 *
 * This is the drop-in boundary: plain C types, caller-owned host buffers that are
 * copied during the call (the library never keeps a host pointer), one handle = one host thread,
 * return 0 on success / negative svbfm_status on failure, no exceptions across the boundary.
 * The host-side learner shells (host/fm_learn_cuda.h, the Python mirror in __init__.py) convert a
 * non-zero status into the reference's `throw std::string` convention (libfm.cpp:521-525).
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference root):
 *   vb.h    = src/libfm/src/fm_learn_vb.h            vbs.h  = src/libfm/src/fm_learn_vb_simultaneous.h
 *   vbo.h   = src/libfm/src/fm_learn_vb_online.h     vbos.h = src/libfm/src/fm_learn_vb_online_simultaneous.h
 *   mcmc.h  = src/libfm/src/fm_learn_mcmc.h          mcmcs.h= src/libfm/src/fm_learn_mcmc_simultaneous.h
 *
 * There is NO CPU fallback: every compute entry point fails with SVBFM_ERR_CUDA when no sm_100 device
 * is usable.
 */
#ifndef SVBFM_H_
#define SVBFM_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SVBFM_ABI_VERSION 1

typedef struct svbfm svbfm_t;

typedef enum svbfm_status {
    SVBFM_OK = 0,
    SVBFM_ERR_ARG = -1,        /* bad argument / call order */
    SVBFM_ERR_CUDA = -2,       /* CUDA runtime failure or no device */
    SVBFM_ERR_NCCL = -3,       /* NCCL missing or failed */
    SVBFM_ERR_DATA = -4,       /* unsupported input (e.g. a feature id repeated inside one row) */
    SVBFM_ERR_OOM = -5
} svbfm_status;

enum { SVBFM_VB = 0, SVBFM_VB_ONLINE = 1, SVBFM_MCMC = 2 };   /* -method vb | vb_online | mcmc (libfm.cpp:297-320) */
enum { SVBFM_TRAIN = 0, SVBFM_TEST = 1,
       SVBFM_TRAIN_SECOND = 2   /* several GPUs only: the shard of the SECOND residual copy (see svbfm_set_csc) */ };

/* What main() wires into the learner before init() (libfm.cpp:259-274, 301-320, 331-336). */
typedef struct svbfm_config {
    uint32_t struct_size;      /* = sizeof(svbfm_config) */
    int32_t  method;           /* SVBFM_VB | SVBFM_VB_ONLINE | SVBFM_MCMC */
    uint32_t num_attribute;    /* D  = fm.num_attribute (libfm.cpp:215,261) */
    int32_t  num_factor;       /* K  = fm.num_factor   (libfm.cpp:270) */
    int32_t  k0, k1;           /* fm.k0, fm.k1         (libfm.cpp:268-269) */
    int32_t  task;             /* 0 = regression; 1 = binary classification (fm_learn.h:67-68), SVBFM_MCMC only: targets are -1 / +1
                                * (the caller maps them like libfm.cpp:337-343); the iteration statistics then carry accuracies:
                                * train_stat = "Train=", test_rmse = "Test=" (running mean), rmse_this = acc_mcmc_this
                                * (fm_learn_mcmc_simultaneous.h:176-221, 262-275); svbfm_predict returns probabilities in [0, 1].
                                * SVBFM_FLAG_MCMC_NO_REPREDICT is ignored (e = yhat - latent target). MAP@k is not computed. */
    double   min_target;       /* fml->min_target / max_target: taken from TRAIN (libfm.cpp:332-333) */
    double   max_target;
    int32_t  device;           /* CUDA device ordinal */
    int32_t  do_sample;        /* mcmc: fm_learn_mcmc::do_sample     (libfm.cpp:304) */
    int32_t  do_multilevel;    /* mcmc: fm_learn_mcmc::do_multilevel (libfm.cpp:305) */
    uint64_t seed;             /* mcmc: key of the counter-based (Philox4x32-10) draw stream */
    double   reg0, regw, regv; /* mcmc: initial w0 / w / v regularisation (libfm.cpp:372-405) */
    uint32_t tile_entries;     /* 0 = default; max column entries handled by one warp-tile */
    uint32_t flags;            /* SVBFM_FLAG_* */
} svbfm_config;

#define SVBFM_FLAG_NO_ROW_REORDER 1u   /* keep train rows in caller order on the device (debug) */
#define SVBFM_FLAG_MCMC_NO_REPREDICT 2u /* mcmc: skip the per-iteration re-prediction of train (mcmcs.h:134 does it) */

/* Per-iteration outputs: what the reference prints / appends to test_rmse_* and free_energy_*
 * (vbs.h:218-222, vb.h:678-680, mcmcs.h:241-245, vbos.h:242-244). */
typedef struct svbfm_iter_stats {
    double test_rmse;          /* value appended to test_rmse_<k0k1K>_<method> */
    double train_stat;         /* "Train=" value on stdout */
    double free_energy;        /* +F as printed; the file holds -F */
    double alpha;
    double rmse_this;          /* mcmc: rmse of this draw (rlog rmse_mcmc_this) */
    int32_t has_free_energy;   /* 0 when update_all returned early on a non-finite alpha (vb.h:456-469) */
    uint32_t nan_inf_count;    /* sum of the reference's nan_ and inf_ counters for this iteration */
    float sweep_ms;            /* device time of the update_all / draw_all equivalent (CUDA events) */
    float predict_ms;          /* device time of the test (and mcmc: train) prediction + evaluation */
    double free_energy_first;  /* vb_online: +F after batch 1 of the epoch (free_energy holds batch B's); else = free_energy */
} svbfm_iter_stats;

/* ---- lifetime ------------------------------------------------------------------------------------ */
/* replaces: `new fm_learn_*_simultaneous()` + the field pokes in libfm.cpp:297-336 */
int svbfm_create(svbfm_t** out, const svbfm_config* cfg);
void svbfm_destroy(svbfm_t* h);
/* last error text of this handle (h may be NULL: last error of svbfm_create on this thread) */
const char* svbfm_last_error(const svbfm_t* h);
int svbfm_abi_version(void);

/* ---- multi-GPU (one process per GPU; rows sharded) -------------------------- */
/* Rank 0 obtains an id, the host distributes it (torch.distributed / file / MPI), every rank joins.
 * Must be called before svbfm_set_csc. NCCL is dlopen()ed; absent NCCL => SVBFM_ERR_NCCL. */
#define SVBFM_COMM_ID_BYTES 128
int svbfm_comm_get_unique_id(uint8_t id[SVBFM_COMM_ID_BYTES]);
int svbfm_comm_init(svbfm_t* h, const uint8_t id[SVBFM_COMM_ID_BYTES], int32_t rank, int32_t world_size);

/* ---- inputs -------------------------------------------------------------------------------------- */
/* replaces: DataMetaInfo::attr_group / num_attr_per_group (Data.h:35-69; libfm.cpp:219-254) */
int svbfm_set_groups(svbfm_t* h, const uint32_t* attr_group /*[D]*/, uint32_t num_groups);

/* replaces: DataSubset::data_t + DataSubset::target handed to learn() (Data.h:87-89, fmatrix.h:36-44).
 * CSC of this rank's cases: column j (feature) holds (case id, x) pairs in ascending case id.
 * num_cols = data_t->getNumRows() (= max feature id of the split + 1); case ids are local, 0..num_cases-1. */
int svbfm_set_csc(svbfm_t* h, int32_t split, uint32_t num_cases, uint32_t num_cols,
                  const uint64_t* colptr /*[num_cols+1]*/, const uint32_t* case_id /*[nnz]*/,
                  const float* x /*[nnz]*/, const float* target /*[num_cases]*/);
/* x may be NULL (here, in svbfm_set_csr and in svbfm_transpose_csr): every value is 1 (one-hot indicator data; the loaders of
 * host/data.h notice it while they parse): a third of the entry bytes neither cross the bus nor get checked on the device. */
/* The same split handed over ROW-wise: CSR of this rank's cases as the reference's `data` holds it (DataSubset::data, Data.h:87;
 * the rows of a binary .x file or of a parsed text file), features in any order inside a case. The transposed matrix the sweeps
 * run on is built on the device (replaces Data::create_data_t, Data.h:457-509, and tools/transpose.cpp:91-162 for this path): the
 * caller neither transposes nor ships the data twice. Same results as svbfm_set_csc on the transposed data. */
int svbfm_set_csr(svbfm_t* h, int32_t split, uint32_t num_cases, uint32_t num_cols,
                  const uint64_t* rowptr /*[num_cases+1]*/, const uint32_t* feature_id /*[nnz]*/,
                  const float* x /*[nnz]*/, const float* target /*[num_cases]*/);
/* The device transpose by itself (no handle): CSR in, CSC out into caller-allocated host arrays; what bin/transpose --device writes
 * as .xt, byte for byte the reference tool's output. Returns SVBFM_ERR_ARG for a feature id >= num_cols. */
int svbfm_transpose_csr(int32_t device, uint32_t num_cases, uint32_t num_cols, const uint64_t* rowptr, const uint32_t* feature_id,
                        const float* x, uint64_t* out_colptr /*[num_cols+1]*/, uint32_t* out_case_id /*[nnz]*/, float* out_x /*[nnz]*/);

/* Cross shards (optional, several GPUs, two complete one-hot fields with x = 1, vb / mcmc regression): after SVBFM_TRAIN
 * -- which must then be a shard by blocks of the FIRST field's columns (rank r holds every case of its users; blocks disjoint and
 * ordered by rank) -- the caller hands over, as split SVBFM_TRAIN_SECOND in the same CSC format, the cases of the same global train
 * set that fall into this rank's block of the SECOND field's columns (every case of its items). The engine then keeps its second
 * residual copy on that shard: every column of either field has all of its entries on one rank, so no column sums are reduced
 * between the ranks; the records of the updated columns travel instead (one allgather per factor and field). Results equal the
 * single-GPU run up to summation order. Every rank must make the call (it is collective); SVBFM_ERR_ARG when the data does not
 * qualify (all ranks fail together). No reference counterpart (the reference is single-process). */


/* replaces: the variational state set up by fm_learn_vb::init (vb.h:693-712) / fm_model (fm_model.h:92-101).
 * vb / vb_online: (mean, var) = (mu', sigma'); mcmc: mean = the parameter, var ignored.
 * v_* are row-major [K][D] like DMatrix::value[f][j]. The host computes them with libc rand() in the
 * reference's order so that iteration 0 starts from the reference's state (host/init_state.h). */
int svbfm_set_state(svbfm_t* h, double w0_mean, double w0_var, const double* w_mean, const double* w_var,
                    const double* v_mean, const double* v_var);
int svbfm_get_state(svbfm_t* h, double* w0_mean, double* w0_var, double* w_mean, double* w_var,
                    double* v_mean, double* v_var);
/* alpha, sigma_0, sigma_w[G], sigma_v[G][K] (vb.h:37-39) / alpha, reg0, w_lambda[G], v_lambda[G][K] (mcmc.h:79-85) */
int svbfm_get_hyper(svbfm_t* h, double* alpha, double* sigma_0, double* sigma_w, double* sigma_v);
int svbfm_set_hyper(svbfm_t* h, double alpha, double sigma_0, const double* sigma_w, const double* sigma_v);

/* ---- the path ------------------------------------------------------------------------------------ */
/* replaces: the pre-loop block of _learn (vbs.h:37-44; mcmcs.h:75-80): initial y-hat, T, e. */
int svbfm_begin(svbfm_t* h);
/* replaces: one pass of the vb loop body: update_all + test predict + evaluation (vbs.h:75-258; vb.h:383-501) */
int svbfm_vb_sweep(svbfm_t* h, svbfm_iter_stats* out);
/* replaces: one pass of the mcmc loop body: draw_all + predict + evaluation (mcmcs.h:96-303; mcmc.h:411-623) */
int svbfm_mcmc_sweep(svbfm_t* h, svbfm_iter_stats* out);
/* replaces: one epoch of the vb_online loop (vbos.h:66-288; vbo.h:354-468). batch_of_case[i] in [0,num_batch)
 * is the batch of local train case i (the host replays std::random_shuffle, vbos.h:74-95). Batches are case
 * subsets of the resident design matrix (the reference writes and re-parses batch files); on two complete fields
 * every batch is swept on its own entries through per-epoch index lists (sharded: per-batch column counts and {A, B}
 * sums are allreduced), otherwise with a batch mask. */
int svbfm_vb_online_epoch(svbfm_t* h, const uint32_t* batch_of_case, uint32_t num_batch, svbfm_iter_stats* out);
/* forget data + state of a handle but keep the device context, communicator and groups (a long-lived service handle:
 * the next learn() starts again at svbfm_set_csc). No reference counterpart: the reference builds a new learner per process. */
int svbfm_reset(svbfm_t* h);
/* n iterations back to back without a host round trip in between; out[n_iter] filled at the end */
int svbfm_run(svbfm_t* h, uint32_t n_iter, svbfm_iter_stats* out);

/* ---- outputs ------------------------------------------------------------------------------------- */
/* replaces: fm_learn::predict(test, out) used by `-out` (libfm.cpp:514-519; mcmc.h:355-379).
 * vb: clamped prediction of the current means (the reference's body is empty, vb.h:321-348; this is the
 * pred_this it would have returned); mcmc: clamped running mean of the clamped draws. */
int svbfm_predict(svbfm_t* h, int32_t split, double* out /*[num_cases]*/);
/* cached residuals e_i in caller case order (vb: y - yhat, mcmc: yhat - y) -- for tests and checkpoints */
int svbfm_get_residuals(svbfm_t* h, double* e /*[num_cases of train]*/);
/* sum_i T_i (vb.h:207-312 gives T_i; only its sum enters alpha and the free energy) */
int svbfm_get_sum_t(svbfm_t* h, double* sum_t);
/* checkpoint restore (after svbfm_begin, one GPU): the cached residuals in caller case order and sum_i T_i as svbfm_get_residuals /
 * svbfm_get_sum_t handed them out. Together with svbfm_set_state / svbfm_set_hyper a vb run continues bit for bit where the saved one
 * stopped (the CLI's -save_model / -load_model). No reference counterpart: the fork has no model files. */
int svbfm_set_residuals(svbfm_t* h, const double* e /*[num_cases of train]*/, double sum_t);
/* stream schedule self-check: max |e2[p] - e[case of p]| between the two residual copies (0.0 expected: the copies
 * are updated with identical arithmetic); 0.0 when the schedule is not in use */
int svbfm_copies_max_diff(svbfm_t* h, double* max_abs_diff);

/* introspection for DESIGN/bench: number of field runs, tiles, kernel launches issued so far */
typedef struct svbfm_info {
    uint32_t num_runs;         /* maximal groups of consecutive, row-disjoint columns (one launch set each) */
    uint32_t num_tiles;
    uint32_t uniform_row_nnz;  /* F when every train row has exactly F entries, else 0 */
    uint32_t all_ones;         /* 1 when every x == 1.0f (values elided on the device) */
    uint64_t kernel_launches;  /* kernels launched by this handle so far */
    uint64_t device_bytes;     /* device memory held */
    uint64_t train_nnz;
    uint32_t rows_reordered;
    uint32_t world_size;
    uint32_t fused_schedule;   /* bit 0: the two-field stream schedule (two residual copies, k_stream) is in use;
                                * bit 1: its records are laid out by popularity rank (SVBFM_REC_RANK=1, experiment);
                                * bit 2: its streams are staged by bulk copies (SVBFM_STREAM_TMA=1, experiment);
                                * bit 3: iterations were replayed from a CUDA graph (SVBFM_GRAPH=1, experiment) */
    uint32_t exclusive_blocks; /* multi-GPU: 1 when the ranks hold disjoint column blocks of the first field (no exchange for it inside the sweep) */
} svbfm_info;
int svbfm_get_info(svbfm_t* h, svbfm_info* out);
/* per-kernel-class device time (CUDA events on the launching stream), for bench.py's roofline block.
 * classes: 0 reduce_v (k_sweep_reduce), 1 combine+finalize_v, 2 apply_v, 3 reduce_w, 4 combine+finalize_w, 5 apply_w,
 *          stream schedule (k_stream): 6 stream_v of the first field, 8 stream_v of the second field, 9 stream_w (both fields),
 *          7 the two flush passes at the end of an iteration; 10 NCCL collectives of the sharded stream schedule (inside the
 *          classes 1 / 4 spans; includes the wait for the slowest rank); 11 the block exchange after the sweep */
#define SVBFM_PROFILE_CLASSES 12
int svbfm_set_profile(svbfm_t* h, int32_t enabled);
int svbfm_get_profile(svbfm_t* h, double ms[SVBFM_PROFILE_CLASSES], uint64_t launches[SVBFM_PROFILE_CLASSES]); /* reads and resets */
/* run on an externally owned CUDA stream (cudaStream_t); NULL restores the handle's own stream */
int svbfm_set_stream(svbfm_t* h, void* cuda_stream);

/* ---- host-side helpers (no device work) ------------------------------------------------------------ */
/* replaces: srand(time(NULL)) + fm_model::init + fm.w.init_normal + fm_learn_vb::init RNG consumption
 * (libfm.cpp:123-124, 273, 298/307/313; fm_model.h:92-101; vb.h:709-712). Calls srand(seed). */
int svbfm_host_init_state(long seed, uint32_t D, int32_t K, double init_stdev, int32_t method, double* w0_mean, double* w0_var,
                          double* w_mean, double* w_var, double* v_mean, double* v_var);
/* replaces: std::random_shuffle on the libc stream (vbos.h:74) */
int svbfm_host_random_shuffle(uint32_t* a, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif /* SVBFM_H_ */
