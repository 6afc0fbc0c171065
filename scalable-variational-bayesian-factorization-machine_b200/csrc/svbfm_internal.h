// csrc/svbfm_internal.h -- shared between svbfm_engine.cu (kernels + C-ABI) and svbfm_ingest.cu (device ingest).
//
// Device layout (DESIGN.md "Data layout in HBM"):
//   * residuals e_i: fp64 SoA, one per train case, in DEVICE case order (cases are re-ordered so that the
//     first field run streams; perm[] maps device order -> caller order). Two complete fields: a second copy e2 in
//     the entry order of the second field, kept bit-identical by the stream schedule (no residual gathers at all).
//   * design matrix twice: CSC per column (colptr u64, case id u32, x f32) for the column sweeps and CSR per
//     case (rowptr u64 or implicit i*F, feature id u32, x f32) for "the other fields of this case".
//     x arrays are elided when every x == 1.0f (one-hot data).
//   * parameters: double2 {mean, var} per (factor, feature), row-major [f][j] like DMatrix::value[f][j]
//     (reference src/util/matrix.h:91-109); one factor row (D*16 B) is L2 resident during its sweep.
//   * no per-case T_i, q_i, S2_i, S3_i arrays: only sum_i T_i is ever consumed (vb.h:446-454, 659-663), and
//     q/S2/S3 minus the own column are re-derived from the case's other features (SURVEY.md section 8d caveat).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/svbfm.h"

// Device memory goes through a size-keyed block cache (svbfm_ingest.cu): a freed block is kept and handed out again
// to the next request of exactly the same size on the same device. A learn() on data of the same shape repeats the same
// sequence of sizes, so after the first pass neither ingest nor the sweeps call the driver's allocator at all (plain
// cudaMalloc/cudaFree of tens of GB cost more than the ingest kernels; the stream-ordered pool showed stalls of up to
// seconds when its free list was fragmented: DESIGN.md section 7). All work of a handle is ordered on one stream, so
// reusing a block right after its release is safe. Blocks return to the driver when the last handle is destroyed or when
// an allocation fails.
namespace svb {
cudaError_t sv_malloc(void** p, size_t bytes);
cudaError_t sv_free(void* p);          // null-tolerant
void sv_cache_release();               // hand every cached block back to the driver
void sv_set_owner(const void* owner);  // this thread now allocates / frees for `owner` (a handle; SV_CUDA and dev_alloc set it)
void sv_owner_release(const void* owner);   // the owner's stream is idle: its cached blocks may go to any handle
}

namespace svb {

struct ColPack;
struct OwnPack;

struct DevSplit {
    uint32_t n = 0;          // local cases
    uint32_t n_cols = 0;     // columns present in this split's data_t (max feature id + 1)
    uint64_t nnz = 0;
    uint64_t* colptr = nullptr;   // [ncols_ext+1] (mcmc train: extended to D with empty columns)
    uint32_t ncols_ext = 0;
    uint32_t* crow = nullptr;     // [nnz] case ids, ascending inside each column (device order)
    float* cval = nullptr;        // [nnz] or null when all_ones
    uint64_t* rowptr = nullptr;   // [n+1] or null when uniformF > 0
    uint32_t* rcol = nullptr;     // [nnz] feature ids, ascending inside each case
    float* rval = nullptr;        // [nnz] or null when all_ones
    uint32_t* cother = nullptr;   // [nnz] F == 2 (train): feature id of the case's other entry, aligned with the CSC entries
    float* cother_val = nullptr;  // [nnz] its x (null when all_ones)
    uint32_t uniformF = 0;
    bool all_ones = false;
    float* y = nullptr;           // [n] device order
    uint32_t* perm = nullptr;     // [n] device order -> caller order, null = identity
    std::vector<uint64_t> h_colptr;   // host copy (tiles, runs)
};

// One maximal group of consecutive, pairwise case-disjoint columns: updating all of them in one launch
// equals the reference's ascending sequential order (fm_learn_vb.h:395-405, 427-438).
struct Run {
    uint32_t col_begin = 0, col_end = 0;
    uint32_t tile_begin = 0, tile_end = 0;      // range in the global tile arrays
    uint32_t heavy_begin = 0, heavy_end = 0;    // range in heavy_cols
    uint64_t nnz = 0;
};

struct RowView {
    const uint64_t* rowptr;
    const uint32_t* rcol;
    const float* rval;
    uint32_t F;
};

// device-resident scalar state; kernels read it, single-thread kernels update it
struct Scalars {
    double alpha, sigma_0;        // vb: alpha, sigma_0          mcmc: alpha, reg0
    double w0_mean, w0_var;       // vb: mu_0', sigma_0'         mcmc: w0, unused
    double sum_t;                 // sum_i T_i (global)
    double n_total, nt_total;     // global number of train / test cases
    double w0_delta;              // pending shift of every e_i
    double red[8];                // reduction results (global after allreduce)
    double min_target, max_target;
    double nat_mu_0, nat_sg_0, rho_0;   // vb_online natural parameters / rate of w0 and the hyper-parameters
    double batch_n;               // vb_online: (global) number of cases in the current batch
    unsigned long long nan_inf;
    uint32_t iter;
    uint32_t t_w0;
    int32_t alpha_ok;
    int32_t pad;
};

// stats slot written on the device once per iteration
struct DevStats {
    double test_rmse, train_stat, free_energy, alpha, rmse_this;
    double has_fe, nan_inf, pad;
};

typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);

struct Engine {
    svbfm_config cfg{};
    std::string err;
    int dev = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;    // second stream of the ingest (H2D of values / targets overlaps the sort of the case ids)
    cudaEvent_t copy_event = nullptr;
    uint32_t D = 0, G = 1;
    int K = 0;
    // groups
    std::vector<uint32_t> h_group, h_n_per_group;
    uint32_t* d_group = nullptr;
    double* d_n_per_group = nullptr;
    // data
    DevSplit tr, te;
    std::vector<Run> runs;
    uint32_t* d_tile_col = nullptr;
    uint64_t* d_tile_begin = nullptr;
    uint32_t* d_tile_len = nullptr;
    uint32_t* d_exec_order = nullptr;  // per run: tile ids in execution order (block-major for cut columns)
    uint32_t* d_col_tile0 = nullptr;   // [ncols_ext+1] first tile of each column
    uint32_t* d_heavy_cols = nullptr;
    uint32_t n_tiles = 0, n_heavy = 0, tile_entries = 1024;
    bool rows_reordered = false;
    bool run0_sequential = false;      // device case order = entry order of run 0 (run 0 holds every case exactly once)
    // two-copy stream schedule (two complete one-hot fields; kernels.cuh k_stream)
    bool streams = false;
    uint32_t ts_shift = 10;            // implicit tile = 2^ts_shift entries (largest power of two <= tile_entries, >= 32)
    bool ts_auto = false;              // no tile size was asked for: small train splits get smaller tiles (svbfm_ingest.cu)
    uint32_t s_ntiles[2] = {0, 0};     // implicit tiles of run 0 / run 1
    uint32_t* d_stile_col0 = nullptr;  // [s_ntiles[0] + s_ntiles[1]] first column of every implicit tile
    uint32_t* d_span_heavy = nullptr;  // columns spanning more than SV_SPAN_LIGHT tiles, run 0 then run 1
    uint32_t span_heavy_n[2] = {0, 0};
    double* d_e2 = nullptr;            // [n] residuals in the entry order of run 1
    struct OwnPack* d_opack = nullptr; // [D] own-side constants of the next pass
    double2* d_ab = nullptr;           // [D] sharded: {A, B} of every column, the allreduce buffer
    double2* d_pvT = nullptr;          // [D][K] transposed factor parameters for the two-field train prediction (k_predict2)
    // vb_online on the stream schedule (single GPU, two complete fields): per epoch, for each field the entries of every
    // batch in column order (idx: positions inside the run) and the batches' own column pointers into idx
    bool vbo_streams = false;
    uint32_t* d_vbo_idx[2] = {nullptr, nullptr};           // [n] each
    unsigned long long* d_vbo_colptr[2] = {nullptr, nullptr};   // [num_batch * ncols(run) + 1] each
    uint32_t* d_vbo_gcnt[2] = {nullptr, nullptr};          // sharded: [num_batch * ncols(run)] GLOBAL number of batch entries per column
    uint32_t* d_vbo_clist[2] = {nullptr, nullptr};         // the non-empty columns of every batch (column ids, batch after batch): what a batch's finalize walks
    std::vector<uint32_t> vbo_clist_off[2];                // [num_batch + 1] first list entry of every batch
    uint32_t* d_vbo_tile_col0 = nullptr;                   // [2][vbo_max_tiles]
    double* d_vbo_partial = nullptr;                       // [2][vbo_max_tiles][2][4]
    uint32_t vbo_max_tiles = 0;
    std::vector<uint64_t> vbo_off;                         // [num_batch + 1] first entry of every batch in idx
    // ... and the batch in flight packed into contiguous streams (k_vbo_pack): a batch is a random 1/num_batch of the cases, so
    // reading its residuals and other-column ids through idx costs a 32-byte sector per 8- / 4-byte value in each of the
    // 2 (K + 1) passes of the batch; packed once per batch, the passes stream them like a whole-run pass does
    double* d_vbo_eb[2] = {nullptr, nullptr};              // [vbo_batch_cap] residuals of the batch in idx order of run 0 / run 1
    uint32_t* d_vbo_ocb[2] = {nullptr, nullptr};           // [vbo_batch_cap] other-column id / record slot of every batch entry
    uint32_t* d_vbo_ownb[2] = {nullptr, nullptr};          // [vbo_batch_cap] own column of every batch entry (k_stream_rows)
    float* d_vbo_xb[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [side][own x, other x][vbo_batch_cap] (x != 1 only)
    uint32_t vbo_batch_cap = 0;
    // ... and, on one GPU, the batch's non-empty columns as a dense id space (kernels.cuh VboCol): per epoch the rank of every
    // (batch, column) among the non-empty pairs, per batch the column table and dense column sums / own constants / d(sum T)
    uint32_t* d_vbo_cpos[2] = {nullptr, nullptr};          // [num_batch * ncols(run) + 1] exclusive scan of the non-empty flags
    struct VboCol* d_vbo_cc = nullptr;                     // [vbo_cols_cap]
    unsigned long long* d_vbo_ccptr = nullptr;             // [vbo_cols_cap + 2]
    struct OwnPack* d_vbo_opack_c = nullptr;               // [vbo_cols_cap]
    double* d_vbo_colsum_c = nullptr;                      // [vbo_cols_cap][4]
    double* d_vbo_dT_c = nullptr;                          // [vbo_cols_cap]
    double2* d_vbo_nextp_c = nullptr;                      // [vbo_cols_cap] carried from one step's finalize to the next (FinalizeArgs)
    double* d_vbo_prevm_c = nullptr;                       // [vbo_cols_cap]
    uint32_t vbo_cols_cap = 0;
    bool vbo_compact = true;                               // SVBFM_VBO_COMPACT=0: global column ids in the batch passes and finalizes
    bool no_predict2 = false, predict2_half = false;      // SVBFM_NO_PREDICT2 / SVBFM_PREDICT2_HALFWARP (read at svbfm_create)
    bool vbo_full_passes = false;                          // SVBFM_VBO_FULL_PASSES=1: masked passes over the whole arrays per batch (comparison)
    bool vbo_rows = true;                                  // SVBFM_VBO_ROWS=0: packed batches go through k_stream instead of k_stream_rows
    bool vbo_pack = true;                                  // SVBFM_VBO_PACK=0: the passes read through idx (round 1 / 2 path, kept for comparison)
    struct BatchView {                                      // what launch_stream / combine_finalize use instead of the whole run
        bool on = false;
        bool lists = false;                                 // prediction / reductions / w0 shift of the batch walk its case list too
        bool compact = false;                               // ... and its columns are a dense id space (d_vbo_cc): first field [0, nl0), second [nl0, nl0 + nl1)
        bool packed = false;                                // the batch's streams are packed (d_vbo_eb / d_vbo_ocb): passes, reductions and shifts run on them
        const uint64_t* colptr[2] = {nullptr, nullptr};
        const uint32_t* gcnt[2] = {nullptr, nullptr};       // sharded: global batch entries per column (same indexing as colptr)
        const uint32_t* clist[2] = {nullptr, nullptr};      // the batch's non-empty columns of run 0 / 1 (null: all columns)
        uint32_t nclist[2] = {0, 0};
        uint64_t entry0 = 0;
        uint32_t n = 0, ntiles = 0;
    } bv;
    uint32_t vbo_ts_shift = 9;          // tile size of the batch passes: small enough that one batch fills the GPU with warps
    // state
    double2* d_pw = nullptr;          // [D]
    double2* d_pv = nullptr;          // [K][D]
    double* d_hyper_w = nullptr;      // vb: sigma_w[G]      mcmc: w_lambda[G]
    double* d_hyper_v = nullptr;      // vb: sigma_v[G][K]   mcmc: v_lambda[G][K]
    double* d_mu_w = nullptr;         // mcmc: w_mu[G]
    double* d_mu_v = nullptr;         // mcmc: v_mu[G][K]
    Scalars* d_sc = nullptr;
    bool have_state = false, begun = false;
    // vb_online per-parameter natural state
    double2* d_nat_w = nullptr;       // [D]   {eta1, eta2}
    double2* d_nat_v = nullptr;       // [K][D]
    uint32_t* d_t_w = nullptr;        // [D] t_wj
    uint32_t* d_t_v = nullptr;        // [D] t_vj
    double* d_col_count = nullptr;    // [D] global count of each feature in train
    uint16_t* d_rbatch = nullptr;     // [n] batch of every case (device order), this epoch
    uint16_t* d_cbatch = nullptr;     // [nnz] batch of the case of every CSC entry
    double* d_cnt_col = nullptr;      // [D] batch entries per column (written by the w pass)
    unsigned long long* d_batch_cnt = nullptr;
    double* d_batch_n = nullptr;      // [num_batch] global batch sizes
    uint32_t batch_cap = 0;
    // work buffers
    double* d_e = nullptr;            // [n]
    double* d_partial = nullptr;      // [n_tiles][4]
    double* d_colsum = nullptr;       // [D][4]
    double* d_delta = nullptr;        // [D]
    struct ColPack* d_cpack = nullptr; // [D] stream schedule: 32-byte per-column records gathered by the other side's pass
    // Rank layout (default; SVBFM_REC_RANK=0 turns it off): the records of the second field's columns are laid out by popularity
    // rank (slot = rank by descending column length) and the cases of every first-field column are ordered by that rank, so
    // that neighbouring lanes of the first field's pass gather neighbouring records (13 instead of 29 distinct 128-byte lines per
    // warp gather at the 200 M shape; first-field pass 1.17 -> 1.08 ms). d_rec_slot[j] = record slot of column j
    // (identity outside the second field); the first field's `cother` entries then hold slots, not column ids.
    uint32_t* d_rec_slot = nullptr;    // [D]
    bool want_rec_rank = true;
    bool use_graph = false;            // SVBFM_GRAPH=1 (experiment): iterations 1.. of svbfm_run replay a CUDA graph of one iteration
    uint64_t graph_replays = 0;
    bool stream_tma = true;            // k_stream's all-ones streams go through a shared-memory ring of bulk copies (SVBFM_STREAM_TMA=0: plain loads)
    bool rec_rank = false;
    int rec_na_mask = 2;               // bit s: the pass over side s gathers its records with L1::no_allocate (kernels.cuh sv_load_record)
    int rec_hot = 2048;                // rank layout: the first field's pass allocates L1 lines only for the rec_hot most popular records (0: for all); 49.8 -> 48.2 ms per iteration (profiles/r02_m_*)
    double* d_dT = nullptr;           // [D]
    double* d_red_partial = nullptr;  // reduction scratch
    double* d_grp_sums = nullptr;     // [(K+1)][G][2]
    double* d_pred_test = nullptr;    // [nt] last prediction (vb: clamped)
    double* d_pred_sum = nullptr;     // [nt] mcmc running sum of clamped predictions
    DevStats* d_stats = nullptr;
    uint32_t stats_cap = 0;
    uint64_t launches = 0;
    uint64_t dev_bytes = 0;
    // profiling (svbfm_set_profile): events around each kernel class of the sweeps
    bool profile = false;
    struct ProfSpan { int cls; cudaEvent_t a, b; };
    std::vector<ProfSpan> prof_spans;
    // multi-GPU
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    uint64_t n_total = 0, nt_total = 0;
    // stream schedule on several GPUs: when every rank's cases cover a disjoint, rank-ordered range of run 0's columns
    // (cases sharded by blocks of the first field), run 0 needs no exchange inside the sweep: each rank updates the
    // columns of its block [blk[rank], blk[rank+1]) and the blocks are broadcast once after the sweep.
    bool excl0 = false;
    std::vector<uint32_t> blk;        // [world + 1]
    double2* d_xchg = nullptr;        // staging of the block exchange: send [rows][maxcnt] + recv [world][rows][maxcnt]
    size_t xchg_cap = 0;
    // Cross shards (several GPUs, stream schedule; svbfm_set_csc(SVBFM_TRAIN_SECOND)): the two residual copies are sharded
    // DIFFERENTLY. Copy 1 (e, entry order of the first field) holds the cases of this rank's block of first-field columns
    // (excl0), copy 2 (e2, entry order of the second field) the cases of this rank's block of SECOND-field columns, handed
    // over as a second shard of the same global train set. Every column of either field then has all of its entries on one
    // rank: no sums are reduced between the ranks and no column is updated twice; what travels is the 32-byte record of
    // every updated column (one in-place ncclAllGather per (step, field) over slot-contiguous blocks of `cpack`) and, once
    // per iteration, the parameter blocks (exchange_blocks). Per-rank work per pass: n / world entries and |field| / world
    // columns, against n / world entries and ALL columns of the second field with the allreduce scheme.
    bool xs = false;
    std::vector<uint32_t> blk1;       // [world + 1] blocks of the second field's columns
    struct SecondShard {
        uint32_t n = 0;                    // cases of the second copy on this rank
        uint64_t* colptr = nullptr;        // [tr.ncols_ext + 1], indexed by global column id: entry pointer into the shard (0 below run 1)
        std::vector<uint64_t> h_colptr;    // host copy of the same
        uint32_t* oc = nullptr;            // [n] record slot of the first-field column of every entry
        uint32_t* rcol = nullptr;          // [n][2] {first-field column, second-field column} of every entry (re-prediction)
        float* y = nullptr;                // [n] targets in entry order
    } sec;
    uint32_t slot_base[2] = {0, 0}, slot_max[2] = {0, 0};   // record slots of field f: [slot_base[f], + slot_max[f]) in column order of the blocks (no padding)
    size_t cpack_cap = 0;             // records allocated in d_cpack
    // {new mean, new var} of the columns finalized in a step, slot order. One device allocation holds 16 flag words and two such
    // stages (one per field). A rank's k_finalize leaves its block in its own stage; when the ranks can map each other's
    // allocation (CUDA IPC over NVLink: p2p) k_records_remote raises / waits on flag words and FETCHES the other ranks' blocks
    // from their memory, coalesced in slot order (fused exchange + record build, no collective); otherwise stage 0 is the buffer
    // of an in-place ncclAllGather and the same kernel reads it locally.
    uint32_t* d_col_of_slot = nullptr; // [slots of both fields] column of every record slot (padding: unused)
    unsigned char* d_xipc = nullptr;  // [256 B flags | stage 0 | stage 1]
    double2* d_xstage = nullptr;      // stage 0 (inside d_xipc)
    size_t xstage_cap = 0;            // columns per stage: the larger field
    bool p2p = false;
    unsigned char* peer_base[16] = {nullptr};   // [world] rank r's allocation as mapped here (own: d_xipc)
    uint64_t xs_epoch = 0;            // exchanges done: the value a rank writes into the others' flag words when its stores of an exchange are out
    // what a pass over side 0 / 1 of the stream schedule reads (set_side_views)
    struct SideView {
        const uint64_t* colptr = nullptr;  // indexed by global column id
        uint64_t entry0 = 0;               // colptr[first column of the run]
        uint32_t n = 0;                    // entries of the side
        const uint32_t* oc = nullptr;      // other-column record slot / id per entry; the side's first entry is oc[entry0]
        const float* xv = nullptr;         // x of the entry / of the other entry, indexed like oc (null: all ones)
        const float* xo = nullptr;
    } side[2];
};
void set_side_views(Engine* E);

// svbfm_ingest.cu
int ingest_split(Engine* E, DevSplit& S, bool is_train, uint32_t num_cases, uint32_t num_cols, const uint64_t* colptr,
                 const uint32_t* case_id, const float* x, const float* target);
void free_split(Engine* E, DevSplit& S);
int transpose_on_device(Engine* E, cudaStream_t st, uint32_t num_cases, uint32_t num_cols, uint64_t nnz, const uint64_t* d_rowptr, const uint32_t* d_col,
                        const float* d_x, uint64_t** d_colptr, uint32_t** d_case, float** d_xt);      // svbfm_ingest.cu: CSR -> CSC on the device
int stream_tile_cols(Engine* E);   // svbfm_engine.cu: first column of every implicit tile (k_tile_col0)
int allreduce(Engine* E, void* buf, size_t count, int dtype /*nccl*/, int op /*nccl*/);
int detect_exclusive_blocks(Engine* E);   // svbfm_engine.cu; collective (every rank calls it after the train split is in)
int ingest_second(Engine* E, uint32_t num_cases, uint32_t num_cols, const uint64_t* colptr, const uint32_t* case_id, const float* x,
                  const float* target);       // svbfm_ingest.cu: the second residual copy's shard (cross shards)
void free_second(Engine* E);
int setup_exchange(Engine* E, size_t columns_per_stage);      // svbfm_engine.cu; collective
int detect_blocks(Engine* E, const Run& r, const std::vector<uint64_t>& h_colptr, std::vector<uint32_t>& blk, bool& exclusive);   // collective
int vbo_stream_prepare(Engine* E, uint32_t num_batch);   // svbfm_ingest.cu: per-epoch batch index lists (needs d_rbatch, d_cbatch, d_batch_cnt)

// error helpers
int fail(Engine* E, int code, const std::string& msg);
#define SV_CUDA(E, call)                                                                                   \
    do {                                                                                                   \
        svb::sv_set_owner(E);       /* the block cache hands out and takes back blocks per handle */       \
        cudaError_t _e = (call);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            return svb::fail((E), SVBFM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));   \
    } while (0)

template <typename T>
int dev_alloc(Engine* E, T** p, size_t count) {
    if (count == 0) count = 1;
    sv_set_owner(E);
    cudaError_t e = sv_malloc((void**)p, count * sizeof(T));
    if (e != cudaSuccess) return fail(E, SVBFM_ERR_OOM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    E->dev_bytes += count * sizeof(T);
    return 0;
}

}  // namespace svb
