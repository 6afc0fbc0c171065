// csrc/svbfm_engine.cu -- C-ABI (include/svbfm.h) and launch scheduling of the VB / MCMC sweep on one B200.
//
// An iteration is a fixed sequence of launches on one stream with every scalar (alpha, sigma_0, w0, sum T,
// hyper-parameters, counters, evaluation sums) resident on the device; the host only enqueues, and reads the
// per-iteration stats slots once at the end of svbfm_run / svbfm_*_sweep. When a communicator is attached,
// the per-run column sums and the few scalar sums go through ncclAllReduce on the same stream.
#include <dlfcn.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include "svbfm_kernels.cuh"

using namespace svb;

static thread_local std::string g_create_error;
static std::atomic<int> g_handles{0};

// ---------------------------------------------------------------------------------------------- NCCL (dlopen)
struct Id128 { char b[SVBFM_COMM_ID_BYTES]; };
namespace {
struct Nccl {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ Id128, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
}  // namespace
static Nccl g_nccl;
static std::once_flag g_nccl_once;

static bool nccl_load() {
    std::call_once(g_nccl_once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !g_nccl.lib; i++) g_nccl.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!g_nccl.lib) return;
        g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.lib, "ncclGetUniqueId");
        g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(g_nccl.lib, "ncclCommInitRank");
        g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
        g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
        g_nccl.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclBroadcast");
        g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllGather");
        g_nccl.GroupStart = (int (*)())dlsym(g_nccl.lib, "ncclGroupStart");
        g_nccl.GroupEnd = (int (*)())dlsym(g_nccl.lib, "ncclGroupEnd");
        g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    });
    return g_nccl.lib && g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy && g_nccl.Broadcast &&
           g_nccl.AllGather && g_nccl.GroupStart && g_nccl.GroupEnd;
}

namespace svb {
int allreduce(Engine* E, void* buf, size_t count, int dtype, int op) {
    if (E->world <= 1 || count == 0) return 0;
    int r = g_nccl.AllReduce(buf, buf, count, dtype, op, E->nccl_comm, E->stream);
    if (r != 0) return fail(E, SVBFM_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    return 0;
}
}  // namespace svb
namespace svb {
int stream_tile_cols(Engine* E);
}
static inline int allreduce_sum_f64(Engine* E, double* buf, size_t n) { return allreduce(E, buf, n, 8 /*ncclDouble*/, 0 /*ncclSum*/); }

// ---------------------------------------------------------------------------------------------- helpers
static inline unsigned nblk(uint64_t n, unsigned t = 256) { return (unsigned)((n + t - 1) / t); }
#define LAUNCHED(E) ((E)->launches++)
// reduction scratch layout (doubles): [0, RGRID*4) block partials | [RGRID*4, +8) final sums | group partials
#define SCR_FINAL ((size_t)SV_RGRID * 4)
#define SCR_GROUP ((size_t)SV_RGRID * 4 + 8)
static cudaError_t copy_sync(Engine* E, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, E->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(E->stream);
}
#define RED(sc, k) (reinterpret_cast<double*>(reinterpret_cast<char*>(sc) + offsetof(Scalars, red)) + (k))

static RowView row_view(const DevSplit& S) { return RowView{S.rowptr, S.rcol, S.rval, S.uniformF}; }
static int fmt_of(const DevSplit& S) { return S.uniformF == 2 ? 2 : (S.uniformF > 0 ? 1 : 0); }

// dispatch on (row format, all-ones)
#define DISPATCH_FMT(S, CALL)                                                            \
    do {                                                                                 \
        int _ft = fmt_of(S);                                                             \
        if ((S).all_ones) {                                                              \
            if (_ft == 2) { CALL(2, true); } else if (_ft == 1) { CALL(1, true); } else { CALL(0, true); }   \
        } else {                                                                         \
            if (_ft == 2) { CALL(2, false); } else if (_ft == 1) { CALL(1, false); } else { CALL(0, false); } \
        }                                                                                \
    } while (0)

static int check_launch(Engine* E, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(E, SVBFM_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return 0;
}

// profiling spans: events on the launching stream around one kernel class
struct ProfScope {
    Engine* E; int cls; cudaEvent_t a = nullptr;
    ProfScope(Engine* E_, int cls_) : E(E_), cls(cls_) {
        if (E->profile) { cudaEventCreate(&a); cudaEventRecord(a, E->stream); }
    }
    ~ProfScope() {
        if (a) { cudaEvent_t b; cudaEventCreate(&b); cudaEventRecord(b, E->stream); E->prof_spans.push_back({cls, a, b}); }
    }
};

namespace svb {
// Do the ranks hold disjoint, rank-ordered ranges of the run's columns? (cases sharded by blocks of that field). Collective.
int detect_blocks(Engine* E, const Run& r0, const std::vector<uint64_t>& cp, std::vector<uint32_t>& blk, bool& exclusive) {
    exclusive = false;
    uint32_t lo = r0.col_end, hi = r0.col_begin;            // first / one past the last non-empty column of this rank
    for (uint32_t j = r0.col_begin; j < r0.col_end; j++)
        if (cp[j + 1] > cp[j]) { if (lo == r0.col_end) lo = j; hi = j + 1; }
    std::vector<uint32_t> v((size_t)E->world * 2, 0u);
    v[(size_t)E->rank * 2] = lo; v[(size_t)E->rank * 2 + 1] = hi;
    uint32_t* d = nullptr;
    SV_CUDA(E, sv_malloc((void**)&d, v.size() * 4));
    SV_CUDA(E, cudaMemcpyAsync(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice, E->stream));
    if (int rc = allreduce(E, d, v.size(), 3 /*ncclUint32*/, 0 /*ncclSum*/)) return rc;
    SV_CUDA(E, cudaMemcpyAsync(v.data(), d, v.size() * 4, cudaMemcpyDeviceToHost, E->stream));
    SV_CUDA(E, cudaStreamSynchronize(E->stream));
    sv_free(d);
    bool ok = true;
    uint32_t prev_hi = r0.col_begin;
    for (int r = 0; r < E->world; r++) {
        uint32_t l = v[(size_t)r * 2], h = v[(size_t)r * 2 + 1];
        if (h <= l) continue;                              // no cases on that rank
        if (l < prev_hi) ok = false;
        prev_hi = h;
    }
    if (!ok) return 0;
    blk.assign((size_t)E->world + 1, r0.col_end);
    for (int r = E->world - 1; r >= 1; r--) {
        uint32_t l = v[(size_t)r * 2], h = v[(size_t)r * 2 + 1];
        blk[r] = (h > l) ? l : blk[r + 1];
    }
    blk[0] = r0.col_begin;
    exclusive = true;
    return 0;
}
int detect_exclusive_blocks(Engine* E) {
    E->excl0 = false;
    if (E->world <= 1 || E->runs.size() != 2 || getenv("SVBFM_NO_EXCL")) return 0;
    return detect_blocks(E, E->runs[0], E->tr.h_colptr, E->blk, E->excl0);
}
}  // namespace svb

// stream schedule, exclusive blocks: after the sweep every rank hands the parameters of its block of run 0 (w and every
// factor) to the others: pack [rows][maxcnt] -> one ncclAllGather -> unpack the other ranks' blocks
static int exchange_blocks_of(Engine* E, const std::vector<uint32_t>& blk) {
    cudaStream_t st = E->stream;
    const uint32_t rows = (E->cfg.k1 ? 1u : 0u) + (uint32_t)E->K;
    uint32_t maxcnt = 0;
    for (int r = 0; r < E->world; r++) maxcnt = std::max(maxcnt, blk[r + 1] - blk[r]);
    if (!rows || !maxcnt) return 0;
    const size_t per_rank = (size_t)rows * maxcnt;                         // double2 elements
    if (E->xchg_cap < per_rank * (size_t)(E->world + 1)) {
        sv_free(E->d_xchg); E->d_xchg = nullptr;
        if (dev_alloc(E, &E->d_xchg, per_rank * (size_t)(E->world + 1))) return SVBFM_ERR_OOM;
        E->xchg_cap = per_rank * (size_t)(E->world + 1);
    }
    double2* send = E->d_xchg;                                             // [rows][maxcnt]
    double2* recv = E->d_xchg + per_rank;                                  // [world][rows][maxcnt]
    BlockXchg bx{};
    bx.pw = E->cfg.k1 ? E->d_pw : nullptr; bx.pv = E->d_pv; bx.D = E->D; bx.rows = rows; bx.maxcnt = maxcnt; bx.world = E->world; bx.rank = E->rank;
    for (int r = 0; r <= E->world && r <= 16; r++) bx.blk[r] = blk[r];
    const uint32_t own = blk[E->rank + 1] - blk[E->rank];
    if (own) { k_xchg_pack<<<dim3(nblk(own), rows), 256, 0, st>>>(bx, send); LAUNCHED(E); }
    int r = g_nccl.AllGather(send, recv, per_rank * 2, 8 /*ncclDouble*/, E->nccl_comm, st);
    if (r != 0) return fail(E, SVBFM_ERR_NCCL, std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    k_xchg_unpack<<<dim3(nblk(maxcnt), rows, E->world), 256, 0, st>>>(bx, recv); LAUNCHED(E);
    return check_launch(E, "exchange_blocks");
}
static int exchange_blocks(Engine* E) {
    if (!E->excl0) return 0;
    ProfScope pc(E, 11);
    if (E->xs) return 0;      // cross shards: every rank already followed every column's update (k_records_remote)
    return exchange_blocks_of(E, E->blk);
}

// cross shards: what a rank has just finalized travels to every rank: {new mean, new var} of its block of columns, 16 bytes per
// column in slot order (equally sized, padded blocks: one in-place allgather, send block = this rank's part of the receive buffer).
// Every rank then forms the other blocks' records itself and updates its parameter table (k_records_remote).
struct RecPlan;
static int exchange_records(Engine* E, int run, const Run& r, double2* pf, const RecPlan* rp, bool mcmc);
// fallback without peer mappings: rank q's block sits at [q][maxcnt] of stage 0; the pointer is shifted so that it is indexed like the
// compact stage (slot - first slot of the field)
static double2* padded_stage(Engine* E, const std::vector<uint32_t>& blk, const Run& r, int q) {
    uint32_t maxcnt = 1;
    for (int k = 0; k < E->world; k++) maxcnt = std::max(maxcnt, blk[k + 1] - blk[k]);
    return E->d_xstage + (size_t)q * maxcnt - (blk[q] - r.col_begin);
}
static double2* stage_of(Engine* E, unsigned char* base, int run) {      // p2p: one stage per field (a fast rank may already fill the next one)
    return reinterpret_cast<double2*>(base + 256) + (E->p2p && run ? E->xstage_cap : 0);
}

namespace svb {
// cross shards: the allocation [flags | stage 0 | stage 1] and, when CUDA IPC works between the ranks, the other ranks' mappings of it
int setup_exchange(Engine* E, size_t cap) {
    cudaStream_t st = E->stream;
    const size_t bytes = std::max<size_t>(65536, 256 + 2 * cap * sizeof(double2));
    if (cudaMalloc((void**)&E->d_xipc, bytes) != cudaSuccess) { cudaGetLastError(); E->d_xipc = nullptr; return fail(E, SVBFM_ERR_OOM, "cudaMalloc: column stages"); }
    SV_CUDA(E, cudaMemsetAsync(E->d_xipc, 0, 256, st));
    E->d_xstage = reinterpret_cast<double2*>(E->d_xipc + 256);
    E->xstage_cap = cap; E->xs_epoch = 0; E->p2p = false;
    for (int r = 0; r < 16; r++) E->peer_base[r] = nullptr;
    E->peer_base[E->rank] = E->d_xipc;
    // every rank publishes the handle of its allocation; any failure anywhere leaves all ranks on the collective
    const int W = E->world;
    constexpr int HW = (int)(sizeof(cudaIpcMemHandle_t) / 4);
    std::vector<uint32_t> hbuf((size_t)W * (HW + 1), 0u);
    cudaIpcMemHandle_t mine;
    bool ok = !getenv("SVBFM_NO_P2P") && cudaIpcGetMemHandle(&mine, E->d_xipc) == cudaSuccess;
    if (!ok) cudaGetLastError();
    if (ok) { memcpy(&hbuf[(size_t)E->rank * (HW + 1)], &mine, sizeof(mine)); hbuf[(size_t)E->rank * (HW + 1) + HW] = 1u; }
    uint32_t* d = nullptr;
    SV_CUDA(E, sv_malloc((void**)&d, hbuf.size() * 4));
    SV_CUDA(E, cudaMemcpyAsync(d, hbuf.data(), hbuf.size() * 4, cudaMemcpyHostToDevice, st));
    int rc = allreduce(E, d, hbuf.size(), 3 /*ncclUint32*/, 0 /*ncclSum*/);
    if (!rc) { cudaMemcpyAsync(hbuf.data(), d, hbuf.size() * 4, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); }
    if (rc) { sv_free(d); return rc; }
    bool all = true;
    for (int r = 0; r < W; r++) all = all && hbuf[(size_t)r * (HW + 1) + HW] == 1u;
    uint32_t opened = 1u;
    if (all) {
        for (int r = 0; r < W && opened; r++) {
            if (r == E->rank) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, &hbuf[(size_t)r * (HW + 1)], sizeof(h));
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = 0u; }
            else E->peer_base[r] = reinterpret_cast<unsigned char*>(p);
        }
    } else opened = 0u;
    cudaMemcpyAsync(d, &opened, 4, cudaMemcpyHostToDevice, st);
    rc = allreduce(E, d, 1, 3 /*ncclUint32*/, 3 /*ncclMin*/);
    if (!rc) { cudaMemcpyAsync(&opened, d, 4, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); }
    sv_free(d);
    if (rc) return rc;
    if (!opened) {
        for (int r = 0; r < W; r++) if (r != E->rank && E->peer_base[r]) { cudaIpcCloseMemHandle(E->peer_base[r]); E->peer_base[r] = nullptr; }
        return 0;
    }
    E->p2p = true;
    return 0;
}
}  // namespace svb

// stream schedule: what a finalize has to leave behind for the passes that follow (kernels.cuh FinalizeArgs)
struct RecPlan {
    int run = -1;                  // 0 / 1: implicit tiles of that run (span layout); -1: explicit tiles
    int rec_mode = 0;
    const double2* p_next = nullptr;
    const double2* p_prev = nullptr;
    bool carry = false;            // vb_online, compact columns: not the first step of the batch
};

// tiles -> column sums (-> allreduce when sharded) -> per-column update
template <int KIND>
static int combine_finalize(Engine* E, const Run& r, int f, int batch = -1, const RecPlan* rp = nullptr) {
    cudaStream_t st = E->stream;
    constexpr bool IS_V = (KIND == KIND_VB_V || KIND == KIND_MC_V || KIND == KIND_VBO_V);
    double2* pf = IS_V ? E->d_pv + (size_t)f * E->D : E->d_pw;
    (void)batch;
    ProfScope ps(E, (IS_V ? 0 : 3) + 1);
    uint32_t ncols = r.col_end - r.col_begin;
    bool from_colsum = false, use_ab = false;
    SpanView sp{nullptr, 0, E->ts_shift, SV_SPAN_LIGHT};
    const double* partial = E->d_partial;
    if (rp && rp->run >= 0 && E->bv.on) {        // one vb_online batch: its own column pointer; any span is summed in k_finalize_vbo
        sp.colptr = E->bv.colptr[rp->run]; sp.entry0 = E->bv.entry0; sp.light_limit = ~0u; sp.ts_shift = E->vbo_ts_shift;
        partial = E->d_vbo_partial + (rp->run ? (size_t)E->vbo_max_tiles * 8 : 0);
        if (E->world > 1) {       // sharded: complete local sums of every column -> {A, B} through the allreduce (C1, C2 stay local)
            k_combine_light_span<KIND == KIND_VBO_V><<<nblk(ncols), 256, 0, st>>>(r.col_begin, r.col_end, sp, partial, E->d_colsum, E->d_ab); LAUNCHED(E);
            ProfScope pc(E, 10);
            if (int rc = allreduce_sum_f64(E, reinterpret_cast<double*>(E->d_ab + r.col_begin), (size_t)ncols * 2)) return rc;
            from_colsum = true; use_ab = true;
        }
    } else if (rp && rp->run >= 0) {
        sp.colptr = E->side[rp->run].colptr; sp.entry0 = E->side[rp->run].entry0;
        partial = E->d_partial + (rp->run ? (size_t)E->s_ntiles[0] * 8 : 0);
        uint32_t nh = E->span_heavy_n[rp->run], h0 = rp->run ? E->span_heavy_n[0] : 0;
        if (nh) { k_combine_span<<<nh, 128, 0, st>>>(E->d_span_heavy, h0, sp, partial, E->d_colsum); LAUNCHED(E); }
        if (E->world > 1 && !(rp->run == 0 && E->excl0) && !E->xs) {
            k_combine_light_span<KIND == KIND_VB_V><<<nblk(ncols), 256, 0, st>>>(r.col_begin, r.col_end, sp, partial, E->d_colsum, E->d_ab); LAUNCHED(E);
            ProfScope pc(E, 10);      // includes the wait for the slowest rank
            if (int rc = allreduce_sum_f64(E, reinterpret_cast<double*>(E->d_ab + r.col_begin), (size_t)ncols * 2)) return rc;
            from_colsum = true; use_ab = true;
        }
    } else {
        uint32_t nheavy = r.heavy_end - r.heavy_begin;
        if (nheavy) { k_combine_heavy<<<nheavy, 128, 0, st>>>(E->d_heavy_cols, r.heavy_begin, E->d_col_tile0, E->d_partial, E->d_colsum); LAUNCHED(E); }
        if (E->world > 1) {
            k_combine_light<<<nblk(ncols), 256, 0, st>>>(r.col_begin, r.col_end, E->d_col_tile0, E->d_partial, E->d_colsum); LAUNCHED(E);
            if (int rc = allreduce_sum_f64(E, E->d_colsum + (size_t)r.col_begin * 4, (size_t)ncols * 4)) return rc;
            from_colsum = true;
        }
    }
    FinalizeArgs fa{};
    fa.c0 = r.col_begin; fa.c1 = r.col_end; fa.f = IS_V ? f : -1; fa.K = E->K;
    if (rp && rp->run == 0 && E->excl0) {      // only the columns of this rank's block (their sums are complete locally)
        fa.c0 = E->blk[E->rank]; fa.c1 = E->blk[E->rank + 1];
        ncols = fa.c1 - fa.c0;
    } else if (rp && rp->run == 1 && E->xs) {  // cross shards: the same for the second field
        fa.c0 = E->blk1[E->rank]; fa.c1 = E->blk1[E->rank + 1];
        ncols = fa.c1 - fa.c0;
    }
    fa.col_tile0 = E->d_col_tile0; fa.partial = partial; fa.colsum = E->d_colsum; fa.from_colsum = from_colsum;
    fa.pf = pf; fa.group = E->d_group; fa.hyper = IS_V ? E->d_hyper_v : E->d_hyper_w;
    fa.hyper_mu = IS_V ? E->d_mu_v : E->d_mu_w; fa.sc = E->d_sc; fa.delta = E->d_delta; fa.dT = E->d_dT;
    fa.seed = E->cfg.seed; fa.do_sample = E->cfg.do_sample;
    fa.span = sp; fa.ab = use_ab ? E->d_ab : nullptr;
    if (rp) { fa.cpack = E->d_cpack; fa.opack = E->d_opack; fa.rec_mode = rp->rec_mode; fa.p_next = rp->p_next; fa.p_prev = rp->p_prev; }
    fa.rec_slot = E->rec_rank ? E->d_rec_slot : nullptr;
    if (rp && rp->run >= 0 && E->xs) {
        fa.stage = E->p2p ? stage_of(E, E->d_xipc, rp->run) : padded_stage(E, rp->run ? E->blk1 : E->blk, r, E->rank);
        fa.stage_base = E->slot_base[rp->run];
    }
    if (rp && rp->run >= 0 && E->bv.on) fa.gcnt = E->bv.gcnt[rp->run];
    if (rp && rp->run >= 0 && E->bv.on && E->bv.compact) {
        fa.cc = E->d_vbo_cc; fa.cid0 = rp->run ? E->bv.nclist[0] : 0u;
        fa.ccptr = reinterpret_cast<const uint64_t*>(E->d_vbo_ccptr) + (rp->run ? E->bv.nclist[0] + 1 : 0u);
        fa.colsum = E->d_vbo_colsum_c; fa.opack = E->d_vbo_opack_c; fa.dT = E->d_vbo_dT_c;
        fa.nextp_c = E->d_vbo_nextp_c; fa.prevm_c = E->d_vbo_prevm_c; fa.carry = rp->carry ? 1 : 0;
    }
    if constexpr (KIND == KIND_VBO_W || KIND == KIND_VBO_V) {
        fa.nat = IS_V ? E->d_nat_v + (size_t)f * E->D : E->d_nat_w;
        fa.t_cnt = IS_V ? E->d_t_v : E->d_t_w;
        fa.col_count = E->d_col_count;
        fa.update_t = IS_V ? (f == E->K - 1) : 1;      // t_vj advances once per batch (vbo.h:399-402), t_wj per w update (:520)
        int update_params = IS_V ? 1 : (E->cfg.k1 != 0);   // with k1 = 0 the w pass only counts the batch entries per column
        uint32_t nthreads = ncols;
        if (rp && rp->run >= 0 && E->bv.on && E->bv.clist[rp->run]) { fa.col_list = E->bv.clist[rp->run]; fa.n_list = E->bv.nclist[rp->run]; nthreads = fa.n_list; }
        if (nthreads) k_finalize_vbo<KIND><<<nblk(nthreads), 256, 0, st>>>(fa, E->d_cnt_col, 0.5, 1u, update_params);
    } else {
        if (ncols) k_finalize<KIND><<<nblk(ncols), 256, 0, st>>>(fa);
    }
    LAUNCHED(E);
    if (rp && rp->run >= 0 && E->xs) if (int rc = exchange_records(E, rp->run, r, pf, rp, KIND == KIND_MC_W || KIND == KIND_MC_V)) return rc;
    return check_launch(E, "combine_finalize");
}

static int exchange_records(Engine* E, int run, const Run& r, double2* pf, const RecPlan* rp, bool mcmc) {
    cudaStream_t st = E->stream;
    const std::vector<uint32_t>& blk = run ? E->blk1 : E->blk;
    RemoteRecArgs a{};
    if (E->p2p) {
        a.flags.n = E->world; a.flags.me = E->rank;
        for (int q = 0; q < E->world; q++) { a.flags.p[q] = reinterpret_cast<unsigned long long*>(E->peer_base[q]); a.src[q] = stage_of(E, E->peer_base[q], run); }
        a.epoch = ++E->xs_epoch;
    } else {
        // no peer mappings: one in-place ncclAllGather over blocks padded to the largest one ([world][maxcnt]; only this fallback
        // pays the padding: a grouped ncclBroadcast of the exact blocks took 22.8 ms per iteration at 8 GPUs against 10.0 ms)
        ProfScope pc(E, 10);
        uint32_t maxcnt = 1;
        for (int q = 0; q < E->world; q++) maxcnt = std::max(maxcnt, blk[q + 1] - blk[q]);
        double* base = reinterpret_cast<double*>(E->d_xstage);
        const size_t per = (size_t)maxcnt * 2;                                  // doubles per rank
        int rc = g_nccl.AllGather(base + per * (size_t)E->rank, base, per, 8 /*ncclDouble*/, E->nccl_comm, st);
        if (rc != 0) return fail(E, SVBFM_ERR_NCCL, std::string("ncclAllGather (columns): ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
        for (int q = 0; q < E->world; q++) a.src[q] = padded_stage(E, blk, r, q);
    }
    ProfScope pc(E, E->p2p ? 10 : 1);       // p2p: flags + fetch from the peers are the exchange
    a.n = r.col_end - r.col_begin; a.world = E->world; a.me = E->rank;
    for (int q = 0; q <= E->world; q++) a.bnd[q] = blk[q] - r.col_begin;
    a.col_of_slot = E->d_col_of_slot; a.stage_base = E->slot_base[run];
    a.pf = pf; a.p_next = rp->p_next; a.p_prev = rp->p_prev; a.rec_mode = rp->rec_mode; a.mcmc = mcmc ? 1 : 0; a.cpack = E->d_cpack;
    const uint32_t n_remote = a.n - (a.bnd[E->rank + 1] - a.bnd[E->rank]);
    // (launched even without remote slots: the flags of an exchange are raised and awaited by every rank)
    if (a.n) { k_records_remote<<<std::max(1u, nblk(n_remote)), 256, 0, st>>>(a); LAUNCHED(E); }
    return 0;
}

// ---------------------------------------------------------------------------------------------- one run sweep
template <int KIND>
static int sweep_run(Engine* E, const Run& r, int f, int batch = -1) {
    const DevSplit& S = E->tr;
    cudaStream_t st = E->stream;
    uint32_t ntiles = r.tile_end - r.tile_begin;
    constexpr bool IS_V = (KIND == KIND_VB_V || KIND == KIND_MC_V || KIND == KIND_VBO_V);
    double2* pf = IS_V ? E->d_pv + (size_t)f * E->D : E->d_pw;
    SweepArgs a{};
    a.tile_col = E->d_tile_col; a.tile_begin = E->d_tile_begin; a.tile_len = E->d_tile_len; a.exec_order = E->d_exec_order; a.colptr = S.colptr; a.crow = S.crow; a.cval = S.cval;
    a.rv = row_view(S); a.ov = OtherView{S.cother, S.cother_val}; a.e = E->d_e; a.pf = pf; a.partial = E->d_partial; a.delta = E->d_delta;
    a.tile0 = r.tile_begin; a.ntiles = ntiles; a.tile_entries = E->tile_entries;
    a.cbatch = batch >= 0 ? E->d_cbatch : nullptr; a.batch = batch >= 0 ? (uint32_t)batch : 0u;
    const int pc = IS_V ? 0 : 3;
    if (ntiles) {
        ProfScope ps(E, pc + 0);
        unsigned grid = (ntiles + 7) / 8;
        if constexpr (IS_V) {
#define CALL_R(FT, ONES) k_sweep_reduce<KIND, FT, ONES><<<grid, 256, 0, st>>>(a)
            DISPATCH_FMT(S, CALL_R);
#undef CALL_R
        } else {
            if (S.all_ones) k_sweep_reduce<KIND, 0, true><<<grid, 256, 0, st>>>(a);
            else k_sweep_reduce<KIND, 0, false><<<grid, 256, 0, st>>>(a);
        }
        LAUNCHED(E);
    }
    if (int rc = combine_finalize<KIND>(E, r, f, batch)) return rc;
    if (ntiles && r.nnz * 4 >= (uint64_t)S.n) {
        // run touches a large share of the cases: streaming pass in case order
        ProfScope ps(E, pc + 2);
        RowApplyArgs ra{};
        ra.rv = row_view(S); ra.n = S.n; ra.c0 = r.col_begin; ra.c1 = r.col_end; ra.e = E->d_e; ra.pf = pf; ra.delta = E->d_delta;
        ra.rbatch = batch >= 0 ? E->d_rbatch : nullptr; ra.batch = batch >= 0 ? (uint32_t)batch : 0u;
        unsigned grid = std::max(1u, std::min<unsigned>(nblk(S.n), 148 * 16));
#define CALL_RA(FT, ONES) k_row_apply<IS_V, FT, ONES><<<grid, 256, 0, st>>>(ra)
        DISPATCH_FMT(S, CALL_RA);
#undef CALL_RA
        LAUNCHED(E);
    } else if (ntiles) {
        ProfScope ps(E, pc + 2);
        unsigned grid = (ntiles + 7) / 8;
        if constexpr (IS_V) {
#define CALL_A(FT, ONES) k_sweep_apply<true, FT, ONES><<<grid, 256, 0, st>>>(a)
            DISPATCH_FMT(S, CALL_A);
#undef CALL_A
        } else {
            if (S.all_ones) k_sweep_apply<false, 0, true><<<grid, 256, 0, st>>>(a);
            else k_sweep_apply<false, 0, false><<<grid, 256, 0, st>>>(a);
        }
        LAUNCHED(E);
    }
    return check_launch(E, "sweep_run");
}

// Two-copy stream schedule (two complete one-hot fields, device case order = run 0): see k_stream.
static bool stream_ok(const Engine* E) { return E->streams; }

// one pass over `side` (0: run 0 / e, 1: run 1 / e2). W: a w step (KIND_*_W), else v. vb_online shares vb's pass arithmetic.
template <bool MCMC, bool W, bool REDUCE>
static void launch_stream(Engine* E, int side, bool has_own, bool own_is_w, bool has_oth, bool oth_is_w) {
    const DevSplit& S = E->tr;
    const Run& r = E->runs[side];
    StreamArgs a{};
    const Engine::SideView& sv = E->side[side];
    a.c0 = r.col_begin; a.c1 = r.col_end; a.real0 = sv.entry0; a.ts_shift = E->ts_shift;
    if (E->bv.on) {
        a.colptr = E->bv.colptr[side]; a.entry0 = E->bv.entry0; a.n = E->bv.n; a.ntiles = E->bv.ntiles; a.ts_shift = E->vbo_ts_shift;
        a.tile_col0 = E->d_vbo_tile_col0 + (side ? E->vbo_max_tiles : 0); a.idx = E->bv.packed ? nullptr : E->d_vbo_idx[side];
        a.partial = E->d_vbo_partial + (side ? (size_t)E->vbo_max_tiles * 8 : 0);
    } else {
        a.colptr = sv.colptr; a.entry0 = a.real0; a.n = sv.n; a.ntiles = E->s_ntiles[side];
        a.tile_col0 = E->d_stile_col0 + (side ? E->s_ntiles[0] : 0); a.idx = nullptr;
        a.partial = E->d_partial + (side ? (size_t)E->s_ntiles[0] * 8 : 0);
    }
    a.oc = sv.oc; a.xv = sv.xv; a.xo = sv.xo;
    a.e = side ? E->d_e2 : E->d_e;
    if (E->bv.on && E->bv.packed) {     // the batch's own contiguous streams (k_vbo_pack): entry k of the pass is element k
        a.real0 = 0; a.oc = E->d_vbo_ocb[side]; a.e = E->d_vbo_eb[side]; a.ownc = E->d_vbo_ownb[side];
        if (sv.xv) { a.xv = E->d_vbo_xb[side][0]; a.xo = E->d_vbo_xb[side][1]; }
    }
    a.rec = E->d_cpack; a.own = E->d_opack;
    a.rec_no_alloc = (E->rec_na_mask >> side) & 1;
    // rank layout, one GPU: the first field's pass keeps the records of the rec_hot most popular second-field columns in L1 and
    // fetches the others past it (a cold record would take a 128-byte line for its 32 bytes and push a hot one out)
    a.rec_hot_end = 0xffffffffu;
    if (side == 0 && E->rec_rank && !E->xs && E->rec_hot > 0) a.rec_hot_end = E->runs[1].col_begin + (uint32_t)E->rec_hot;
    a.has_own = has_own; a.own_is_w = own_is_w; a.has_oth = has_oth; a.oth_is_w = oth_is_w;
    a.colsum = E->d_colsum;
    if (E->bv.on && E->bv.compact) { a.own = E->d_vbo_opack_c; a.colsum = E->d_vbo_colsum_c; }      // own-column ids of the packed batch are compact
    constexpr int KIND = MCMC ? (W ? KIND_MC_W : KIND_MC_V) : (W ? KIND_VB_W : KIND_VB_V);
    constexpr unsigned SW = SV_STREAM_WARPS, ST = 32 * SV_STREAM_WARPS;
    unsigned grid = (a.ntiles + SW - 1) / SW;
    if (!a.ntiles) return;
    const bool steady = !W && has_own && !own_is_w && has_oth && !oth_is_w;
#define CALL_S(ONES, STEADY)                                                                                         \
    do {                                                                                                             \
        if (a.idx) { if constexpr (!MCMC) k_stream<KIND, ONES, REDUCE, STEADY, true><<<grid, ST, 0, E->stream>>>(a); } \
        else k_stream<KIND, ONES, REDUCE, STEADY, false><<<grid, ST, 0, E->stream>>>(a);                           \
    } while (0)
    if constexpr (!MCMC) {
        if (E->bv.on && E->bv.packed && E->vbo_rows) {      // a packed vb_online batch: rows of 32 entries reduced at once (k_stream_rows)
#define CALL_R(ONES, STEADY) k_stream_rows<KIND, ONES, REDUCE, STEADY><<<grid, ST, 0, E->stream>>>(a)
            if (steady) { if (S.all_ones) CALL_R(true, true); else CALL_R(false, true); }
            else { if (S.all_ones) CALL_R(true, false); else CALL_R(false, false); }
#undef CALL_R
            LAUNCHED(E);
            return;
        }
    }
    // all-ones streams are staged through shared memory by bulk copies (SVBFM_STREAM_TMA=0: plain loads); needs 16-byte aligned stream starts
    const bool tma = E->stream_tma && S.all_ones && !a.idx && (a.real0 % 4 == 0);
    if (tma) {
        if (steady) k_stream<KIND, true, REDUCE, true, false, true><<<grid, ST, 0, E->stream>>>(a);
        else k_stream<KIND, true, REDUCE, false, false, true><<<grid, ST, 0, E->stream>>>(a);
    }
    else if (steady) { if (S.all_ones) CALL_S(true, true); else CALL_S(false, true); }
    else { if (S.all_ones) CALL_S(true, false); else CALL_S(false, false); }
#undef CALL_S
    LAUNCHED(E);
}

// update_w + update_v of one iteration (vb.h:390-440 / mcmc.h:465-623 without the hyper-parameter draws), or of one
// vb_online batch (vbo.h:360-408; E->bv set). FLAVOR: 0 vb, 1 mcmc, 2 vb_online
template <int FLAVOR>
static int sweep_streams(Engine* E) {
    constexpr bool MCMC = (FLAVOR == 1);
    const Run &r0 = E->runs[0], &r1 = E->runs[1];
    cudaStream_t st = E->stream;
    std::vector<int> steps;                       // -1 = w, f = v_f
    if (E->cfg.k1) steps.push_back(-1);
    for (int f = 0; f < E->K; f++) steps.push_back(f);
    if (steps.empty()) return 0;
    auto table = [&](int s) -> double2* { return s < 0 ? E->d_pw : E->d_pv + (size_t)s * E->D; };
    k_pack_init<<<nblk(r1.col_end - r0.col_begin), 256, 0, st>>>(r0.col_begin, r0.col_end, r1.col_begin, r1.col_end, table(steps[0]), E->d_cpack, E->d_opack,
                                                                          E->rec_rank ? E->d_rec_slot : nullptr);
    LAUNCHED(E);
    if (E->bv.on && E->bv.compact) {
        const uint32_t nc = E->bv.nclist[0] + E->bv.nclist[1];
        if (nc) { k_vbo_opack_init<<<nblk(nc), 256, 0, st>>>(E->d_vbo_cc, E->bv.nclist[0], nc, table(steps[0]), E->d_vbo_opack_c); LAUNCHED(E); }
    }
    constexpr int KW = FLAVOR == 1 ? KIND_MC_W : (FLAVOR == 2 ? KIND_VBO_W : KIND_VB_W);
    constexpr int KV = FLAVOR == 1 ? KIND_MC_V : (FLAVOR == 2 ? KIND_VBO_V : KIND_VB_V);
    const int batch = E->bv.on ? 0 : -1;
    for (size_t k = 0; k < steps.size(); k++) {
        const int s = steps[k];
        const bool first = (k == 0), prev_w = (!first && steps[k - 1] < 0), w = (s < 0);
        RecPlan rp;
        rp.carry = !first;
        rp.p_next = (k + 1 < steps.size()) ? table(steps[k + 1]) : nullptr;
        {   // first field: pending U(s-1), I(s-1) + pass 1
            ProfScope ps(E, w ? 9 : 6);
            if (w) launch_stream<MCMC, true, true>(E, 0, !first, prev_w, !first, prev_w);
            else launch_stream<MCMC, false, true>(E, 0, !first, prev_w, !first, prev_w);
        }
        rp.run = 0; rp.rec_mode = 1; rp.p_prev = first ? nullptr : table(steps[k - 1]);
        if (int rc = w ? combine_finalize<KW>(E, r0, s, batch, &rp) : combine_finalize<KV>(E, r0, s, batch, &rp)) return rc;
        {   // second field: pending I(s-1), U(s) + pass 1
            ProfScope ps(E, w ? 9 : 8);
            if (w) launch_stream<MCMC, true, true>(E, 1, !first, prev_w, true, true);
            else launch_stream<MCMC, false, true>(E, 1, !first, prev_w, true, false);
        }
        rp.run = 1; rp.rec_mode = 2; rp.p_prev = nullptr;
        if (int rc = w ? combine_finalize<KW>(E, r1, s, batch, &rp) : combine_finalize<KV>(E, r1, s, batch, &rp)) return rc;
    }
    {   // flush the last step's updates into both copies
        ProfScope ps(E, 7);
        const bool lw = steps.back() < 0;
        launch_stream<MCMC, false, false>(E, 0, true, lw, true, lw);
        if (E->xs) {      // the first field's final means of the other ranks' blocks are not in the parameter tables yet: they are the records' own means
            const uint32_t ns = E->slot_max[0];
            k_pack_h4_self<<<nblk(ns), 256, 0, st>>>(E->slot_base[0], E->slot_base[0] + ns, E->d_cpack); LAUNCHED(E);
        } else { k_pack_h4<<<nblk(r0.col_end - r0.col_begin), 256, 0, st>>>(r0.col_begin, r0.col_end, table(steps.back()), E->d_cpack); LAUNCHED(E); }
        launch_stream<MCMC, false, false>(E, 1, true, lw, false, false);
    }
    return check_launch(E, "sweep_streams");
}

namespace svb {
int stream_tile_cols(Engine* E) {
    uint32_t off = 0;
    for (int ri = 0; ri < 2; ri++) {
        const Run& r = E->runs[ri];
        uint32_t nt = E->s_ntiles[ri];
        if (nt) { k_tile_col0<<<(nt + 255) / 256, 256, 0, E->stream>>>(E->side[ri].colptr, r.col_begin, r.col_end, nt, E->ts_shift, E->d_stile_col0 + off); LAUNCHED(E); }
        off += nt;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(E, SVBFM_ERR_CUDA, std::string("k_tile_col0: ") + cudaGetErrorString(e));
    return 0;
}
}  // namespace svb

// every e_i += w0_delta, on both copies
static void shift_e(Engine* E) {
    if (E->bv.on && E->bv.packed) {
        const uint32_t n = E->bv.n, grid = std::max(1u, std::min<unsigned>(nblk(n), SV_RGRID));
        k_shift_e<<<grid, 256, 0, E->stream>>>(E->d_vbo_eb[0], n, E->d_sc); LAUNCHED(E);
        k_shift_e<<<grid, 256, 0, E->stream>>>(E->d_vbo_eb[1], n, E->d_sc); LAUNCHED(E);
        return;
    }
    if (E->bv.on && E->bv.lists) {      // vb_online batch on the stream schedule: only the entries of the batch are live
        const uint32_t n = E->bv.n, grid = std::max(1u, std::min<unsigned>(nblk(n), SV_RGRID));
        k_shift_e_list<<<grid, 256, 0, E->stream>>>(E->d_e, E->d_vbo_idx[0] + E->bv.entry0, n, E->d_sc); LAUNCHED(E);
        k_shift_e_list<<<grid, 256, 0, E->stream>>>(E->d_e2, E->d_vbo_idx[1] + E->bv.entry0, n, E->d_sc); LAUNCHED(E);
        return;
    }
    k_shift_e<<<SV_RGRID, 256, 0, E->stream>>>(E->d_e, E->tr.n, E->d_sc); LAUNCHED(E);
    if (E->d_e2) { k_shift_e<<<SV_RGRID, 256, 0, E->stream>>>(E->d_e2, E->xs ? E->sec.n : E->tr.n, E->d_sc); LAUNCHED(E); }
}
// (re)build the second copy from the first
static int predict_second(Engine* E);
static void sync_e2(Engine* E) {
    if (E->xs) { predict_second(E); return; }      // cross shards: the second copy lives on other cases: predicted from the parameters
    if (!E->streams || !E->tr.n) return;
    const Run& r1 = E->runs[1];
    k_gather_e<<<nblk(E->tr.n), 256, 0, E->stream>>>(E->d_e, E->tr.crow + E->tr.h_colptr[r1.col_begin], E->tr.n, E->d_e2); LAUNCHED(E);
}

// sum e, sum e^2, sum clamp(e)^2 -> red[0..2] (global)
static int reduce_e(Engine* E, int batch = -1) {
    cudaStream_t st = E->stream;
    if (E->bv.on && E->bv.packed) {     // packed batch: its residuals are one contiguous array
        const uint32_t n = E->bv.n, grid = std::max(1u, std::min<unsigned>(nblk(n), SV_RGRID));
        k_reduce_e<<<grid, 256, 0, st>>>(E->d_vbo_eb[0], n, E->d_sc, E->d_red_partial, nullptr, 0u); LAUNCHED(E);
        k_reduce_final<<<1, 256, 0, st>>>(E->d_red_partial, grid, 3, RED(E->d_sc, 0), 0); LAUNCHED(E);
        return allreduce_sum_f64(E, RED(E->d_sc, 0), 3);
    }
    if (E->bv.on && E->bv.lists) {      // vb_online batch on the stream schedule: the batch's own case list (positions in e = device case ids)
        const uint32_t n = E->bv.n, grid = std::max(1u, std::min<unsigned>(nblk(n), SV_RGRID));
        k_reduce_e_list<<<grid, 256, 0, st>>>(E->d_e, E->d_vbo_idx[0] + E->bv.entry0, n, E->d_sc, E->d_red_partial); LAUNCHED(E);
        k_reduce_final<<<1, 256, 0, st>>>(E->d_red_partial, grid, 3, RED(E->d_sc, 0), 0); LAUNCHED(E);
        return allreduce_sum_f64(E, RED(E->d_sc, 0), 3);
    }
    k_reduce_e<<<SV_RGRID, 256, 0, st>>>(E->d_e, E->tr.n, E->d_sc, E->d_red_partial, batch >= 0 ? E->d_rbatch : nullptr, batch >= 0 ? (uint32_t)batch : 0u); LAUNCHED(E);
    k_reduce_final<<<1, 256, 0, st>>>(E->d_red_partial, SV_RGRID, 3, RED(E->d_sc, 0), 0); LAUNCHED(E);
    return allreduce_sum_f64(E, RED(E->d_sc, 0), 3);
}

static int group_sums(Engine* E, bool mcmc) {
    cudaStream_t st = E->stream;
    double* partial = E->d_red_partial + SCR_GROUP;
    dim3 grid(SV_GGRID, E->K + 1);
    if (mcmc) k_group_sums<true><<<grid, 256, 0, st>>>(E->d_pw, E->d_pv, E->D, E->G, E->d_group, partial);
    else k_group_sums<false><<<grid, 256, 0, st>>>(E->d_pw, E->d_pv, E->D, E->G, E->d_group, partial);
    LAUNCHED(E);
    uint32_t nrows = (uint32_t)(E->K + 1) * E->G * 2;
    k_group_sums_final<<<nrows, 64, 0, st>>>(partial, nrows, E->d_grp_sums); LAUNCHED(E);
    return check_launch(E, "group_sums");
}

template <int MODE>
static int predict(Engine* E, const DevSplit& S, double* e_out, int red_slot, int nred, int batch = -1) {
    cudaStream_t st = E->stream;
    PredictArgs a{};
    a.rv = row_view(S); a.y = S.y; a.n = S.n; a.pw = E->d_pw; a.pv = E->d_pv; a.D = E->D; a.K = E->K;
    a.k0 = E->cfg.k0; a.k1 = E->cfg.k1; a.sc = E->d_sc; a.e = e_out; a.pred = E->d_pred_test; a.pred_sum = E->d_pred_sum;
    a.partial = E->d_red_partial;
    a.rbatch = batch >= 0 ? E->d_rbatch : nullptr; a.batch = batch >= 0 ? (uint32_t)batch : 0u;
    unsigned grid = std::max(1u, std::min<unsigned>(nblk(S.n), SV_RGRID));
    bool listed = false;
    if constexpr (MODE == PRED_VB_TRAIN) {
        if (E->bv.on && E->bv.lists) {     // vb_online batch on the stream schedule: walk the batch's own case list
            listed = true;
            a.list = E->d_vbo_idx[0] + E->bv.entry0; a.nlist = E->bv.n;
            grid = std::max(1u, std::min<unsigned>(nblk(a.nlist), SV_RGRID));
#define CALL_PL(FT, ONES) k_predict<MODE, FT, ONES, true><<<grid, 256, 0, st>>>(a)
            DISPATCH_FMT(S, CALL_PL);
#undef CALL_PL
        }
    }
    if (!listed) {
#define CALL_P(FT, ONES) k_predict<MODE, FT, ONES><<<grid, 256, 0, st>>>(a)
        DISPATCH_FMT(S, CALL_P);
#undef CALL_P
    }
    LAUNCHED(E);
    k_reduce_final<<<1, 256, 0, st>>>(E->d_red_partial, grid, 4, E->d_red_partial + SCR_FINAL, 0); LAUNCHED(E);
    // move the first nred sums into red[red_slot..]
    SV_CUDA(E, cudaMemcpyAsync(RED(E->d_sc, red_slot), E->d_red_partial + SCR_FINAL, sizeof(double) * nred, cudaMemcpyDeviceToDevice, st));
    if (int rc = allreduce_sum_f64(E, RED(E->d_sc, red_slot), nred)) return rc;
    return check_launch(E, "predict");
}

// train prediction, two complete fields: transposed parameters + warp per case (k_predict2). Falls back to k_predict.
template <int MODE>   // PRED_VB_TRAIN or PRED_MC_TRAIN
static int predict_train(Engine* E, int red_slot) {
    const DevSplit& S = E->tr;
    if (red_slot >= 0 && (!E->streams || E->K < 1 || E->K > 256 || E->no_predict2)) return predict<MODE>(E, S, E->d_e, red_slot, 1);
    cudaStream_t st = E->stream;
    if (!E->d_pvT && dev_alloc(E, &E->d_pvT, (size_t)E->K * E->D)) return SVBFM_ERR_OOM;
    constexpr bool MC = (MODE == PRED_MC_TRAIN);
    if (MC) { k_transpose_means<<<dim3((E->D + 31) / 32, (E->K + 31) / 32), dim3(32, 8), 0, st>>>(E->d_pv, E->D, E->K, reinterpret_cast<double*>(E->d_pvT)); LAUNCHED(E); }
    else { k_transpose_params<<<dim3((E->D + 31) / 32, (E->K + 31) / 32), dim3(32, 8), 0, st>>>(E->d_pv, E->D, E->K, E->d_pvT); LAUNCHED(E); }
    Predict2Args a{};
    a.rcol = S.rcol; a.rval = S.rval; a.y = S.y; a.n = S.n; a.pw = E->d_pw; a.pvT = E->d_pvT; a.pvTm = reinterpret_cast<const double*>(E->d_pvT);
    a.K = E->K; a.k0 = E->cfg.k0; a.k1 = E->cfg.k1;
    a.sc = E->d_sc; a.e = E->d_e; a.partial = E->d_red_partial;
    if (red_slot < 0) {      // cross shards: the residuals of the second copy's cases; their sums are not used (the first copy holds every case once)
        a.rcol = E->sec.rcol; a.rval = nullptr; a.y = E->sec.y; a.n = E->sec.n; a.e = E->d_e2;
    }
    const unsigned grid = std::max(1u, std::min<unsigned>((a.n + 255) / 256, SV_RGRID / 2));   // partial[] holds grid * 8 <= SV_RGRID * 4 sums
    // a warp per case. Half a warp per case (two cases share the fixed cost of a warp step) is built too, but measured slower
    // (begin 57 ms instead of 40 ms at 200 M cases, K = 50): opt-in for experiments
    const bool half = E->K <= 128 && E->predict2_half;
    const int per = half ? 16 : 32;
    const int ns = E->K <= per ? 1 : (E->K <= 2 * per ? 2 : (E->K <= 4 * per ? 4 : 8));
#define CALL_P2(NS)                                                                                  \
    do {                                                                                             \
        if (half) {                                                                                  \
            if (S.all_ones) k_predict2<MC, true, NS, 16><<<grid, 256, 0, st>>>(a);                   \
            else k_predict2<MC, false, NS, 16><<<grid, 256, 0, st>>>(a);                             \
        } else {                                                                                     \
            if (S.all_ones) k_predict2<MC, true, NS, 32><<<grid, 256, 0, st>>>(a);                   \
            else k_predict2<MC, false, NS, 32><<<grid, 256, 0, st>>>(a);                             \
        }                                                                                            \
    } while (0)
    if (ns == 1) CALL_P2(1); else if (ns == 2) CALL_P2(2); else if (ns == 4) CALL_P2(4); else CALL_P2(8);
#undef CALL_P2
    LAUNCHED(E);
    if (red_slot < 0) return check_launch(E, "predict_second");
    k_reduce_final<<<1, 256, 0, st>>>(E->d_red_partial, grid * 8, 1, RED(E->d_sc, red_slot), 0); LAUNCHED(E);
    if (int rc = allreduce_sum_f64(E, RED(E->d_sc, red_slot), 1)) return rc;
    return check_launch(E, "predict_train");
}
static int predict_second(Engine* E) {
    const bool mc = E->cfg.method == SVBFM_MCMC;
    if (E->K < 1 || E->K > 256 || E->no_predict2) {      // the case-wise kernel on a view of the shard (its sums go to a scratch slot)
        DevSplit V;
        V.n = E->sec.n; V.nnz = 2ull * E->sec.n; V.rcol = E->sec.rcol; V.uniformF = 2; V.all_ones = true; V.y = E->sec.y;
        return mc ? predict<PRED_MC_TRAIN>(E, V, E->d_e2, 7, 1) : predict<PRED_VB_TRAIN>(E, V, E->d_e2, 7, 1);
    }
    if (!E->sec.n) return 0;
    return mc ? predict_train<PRED_MC_TRAIN>(E, -1) : predict_train<PRED_VB_TRAIN>(E, -1);
}

// ---------------------------------------------------------------------------------------------- iterations
struct IterEvents { cudaEvent_t t0, t1, t2; };

static int vb_iteration(Engine* E, uint32_t slot, IterEvents* ev) {
    cudaStream_t st = E->stream;
    DevStats* stp = E->d_stats + slot;
    if (ev) cudaEventRecord(ev->t0, st);
    if (E->cfg.k0) {                                               // update_w0 (vb.h:385-387)
        if (int rc = reduce_e(E)) return rc;
        k_vb_w0<<<1, 1, 0, st>>>(E->d_sc); LAUNCHED(E);
        shift_e(E);
    }
    if (stream_ok(E)) {                                            // update_w + update_v (vb.h:390-440)
        if (int rc = sweep_streams<0>(E)) return rc;
        if (int rc = exchange_blocks(E)) return rc;
    } else {
        if (E->cfg.k1)                                             // update_w, all columns (vb.h:390-406)
            for (const Run& r : E->runs)
                if (int rc = sweep_run<KIND_VB_W>(E, r, -1)) return rc;
        for (int f = 0; f < E->K; f++)                             // update_v (vb.h:409-440)
            for (const Run& r : E->runs)
                if (int rc = sweep_run<KIND_VB_V>(E, r, f)) return rc;
    }
    // hyper-parameters + free energy (vb.h:446-500)
    if (int rc = reduce_e(E)) return rc;
    k_reduce_dT<<<SV_GGRID, 256, 0, st>>>(E->d_dT, E->D, E->d_red_partial); LAUNCHED(E);
    if (E->streams && E->world > 1) {
        // sharded stream schedule: every rank holds its own share of d(sum T) (columns of its block; local C1, C2 elsewhere)
        k_reduce_final<<<1, 256, 0, st>>>(E->d_red_partial, SV_GGRID, 1, RED(E->d_sc, 6), 0); LAUNCHED(E);
        if (int rc = allreduce_sum_f64(E, RED(E->d_sc, 6), 1)) return rc;
        k_add_scalar<<<1, 1, 0, st>>>(&E->d_sc->sum_t, RED(E->d_sc, 6)); LAUNCHED(E);
    } else {
        k_reduce_final<<<1, 256, 0, st>>>(E->d_red_partial, SV_GGRID, 1, &E->d_sc->sum_t, 1); LAUNCHED(E);
    }
    if (int rc = group_sums(E, false)) return rc;
    k_vb_hyper<<<1, 1, 0, st>>>(E->d_sc, E->d_grp_sums, E->d_n_per_group, E->G, E->K, E->d_hyper_w, E->d_hyper_v, stp); LAUNCHED(E);
    if (ev) cudaEventRecord(ev->t1, st);
    // test prediction + evaluation (vbs.h:125-222)
    if (int rc = predict<PRED_VB_TEST>(E, E->te, nullptr, 3, 1)) return rc;
    k_finish_iter<<<1, 1, 0, st>>>(E->d_sc, stp, SVBFM_VB); LAUNCHED(E);
    if (ev) cudaEventRecord(ev->t2, st);
    return check_launch(E, "vb_iteration");
}

static int mcmc_iteration(Engine* E, uint32_t slot, IterEvents* ev) {
    cudaStream_t st = E->stream;
    DevStats* stp = E->d_stats + slot;
    if (ev) cudaEventRecord(ev->t0, st);
    if (int rc = reduce_e(E)) return rc;                           // sum e, sum e^2 for alpha and w0
    if (int rc = group_sums(E, true)) return rc;
    k_mcmc_hyper<<<1, 1, 0, st>>>(E->d_sc, E->d_grp_sums, E->d_n_per_group, E->G, E->K, E->cfg.k0, E->cfg.k1, E->d_hyper_w, E->d_mu_w,
                                  E->d_hyper_v, E->d_mu_v, E->cfg.seed, E->cfg.do_sample, E->cfg.do_multilevel); LAUNCHED(E);
    if (E->cfg.k0) shift_e(E);
    if (stream_ok(E)) {
        if (int rc = sweep_streams<1>(E)) return rc;
        if (int rc = exchange_blocks(E)) return rc;
    } else {
        if (E->cfg.k1)
            for (const Run& r : E->runs)
                if (int rc = sweep_run<KIND_MC_W>(E, r, -1)) return rc;
        for (int f = 0; f < E->K; f++)
            for (const Run& r : E->runs)
                if (int rc = sweep_run<KIND_MC_V>(E, r, f)) return rc;
    }
    if (ev) cudaEventRecord(ev->t1, st);
    // re-prediction of train and test (mcmcs.h:134-174)
    const bool classify = E->cfg.task == 1;
    if (classify) {
        // classification: e = yhat - latent target, so yhat cannot be recovered from e: always re-predict (the NO_REPREDICT flag
        // is ignored), then train accuracy + new latent targets in one more streaming pass (mcmcs.h:188-221)
        if (int rc = predict_train<PRED_MC_TRAIN>(E, 5)) return rc;
        unsigned grid = std::max(1u, std::min<unsigned>(nblk(E->tr.n), SV_RGRID));
        k_mc_class_targets<<<grid, 256, 0, st>>>(E->d_e, E->tr.y, E->tr.perm, E->tr.n, E->cfg.seed, (uint32_t)E->rank, E->cfg.do_sample, E->d_sc,
                                                 E->d_red_partial); LAUNCHED(E);
        k_reduce_final<<<1, 256, 0, st>>>(E->d_red_partial, grid, 1, RED(E->d_sc, 5), 0); LAUNCHED(E);
        if (int rc = allreduce_sum_f64(E, RED(E->d_sc, 5), 1)) return rc;
        sync_e2(E);
    } else if (E->cfg.flags & SVBFM_FLAG_MCMC_NO_REPREDICT) {
        unsigned grid = std::max(1u, std::min<unsigned>(nblk(E->tr.n), SV_RGRID));
        k_train_sse_from_e<<<grid, 256, 0, st>>>(E->d_e, E->tr.y, E->tr.n, E->d_sc, E->d_red_partial); LAUNCHED(E);
        k_reduce_final<<<1, 256, 0, st>>>(E->d_red_partial, grid, 4, E->d_red_partial + SCR_FINAL, 0); LAUNCHED(E);
        SV_CUDA(E, cudaMemcpyAsync(RED(E->d_sc, 5), E->d_red_partial + SCR_FINAL, sizeof(double), cudaMemcpyDeviceToDevice, st));
        if (int rc = allreduce_sum_f64(E, RED(E->d_sc, 5), 1)) return rc;
    } else {
        if (int rc = predict_train<PRED_MC_TRAIN>(E, 5)) return rc;
        sync_e2(E);
    }
    if (classify) { if (int rc = predict<PRED_MC_TEST_CLASS>(E, E->te, nullptr, 3, 2)) return rc; }
    else if (int rc = predict<PRED_MC_TEST>(E, E->te, nullptr, 3, 2)) return rc;
    k_finish_iter<<<1, 1, 0, st>>>(E->d_sc, stp, SVBFM_MCMC, E->cfg.task); LAUNCHED(E);
    if (ev) cudaEventRecord(ev->t2, st);
    return check_launch(E, "mcmc_iteration");
}

static int ensure_stats(Engine* E, uint32_t n) {
    if (E->stats_cap >= n) return 0;
    sv_free(E->d_stats); E->d_stats = nullptr;
    if (dev_alloc(E, &E->d_stats, n)) return SVBFM_ERR_OOM;
    E->stats_cap = n;
    return 0;
}

static int run_iterations(Engine* E, uint32_t n_iter, svbfm_iter_stats* out) {
    if (!E->begun) return fail(E, SVBFM_ERR_ARG, "svbfm_begin must be called first");
    if (E->cfg.method == SVBFM_VB_ONLINE) return fail(E, SVBFM_ERR_ARG, "use svbfm_vb_online_epoch for vb_online");
    if (n_iter == 0) return 0;
    if (int rc = ensure_stats(E, n_iter + 1)) return rc;      // slot n_iter: where a replayed graph leaves its statistics
    SV_CUDA(E, cudaMemsetAsync(E->d_stats, 0, sizeof(DevStats) * (n_iter + 1), E->stream));
    std::vector<IterEvents> ev(n_iter);
    for (auto& e : ev) { cudaEventCreate(&e.t0); cudaEventCreate(&e.t1); cudaEventCreate(&e.t2); }
    int rc = 0;
    uint32_t it = 0;
    auto iterate = [&](uint32_t slot, IterEvents* e) { return (E->cfg.method == SVBFM_VB) ? vb_iteration(E, slot, e) : mcmc_iteration(E, slot, e); };
    // SVBFM_GRAPH=1 (experiment, one GPU): an iteration is a fixed sequence of launches whose arguments do not change (every
    // per-iteration quantity lives in device memory), so iteration 1 is captured into a CUDA graph and replayed for the others:
    // one graph launch instead of ~3 (K + 1) x 2 kernel launches. Matters when the kernels are short (small data); at 200 M
    // ratings a launch is 0.4 ms of work. Anything that goes wrong with the capture falls back to plain launches.
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    if (E->use_graph && E->world == 1 && !E->profile && n_iter >= 3) {
        cudaStream_t st = E->stream;
        rc = iterate(0, &ev[0]);             // plain: whatever is allocated lazily exists afterwards
        it = 1;
        if (!rc) {
            const uint64_t launches_before = E->launches;
            cudaError_t cb = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            int crc = (cb == cudaSuccess) ? iterate(n_iter, nullptr) : -1;
            cudaError_t ce = (cb == cudaSuccess) ? cudaStreamEndCapture(st, &graph) : cb;
            const uint64_t per_iter = E->launches - launches_before;
            E->launches = launches_before;
            if (cb == cudaSuccess && crc == 0 && ce == cudaSuccess && graph && cudaGraphInstantiate(&graph_exec, graph, 0) == cudaSuccess) {
                for (; it < n_iter; it++) {
                    cudaEventRecord(ev[it].t0, st);
                    if (cudaGraphLaunch(graph_exec, st) != cudaSuccess) { rc = fail(E, SVBFM_ERR_CUDA, "cudaGraphLaunch failed"); break; }
                    cudaEventRecord(ev[it].t1, st);      // the whole iteration counts as sweep time in this mode
                    cudaMemcpyAsync(E->d_stats + it, E->d_stats + n_iter, sizeof(DevStats), cudaMemcpyDeviceToDevice, st);
                    cudaEventRecord(ev[it].t2, st);
                    E->launches += per_iter;
                    E->graph_replays++;
                }
            } else {
                cudaGetLastError();          // capture refused or invalidated: nothing of iteration 1 has run, carry on with plain launches
                E->err.clear();
                graph_exec = nullptr;
            }
        }
    }
    for (; it < n_iter && !rc; it++) rc = iterate(it, &ev[it]);
    std::vector<DevStats> hs(n_iter);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(hs.data(), E->d_stats, sizeof(DevStats) * n_iter, cudaMemcpyDeviceToHost, E->stream);
        unsigned long long timeouts = 0;
        if (e == cudaSuccess && E->p2p) e = cudaMemcpyAsync(&timeouts, E->d_xipc + 16 * 8, 8, cudaMemcpyDeviceToHost, E->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(E->stream);
        if (e != cudaSuccess) rc = fail(E, SVBFM_ERR_CUDA, std::string("iteration: ") + cudaGetErrorString(e));
        else if (timeouts) rc = fail(E, SVBFM_ERR_NCCL, "peer exchange: a rank did not raise its flag (timed out); the results of this run are invalid");
    }
    if (!rc && out)
        for (uint32_t it = 0; it < n_iter; it++) {
            svbfm_iter_stats& o = out[it];
            memset(&o, 0, sizeof(o));
            o.test_rmse = hs[it].test_rmse; o.train_stat = hs[it].train_stat; o.free_energy = hs[it].free_energy;
            o.alpha = hs[it].alpha; o.rmse_this = hs[it].rmse_this; o.has_free_energy = hs[it].has_fe != 0.0;
            o.nan_inf_count = (uint32_t)hs[it].nan_inf; o.free_energy_first = hs[it].free_energy;
            cudaEventElapsedTime(&o.sweep_ms, ev[it].t0, ev[it].t1);
            cudaEventElapsedTime(&o.predict_ms, ev[it].t1, ev[it].t2);
        }
    for (auto& e : ev) { cudaEventDestroy(e.t0); cudaEventDestroy(e.t1); cudaEventDestroy(e.t2); }
    if (rc) cudaStreamSynchronize(E->stream);
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (graph) cudaGraphDestroy(graph);
    return rc;
}

// ---------------------------------------------------------------------------------------------- C-ABI
extern "C" {

int svbfm_abi_version(void) { return SVBFM_ABI_VERSION; }

const char* svbfm_last_error(const svbfm_t* h) {
    if (!h) return g_create_error.c_str();
    return reinterpret_cast<const Engine*>(h)->err.c_str();
}

int svbfm_create(svbfm_t** out, const svbfm_config* cfg) {
    if (!out || !cfg) { g_create_error = "svbfm_create: null argument"; return SVBFM_ERR_ARG; }
    *out = nullptr;
    if (cfg->struct_size != sizeof(svbfm_config)) { g_create_error = "svbfm_create: struct_size mismatch (ABI)"; return SVBFM_ERR_ARG; }
    if (cfg->method < SVBFM_VB || cfg->method > SVBFM_MCMC) { g_create_error = "svbfm_create: unknown method"; return SVBFM_ERR_ARG; }
    if (cfg->task != 0 && !(cfg->task == 1 && cfg->method == SVBFM_MCMC)) {
        g_create_error = "svbfm_create: task 0 (regression), or task 1 (binary classification, targets -1 / +1) with the mcmc method";
        return SVBFM_ERR_ARG;
    }
    if (cfg->num_factor < 0 || cfg->num_attribute == 0) { g_create_error = "svbfm_create: bad dimensions"; return SVBFM_ERR_ARG; }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        g_create_error = std::string("svbfm_create: no CUDA device (") + cudaGetErrorString(ce) + "); there is no CPU fallback";
        return SVBFM_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "svbfm_create: device ordinal out of range"; return SVBFM_ERR_ARG; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, cfg->device);
    if (prop.major < 10) {
        g_create_error = std::string("svbfm_create: device '") + prop.name + "' is not sm_100 (kernels are built for sm_100a only)";
        return SVBFM_ERR_CUDA;
    }
    ce = cudaSetDevice(cfg->device);
    if (ce != cudaSuccess) { g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(ce); return SVBFM_ERR_CUDA; }
    {
        // The sweeps of every field but the first gather 8-16 B per case at random; with the default L2 fetch
        // granularity each miss moves 128 B from HBM (ncu: 8 sectors per entry, profiles/r01_*). Ask for 32 B.
        size_t gran = 32;
        if (const char* s = getenv("SVBFM_L2_FETCH")) gran = (size_t)atoi(s);
        if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    }
    Engine* E = new Engine();
    ++g_handles;
    E->cfg = *cfg; E->dev = cfg->device; E->D = cfg->num_attribute; E->K = cfg->num_factor;
    E->tile_entries = cfg->tile_entries ? cfg->tile_entries : 1024;
    if (const char* tm = getenv("SVBFM_STREAM_TMA")) E->stream_tma = atoi(tm) != 0;      // default on; 0: plain loads (A/B runs, tests)
    if (const char* rr = getenv("SVBFM_REC_RANK")) E->want_rec_rank = atoi(rr) != 0;    // default on; 0: records in column order
    if (const char* gr = getenv("SVBFM_GRAPH")) E->use_graph = atoi(gr) != 0;
    if (const char* na = getenv("SVBFM_REC_NA")) E->rec_na_mask = atoi(na);              // bit s: side s gathers its records past L1 (default 2)
    E->no_predict2 = getenv("SVBFM_NO_PREDICT2") != nullptr;          // knobs of the per-iteration paths are read once, here
    E->predict2_half = getenv("SVBFM_PREDICT2_HALFWARP") != nullptr;
    E->vbo_full_passes = getenv("SVBFM_VBO_FULL_PASSES") != nullptr;
    if (const char* vc = getenv("SVBFM_VBO_COMPACT")) E->vbo_compact = atoi(vc) != 0;    // default on; 0: global column ids in the batch passes / finalizes
    if (const char* vr = getenv("SVBFM_VBO_ROWS")) E->vbo_rows = atoi(vr) != 0;          // default on; 0: packed batches through k_stream
    if (const char* vp = getenv("SVBFM_VBO_PACK")) E->vbo_pack = atoi(vp) != 0;          // default on; 0: the batch passes read through the index lists
    if (const char* rh = getenv("SVBFM_REC_HOT")) E->rec_hot = atoi(rh);                // first-field pass: records of the rank layout kept in L1 (0: all)
    if (const char* te = getenv("SVBFM_TILE_ENTRIES")) if (atoi(te) >= 32) E->tile_entries = (uint32_t)atoi(te);   // tuning knob
    // implicit tiles of the stream schedule: 4096 entries unless the caller (or the knob) says otherwise
    uint32_t ts = (cfg->tile_entries || getenv("SVBFM_TILE_ENTRIES")) ? E->tile_entries : 4096u;
    E->ts_auto = !(cfg->tile_entries || getenv("SVBFM_TILE_ENTRIES"));     // ingest may shrink the tiles of a small train split
    E->ts_shift = 5;
    while (E->ts_shift < 20 && (2u << E->ts_shift) <= ts) E->ts_shift++;
    ce = cudaStreamCreateWithFlags(&E->own_stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(ce); delete E; --g_handles; return SVBFM_ERR_CUDA; }
    E->stream = E->own_stream;
    if (!getenv("SVBFM_NO_COPY_STREAM") && cudaStreamCreateWithFlags(&E->copy_stream, cudaStreamNonBlocking) == cudaSuccess) {
        if (cudaEventCreateWithFlags(&E->copy_event, cudaEventDisableTiming) != cudaSuccess) { cudaStreamDestroy(E->copy_stream); E->copy_stream = nullptr; }
    } else E->copy_stream = nullptr;
    E->G = 1;
    E->h_group.assign(E->D, 0);
    E->h_n_per_group.assign(1, E->D);
    *out = reinterpret_cast<svbfm_t*>(E);
    // device state
    size_t D = E->D, K = (size_t)E->K;
    int rc = 0;
    rc |= dev_alloc(E, &E->d_group, D);
    rc |= dev_alloc(E, &E->d_pw, D);
    rc |= dev_alloc(E, &E->d_pv, K * D);
    rc |= dev_alloc(E, &E->d_sc, 1);
    rc |= dev_alloc(E, &E->d_colsum, D * 4);
    rc |= dev_alloc(E, &E->d_delta, D);
    rc |= dev_alloc(E, &E->d_cpack, D); E->cpack_cap = D;
    rc |= dev_alloc(E, &E->d_opack, D);
    rc |= dev_alloc(E, &E->d_ab, D);
    rc |= dev_alloc(E, &E->d_dT, D);
    rc |= dev_alloc(E, &E->d_red_partial, SCR_GROUP + (K + 1) * 64 /*max groups*/ * 2 * SV_GGRID);
    if (rc) { g_create_error = E->err; svbfm_destroy(*out); *out = nullptr; return SVBFM_ERR_OOM; }
    cudaMemsetAsync(E->d_group, 0, D * 4, E->stream);
    cudaMemsetAsync(E->d_dT, 0, D * 8, E->stream);
    cudaMemsetAsync(E->d_delta, 0, D * 8, E->stream);
    cudaMemsetAsync(E->d_colsum, 0, D * 32, E->stream);
    cudaMemsetAsync(E->d_sc, 0, sizeof(Scalars), E->stream);
    cudaStreamSynchronize(E->stream);
    // default groups
    std::vector<uint32_t> g0(E->D, 0);
    if (svbfm_set_groups(*out, g0.data(), 1)) { g_create_error = E->err; svbfm_destroy(*out); *out = nullptr; return SVBFM_ERR_CUDA; }
    return SVBFM_OK;
}

void svbfm_destroy(svbfm_t* h) {
    if (!h) return;
    Engine* E = reinterpret_cast<Engine*>(h);
    cudaSetDevice(E->dev);
    cudaStreamSynchronize(E->stream);
    if (E->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(E->nccl_comm);
    free_split(E, E->tr); free_split(E, E->te); free_second(E);
    void* ptrs[] = {E->d_group, E->d_n_per_group, E->d_tile_col, E->d_tile_begin, E->d_tile_len, E->d_exec_order, E->d_col_tile0, E->d_heavy_cols, E->d_pw, E->d_pv,
                    E->d_hyper_w, E->d_hyper_v, E->d_mu_w, E->d_mu_v, E->d_sc, E->d_nat_w, E->d_nat_v, E->d_t_w, E->d_t_v, E->d_col_count,
                    E->d_e, E->d_partial, E->d_colsum, E->d_delta, E->d_dT, E->d_red_partial, E->d_grp_sums, E->d_pred_test,
                    E->d_pred_sum, E->d_stats, E->d_cpack, E->d_opack, E->d_ab, E->d_xchg, E->d_pvT, E->d_vbo_idx[0], E->d_vbo_idx[1], E->d_vbo_colptr[0], E->d_vbo_colptr[1], E->d_vbo_gcnt[0], E->d_vbo_gcnt[1], E->d_vbo_clist[0], E->d_vbo_clist[1], E->d_vbo_tile_col0, E->d_vbo_partial, E->d_vbo_eb[0], E->d_vbo_eb[1], E->d_vbo_ocb[0], E->d_vbo_ocb[1], E->d_vbo_ownb[0], E->d_vbo_ownb[1], E->d_vbo_cpos[0], E->d_vbo_cpos[1], E->d_vbo_cc, E->d_vbo_ccptr, E->d_vbo_opack_c, E->d_vbo_colsum_c, E->d_vbo_dT_c, E->d_vbo_nextp_c, E->d_vbo_prevm_c, E->d_vbo_xb[0][0], E->d_vbo_xb[0][1], E->d_vbo_xb[1][0], E->d_vbo_xb[1][1], E->d_e2, E->d_stile_col0, E->d_span_heavy, E->d_rec_slot, E->d_rbatch, E->d_cbatch, E->d_cnt_col, E->d_batch_cnt, E->d_batch_n};
    for (void* p : ptrs) sv_free(p);
    cudaStreamSynchronize(E->stream);
    if (E->copy_stream) { cudaStreamSynchronize(E->copy_stream); cudaStreamDestroy(E->copy_stream); }
    sv_owner_release(E);      // everything this handle cached is idle: the next handle may take it
    if (E->copy_event) cudaEventDestroy(E->copy_event);
    if (E->own_stream) cudaStreamDestroy(E->own_stream);
    delete E;
    if (--g_handles == 0) sv_cache_release();      // the last handle gives the cached device blocks back to the driver
}

int svbfm_set_stream(svbfm_t* h, void* cuda_stream) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E) return SVBFM_ERR_ARG;
    cudaStreamSynchronize(E->stream);
    sv_owner_release(E);      // blocks this handle cached were last used on the stream it leaves: idle now
    E->stream = cuda_stream ? (cudaStream_t)cuda_stream : E->own_stream;
    return SVBFM_OK;
}

int svbfm_comm_get_unique_id(uint8_t id[SVBFM_COMM_ID_BYTES]) {
    if (!nccl_load()) { g_create_error = "NCCL (libnccl.so.2) could not be loaded"; return SVBFM_ERR_NCCL; }
    int r = g_nccl.GetUniqueId(id);
    if (r != 0) { g_create_error = "ncclGetUniqueId failed"; return SVBFM_ERR_NCCL; }
    return SVBFM_OK;
}

int svbfm_comm_init(svbfm_t* h, const uint8_t id[SVBFM_COMM_ID_BYTES], int32_t rank, int32_t world_size) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E || !id || world_size < 1 || rank < 0 || rank >= world_size) return fail(E, SVBFM_ERR_ARG, "svbfm_comm_init: bad arguments");
    if (E->tr.n || E->te.n) return fail(E, SVBFM_ERR_ARG, "svbfm_comm_init must precede svbfm_set_csc");
    if (world_size == 1) { E->rank = 0; E->world = 1; return SVBFM_OK; }
    if (world_size > 16) return fail(E, SVBFM_ERR_ARG, "svbfm_comm_init: at most 16 ranks (one node)");
    if (!nccl_load()) return fail(E, SVBFM_ERR_NCCL, "NCCL (libnccl.so.2) could not be loaded");
    SV_CUDA(E, cudaSetDevice(E->dev));
    Id128 uid;
    memcpy(uid.b, id, SVBFM_COMM_ID_BYTES);
    int r = g_nccl.CommInitRank(&E->nccl_comm, world_size, uid, rank);
    if (r != 0) return fail(E, SVBFM_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    E->rank = rank; E->world = world_size;
    return SVBFM_OK;
}

int svbfm_set_groups(svbfm_t* h, const uint32_t* attr_group, uint32_t num_groups) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E || !attr_group || num_groups == 0) return fail(E, SVBFM_ERR_ARG, "svbfm_set_groups: bad arguments");
    if (num_groups > 64) return fail(E, SVBFM_ERR_ARG, "svbfm_set_groups: more than 64 attribute groups are not supported");
    if (E->begun) return fail(E, SVBFM_ERR_ARG, "svbfm_set_groups after svbfm_begin");
    SV_CUDA(E, cudaSetDevice(E->dev));
    std::vector<double> npg(num_groups, 0.0);
    for (uint32_t j = 0; j < E->D; j++) {
        if (attr_group[j] >= num_groups) return fail(E, SVBFM_ERR_ARG, "svbfm_set_groups: group id out of range");
        npg[attr_group[j]] += 1.0;                       // DataMetaInfo::num_attr_per_group counts ALL D attributes (Data.h:58-60)
    }
    E->G = num_groups;
    E->h_group.assign(attr_group, attr_group + E->D);
    E->h_n_per_group.resize(num_groups);
    for (uint32_t g = 0; g < num_groups; g++) E->h_n_per_group[g] = (uint32_t)npg[g];
    SV_CUDA(E, copy_sync(E, E->d_group, attr_group, (size_t)E->D * 4, cudaMemcpyHostToDevice));
    void* old[] = {E->d_n_per_group, E->d_hyper_w, E->d_hyper_v, E->d_mu_w, E->d_mu_v, E->d_grp_sums};
    for (void* p : old) sv_free(p);
    E->d_n_per_group = E->d_hyper_w = E->d_hyper_v = E->d_mu_w = E->d_mu_v = E->d_grp_sums = nullptr;
    size_t G = num_groups, K = (size_t)E->K;
    int rc = 0;
    rc |= dev_alloc(E, &E->d_n_per_group, G);
    rc |= dev_alloc(E, &E->d_hyper_w, G);
    rc |= dev_alloc(E, &E->d_hyper_v, G * std::max<size_t>(K, 1));
    rc |= dev_alloc(E, &E->d_mu_w, G);
    rc |= dev_alloc(E, &E->d_mu_v, G * std::max<size_t>(K, 1));
    rc |= dev_alloc(E, &E->d_grp_sums, (K + 1) * G * 2);
    if (rc) return SVBFM_ERR_OOM;
    SV_CUDA(E, copy_sync(E, E->d_n_per_group, npg.data(), G * 8, cudaMemcpyHostToDevice));
    // initial hyper-parameters: vb sigma_w = sigma_v = 1 (vb.h:707-708); mcmc lambda = reg, mu = 0 (mcmc.h:1109-1117, libfm.cpp:372-405)
    bool mc = E->cfg.method == SVBFM_MCMC;
    std::vector<double> hw(G, mc ? E->cfg.regw : 1.0), hv(G * std::max<size_t>(K, 1), mc ? E->cfg.regv : 1.0), z(G * std::max<size_t>(K, 1), 0.0);
    SV_CUDA(E, copy_sync(E, E->d_hyper_w, hw.data(), G * 8, cudaMemcpyHostToDevice));
    SV_CUDA(E, copy_sync(E, E->d_hyper_v, hv.data(), hv.size() * 8, cudaMemcpyHostToDevice));
    SV_CUDA(E, copy_sync(E, E->d_mu_w, z.data(), G * 8, cudaMemcpyHostToDevice));
    SV_CUDA(E, copy_sync(E, E->d_mu_v, z.data(), z.size() * 8, cudaMemcpyHostToDevice));
    return SVBFM_OK;
}

int svbfm_set_csc(svbfm_t* h, int32_t split, uint32_t num_cases, uint32_t num_cols, const uint64_t* colptr, const uint32_t* case_id,
                  const float* x, const float* target) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E || !colptr || (split != SVBFM_TRAIN && split != SVBFM_TEST && split != SVBFM_TRAIN_SECOND)) return fail(E, SVBFM_ERR_ARG, "svbfm_set_csc: bad arguments");
    if (E->begun) return fail(E, SVBFM_ERR_ARG, "svbfm_set_csc after svbfm_begin");
    SV_CUDA(E, cudaSetDevice(E->dev));
    if (split == SVBFM_TRAIN_SECOND) return ingest_second(E, num_cases, num_cols, colptr, case_id, x, target);      // validates its arguments collectively
    if (colptr[num_cols] > 0 && !case_id) return fail(E, SVBFM_ERR_ARG, "svbfm_set_csc: null entry arrays");
    if (num_cases > 0 && !target) return fail(E, SVBFM_ERR_ARG, "svbfm_set_csc: null target");
    bool is_train = split == SVBFM_TRAIN;
    DevSplit& S = is_train ? E->tr : E->te;
    int rc = ingest_split(E, S, is_train, num_cases, num_cols, colptr, case_id, x, target);
    if (rc) return rc;
    // global case counts
    double cnt = (double)num_cases, *d_cnt = RED(E->d_sc, 7);
    SV_CUDA(E, cudaMemcpyAsync(d_cnt, &cnt, 8, cudaMemcpyHostToDevice, E->stream));
    if (int r2 = allreduce_sum_f64(E, d_cnt, 1)) return r2;
    SV_CUDA(E, cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, E->stream));
    SV_CUDA(E, cudaStreamSynchronize(E->stream));
    if (is_train) {
        E->n_total = (uint64_t)cnt;
        sv_free(E->d_e); sv_free(E->d_partial); sv_free(E->d_e2);
        E->d_e = nullptr; E->d_partial = nullptr; E->d_e2 = nullptr;
        if (dev_alloc(E, &E->d_e, num_cases)) return SVBFM_ERR_OOM;
        if (E->streams) {
            if (dev_alloc(E, &E->d_e2, num_cases)) return SVBFM_ERR_OOM;
            if (dev_alloc(E, &E->d_partial, ((size_t)E->s_ntiles[0] + E->s_ntiles[1]) * 8)) return SVBFM_ERR_OOM;
        } else if (dev_alloc(E, &E->d_partial, (size_t)E->n_tiles * 4)) return SVBFM_ERR_OOM;
    } else {
        E->nt_total = (uint64_t)cnt;
        sv_free(E->d_pred_test); sv_free(E->d_pred_sum); E->d_pred_test = nullptr; E->d_pred_sum = nullptr;
        if (dev_alloc(E, &E->d_pred_test, num_cases)) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->d_pred_sum, num_cases)) return SVBFM_ERR_OOM;
        SV_CUDA(E, cudaMemsetAsync(E->d_pred_sum, 0, std::max<size_t>(num_cases, 1) * 8, E->stream));
        SV_CUDA(E, cudaMemsetAsync(E->d_pred_test, 0, std::max<size_t>(num_cases, 1) * 8, E->stream));
        SV_CUDA(E, cudaStreamSynchronize(E->stream));
    }
    return SVBFM_OK;
}

int svbfm_set_csr(svbfm_t* h, int32_t split, uint32_t num_cases, uint32_t num_cols, const uint64_t* rowptr, const uint32_t* feature_id, const float* x,
                  const float* target) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E || !rowptr || (split != SVBFM_TRAIN && split != SVBFM_TEST && split != SVBFM_TRAIN_SECOND)) return fail(E, SVBFM_ERR_ARG, "svbfm_set_csr: bad arguments");
    if (E->begun) return fail(E, SVBFM_ERR_ARG, "svbfm_set_csr after svbfm_begin");
    const uint64_t nnz = rowptr[num_cases];
    // local argument errors would leave the other ranks waiting in set_csc's collectives: hand an impossible shard on instead
    std::string why;
    if (rowptr[0] != 0) why = "svbfm_set_csr: rowptr[0] != 0";
    else if (nnz >= (1ull << 32)) why = "svbfm_set_csr: more than 2^32-1 entries per rank are not supported";
    else if (nnz > 0 && !feature_id) why = "svbfm_set_csr: null entry arrays";
    else if (num_cases > 0 && !target) why = "svbfm_set_csr: null target";
    else for (uint32_t i = 0; i < num_cases && why.empty(); i++) if (rowptr[i + 1] < rowptr[i]) why = "svbfm_set_csr: rowptr not monotone";
    if (!why.empty() && E->world <= 1) return fail(E, SVBFM_ERR_ARG, why);
    SV_CUDA(E, cudaSetDevice(E->dev));
    cudaStream_t st = E->stream;
    uint64_t *d_rowptr = nullptr, *d_colptr = nullptr;
    uint32_t *d_col = nullptr, *d_case = nullptr;
    float *d_x = nullptr, *d_xt = nullptr, *d_y = nullptr;
    struct Temps { std::vector<void**> v; ~Temps() { for (void** p : v) { sv_free(*p); *p = nullptr; } } } temps;
    for (void** p : {(void**)&d_rowptr, (void**)&d_colptr, (void**)&d_col, (void**)&d_case, (void**)&d_x, (void**)&d_xt, (void**)&d_y}) temps.v.push_back(p);
    std::vector<uint64_t> h_colptr((size_t)num_cols + 1, 0);
    int rc = 0;
    if (why.empty()) {
        SV_CUDA(E, sv_malloc((void**)&d_rowptr, ((size_t)num_cases + 1) * 8));
        SV_CUDA(E, sv_malloc((void**)&d_col, std::max<uint64_t>(nnz, 1) * 4));
        if (x) SV_CUDA(E, sv_malloc((void**)&d_x, std::max<uint64_t>(nnz, 1) * 4));
        SV_CUDA(E, sv_malloc((void**)&d_y, std::max<size_t>(num_cases, 1) * 4));
        SV_CUDA(E, cudaMemcpyAsync(d_rowptr, rowptr, ((size_t)num_cases + 1) * 8, cudaMemcpyHostToDevice, st));
        SV_CUDA(E, cudaMemcpyAsync(d_col, feature_id, nnz * 4, cudaMemcpyHostToDevice, st));
        if (x) SV_CUDA(E, cudaMemcpyAsync(d_x, x, nnz * 4, cudaMemcpyHostToDevice, st));
        SV_CUDA(E, cudaMemcpyAsync(d_y, target, (size_t)num_cases * 4, cudaMemcpyHostToDevice, st));
        rc = transpose_on_device(E, st, num_cases, num_cols, nnz, d_rowptr, d_col, d_x, &d_colptr, &d_case, &d_xt);
        if (rc) why = E->err;
        else {
            SV_CUDA(E, cudaMemcpyAsync(h_colptr.data(), d_colptr, ((size_t)num_cols + 1) * 8, cudaMemcpyDeviceToHost, st));
            SV_CUDA(E, cudaStreamSynchronize(st));
        }
        sv_free(d_rowptr); d_rowptr = nullptr; sv_free(d_col); d_col = nullptr; sv_free(d_x); d_x = nullptr;
    }
    if (!why.empty()) {
        if (E->world <= 1) return fail(E, SVBFM_ERR_ARG, why);
        h_colptr.assign((size_t)num_cols + 1, 0); h_colptr[0] = 1;       // set_csc rejects it (colptr[0] != 0) together with the other ranks
        int r2 = svbfm_set_csc(h, split, 0, num_cols, h_colptr.data(), nullptr, nullptr, nullptr);
        (void)r2;
        return fail(E, SVBFM_ERR_ARG, why);
    }
    // the CSC arrays are on the device already: svbfm_set_csc's copies find that out themselves (cudaMemcpyDefault)
    return svbfm_set_csc(h, split, num_cases, num_cols, h_colptr.data(), d_case, d_xt, d_y);
}

int svbfm_transpose_csr(int32_t device, uint32_t num_cases, uint32_t num_cols, const uint64_t* rowptr, const uint32_t* feature_id, const float* x,
                        uint64_t* out_colptr, uint32_t* out_case_id, float* out_x) {
    if (!rowptr || !out_colptr) { g_create_error = "svbfm_transpose_csr: null argument"; return SVBFM_ERR_ARG; }
    const uint64_t nnz = rowptr[num_cases];
    if (nnz >= (1ull << 32) || (nnz && (!feature_id || !out_case_id || (x && !out_x)))) { g_create_error = "svbfm_transpose_csr: bad arguments"; return SVBFM_ERR_ARG; }
    svbfm_config c;
    memset(&c, 0, sizeof(c));
    c.struct_size = sizeof(c); c.method = SVBFM_VB; c.num_attribute = 1; c.num_factor = 0; c.device = device;
    svbfm_t* h = nullptr;
    int rc = svbfm_create(&h, &c);        // device checks, stream, block cache
    if (rc) return rc;
    Engine* E = reinterpret_cast<Engine*>(h);
    cudaStream_t st = E->stream;
    uint64_t *d_rowptr = nullptr, *d_colptr = nullptr;
    uint32_t *d_col = nullptr, *d_case = nullptr;
    float *d_x = nullptr, *d_xt = nullptr;
    auto body = [&]() -> int {
        SV_CUDA(E, sv_malloc((void**)&d_rowptr, ((size_t)num_cases + 1) * 8));
        SV_CUDA(E, sv_malloc((void**)&d_col, std::max<uint64_t>(nnz, 1) * 4));
        if (x) SV_CUDA(E, sv_malloc((void**)&d_x, std::max<uint64_t>(nnz, 1) * 4));
        SV_CUDA(E, cudaMemcpyAsync(d_rowptr, rowptr, ((size_t)num_cases + 1) * 8, cudaMemcpyHostToDevice, st));
        SV_CUDA(E, cudaMemcpyAsync(d_col, feature_id, nnz * 4, cudaMemcpyHostToDevice, st));
        if (x) SV_CUDA(E, cudaMemcpyAsync(d_x, x, nnz * 4, cudaMemcpyHostToDevice, st));
        if (int r = transpose_on_device(E, st, num_cases, num_cols, nnz, d_rowptr, d_col, d_x, &d_colptr, &d_case, &d_xt)) return r;
        SV_CUDA(E, cudaMemcpyAsync(out_colptr, d_colptr, ((size_t)num_cols + 1) * 8, cudaMemcpyDeviceToHost, st));
        if (nnz) {
            SV_CUDA(E, cudaMemcpyAsync(out_case_id, d_case, nnz * 4, cudaMemcpyDeviceToHost, st));
            if (x) SV_CUDA(E, cudaMemcpyAsync(out_x, d_xt, nnz * 4, cudaMemcpyDeviceToHost, st));
        }
        SV_CUDA(E, cudaStreamSynchronize(st));
        return 0;
    };
    rc = body();
    if (rc) g_create_error = E->err;
    for (void* p : {(void*)d_rowptr, (void*)d_colptr, (void*)d_col, (void*)d_case, (void*)d_x, (void*)d_xt}) sv_free(p);
    svbfm_destroy(h);
    return rc;
}

int svbfm_set_state(svbfm_t* h, double w0_mean, double w0_var, const double* w_mean, const double* w_var, const double* v_mean,
                    const double* v_var) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E || !w_mean || (E->K > 0 && !v_mean)) return fail(E, SVBFM_ERR_ARG, "svbfm_set_state: null argument");
    SV_CUDA(E, cudaSetDevice(E->dev));
    size_t D = E->D, KD = (size_t)E->K * D;
    double *tm = nullptr, *tv = nullptr;
    SV_CUDA(E, sv_malloc((void**)&tm, std::max<size_t>(KD, D) * 8));
    SV_CUDA(E, sv_malloc((void**)&tv, std::max<size_t>(KD, D) * 8));
    cudaStream_t st = E->stream;
    SV_CUDA(E, cudaMemcpyAsync(tm, w_mean, D * 8, cudaMemcpyHostToDevice, st));
    if (w_var) SV_CUDA(E, cudaMemcpyAsync(tv, w_var, D * 8, cudaMemcpyHostToDevice, st));
    k_pack<<<nblk(D), 256, 0, st>>>(tm, w_var ? tv : nullptr, D, E->d_pw);
    SV_CUDA(E, cudaStreamSynchronize(st));
    if (KD) {
        SV_CUDA(E, cudaMemcpyAsync(tm, v_mean, KD * 8, cudaMemcpyHostToDevice, st));
        if (v_var) SV_CUDA(E, cudaMemcpyAsync(tv, v_var, KD * 8, cudaMemcpyHostToDevice, st));
        k_pack<<<nblk(KD), 256, 0, st>>>(tm, v_var ? tv : nullptr, KD, E->d_pv);
        SV_CUDA(E, cudaStreamSynchronize(st));
    }
    sv_free(tm); sv_free(tv);
    Scalars sc;
    SV_CUDA(E, copy_sync(E, &sc, E->d_sc, sizeof(sc), cudaMemcpyDeviceToHost));
    sc.w0_mean = w0_mean; sc.w0_var = w0_var;
    if (!E->have_state) {
        sc.alpha = 1.0;                                              // vb.h:693 / mcmc.h:1105
        sc.sigma_0 = (E->cfg.method == SVBFM_MCMC) ? E->cfg.reg0 : 1.0;   // vb.h:694 / libfm.cpp:373
        sc.min_target = E->cfg.min_target; sc.max_target = E->cfg.max_target;
        sc.iter = 0; sc.nan_inf = 0; sc.alpha_ok = 1;
        sc.nat_mu_0 = 0.0; sc.nat_sg_0 = (w0_var != 0.0) ? 1.0 / w0_var : 0.0; sc.rho_0 = 1.0; sc.t_w0 = 0;   // vbo.h:683-701
    }
    SV_CUDA(E, copy_sync(E, E->d_sc, &sc, sizeof(sc), cudaMemcpyHostToDevice));
    E->have_state = true;
    return check_launch(E, "set_state");
}

int svbfm_get_state(svbfm_t* h, double* w0_mean, double* w0_var, double* w_mean, double* w_var, double* v_mean, double* v_var) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E) return SVBFM_ERR_ARG;
    SV_CUDA(E, cudaSetDevice(E->dev));
    size_t D = E->D, KD = (size_t)E->K * D;
    cudaStream_t st = E->stream;
    double *tm = nullptr, *tv = nullptr;
    SV_CUDA(E, sv_malloc((void**)&tm, std::max<size_t>(KD, D) * 8));
    SV_CUDA(E, sv_malloc((void**)&tv, std::max<size_t>(KD, D) * 8));
    k_unpack<<<nblk(D), 256, 0, st>>>(E->d_pw, D, tm, tv);
    if (w_mean) SV_CUDA(E, cudaMemcpyAsync(w_mean, tm, D * 8, cudaMemcpyDeviceToHost, st));
    if (w_var) SV_CUDA(E, cudaMemcpyAsync(w_var, tv, D * 8, cudaMemcpyDeviceToHost, st));
    SV_CUDA(E, cudaStreamSynchronize(st));
    if (KD) {
        k_unpack<<<nblk(KD), 256, 0, st>>>(E->d_pv, KD, tm, tv);
        if (v_mean) SV_CUDA(E, cudaMemcpyAsync(v_mean, tm, KD * 8, cudaMemcpyDeviceToHost, st));
        if (v_var) SV_CUDA(E, cudaMemcpyAsync(v_var, tv, KD * 8, cudaMemcpyDeviceToHost, st));
        SV_CUDA(E, cudaStreamSynchronize(st));
    }
    sv_free(tm); sv_free(tv);
    Scalars sc;
    SV_CUDA(E, copy_sync(E, &sc, E->d_sc, sizeof(sc), cudaMemcpyDeviceToHost));
    if (w0_mean) *w0_mean = sc.w0_mean;
    if (w0_var) *w0_var = sc.w0_var;
    return check_launch(E, "get_state");
}

int svbfm_get_hyper(svbfm_t* h, double* alpha, double* sigma_0, double* sigma_w, double* sigma_v) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E) return SVBFM_ERR_ARG;
    SV_CUDA(E, cudaSetDevice(E->dev));
    SV_CUDA(E, cudaStreamSynchronize(E->stream));
    Scalars sc;
    SV_CUDA(E, copy_sync(E, &sc, E->d_sc, sizeof(sc), cudaMemcpyDeviceToHost));
    if (alpha) *alpha = sc.alpha;
    if (sigma_0) *sigma_0 = sc.sigma_0;
    if (sigma_w) SV_CUDA(E, copy_sync(E, sigma_w, E->d_hyper_w, (size_t)E->G * 8, cudaMemcpyDeviceToHost));
    if (sigma_v && E->K) SV_CUDA(E, copy_sync(E, sigma_v, E->d_hyper_v, (size_t)E->G * E->K * 8, cudaMemcpyDeviceToHost));
    return SVBFM_OK;
}

int svbfm_set_hyper(svbfm_t* h, double alpha, double sigma_0, const double* sigma_w, const double* sigma_v) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E) return SVBFM_ERR_ARG;
    SV_CUDA(E, cudaSetDevice(E->dev));
    SV_CUDA(E, cudaStreamSynchronize(E->stream));
    Scalars sc;
    SV_CUDA(E, copy_sync(E, &sc, E->d_sc, sizeof(sc), cudaMemcpyDeviceToHost));
    sc.alpha = alpha; sc.sigma_0 = sigma_0;
    SV_CUDA(E, copy_sync(E, E->d_sc, &sc, sizeof(sc), cudaMemcpyHostToDevice));
    if (sigma_w) SV_CUDA(E, copy_sync(E, E->d_hyper_w, sigma_w, (size_t)E->G * 8, cudaMemcpyHostToDevice));
    if (sigma_v && E->K) SV_CUDA(E, copy_sync(E, E->d_hyper_v, sigma_v, (size_t)E->G * E->K * 8, cudaMemcpyHostToDevice));
    return SVBFM_OK;
}

int svbfm_begin(svbfm_t* h) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E) return SVBFM_ERR_ARG;
    if (!E->have_state) return fail(E, SVBFM_ERR_ARG, "svbfm_begin: svbfm_set_state must be called first");
    if (!E->d_e || !E->d_pred_test) return fail(E, SVBFM_ERR_ARG, "svbfm_begin: train and test data must be set first");
    SV_CUDA(E, cudaSetDevice(E->dev));
    cudaStream_t st = E->stream;
    Scalars sc;
    SV_CUDA(E, copy_sync(E, &sc, E->d_sc, sizeof(sc), cudaMemcpyDeviceToHost));
    sc.n_total = (double)E->n_total; sc.nt_total = (double)E->nt_total; sc.sum_t = 0.0; sc.w0_delta = 0.0;
    SV_CUDA(E, copy_sync(E, E->d_sc, &sc, sizeof(sc), cudaMemcpyHostToDevice));
    if (int rc = ensure_stats(E, 16)) return rc;
    if (E->cfg.method == SVBFM_VB) {
        // initial y-hat and T over train (vbs.h:37-44): e_i = y_i - yhat_i ; sum_t = sum_i T_i
        if (int rc = predict_train<PRED_VB_TRAIN>(E, 6)) return rc;
        SV_CUDA(E, cudaMemcpyAsync(&E->d_sc->sum_t, RED(E->d_sc, 6), 8, cudaMemcpyDeviceToDevice, st));
    } else if (E->cfg.method == SVBFM_MCMC) {
        if (int rc = predict_train<PRED_MC_TRAIN>(E, 6)) return rc;                 // e = yhat - y (mcmcs.h:75-80)
    }
    if (E->cfg.method != SVBFM_VB_ONLINE) sync_e2(E);
    else {   // vb_online: global count of every feature in the training data (vbo.h:704-726) and the natural parameters
        if (!E->d_col_count) {
            if (dev_alloc(E, &E->d_col_count, E->D)) return SVBFM_ERR_OOM;
            if (dev_alloc(E, &E->d_nat_w, E->D)) return SVBFM_ERR_OOM;
            if (dev_alloc(E, &E->d_nat_v, (size_t)E->K * E->D)) return SVBFM_ERR_OOM;
            if (dev_alloc(E, &E->d_t_w, E->D)) return SVBFM_ERR_OOM;
            if (dev_alloc(E, &E->d_t_v, E->D)) return SVBFM_ERR_OOM;
        }
        k_col_counts<<<nblk(E->D), 256, 0, st>>>(E->tr.colptr, E->tr.ncols_ext, E->D, E->d_col_count); LAUNCHED(E);
        if (int rc = allreduce_sum_f64(E, E->d_col_count, E->D)) return rc;
        k_nat_from_params<<<nblk(E->D), 256, 0, st>>>(E->d_pw, E->D, E->d_nat_w); LAUNCHED(E);
        if (E->K) { k_nat_from_params<<<nblk((size_t)E->K * E->D), 256, 0, st>>>(E->d_pv, (size_t)E->K * E->D, E->d_nat_v); LAUNCHED(E); }
        SV_CUDA(E, cudaMemsetAsync(E->d_t_w, 0, (size_t)E->D * 4, st));
        SV_CUDA(E, cudaMemsetAsync(E->d_t_v, 0, (size_t)E->D * 4, st));
    }
    SV_CUDA(E, cudaStreamSynchronize(st));
    E->begun = true;
    return check_launch(E, "begin");
}

int svbfm_reset(svbfm_t* h) {
    // back to the state right after svbfm_create (+ comm_init, set_groups): new data and a new initial state may follow.
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E) return SVBFM_ERR_ARG;
    SV_CUDA(E, cudaSetDevice(E->dev));
    SV_CUDA(E, cudaStreamSynchronize(E->stream));
    free_split(E, E->tr); free_split(E, E->te); free_second(E);
    E->runs.clear();
    E->begun = false; E->have_state = false; E->rows_reordered = false; E->run0_sequential = false; E->streams = false; E->excl0 = false; E->vbo_streams = false; E->bv.on = false;
    SV_CUDA(E, cudaMemsetAsync(E->d_dT, 0, (size_t)E->D * 8, E->stream));
    SV_CUDA(E, cudaMemsetAsync(E->d_sc, 0, sizeof(Scalars), E->stream));
    SV_CUDA(E, cudaStreamSynchronize(E->stream));
    std::vector<uint32_t> g = E->h_group;
    return svbfm_set_groups(h, g.data(), E->G);     // re-initialises the hyper-parameters (vb.h:707-708 / mcmc.h:1109-1117)
}

int svbfm_run(svbfm_t* h, uint32_t n_iter, svbfm_iter_stats* out) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E) return SVBFM_ERR_ARG;
    SV_CUDA(E, cudaSetDevice(E->dev));
    return run_iterations(E, n_iter, out);
}
int svbfm_vb_sweep(svbfm_t* h, svbfm_iter_stats* out) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E) return SVBFM_ERR_ARG;
    if (E->cfg.method != SVBFM_VB) return fail(E, SVBFM_ERR_ARG, "svbfm_vb_sweep on a non-vb handle");
    return svbfm_run(h, 1, out);
}
int svbfm_mcmc_sweep(svbfm_t* h, svbfm_iter_stats* out) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E) return SVBFM_ERR_ARG;
    if (E->cfg.method != SVBFM_MCMC) return fail(E, SVBFM_ERR_ARG, "svbfm_mcmc_sweep on a non-mcmc handle");
    return svbfm_run(h, 1, out);
}
int svbfm_vb_online_epoch(svbfm_t* h, const uint32_t* batch_of_case, uint32_t num_batch, svbfm_iter_stats* out) {
    // One epoch of fm_learn_vb_online_simultaneous::_learn (vbos.h:66-288). Batches are case subsets of the resident
    // design matrix, selected by a 16-bit batch id per case / per CSC entry; nothing goes through the disk.
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E || !batch_of_case || num_batch == 0) return fail(E, SVBFM_ERR_ARG, "svbfm_vb_online_epoch: bad arguments");
    if (E->cfg.method != SVBFM_VB_ONLINE) return fail(E, SVBFM_ERR_ARG, "svbfm_vb_online_epoch on a non vb_online handle");
    if (!E->begun) return fail(E, SVBFM_ERR_ARG, "svbfm_begin must be called first");
    if (num_batch > 65535) return fail(E, SVBFM_ERR_ARG, "svbfm_vb_online_epoch: at most 65535 batches");
    SV_CUDA(E, cudaSetDevice(E->dev));
    cudaStream_t st = E->stream;
    const DevSplit& S = E->tr;
    if (!E->d_rbatch) {
        if (dev_alloc(E, &E->d_rbatch, S.n)) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->d_cbatch, S.nnz)) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->d_cnt_col, E->D)) return SVBFM_ERR_OOM;
    }
    if (E->batch_cap < num_batch) {
        sv_free(E->d_batch_cnt); sv_free(E->d_batch_n); E->d_batch_cnt = nullptr; E->d_batch_n = nullptr;
        if (dev_alloc(E, &E->d_batch_cnt, num_batch)) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->d_batch_n, num_batch)) return SVBFM_ERR_OOM;
        E->batch_cap = num_batch;
    }
    uint32_t* d_boc = nullptr;
    SV_CUDA(E, sv_malloc((void**)&d_boc, std::max<size_t>(S.n, 1) * 4));
    SV_CUDA(E, cudaMemcpyAsync(d_boc, batch_of_case, (size_t)S.n * 4, cudaMemcpyHostToDevice, st));
    if (S.n) {      // ids out of range are rejected before anything is indexed by them
        uint32_t* d_bad = nullptr;
        uint32_t h_bad = 0;
        cudaError_t ce = sv_malloc((void**)&d_bad, 4);
        if (ce == cudaSuccess) ce = cudaMemsetAsync(d_bad, 0, 4, st);
        if (ce == cudaSuccess) { k_vbo_check_batch_ids<<<nblk(S.n), 256, 0, st>>>(d_boc, S.n, num_batch, d_bad); LAUNCHED(E); ce = cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, st); }
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        sv_free(d_bad);
        if (ce != cudaSuccess || h_bad) sv_free(d_boc);
        SV_CUDA(E, ce);
        if (h_bad) return fail(E, SVBFM_ERR_ARG, "svbfm_vb_online_epoch: batch id out of range");
    }
    SV_CUDA(E, cudaMemsetAsync(E->d_batch_cnt, 0, (size_t)num_batch * 8, st));
    if (S.n) {
        k_vbo_set_rbatch<<<nblk(S.n), 256, 0, st>>>(d_boc, S.perm, S.n, E->d_rbatch); LAUNCHED(E);
        k_vbo_batch_counts<<<nblk(S.n), 256, 0, st>>>(E->d_rbatch, S.n, E->d_batch_cnt); LAUNCHED(E);
    }
    if (S.nnz) { k_vbo_set_cbatch<<<nblk(S.nnz), 256, 0, st>>>(S.crow, S.nnz, E->d_rbatch, E->d_cbatch); LAUNCHED(E); }
    k_u64_to_f64<<<nblk(num_batch), 256, 0, st>>>(E->d_batch_cnt, num_batch, E->d_batch_n); LAUNCHED(E);
    if (int rc = allreduce_sum_f64(E, E->d_batch_n, num_batch)) return rc;
    if (int rc = ensure_stats(E, 1)) return rc;
    SV_CUDA(E, cudaMemsetAsync(E->d_stats, 0, sizeof(DevStats), st));
    // two complete fields on one GPU: every batch is swept by the stream schedule on its own entries (batch index lists built once
    // per epoch) instead of scanning the whole design matrix with a batch mask per (batch, factor, field)
    // on the stream schedule the prediction, the reductions and the w0 shift of a batch walk the batch's own case list
    // (E->bv) instead of masking the whole arrays; SVBFM_VBO_FULL_PASSES=1 keeps the masked passes (for comparison)
    const bool full_passes = E->vbo_full_passes;
    const bool use_streams = E->vbo_streams && (S.n > 0 || E->world > 1) &&      // sharded: the same decision on every rank, cases or not
                             (uint64_t)num_batch * std::max(E->runs[0].col_end - E->runs[0].col_begin, E->runs[1].col_end - E->runs[1].col_begin) < (1ull << 31);
    if (use_streams) {
        if (!E->d_e2 && dev_alloc(E, &E->d_e2, S.n)) return SVBFM_ERR_OOM;
        if (int rc = vbo_stream_prepare(E, num_batch)) return rc;
    }
    cudaEvent_t t0, t1, t2;
    cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventCreate(&t2);
    cudaEventRecord(t0, st);
    for (uint32_t b = 0; b < num_batch; b++) {
        k_vbo_batch_begin<<<1, 1, 0, st>>>(E->d_sc, E->d_batch_n, b); LAUNCHED(E);
        const uint32_t nb_cases = use_streams ? (uint32_t)(E->vbo_off[b + 1] - E->vbo_off[b]) : 0u;
        struct BvGuard { Engine* E; ~BvGuard() { E->bv.on = false; E->bv.packed = false; E->bv.compact = false; } } bv_guard{E};      // every exit leaves the whole-run views in force
        const bool batch_streams = use_streams && (nb_cases || E->world > 1);      // sharded: a rank without cases in the batch still takes part in the collectives
        if (batch_streams) { E->bv.on = true; E->bv.lists = !full_passes; E->bv.packed = false; E->bv.compact = false; E->bv.entry0 = E->vbo_off[b]; E->bv.n = nb_cases; }
        // fresh y-hat, T for the cases of the batch (vbos.h:120-127)
        if (int rc = predict<PRED_VB_TRAIN>(E, S, E->d_e, 6, 1, (int)b)) return rc;
        SV_CUDA(E, cudaMemcpyAsync(&E->d_sc->sum_t, RED(E->d_sc, 6), 8, cudaMemcpyDeviceToDevice, st));
        if (batch_streams) {
            const Run &r0 = E->runs[0], &r1 = E->runs[1];
            E->bv.ntiles = (uint32_t)(((uint64_t)nb_cases + (1ull << E->vbo_ts_shift) - 1) >> E->vbo_ts_shift);
            for (int ri = 0; ri < 2; ri++) {
                const Run& r = ri ? r1 : r0;
                const uint32_t nc = r.col_end - r.col_begin;
                // view indexed by global column id: view[j] = first position in idx of column j of batch b
                E->bv.colptr[ri] = reinterpret_cast<const uint64_t*>(E->d_vbo_colptr[ri]) + (size_t)b * nc - r.col_begin;
                E->bv.gcnt[ri] = E->d_vbo_gcnt[ri] ? E->d_vbo_gcnt[ri] + (size_t)b * nc - r.col_begin : nullptr;
                E->bv.clist[ri] = E->d_vbo_clist[ri] ? E->d_vbo_clist[ri] + E->vbo_clist_off[ri][b] : nullptr;
                E->bv.nclist[ri] = E->d_vbo_clist[ri] ? E->vbo_clist_off[ri][b + 1] - E->vbo_clist_off[ri][b] : 0u;
                if (E->bv.ntiles && !(E->vbo_pack && E->vbo_rows && E->bv.lists)) {      // (k_stream_rows needs no first column per tile)
                    k_tile_col0<<<nblk(E->bv.ntiles), 256, 0, st>>>(E->bv.colptr[ri], r.col_begin, r.col_end, E->bv.ntiles, E->vbo_ts_shift,
                                                                   E->d_vbo_tile_col0 + (ri ? E->vbo_max_tiles : 0)); LAUNCHED(E);
                }
            }
            if (E->vbo_pack && E->bv.lists) {
                // the batch's non-empty columns as a dense id space (one GPU, k_stream_rows)
                const bool compact = E->vbo_compact && E->vbo_rows && E->world == 1 && E->d_vbo_cc && E->bv.clist[0] && E->bv.clist[1];
                if (compact && E->bv.nclist[0] + E->bv.nclist[1]) {
                    VboColsArgs ca{};
                    for (int ri = 0; ri < 2; ri++) { ca.clist[ri] = E->bv.clist[ri]; ca.nl[ri] = E->bv.nclist[ri]; ca.colptr[ri] = E->bv.colptr[ri]; }
                    ca.rec_slot = E->rec_rank ? E->d_rec_slot : nullptr; ca.group = E->d_group; ca.t_v = E->d_t_v; ca.col_count = E->d_col_count;
                    ca.cc = E->d_vbo_cc; ca.ccptr = reinterpret_cast<uint64_t*>(E->d_vbo_ccptr); ca.dT_c = E->d_vbo_dT_c;
                    k_vbo_cols<<<nblk(ca.nl[0] + ca.nl[1]), 256, 0, st>>>(ca); LAUNCHED(E);
                }
                // the batch's residuals (both entry orders), other-column ids and x values as contiguous streams
                if (nb_cases) {
                    VboPackArgs pa{};
                    if (compact) {
                        for (int ri = 0; ri < 2; ri++) {
                            const Run& r = ri ? r1 : r0;
                            pa.cpos[ri] = E->d_vbo_cpos[ri] + (size_t)b * (r.col_end - r.col_begin) - r.col_begin;
                            pa.csub[ri] = E->vbo_clist_off[ri][b] - (ri ? E->bv.nclist[0] : 0u);
                        }
                    }
                    pa.e = E->d_e; pa.rcol = S.rcol; pa.crow1 = S.crow + S.h_colptr[r1.col_begin]; pa.n = nb_cases;
                    for (int ri = 0; ri < 2; ri++) {
                        const Engine::SideView& sv = E->side[ri];
                        pa.idx[ri] = E->d_vbo_idx[ri] + E->vbo_off[b];
                        pa.oc[ri] = sv.oc + sv.entry0;
                        pa.xv[ri] = sv.xv ? sv.xv + sv.entry0 : nullptr; pa.xo[ri] = sv.xo ? sv.xo + sv.entry0 : nullptr;
                        pa.eb[ri] = E->d_vbo_eb[ri]; pa.ocb[ri] = E->d_vbo_ocb[ri]; pa.ownb[ri] = E->d_vbo_ownb[ri];
                        pa.xvb[ri] = E->d_vbo_xb[ri][0]; pa.xob[ri] = E->d_vbo_xb[ri][1];
                    }
                    k_vbo_pack<<<nblk(nb_cases), 256, 0, st>>>(pa); LAUNCHED(E);
                }
                E->bv.packed = true; E->bv.compact = compact;
            } else if (nb_cases) {
                // second residual copy for the entries of the batch
                k_gather_e_idx<<<nblk(nb_cases), 256, 0, st>>>(E->d_e, S.crow + S.h_colptr[r1.col_begin], E->d_vbo_idx[1], (uint32_t)E->vbo_off[b],
                                                                (uint32_t)E->vbo_off[b + 1], E->d_e2); LAUNCHED(E);
            }
        }
        if (E->cfg.k0) {                                           // update_w0 (vbo.h:356-358)
            if (int rc = reduce_e(E, (int)b)) return rc;
            k_vbo_w0<<<1, 1, 0, st>>>(E->d_sc); LAUNCHED(E);
            shift_e(E);
        }
        if (use_streams) {
            if (batch_streams) if (int rc = sweep_streams<2>(E)) return rc;
        } else {
            for (const Run& r : E->runs)                           // update_w; also counts |Omega_j^b| (vbo.h:360-373)
                if (int rc = sweep_run<KIND_VBO_W>(E, r, -1, (int)b)) return rc;
            for (int f = 0; f < E->K; f++)                         // update_v (vbo.h:375-408)
                for (const Run& r : E->runs)
                    if (int rc = sweep_run<KIND_VBO_V>(E, r, f, (int)b)) return rc;
        }
        if (int rc = reduce_e(E, (int)b)) return rc;
        if (E->bv.on && E->bv.packed && nb_cases) {      // the batch's residuals back into e / e2
            k_vbo_unpack<<<nblk(nb_cases), 256, 0, st>>>(E->d_vbo_eb[0], E->d_vbo_eb[1], E->d_vbo_idx[0] + E->vbo_off[b], E->d_vbo_idx[1] + E->vbo_off[b], nb_cases,
                                                        E->d_e, E->d_e2); LAUNCHED(E);
        }
        if (E->bv.on && E->bv.compact) { k_reduce_dT<<<SV_GGRID, 256, 0, st>>>(E->d_vbo_dT_c, E->bv.nclist[0] + E->bv.nclist[1], E->d_red_partial); LAUNCHED(E); }
        else { k_reduce_dT<<<SV_GGRID, 256, 0, st>>>(E->d_dT, E->D, E->d_red_partial); LAUNCHED(E); }
        if (use_streams && E->world > 1) {       // sharded stream schedule: every rank holds its own share of d(sum T)
            k_reduce_final<<<1, 256, 0, st>>>(E->d_red_partial, SV_GGRID, 1, RED(E->d_sc, 6), 0); LAUNCHED(E);
            if (int rc = allreduce_sum_f64(E, RED(E->d_sc, 6), 1)) return rc;
            k_add_scalar<<<1, 1, 0, st>>>(&E->d_sc->sum_t, RED(E->d_sc, 6)); LAUNCHED(E);
        } else {
            k_reduce_final<<<1, 256, 0, st>>>(E->d_red_partial, SV_GGRID, 1, &E->d_sc->sum_t, 1); LAUNCHED(E);
        }
        if (int rc = group_sums(E, false)) return rc;
        int want_fe = (b == 0 || b + 1 == num_batch);              // vbos.h:143-146
        k_vbo_hyper<<<1, 1, 0, st>>>(E->d_sc, E->d_grp_sums, E->d_n_per_group, E->G, E->K, E->d_hyper_w, E->d_hyper_v, E->d_stats, want_fe, b == 0, 0.5);
        LAUNCHED(E);
    }
    cudaEventRecord(t1, st);
    if (int rc = predict<PRED_VB_TEST>(E, E->te, nullptr, 3, 1)) return rc;   // vbos.h:190-244
    k_finish_iter<<<1, 1, 0, st>>>(E->d_sc, E->d_stats, SVBFM_VB_ONLINE); LAUNCHED(E);
    cudaEventRecord(t2, st);
    DevStats hs;
    cudaError_t ce = cudaMemcpyAsync(&hs, E->d_stats, sizeof(hs), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    sv_free(d_boc);
    if (ce != cudaSuccess) return fail(E, SVBFM_ERR_CUDA, std::string("vb_online epoch: ") + cudaGetErrorString(ce));
    if (out) {
        memset(out, 0, sizeof(*out));
        out->test_rmse = hs.test_rmse; out->rmse_this = hs.test_rmse; out->train_stat = NAN; out->free_energy = hs.free_energy;
        out->free_energy_first = hs.pad; out->alpha = hs.alpha; out->has_free_energy = hs.has_fe != 0.0; out->nan_inf_count = (uint32_t)hs.nan_inf;
        cudaEventElapsedTime(&out->sweep_ms, t0, t1);
        cudaEventElapsedTime(&out->predict_ms, t1, t2);
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1); cudaEventDestroy(t2);
    return check_launch(E, "vb_online_epoch");
}

int svbfm_predict(svbfm_t* h, int32_t split, double* out) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E || !out) return SVBFM_ERR_ARG;
    if (split != SVBFM_TEST) return fail(E, SVBFM_ERR_ARG, "svbfm_predict: only the test split holds predictions");
    SV_CUDA(E, cudaSetDevice(E->dev));
    SV_CUDA(E, cudaStreamSynchronize(E->stream));
    size_t n = E->te.n;
    if (E->cfg.method == SVBFM_MCMC) {
        // mcmc.h:355-379: the mean of the draws when sampling, the last prediction otherwise (als), clamped either way
        std::vector<double> s(n);
        SV_CUDA(E, copy_sync(E, s.data(), E->cfg.do_sample ? E->d_pred_sum : E->d_pred_test, n * 8, cudaMemcpyDeviceToHost));
        Scalars sc;
        SV_CUDA(E, copy_sync(E, &sc, E->d_sc, sizeof(sc), cudaMemcpyDeviceToHost));
        double it = (E->cfg.do_sample && sc.iter) ? (double)sc.iter : 1.0;
        const double lo = E->cfg.task == 1 ? 0.0 : sc.min_target, hi = E->cfg.task == 1 ? 1.0 : sc.max_target;    // classification: a probability
        for (size_t i = 0; i < n; i++) out[i] = std::fmax(lo, std::fmin(hi, s[i] / it));
    } else {
        SV_CUDA(E, copy_sync(E, out, E->d_pred_test, n * 8, cudaMemcpyDeviceToHost));
    }
    return SVBFM_OK;
}

int svbfm_get_residuals(svbfm_t* h, double* e) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E || !e || !E->d_e) return SVBFM_ERR_ARG;
    SV_CUDA(E, cudaSetDevice(E->dev));
    uint32_t n = E->tr.n;
    double* tmp = nullptr;
    SV_CUDA(E, sv_malloc((void**)&tmp, std::max<size_t>(n, 1) * 8));
    k_unpermute<<<nblk(n), 256, 0, E->stream>>>(E->d_e, E->tr.perm, n, tmp);
    SV_CUDA(E, cudaMemcpyAsync(e, tmp, (size_t)n * 8, cudaMemcpyDeviceToHost, E->stream));
    SV_CUDA(E, cudaStreamSynchronize(E->stream));
    sv_free(tmp);
    return check_launch(E, "get_residuals");
}

int svbfm_set_residuals(svbfm_t* h, const double* e, double sum_t) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E || !e) return fail(E, SVBFM_ERR_ARG, "svbfm_set_residuals: null argument");
    if (!E->begun || !E->d_e) return fail(E, SVBFM_ERR_ARG, "svbfm_set_residuals: svbfm_begin must be called first");
    if (E->world > 1) return fail(E, SVBFM_ERR_ARG, "svbfm_set_residuals: one GPU only");
    SV_CUDA(E, cudaSetDevice(E->dev));
    const uint32_t n = E->tr.n;
    double* tmp = nullptr;
    SV_CUDA(E, sv_malloc((void**)&tmp, std::max<size_t>(n, 1) * 8));
    SV_CUDA(E, cudaMemcpyAsync(tmp, e, (size_t)n * 8, cudaMemcpyHostToDevice, E->stream));
    if (n) { k_permute<<<nblk(n), 256, 0, E->stream>>>(tmp, E->tr.perm, n, E->d_e); LAUNCHED(E); }
    if (E->cfg.method != SVBFM_VB_ONLINE) sync_e2(E);
    SV_CUDA(E, cudaMemcpyAsync(&E->d_sc->sum_t, &sum_t, 8, cudaMemcpyHostToDevice, E->stream));
    SV_CUDA(E, cudaStreamSynchronize(E->stream));
    sv_free(tmp);
    return check_launch(E, "set_residuals");
}

int svbfm_get_sum_t(svbfm_t* h, double* sum_t) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E || !sum_t) return SVBFM_ERR_ARG;
    SV_CUDA(E, cudaSetDevice(E->dev));
    SV_CUDA(E, cudaStreamSynchronize(E->stream));
    SV_CUDA(E, copy_sync(E, sum_t, &E->d_sc->sum_t, 8, cudaMemcpyDeviceToHost));
    return SVBFM_OK;
}

int svbfm_copies_max_diff(svbfm_t* h, double* max_abs_diff) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E || !max_abs_diff) return SVBFM_ERR_ARG;
    *max_abs_diff = 0.0;
    if (!E->streams || !E->d_e2 || !E->tr.n || E->xs) return SVBFM_OK;      // cross shards: the two copies hold different cases on a rank
    SV_CUDA(E, cudaSetDevice(E->dev));
    unsigned long long* d = nullptr;
    SV_CUDA(E, sv_malloc((void**)&d, 8));
    SV_CUDA(E, cudaMemsetAsync(d, 0, 8, E->stream));
    const Run& r1 = E->runs[1];
    k_copies_max_diff<<<nblk(E->tr.n), 256, 0, E->stream>>>(E->d_e, E->tr.crow + E->tr.h_colptr[r1.col_begin], E->tr.n, E->d_e2, d);
    unsigned long long bits = 0;
    SV_CUDA(E, copy_sync(E, &bits, d, 8, cudaMemcpyDeviceToHost));
    sv_free(d);
    memcpy(max_abs_diff, &bits, 8);
    return check_launch(E, "copies_max_diff");
}

int svbfm_set_profile(svbfm_t* h, int32_t enabled) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E) return SVBFM_ERR_ARG;
    E->profile = enabled != 0;
    return SVBFM_OK;
}

int svbfm_get_profile(svbfm_t* h, double ms[SVBFM_PROFILE_CLASSES], uint64_t launches[SVBFM_PROFILE_CLASSES]) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E || !ms || !launches) return SVBFM_ERR_ARG;
    SV_CUDA(E, cudaSetDevice(E->dev));
    SV_CUDA(E, cudaStreamSynchronize(E->stream));
    for (int c = 0; c < SVBFM_PROFILE_CLASSES; c++) { ms[c] = 0.0; launches[c] = 0; }
    for (auto& sp : E->prof_spans) {
        float t = 0.f;
        cudaEventElapsedTime(&t, sp.a, sp.b);
        ms[sp.cls] += t; launches[sp.cls] += 1;
        cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);
    }
    E->prof_spans.clear();
    return SVBFM_OK;
}

int svbfm_get_info(svbfm_t* h, svbfm_info* out) {
    Engine* E = reinterpret_cast<Engine*>(h);
    if (!E || !out) return SVBFM_ERR_ARG;
    memset(out, 0, sizeof(*out));
    out->num_runs = (uint32_t)E->runs.size();
    out->num_tiles = E->streams ? E->s_ntiles[0] + E->s_ntiles[1] : E->n_tiles;
    out->uniform_row_nnz = E->tr.uniformF;
    out->all_ones = E->tr.all_ones;
    out->kernel_launches = E->launches;
    out->device_bytes = E->dev_bytes;
    out->train_nnz = E->tr.nnz;
    out->rows_reordered = E->rows_reordered;
    out->world_size = (uint32_t)E->world;
    out->fused_schedule = ((stream_ok(E) || E->vbo_streams) ? 1u : 0u) | (E->rec_rank ? 2u : 0u) | ((E->stream_tma && stream_ok(E)) ? 4u : 0u) |
                          (E->graph_replays ? 8u : 0u);
    out->exclusive_blocks = (E->excl0 ? 1u : 0u) | (E->xs ? 2u : 0u) | (E->p2p ? 4u : 0u);
    return SVBFM_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- host helpers
// Exported for the Python mirror of the learner shells (the C++ shells include host/init_state.h directly).
#include "../host/init_state.h"
extern "C" {
// replaces: srand(time) + fm.init() + fm.w.init_normal + fm_learn_vb::init draws (see host/init_state.h)
int svbfm_host_init_state(long seed, uint32_t D, int32_t K, double init_stdev, int32_t method, double* w0_mean, double* w0_var,
                          double* w_mean, double* w_var, double* v_mean, double* v_var) {
    svbfm_host::InitialState s;
    svbfm_host::init_state(seed, D, K, init_stdev, method, s);
    if (w0_mean) *w0_mean = s.w0_mean;
    if (w0_var) *w0_var = s.w0_var;
    size_t KD = (size_t)K * D;
    if (w_mean) memcpy(w_mean, s.w_mean.data(), (size_t)D * 8);
    if (w_var) memcpy(w_var, s.w_var.data(), (size_t)D * 8);
    if (v_mean) memcpy(v_mean, s.v_mean.data(), KD * 8);
    if (v_var) memcpy(v_var, s.v_var.data(), KD * 8);
    return 0;
}
// replaces: std::random_shuffle(shuffle, shuffle + n) on the libc stream (fm_learn_vb_online_simultaneous.h:74)
int svbfm_host_random_shuffle(uint32_t* a, uint32_t n) { svbfm_host::libc_random_shuffle(a, n); return 0; }
}
