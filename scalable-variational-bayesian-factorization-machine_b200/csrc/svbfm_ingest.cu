// csrc/svbfm_ingest.cu -- device ingest of one data split handed over as CSC (reference DataSubset::data_t).
//
//   CSC (caller) --H2D--> stable radix sort by case id --> CSR (features ascending inside each case)
//   train only:  need[j] = max "previous feature of the same case" over column j  --> field runs (host scan)
//                case re-ordering so that run 0 streams, CSC rebuilt by a stable sort by feature id
//                warp tiles (<= tile_entries CSC entries of ONE column) + list of heavy columns
//
// Replaces, for the device side, Data::create_data_t (reference src/libfm/src/Data.h:457-509) -- here in the
// CSC -> CSR direction -- and adds the field-run scheduler the reference does not need (it sweeps columns one
// by one, fm_learn_vb.h:395, 427). CUB is used for the two radix sorts and one scan: ingest plumbing, not the
// sweep.
#include <cub/cub.cuh>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include "svbfm_internal.h"

#include <map>
#include <tuple>
#include <mutex>
#include <unordered_map>

namespace svb {

// ---- size-keyed block cache (see svbfm_internal.h) -----------------------------------------------------------------
namespace {
struct BlockCache {
    std::mutex mu;
    // (device, bytes, owner) -> block. A freed block goes back under the handle that allocated it: kernels of that handle's stream may
    // still be using it, and only work queued on the same stream afterwards is ordered behind them. Owner null: idle blocks anyone may
    // take (their owner synchronized its stream: svbfm_destroy, svbfm_set_stream).
    std::multimap<std::tuple<int, size_t, const void*>, void*> free_blocks;
    struct Live { int dev; size_t bytes; const void* owner; };
    std::unordered_map<void*, Live> live;
};
BlockCache& cache() { static BlockCache c; return c; }
thread_local const void* t_owner = nullptr;
}  // namespace

void sv_set_owner(const void* owner) { t_owner = owner; }
void sv_owner_release(const void* owner) {
    if (!owner) return;
    BlockCache& c = cache();
    std::lock_guard<std::mutex> g(c.mu);
    std::vector<std::pair<std::tuple<int, size_t, const void*>, void*>> moved;
    for (auto f = c.free_blocks.begin(); f != c.free_blocks.end();)
        if (std::get<2>(f->first) == owner) { moved.push_back({{std::get<0>(f->first), std::get<1>(f->first), nullptr}, f->second}); f = c.free_blocks.erase(f); } else ++f;
    for (auto& m : moved) c.free_blocks.insert(m);
}

cudaError_t sv_malloc(void** p, size_t bytes) {
#if defined(SVBFM_EMULATED)
    bytes = (std::max<size_t>(bytes, 1) + 31) & ~(size_t)31;        // tests/emu: the guard page of SVBFM_EMU_GUARD=1 sits at the buffer's own end
#else
    bytes = (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
#endif
    int dev = 0;
    cudaGetDevice(&dev);
    BlockCache& c = cache();
    std::lock_guard<std::mutex> g(c.mu);
    auto it = c.free_blocks.find({dev, bytes, t_owner});
    if (it == c.free_blocks.end() && t_owner) it = c.free_blocks.find({dev, bytes, nullptr});
    if (it != c.free_blocks.end()) {
        *p = it->second;
        c.free_blocks.erase(it);
        c.live[*p] = {dev, bytes, t_owner};
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {                      // memory pressure: give the cached blocks of this device back and retry once
        cudaGetLastError();
        cudaDeviceSynchronize();
        for (auto f = c.free_blocks.begin(); f != c.free_blocks.end();)
            if (std::get<0>(f->first) == dev) { cudaFree(f->second); f = c.free_blocks.erase(f); } else ++f;
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) c.live[*p] = {dev, bytes, t_owner};
    return e;
}
cudaError_t sv_free(void* p) {
    if (!p) return cudaSuccess;
    BlockCache& c = cache();
    std::lock_guard<std::mutex> g(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) return cudaFree(p);
    c.free_blocks.insert({{it->second.dev, it->second.bytes, it->second.owner}, p});
    c.live.erase(it);
    return cudaSuccess;
}
void sv_cache_release() {
    BlockCache& c = cache();
    std::lock_guard<std::mutex> g(c.mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& f : c.free_blocks) { cudaSetDevice(std::get<0>(f.first)); cudaFree(f.second); }
    c.free_blocks.clear();
    cudaSetDevice(cur);
}

int fail(Engine* E, int code, const std::string& msg) {
    if (E) E->err = msg;
    return code;
}

static __global__ void k_any_not_one(const float* __restrict__ x, uint64_t n, uint32_t* flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < n; i += stride) bad |= (x[i] != 1.0f);
    if (bad) *flag = 1;
}

// feature id of every CSC entry: upper_bound over colptr
static __global__ void k_col_of_entry(const uint64_t* __restrict__ colptr, uint32_t ncols, uint64_t nnz, uint32_t* __restrict__ out) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    uint32_t lo = 0, hi = ncols;   // find j with colptr[j] <= p < colptr[j+1]
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (colptr[mid] <= p) lo = mid; else hi = mid;
    }
    out[p] = lo;
}

static __global__ void k_iota(uint32_t* a, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (uint32_t)i;
}

static __global__ void k_check_case_ids(const uint32_t* __restrict__ ids, uint64_t nnz, uint32_t n, uint32_t* flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz && ids[i] >= n) *flag = 1;
}

// rowptr[i] = first position in sorted_keys with key >= i   (i in [0, n])
static __global__ void k_rowptr_from_sorted(const uint32_t* __restrict__ keys, uint64_t nnz, uint32_t n, uint64_t* __restrict__ rowptr) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint64_t lo = 0, hi = nnz;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (keys[mid] < (uint32_t)i) lo = mid + 1; else hi = mid;
    }
    rowptr[i] = lo;
}

static __global__ void k_gather_u32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t n, uint32_t* __restrict__ dst) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
static __global__ void k_gather_f32(const float* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t n, float* __restrict__ dst) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// per case: duplicate feature ids, row-length uniformity, and need[] for the run scheduler
static __global__ void k_scan_rows(const uint64_t* __restrict__ rowptr, const uint32_t* __restrict__ rcol, uint32_t n,
                                   uint32_t* __restrict__ need /*[ncols] or null*/, uint32_t* flags /*[0]=dup [1]=non-uniform*/) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t b = rowptr[i], e = rowptr[i + 1];
    if ((e - b) != (rowptr[1] - rowptr[0])) flags[1] = 1;
    for (uint64_t k = b + 1; k < e; k++) {
        uint32_t c0 = rcol[k - 1], c1 = rcol[k];
        if (c0 == c1) flags[0] = 1;
        if (need) atomicMax(&need[c1], c0 + 1);   // column c1 must start a new run if c0 is inside the current run
    }
}

static __global__ void k_invert_perm(const uint32_t* __restrict__ perm, uint32_t n, uint32_t* __restrict__ inv, uint32_t* not_identity) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    uint32_t o = perm[d];
    inv[o] = d;
    if (o != d) *not_identity = 1;
}

static __global__ void k_row_lengths_perm(const uint64_t* __restrict__ rowptr, const uint32_t* __restrict__ perm, uint32_t n, uint64_t* __restrict__ len) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) { uint32_t o = perm[d]; len[d] = rowptr[o + 1] - rowptr[o]; }
    if (d == n) len[d] = 0;
}

static __global__ void k_permute_rows(const uint64_t* __restrict__ old_ptr, const uint32_t* __restrict__ old_col, const float* __restrict__ old_val,
                                      const uint32_t* __restrict__ perm, uint32_t n, const uint64_t* __restrict__ new_ptr,
                                      uint32_t* __restrict__ new_col, float* __restrict__ new_val, uint32_t* __restrict__ new_row_of_entry) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    uint32_t o = perm[d];
    uint64_t src = old_ptr[o], dst = new_ptr[d], len = old_ptr[o + 1] - src;
    for (uint64_t k = 0; k < len; k++) {
        new_col[dst + k] = old_col[src + k];
        if (old_val) new_val[dst + k] = old_val[src + k];
        new_row_of_entry[dst + k] = d;
    }
}

static __global__ void k_permute_f32(const float* __restrict__ src, const uint32_t* __restrict__ perm, uint32_t n, float* __restrict__ dst) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) dst[d] = src[perm[d]];
}

// F == 2: the other entry of the case of every CSC entry (entry-aligned, streamed by the sweeps)
static __global__ void k_other_of_entry(const uint64_t* __restrict__ colptr, uint32_t ncols, uint64_t nnz, const uint32_t* __restrict__ crow,
                                        const uint32_t* __restrict__ rcol, const float* __restrict__ rval, uint32_t* __restrict__ ocol,
                                        float* __restrict__ oval) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    uint32_t lo = 0, hi = ncols;
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (colptr[mid] <= p) lo = mid; else hi = mid;
    }
    uint32_t i = crow[p];
    uint2 c = reinterpret_cast<const uint2*>(rcol)[i];
    bool first = (c.x == lo);
    ocol[p] = first ? c.y : c.x;
    if (oval) { float2 x = reinterpret_cast<const float2*>(rval)[i]; oval[p] = first ? x.y : x.x; }
}

// cuts[c][k] = first entry of column cols[c] whose case id is >= k * block_cases  (k = 0..NB)
static __global__ void k_block_cuts(const uint32_t* __restrict__ cols, const uint64_t* __restrict__ colptr, const uint32_t* __restrict__ crow,
                                    uint32_t block_cases, uint32_t NB, uint64_t* __restrict__ cuts) {
    uint32_t c = blockIdx.x, k = blockIdx.y * blockDim.x + threadIdx.x;
    if (k > NB) return;
    uint32_t j = cols[c];
    uint64_t lo = colptr[j], hi = colptr[j + 1];
    uint64_t key = (uint64_t)k * block_cases;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if ((uint64_t)crow[mid] < key) lo = mid + 1; else hi = mid;
    }
    cuts[(size_t)c * (NB + 1) + k] = lo;
}

// ---- SVBFM_REC_RANK (experiment): record slots of the second field's columns by popularity rank
// key = descending column length, value = column id (sorted by key: rank order)
static __global__ void k_slot_keys(const uint64_t* __restrict__ colptr, uint32_t c0, uint32_t nc, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nc) return;
    uint64_t len = colptr[c0 + r + 1] - colptr[c0 + r];
    keys[r] = 0xffffffffu - (uint32_t)(len > 0xffffffffull ? 0xffffffffull : len);
    vals[r] = c0 + r;
}
static __global__ void k_slot_scatter(const uint32_t* __restrict__ cols_by_rank, uint32_t c0, uint32_t nc, uint32_t* __restrict__ slot) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nc) slot[cols_by_rank[r]] = c0 + r;
}
// sort key of the case at position k of `cases`: which = 1 -> slot rank of its second-field column, 0 -> its first-field column
static __global__ void k_case_keys(const uint32_t* __restrict__ cases, const uint32_t* __restrict__ rcol, const uint32_t* __restrict__ slot, uint32_t base,
                                   uint32_t n, int which, uint32_t* __restrict__ keys) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint2 c = reinterpret_cast<const uint2*>(rcol)[cases[k]];
    keys[k] = (which ? slot[c.y] : c.x) - base;
}
static __global__ void k_map_slots(uint32_t* __restrict__ oc, uint64_t n, const uint32_t* __restrict__ slot) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) oc[p] = slot[oc[p]];
}

static inline unsigned nblk(uint64_t n, unsigned t = 256) { return (unsigned)((n + t - 1) / t); }

static int bits_for(uint64_t n) {
    int b = 1;
    while (b < 32 && (1ull << b) < n) b++;
    return b;
}

// stable sort of (key, value) pairs by key; returns device arrays (caller frees with cudaFree)
static int sort_pairs(Engine* E, const uint32_t* keys_in, const uint32_t* vals_in, uint64_t n, uint64_t key_range,
                      uint32_t** keys_out, uint32_t** vals_out) {
    SV_CUDA(E, sv_malloc((void**)keys_out, std::max<uint64_t>(n, 1) * 4));
    SV_CUDA(E, sv_malloc((void**)vals_out, std::max<uint64_t>(n, 1) * 4));
    if (n == 0) return 0;
    size_t tmp_bytes = 0;
    int end_bit = bits_for(key_range);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, *keys_out, vals_in, *vals_out, (int64_t)n, 0, end_bit, E->stream);
    void* tmp = nullptr;
    SV_CUDA(E, sv_malloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, *keys_out, vals_in, *vals_out, (int64_t)n, 0, end_bit, E->stream);
    cudaError_t e2 = cudaStreamSynchronize(E->stream);
    sv_free(tmp);
    if (e != cudaSuccess || e2 != cudaSuccess) return fail(E, SVBFM_ERR_CUDA, std::string("radix sort: ") + cudaGetErrorString(e != cudaSuccess ? e : e2));
    return 0;
}

void free_split(Engine* E, DevSplit& S) {
    (void)E;
    sv_free(S.colptr); sv_free(S.crow); sv_free(S.cval); sv_free(S.rowptr); sv_free(S.rcol); sv_free(S.rval);
    sv_free(S.y); sv_free(S.perm); sv_free(S.cother); sv_free(S.cother_val);
    S = DevSplit();
}

int ingest_split(Engine* E, DevSplit& S, bool is_train, uint32_t n, uint32_t ncols, const uint64_t* colptr, const uint32_t* case_id,
                 const float* x, const float* target) {
    cudaStream_t st = E->stream;
    free_split(E, S);
    if (is_train) {
        free_second(E);
        // per-case / per-entry buffers of vb_online are sized by the train split: a new split (svbfm_reset + svbfm_set_csc on a
        // long-lived handle) must not inherit the old ones
        sv_free(E->d_rbatch); sv_free(E->d_cbatch); sv_free(E->d_cnt_col);
        E->d_rbatch = nullptr; E->d_cbatch = nullptr; E->d_cnt_col = nullptr;
    }
    // temporaries of this function: whatever is still owned when it returns (early, on an error) goes back to the block cache
    struct Temps {
        std::vector<void**> slots;
        cudaStream_t side = nullptr;       // the copy stream: its transfers into these blocks must have landed before they are recycled
        void own(void** p) { slots.push_back(p); }
        ~Temps() {
            bool any = false;
            for (void** s : slots) any = any || (*s != nullptr);
            if (any && side) cudaStreamSynchronize(side);
            for (void** s : slots) { sv_free(*s); *s = nullptr; }
        }
    } temps;
    auto drop = [](auto*& p) { sv_free(p); p = nullptr; };
    const bool timing = getenv("SVBFM_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!timing) return;
        cudaStreamSynchronize(st);
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[svbfm ingest %s] %-28s %8.2f ms\n", is_train ? "train" : "test", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    // Several ranks: a rank that rejects its shard must not leave the others waiting in the collectives below, and the allreduce
    // counts of the train ingest follow num_cols: the ranks first agree on {any rank failed, min and max of num_cols} (agree()).
    auto agree = [&](const std::string& why) -> int {
        if (E->world <= 1) return why.empty() ? 0 : fail(E, SVBFM_ERR_ARG, why);
        uint32_t h[3] = {why.empty() ? 0u : 1u, ncols, ~ncols};           // max-reduced: [1] = max num_cols, ~[2] = min num_cols
        uint32_t* d = nullptr;
        if (sv_malloc((void**)&d, 12) != cudaSuccess) return fail(E, SVBFM_ERR_OOM, "cudaMalloc");
        cudaMemcpyAsync(d, h, 12, cudaMemcpyHostToDevice, st);
        int rc = allreduce(E, d, 3, 3 /*ncclUint32*/, 2 /*ncclMax*/);
        if (!rc) { cudaMemcpyAsync(h, d, 12, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); }
        sv_free(d);
        if (rc) return rc;
        if (!why.empty()) return fail(E, SVBFM_ERR_ARG, why);
        if (h[0]) return fail(E, SVBFM_ERR_ARG, "set_csc: another rank rejected its shard");
        if (is_train && h[1] != ~h[2]) return fail(E, SVBFM_ERR_ARG, "set_csc: num_cols of the train split differs between the ranks");
        return 0;
    };
    std::string why;
    if (ncols > E->D) why = "set_csc: num_cols exceeds num_attribute";
    else if (colptr[0] != 0) why = "set_csc: colptr[0] != 0";
    else {
        for (uint32_t j = 0; j < ncols && why.empty(); j++)
            if (colptr[j + 1] < colptr[j]) why = "set_csc: colptr not monotone";
        if (why.empty() && colptr[ncols] >= (1ull << 32)) why = "set_csc: more than 2^32-1 entries per rank are not supported";
    }
    if (int rc = agree(why)) return rc;
    uint64_t nnz = colptr[ncols];
    S.n = n; S.n_cols = ncols; S.nnz = nnz;
    // mcmc also draws the attributes that never occur in train (reference fm_learn_mcmc.h:449-457, 568-577):
    // they become empty trailing columns of the last run.
    S.ncols_ext = (is_train && E->cfg.method == SVBFM_MCMC) ? E->D : ncols;
    S.h_colptr.assign(colptr, colptr + ncols + 1);
    S.h_colptr.resize((size_t)S.ncols_ext + 1, nnz);

    if (dev_alloc(E, &S.colptr, (size_t)S.ncols_ext + 1)) return SVBFM_ERR_OOM;
    if (dev_alloc(E, &S.crow, nnz)) return SVBFM_ERR_OOM;
    if (dev_alloc(E, &S.y, n)) return SVBFM_ERR_OOM;
    float* d_x = nullptr;
    temps.own((void**)&d_x);
    if (x) SV_CUDA(E, sv_malloc((void**)&d_x, std::max<uint64_t>(nnz, 1) * 4));       // x == null: every value is 1 (nothing to ship or to check)
    // The values and the targets are not needed before the CSR gather / the case re-ordering: they travel on a second stream
    // while the main stream sorts the case ids (pinned host buffers; with pageable memory the copies serialise anyway).
    cudaStream_t cs = E->copy_stream ? E->copy_stream : st;
    temps.side = (cs != st) ? cs : nullptr;
    uint32_t* d_flags = nullptr;   // [0] any x != 1  [1] case id out of range  [2] duplicate feature in a case  [3] non-uniform  [4] perm not identity
    temps.own((void**)&d_flags);
    SV_CUDA(E, sv_malloc((void**)&d_flags, 8 * 4));
    SV_CUDA(E, cudaMemsetAsync(d_flags, 0, 8 * 4, st));
    if (cs != st) {                // everything queued on the main stream so far (earlier users of the recycled blocks, the memset) first
        SV_CUDA(E, cudaEventRecord(E->copy_event, st));
        SV_CUDA(E, cudaStreamWaitEvent(cs, E->copy_event, 0));
    }
    SV_CUDA(E, cudaMemcpyAsync(S.colptr, S.h_colptr.data(), ((size_t)S.ncols_ext + 1) * 8, cudaMemcpyHostToDevice, st));
    SV_CUDA(E, cudaMemcpyAsync(S.crow, case_id, nnz * 4, cudaMemcpyDefault, st));
    if (x) {
        SV_CUDA(E, cudaMemcpyAsync(d_x, x, nnz * 4, cudaMemcpyDefault, cs));
        if (nnz) k_any_not_one<<<std::min<unsigned>(nblk(nnz), 148 * 16), 256, 0, cs>>>(d_x, nnz, d_flags + 0);
    }
    SV_CUDA(E, cudaMemcpyAsync(S.y, target, (size_t)n * 4, cudaMemcpyDefault, cs));
    if (nnz) k_check_case_ids<<<nblk(nnz), 256, 0, st>>>(S.crow, nnz, n, d_flags + 1);
    // CSC -> CSR: feature id per entry, stable sort by case id
    uint32_t *d_colof = nullptr, *d_idx = nullptr, *d_skeys = nullptr, *d_sidx = nullptr;
    temps.own((void**)&d_colof); temps.own((void**)&d_idx); temps.own((void**)&d_skeys); temps.own((void**)&d_sidx);
    SV_CUDA(E, sv_malloc((void**)&d_colof, std::max<uint64_t>(nnz, 1) * 4));
    SV_CUDA(E, sv_malloc((void**)&d_idx, std::max<uint64_t>(nnz, 1) * 4));
    if (nnz) {
        k_col_of_entry<<<nblk(nnz), 256, 0, st>>>(S.colptr, ncols, nnz, d_colof);
        k_iota<<<nblk(nnz), 256, 0, st>>>(d_idx, nnz);
    }
    uint32_t h_flags[8];
    SV_CUDA(E, cudaMemcpyAsync(h_flags + 1, d_flags + 1, 4, cudaMemcpyDeviceToHost, st));
    SV_CUDA(E, cudaStreamSynchronize(st));
    if (int rc = agree(h_flags[1] ? "set_csc: case id out of range" : "")) {
        cudaStreamSynchronize(cs);
        drop(d_x); drop(d_flags); drop(d_colof); drop(d_idx);
        return rc;
    }
    mark("H2D of the case ids + col_of_entry");
    if (int r = sort_pairs(E, S.crow, d_idx, nnz, std::max<uint32_t>(n, 1), &d_skeys, &d_sidx)) { cudaStreamSynchronize(cs); return r; }
    mark("  sort pairs by case");
    drop(d_idx);
    uint64_t* d_rowptr = nullptr;
    temps.own((void**)&d_rowptr);
    SV_CUDA(E, sv_malloc((void**)&d_rowptr, ((size_t)n + 1) * 8));
    k_rowptr_from_sorted<<<nblk((uint64_t)n + 1), 256, 0, st>>>(d_skeys, nnz, n, d_rowptr);
    mark("  rowptr");
    drop(d_skeys);
    uint32_t* d_rcol = nullptr; float* d_rval = nullptr;
    temps.own((void**)&d_rcol); temps.own((void**)&d_rval);
    SV_CUDA(E, sv_malloc((void**)&d_rcol, std::max<uint64_t>(nnz, 1) * 4));
    if (nnz) k_gather_u32<<<nblk(nnz), 256, 0, st>>>(d_colof, d_sidx, nnz, d_rcol);
    drop(d_colof);
    mark("sort by case + CSR gather");

    // per-case scan: duplicates, uniform length, need[]
    uint32_t* d_need = nullptr;
    temps.own((void**)&d_need);
    if (is_train) {
        SV_CUDA(E, sv_malloc((void**)&d_need, std::max<uint32_t>(S.ncols_ext, 1) * 4));
        SV_CUDA(E, cudaMemsetAsync(d_need, 0, std::max<uint32_t>(S.ncols_ext, 1) * 4, st));
    }
    if (n) k_scan_rows<<<nblk(n), 256, 0, st>>>(d_rowptr, d_rcol, n, d_need, d_flags + 2);
    // the values and targets have had the sort, the CSR gather and the row scan to arrive
    SV_CUDA(E, cudaMemcpyAsync(h_flags, d_flags, 4, cudaMemcpyDeviceToHost, cs));
    SV_CUDA(E, cudaStreamSynchronize(cs));
    S.all_ones = (h_flags[0] == 0);
    if (!S.all_ones) {
        SV_CUDA(E, sv_malloc((void**)&d_rval, std::max<uint64_t>(nnz, 1) * 4));
        if (nnz) k_gather_f32<<<nblk(nnz), 256, 0, st>>>(d_x, d_sidx, nnz, d_rval);
    }
    drop(d_sidx);
    mark("  H2D of values and targets (overlapped) + row scan");
    SV_CUDA(E, cudaMemcpyAsync(h_flags, d_flags, 8 * 4, cudaMemcpyDeviceToHost, st));
    SV_CUDA(E, cudaStreamSynchronize(st));
    if (int rc = agree(h_flags[2] ? "set_csc: a feature id occurs twice in one case; not supported" : "")) {
        drop(d_x); drop(d_flags); drop(d_rowptr); drop(d_rcol); drop(d_rval); drop(d_need);
        if (h_flags[2]) return fail(E, SVBFM_ERR_DATA, "set_csc: a feature id occurs twice in one case; not supported");
        return rc;
    }
    bool uniform = (n > 0) && (h_flags[3] == 0);
    uint32_t F = uniform ? (uint32_t)(nnz / n) : 0;
    if (E->world > 1) {   // all ranks must agree on the kernel variant? no: variants are rank-local; only runs must agree
    }

    if (is_train) {
        // ---- field runs (global when sharded: need[] is max-reduced over ranks)
        if (int r = allreduce(E, d_need, S.ncols_ext, 3 /*ncclUint32*/, 2 /*ncclMax*/)) return r;
        std::vector<uint32_t> need(S.ncols_ext);
        SV_CUDA(E, cudaMemcpyAsync(need.data(), d_need, (size_t)S.ncols_ext * 4, cudaMemcpyDeviceToHost, st));
        SV_CUDA(E, cudaStreamSynchronize(st));
        drop(d_need);
        E->runs.clear();
        uint32_t run_start = 0;
        for (uint32_t j = 0; j < S.ncols_ext; j++) {
            if (need[j] > run_start) {   // a case of column j already has a feature inside [run_start, j)
                Run r; r.col_begin = run_start; r.col_end = j;
                E->runs.push_back(r);
                run_start = j;
            }
        }
        if (S.ncols_ext > 0) { Run r; r.col_begin = run_start; r.col_end = S.ncols_ext; E->runs.push_back(r); }
        for (auto& r : E->runs) r.nnz = S.h_colptr[r.col_end] - S.h_colptr[r.col_begin];
        mark("row scan + field runs");

        // ---- case re-ordering: device order = order of the cases inside run 0 (when run 0 holds every case once)
        bool reorder = !(E->cfg.flags & SVBFM_FLAG_NO_ROW_REORDER) && !E->runs.empty() && E->runs[0].nnz == n && n > 0 && nnz > 0;
        // two complete fields: record slots of the second field by popularity rank (Engine::d_rec_slot)
        sv_free(E->d_rec_slot); E->d_rec_slot = nullptr; E->rec_rank = false;
        const bool want_rank = reorder && E->want_rec_rank && !getenv("SVBFM_NO_FUSE") && E->runs.size() == 2 && uniform && F == 2 && E->runs[1].nnz == n;
        if (want_rank) {
            const Run& r1 = E->runs[1];
            const uint32_t nc1 = r1.col_end - r1.col_begin;
            uint32_t *d_k = nullptr, *d_v = nullptr, *d_ks = nullptr, *d_vs = nullptr;
            SV_CUDA(E, sv_malloc((void**)&d_k, (size_t)nc1 * 4));
            SV_CUDA(E, sv_malloc((void**)&d_v, (size_t)nc1 * 4));
            if (dev_alloc(E, &E->d_rec_slot, E->D)) return SVBFM_ERR_OOM;
            k_iota<<<nblk(E->D), 256, 0, st>>>(E->d_rec_slot, E->D);
            k_slot_keys<<<nblk(nc1), 256, 0, st>>>(S.colptr, r1.col_begin, nc1, d_k, d_v);
            if (int r = sort_pairs(E, d_k, d_v, nc1, 1ull << 32, &d_ks, &d_vs)) return r;
            k_slot_scatter<<<nblk(nc1), 256, 0, st>>>(d_vs, r1.col_begin, nc1, E->d_rec_slot);
            SV_CUDA(E, cudaStreamSynchronize(st));
            sv_free(d_k); sv_free(d_v); sv_free(d_ks); sv_free(d_vs);
        }
        if (reorder) {
            uint32_t *d_perm = nullptr, *d_inv = nullptr;
            SV_CUDA(E, sv_malloc((void**)&d_perm, (size_t)n * 4));
            SV_CUDA(E, sv_malloc((void**)&d_inv, (size_t)n * 4));
            SV_CUDA(E, cudaMemcpyAsync(d_perm, S.crow + S.h_colptr[E->runs[0].col_begin], (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
            if (want_rank) {
                // cases of one first-field column in the rank order of their second-field column: stable sort by rank, then by column
                const Run &r0 = E->runs[0], &r1 = E->runs[1];
                for (int which = 1; which >= 0; which--) {
                    uint32_t *d_k = nullptr, *d_ks = nullptr, *d_vs = nullptr;
                    SV_CUDA(E, sv_malloc((void**)&d_k, (size_t)n * 4));
                    k_case_keys<<<nblk(n), 256, 0, st>>>(d_perm, d_rcol, E->d_rec_slot, which ? r1.col_begin : r0.col_begin, n, which, d_k);
                    if (int r = sort_pairs(E, d_k, d_perm, n, std::max<uint32_t>(which ? r1.col_end - r1.col_begin : r0.col_end - r0.col_begin, 1), &d_ks, &d_vs)) return r;
                    sv_free(d_k); sv_free(d_ks); sv_free(d_perm);
                    d_perm = d_vs;
                }
            }
            k_invert_perm<<<nblk(n), 256, 0, st>>>(d_perm, n, d_inv, d_flags + 4);
            SV_CUDA(E, cudaMemcpyAsync(h_flags, d_flags, 8 * 4, cudaMemcpyDeviceToHost, st));
            SV_CUDA(E, cudaStreamSynchronize(st));
            sv_free(d_inv);
            if (h_flags[4]) {
                // new CSR = cases gathered in device order
                uint64_t *d_len = nullptr, *d_newptr = nullptr;
                SV_CUDA(E, sv_malloc((void**)&d_len, ((size_t)n + 1) * 8));
                SV_CUDA(E, sv_malloc((void**)&d_newptr, ((size_t)n + 1) * 8));
                k_row_lengths_perm<<<nblk((uint64_t)n + 1), 256, 0, st>>>(d_rowptr, d_perm, n, d_len);
                size_t tmp_bytes = 0;
                cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_len, d_newptr, (int64_t)n + 1, st);
                void* tmp = nullptr;
                SV_CUDA(E, sv_malloc(&tmp, tmp_bytes ? tmp_bytes : 1));
                cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_len, d_newptr, (int64_t)n + 1, st);
                uint32_t *d_ncol = nullptr, *d_nrow = nullptr; float* d_nval = nullptr;
                SV_CUDA(E, sv_malloc((void**)&d_ncol, nnz * 4));
                SV_CUDA(E, sv_malloc((void**)&d_nrow, nnz * 4));
                if (!S.all_ones) SV_CUDA(E, sv_malloc((void**)&d_nval, nnz * 4));
                k_permute_rows<<<nblk(n), 256, 0, st>>>(d_rowptr, d_rcol, d_rval, d_perm, n, d_newptr, d_ncol, d_nval, d_nrow);
                SV_CUDA(E, cudaStreamSynchronize(st));
                sv_free(tmp); sv_free(d_len);
                drop(d_rowptr); drop(d_rcol); drop(d_rval);
                d_rowptr = d_newptr; d_rcol = d_ncol; d_rval = d_nval;
                // new CSC = stable sort of the new CSR entries by feature id (case ids stay ascending per column)
                uint32_t *d_eidx = nullptr, *d_k2 = nullptr, *d_v2 = nullptr;
                SV_CUDA(E, sv_malloc((void**)&d_eidx, nnz * 4));
                k_iota<<<nblk(nnz), 256, 0, st>>>(d_eidx, nnz);
                if (int r = sort_pairs(E, d_rcol, d_eidx, nnz, std::max<uint32_t>(ncols, 1), &d_k2, &d_v2)) return r;
                sv_free(d_eidx); sv_free(d_k2);
                k_gather_u32<<<nblk(nnz), 256, 0, st>>>(d_nrow, d_v2, nnz, S.crow);
                if (!S.all_ones) {
                    float* d_cv = nullptr;
                    SV_CUDA(E, sv_malloc((void**)&d_cv, nnz * 4));
                    k_gather_f32<<<nblk(nnz), 256, 0, st>>>(d_rval, d_v2, nnz, d_cv);
                    SV_CUDA(E, cudaStreamSynchronize(st));
                    drop(d_x); d_x = d_cv;
                }
                SV_CUDA(E, cudaStreamSynchronize(st));
                sv_free(d_v2); sv_free(d_nrow);
                float* d_y2 = nullptr;
                SV_CUDA(E, sv_malloc((void**)&d_y2, (size_t)n * 4));
                k_permute_f32<<<nblk(n), 256, 0, st>>>(S.y, d_perm, n, d_y2);
                SV_CUDA(E, cudaStreamSynchronize(st));
                sv_free(S.y); S.y = d_y2;
                S.perm = d_perm; E->dev_bytes += (size_t)n * 4;
                E->rows_reordered = true;
            } else {
                sv_free(d_perm);     // the caller's order already is run 0's order
            }
            E->run0_sequential = true;
        }
        const bool two_fields = E->run0_sequential && !getenv("SVBFM_NO_FUSE") && E->runs.size() == 2 && F == 2 && E->runs[1].nnz == n;
        E->streams = two_fields && E->cfg.method != SVBFM_VB_ONLINE;
        E->vbo_streams = two_fields && E->cfg.method == SVBFM_VB_ONLINE && !getenv("SVBFM_NO_VBO_STREAM") && (E->world == 1 || !getenv("SVBFM_NO_VBO_STREAM_SHARDED"));
        E->excl0 = false;
        if (E->world > 1) {      // one schedule for all ranks
            uint32_t* d_ok = d_flags + 6;
            uint32_t ok = (E->streams ? 1u : 0u) | (E->vbo_streams ? 2u : 0u);
            uint32_t ok2[2] = {ok & 1u, (ok >> 1) & 1u};
            SV_CUDA(E, cudaMemcpyAsync(d_ok, ok2, 8, cudaMemcpyHostToDevice, st));
            if (int r = allreduce(E, d_ok, 2, 3 /*ncclUint32*/, 3 /*ncclMin*/)) return r;
            SV_CUDA(E, cudaMemcpyAsync(ok2, d_ok, 8, cudaMemcpyDeviceToHost, st));
            SV_CUDA(E, cudaStreamSynchronize(st));
            E->streams = ok2[0] != 0;
            E->vbo_streams = ok2[1] != 0;
            if (E->streams) if (int r = detect_exclusive_blocks(E)) return r;
        }
    }
    mark("case re-ordering + CSC rebuild");
    drop(d_flags);
    // keep
    S.rcol = d_rcol; d_rcol = nullptr; E->dev_bytes += nnz * 4;
    S.rval = d_rval; d_rval = nullptr; if (S.rval) E->dev_bytes += nnz * 4;
    if (S.all_ones) { drop(d_x); S.cval = nullptr; } else { S.cval = d_x; d_x = nullptr; E->dev_bytes += nnz * 4; }
    S.uniformF = F;
    if (F > 0) { drop(d_rowptr); S.rowptr = nullptr; } else { S.rowptr = d_rowptr; d_rowptr = nullptr; E->dev_bytes += ((size_t)n + 1) * 8; }

    if (is_train && F == 2 && nnz) {
        if (dev_alloc(E, &S.cother, nnz)) return SVBFM_ERR_OOM;
        if (!S.all_ones && dev_alloc(E, &S.cother_val, nnz)) return SVBFM_ERR_OOM;
        k_other_of_entry<<<nblk(nnz), 256, 0, st>>>(S.colptr, ncols, nnz, S.crow, S.rcol, S.rval, S.cother, S.cother_val);
        if (E->d_rec_slot && (E->streams || E->vbo_streams)) {      // the first field's entries gather by record slot
            const Run& r0 = E->runs[0];
            k_map_slots<<<nblk(r0.nnz), 256, 0, st>>>(S.cother + S.h_colptr[r0.col_begin], r0.nnz, E->d_rec_slot);
            E->rec_rank = true;
        }
        SV_CUDA(E, cudaStreamSynchronize(st));
    }
    if (is_train && !E->rec_rank) { sv_free(E->d_rec_slot); E->d_rec_slot = nullptr; }
    if (is_train) set_side_views(E);
    mark("other-feature arrays");
    if (is_train && E->streams) {
        // ---- stream schedule: implicit tiles of 2^ts_shift entries per run; only their first column and the list of the
        // columns that span many tiles are materialised (kernels.cuh k_stream / k_combine_span)
        if (E->ts_auto) {
            // A pass is one warp per tile: 148 SMs x 7 CTAs x 4 warps = 4144 warps are resident at a time. 4096-entry tiles (the best
            // size at 200 M entries) leave most SMs idle below ~15 M entries per field (1 M ratings: 244 warps on
            // 31 SMs). Aim at two waves of warps on one GPU (ML-10M shape, profiles/r02_s_*: 256 / 512 / 1024 / 2048 entries 9.85 / 8.62 /
            // 8.32 / 8.84 ms per iteration; ML-1M shape: 128 / 256 / 512 entries 1.24 / 1.11 / 1.17 ms), at four on a shard of several
            // GPUs (the sizes the strong-scaling runs of DESIGN.md section 5 were measured with), with tiles between 256 and 4096 entries.
            const uint64_t want = (uint64_t)n / (4144ull * (E->world > 1 ? 4 : 2));
            E->ts_shift = 8;
            while (E->ts_shift < 12 && (2ull << E->ts_shift) <= want) E->ts_shift++;
        }
        const uint64_t TS = 1ull << E->ts_shift;
        sv_free(E->d_stile_col0); sv_free(E->d_span_heavy);
        E->d_stile_col0 = nullptr; E->d_span_heavy = nullptr;
        std::vector<uint32_t> heavy;
        for (int ri = 0; ri < 2; ri++) {
            Run& r = E->runs[ri];
            E->s_ntiles[ri] = (uint32_t)((r.nnz + TS - 1) / TS);
            uint64_t e0 = S.h_colptr[r.col_begin];
            size_t h0 = heavy.size();
            for (uint32_t j = r.col_begin; j < r.col_end; j++) {
                uint64_t b = S.h_colptr[j], e = S.h_colptr[j + 1];
                if (e > b && (e - 1 - e0) / TS - (b - e0) / TS > 8 /* SV_SPAN_LIGHT */) heavy.push_back(j);
            }
            E->span_heavy_n[ri] = (uint32_t)(heavy.size() - h0);
            r.tile_begin = r.tile_end = 0; r.heavy_begin = r.heavy_end = 0;
        }
        if (dev_alloc(E, &E->d_stile_col0, (size_t)E->s_ntiles[0] + E->s_ntiles[1])) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->d_span_heavy, heavy.size())) return SVBFM_ERR_OOM;
        SV_CUDA(E, cudaMemcpyAsync(E->d_span_heavy, heavy.data(), heavy.size() * 4, cudaMemcpyHostToDevice, st));
        set_side_views(E);
        if (int r = stream_tile_cols(E)) return r;
        SV_CUDA(E, cudaStreamSynchronize(st));
        E->n_tiles = 0; E->n_heavy = 0;
        mark("implicit tiles");
    } else if (is_train) {
        // ---- warp tiles (host). A tile = <= T consecutive CSC entries of ONE column; a column's tiles are consecutive
        // tile ids (fixed summation order). Runs whose case ids are not sequential ("gather runs") read e_i at random:
        // every miss moves a 128 B HBM line for 8 useful bytes (profiles/r01_v2_*). For their big columns the entry list
        // is cut at case-block boundaries (blocks of `block_cases` cases, e-block sized for L2) and the tiles are EXECUTED
        // block-major (exec_order), so that the warps in flight gather from one L2-resident block of e.
        uint32_t T = E->tile_entries;
        uint32_t block_cases = 4u << 20;
        if (const char* sb = getenv("SVBFM_BLOCK_CASES")) block_cases = (uint32_t)atol(sb);
        uint32_t NB = (block_cases && n) ? (uint32_t)(((uint64_t)n + block_cases - 1) / block_cases) : 1;
        uint64_t cut_min = 16;   // a column is cut when it has at least cut_min entries per block on average
        if (const char* sc = getenv("SVBFM_CUT_MIN")) cut_min = (uint64_t)atol(sc);
        std::vector<uint32_t> tile_col, tile_len, col_tile0((size_t)S.ncols_ext + 1), heavy, exec_order, tile_block;
        std::vector<uint64_t> tile_begin;
        // columns to cut: gather runs only, and only when there is more than one block
        std::vector<uint32_t> cut_cols;
        for (size_t ri = 0; ri < E->runs.size(); ri++) {
            const Run& r = E->runs[ri];
            bool sequential = (ri == 0 && E->rows_reordered);
            if (sequential || NB <= 1) continue;
            for (uint32_t j = r.col_begin; j < r.col_end; j++)
                if (S.h_colptr[j + 1] - S.h_colptr[j] >= cut_min * NB) cut_cols.push_back(j);
        }
        std::vector<uint64_t> cuts;   // [cut_cols][NB+1] absolute entry positions
        if (!cut_cols.empty()) {
            uint32_t* d_cc = nullptr; uint64_t* d_cuts = nullptr;
            SV_CUDA(E, sv_malloc((void**)&d_cc, cut_cols.size() * 4));
            SV_CUDA(E, sv_malloc((void**)&d_cuts, cut_cols.size() * (size_t)(NB + 1) * 8));
            SV_CUDA(E, cudaMemcpyAsync(d_cc, cut_cols.data(), cut_cols.size() * 4, cudaMemcpyHostToDevice, st));
            dim3 grid((unsigned)cut_cols.size(), (NB + 1 + 63) / 64);
            k_block_cuts<<<grid, 64, 0, st>>>(d_cc, S.colptr, S.crow, block_cases, NB, d_cuts);
            cuts.resize(cut_cols.size() * (size_t)(NB + 1));
            SV_CUDA(E, cudaMemcpyAsync(cuts.data(), d_cuts, cuts.size() * 8, cudaMemcpyDeviceToHost, st));
            SV_CUDA(E, cudaStreamSynchronize(st));
            sv_free(d_cc); sv_free(d_cuts);
        }
        size_t cc = 0;
        for (auto& r : E->runs) {
            r.tile_begin = (uint32_t)tile_col.size();
            r.heavy_begin = (uint32_t)heavy.size();
            for (uint32_t j = r.col_begin; j < r.col_end; j++) {
                col_tile0[j] = (uint32_t)tile_col.size();
                uint64_t b = S.h_colptr[j], e = S.h_colptr[j + 1];
                if (cc < cut_cols.size() && cut_cols[cc] == j) {
                    const uint64_t* cp = &cuts[cc * (size_t)(NB + 1)];
                    for (uint32_t k = 0; k < NB; k++)
                        for (uint64_t p = cp[k]; p < cp[k + 1]; p += T) {
                            tile_col.push_back(j); tile_begin.push_back(p); tile_len.push_back((uint32_t)std::min<uint64_t>(T, cp[k + 1] - p));
                            tile_block.push_back(k);
                        }
                    cc++;
                } else {
                    for (uint64_t p = b; p < e; p += T) {
                        tile_col.push_back(j); tile_begin.push_back(p); tile_len.push_back((uint32_t)std::min<uint64_t>(T, e - p));
                        tile_block.push_back(NB);   // not cut: executed after the block-major part, in column order
                    }
                }
                if (tile_col.size() - col_tile0[j] > 8) heavy.push_back(j);
            }
            r.tile_end = (uint32_t)tile_col.size();
            r.heavy_end = (uint32_t)heavy.size();
            // execution order of this run: stable by block id
            std::vector<uint32_t> ord(r.tile_end - r.tile_begin);
            for (uint32_t t = 0; t < ord.size(); t++) ord[t] = r.tile_begin + t;
            std::stable_sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) { return tile_block[x] < tile_block[y]; });
            exec_order.insert(exec_order.end(), ord.begin(), ord.end());
        }
        col_tile0[S.ncols_ext] = (uint32_t)tile_col.size();
        E->n_tiles = (uint32_t)tile_col.size();
        E->n_heavy = (uint32_t)heavy.size();
        sv_free(E->d_tile_col); sv_free(E->d_tile_begin); sv_free(E->d_col_tile0); sv_free(E->d_heavy_cols);
        sv_free(E->d_tile_len); sv_free(E->d_exec_order);
        E->d_tile_col = nullptr; E->d_tile_begin = nullptr; E->d_col_tile0 = nullptr; E->d_heavy_cols = nullptr;
        E->d_tile_len = nullptr; E->d_exec_order = nullptr;
        if (dev_alloc(E, &E->d_tile_col, tile_col.size())) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->d_tile_begin, tile_begin.size())) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->d_tile_len, tile_len.size())) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->d_exec_order, exec_order.size())) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->d_col_tile0, col_tile0.size())) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->d_heavy_cols, heavy.size())) return SVBFM_ERR_OOM;
        SV_CUDA(E, cudaMemcpyAsync(E->d_tile_col, tile_col.data(), tile_col.size() * 4, cudaMemcpyHostToDevice, st));
        SV_CUDA(E, cudaMemcpyAsync(E->d_tile_begin, tile_begin.data(), tile_begin.size() * 8, cudaMemcpyHostToDevice, st));
        SV_CUDA(E, cudaMemcpyAsync(E->d_tile_len, tile_len.data(), tile_len.size() * 4, cudaMemcpyHostToDevice, st));
        SV_CUDA(E, cudaMemcpyAsync(E->d_exec_order, exec_order.data(), exec_order.size() * 4, cudaMemcpyHostToDevice, st));
        SV_CUDA(E, cudaMemcpyAsync(E->d_col_tile0, col_tile0.data(), col_tile0.size() * 4, cudaMemcpyHostToDevice, st));
        SV_CUDA(E, cudaMemcpyAsync(E->d_heavy_cols, heavy.data(), heavy.size() * 4, cudaMemcpyHostToDevice, st));
        SV_CUDA(E, cudaStreamSynchronize(st));
    }
    mark("tiles + exec order");
    SV_CUDA(E, cudaGetLastError());
    return 0;
}

// ---- device transpose: CSR (cases x features, the reference's `data`) -> CSC (features x cases, `data_t`) ------------------
// Replaces Data::create_data_t (Data.h:457-509) and tools/transpose.cpp:91-162 on the device: entry -> case id (search over the
// row pointer), one STABLE radix sort of the entries by feature id (cases stay ascending inside a feature, duplicates keep their
// order: the same bytes as the reference's counting transpose), column pointer by a search over the sorted keys.
static __global__ void k_row_of_entry(const uint64_t* __restrict__ rowptr, uint32_t n, uint64_t nnz, uint32_t* __restrict__ out) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    uint32_t lo = 0, hi = n;          // rowptr[lo] <= p < rowptr[hi]
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (rowptr[mid] <= p) lo = mid; else hi = mid;
    }
    out[p] = lo;
}
static __global__ void k_check_below(const uint32_t* __restrict__ v, uint64_t n, uint32_t bound, uint32_t* flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && v[i] >= bound) *flag = 1;
}

// d_rowptr / d_col / d_x: device CSR. Outputs (device, sv_malloc'ed, owned by the caller): column pointer [ncols + 1], case ids, values
int transpose_on_device(Engine* E, cudaStream_t st, uint32_t n, uint32_t ncols, uint64_t nnz, const uint64_t* d_rowptr, const uint32_t* d_col, const float* d_x,
                        uint64_t** d_colptr, uint32_t** d_case, float** d_xt) {
    *d_colptr = nullptr; *d_case = nullptr; *d_xt = nullptr;
    uint32_t *d_row = nullptr, *d_idx = nullptr, *d_skeys = nullptr, *d_sidx = nullptr, *d_flag = nullptr;
    struct Temps { std::vector<void**> v; ~Temps() { for (void** p : v) { sv_free(*p); *p = nullptr; } } } temps;
    for (void** p : {(void**)&d_row, (void**)&d_idx, (void**)&d_skeys, (void**)&d_sidx, (void**)&d_flag}) temps.v.push_back(p);
    cudaStream_t keep = E->stream;
    struct StreamGuard { Engine* E; cudaStream_t s; ~StreamGuard() { E->stream = s; } } guard{E, keep};
    E->stream = st;                                    // sort_pairs works on the engine's stream
    SV_CUDA(E, sv_malloc((void**)d_colptr, ((size_t)ncols + 1) * 8));
    SV_CUDA(E, sv_malloc((void**)d_case, std::max<uint64_t>(nnz, 1) * 4));
    SV_CUDA(E, sv_malloc((void**)d_xt, std::max<uint64_t>(nnz, 1) * 4));
    SV_CUDA(E, sv_malloc((void**)&d_flag, 4));
    SV_CUDA(E, cudaMemsetAsync(d_flag, 0, 4, st));
    if (nnz) {
        SV_CUDA(E, sv_malloc((void**)&d_row, nnz * 4));
        SV_CUDA(E, sv_malloc((void**)&d_idx, nnz * 4));
        k_row_of_entry<<<nblk(nnz), 256, 0, st>>>(d_rowptr, n, nnz, d_row);
        k_check_below<<<nblk(nnz), 256, 0, st>>>(d_col, nnz, ncols, d_flag);
        k_iota<<<nblk(nnz), 256, 0, st>>>(d_idx, nnz);
        uint32_t bad = 0;
        SV_CUDA(E, cudaMemcpyAsync(&bad, d_flag, 4, cudaMemcpyDeviceToHost, st));
        SV_CUDA(E, cudaStreamSynchronize(st));
        if (bad) return fail(E, SVBFM_ERR_ARG, "feature id out of range in transpose");
        if (int rc = sort_pairs(E, d_col, d_idx, nnz, std::max<uint32_t>(ncols, 1), &d_skeys, &d_sidx)) return rc;
        k_gather_u32<<<nblk(nnz), 256, 0, st>>>(d_row, d_sidx, nnz, *d_case);
        if (d_x) k_gather_f32<<<nblk(nnz), 256, 0, st>>>(d_x, d_sidx, nnz, *d_xt);
        else { sv_free(*d_xt); *d_xt = nullptr; }                                  // every value is 1
        k_rowptr_from_sorted<<<nblk((uint64_t)ncols + 1), 256, 0, st>>>(d_skeys, nnz, ncols, *d_colptr);
    } else SV_CUDA(E, cudaMemsetAsync(*d_colptr, 0, ((size_t)ncols + 1) * 8, st));
    SV_CUDA(E, cudaStreamSynchronize(st));
    SV_CUDA(E, cudaGetLastError());
    return 0;
}

// ---- cross shards: the second residual copy on its own shard (svbfm_set_csc(SVBFM_TRAIN_SECOND); svbfm_internal.h Engine::xs) ----
void set_side_views(Engine* E) {
    const DevSplit& S = E->tr;
    for (int side = 0; side < 2; side++) {
        Engine::SideView& v = E->side[side];
        v = Engine::SideView();
        if (E->runs.size() != 2) continue;
        const Run& r = E->runs[side];
        const uint64_t e0 = S.h_colptr.empty() ? 0 : S.h_colptr[r.col_begin];
        if (side == 1 && E->xs) {
            v.colptr = E->sec.colptr; v.entry0 = 0; v.n = E->sec.n; v.oc = E->sec.oc;
        } else {
            v.colptr = S.colptr; v.entry0 = e0; v.n = (uint32_t)r.nnz;
            v.oc = S.cother; v.xv = S.cval; v.xo = S.cother_val;
        }
    }
}

void free_second(Engine* E) {
    for (int r = 0; r < 16; r++) {
        if (E->peer_base[r] && E->peer_base[r] != E->d_xipc) cudaIpcCloseMemHandle(E->peer_base[r]);
        E->peer_base[r] = nullptr;
    }
    if (E->d_xipc) { cudaStreamSynchronize(E->stream); cudaFree(E->d_xipc); }
    E->d_xipc = nullptr; E->d_xstage = nullptr; E->xstage_cap = 0; E->p2p = false; E->xs_epoch = 0;
    sv_free(E->d_col_of_slot); E->d_col_of_slot = nullptr;
    sv_free(E->sec.colptr); sv_free(E->sec.oc); sv_free(E->sec.rcol); sv_free(E->sec.y);
    E->sec = Engine::SecondShard();
    E->xs = false;
    E->blk1.clear();
}

// common record slots of the cross-shard layout: field f, column j -> base[f] + (j - first column of the field): the blocks of a
// field are contiguous column ranges, so every rank's block is one contiguous run of slots (no padding). Second field with
// `ranked`: the columns OUTSIDE this rank's block are zeroed here, the own block is ordered by popularity (k_slot_scatter) and the
// ranks' pieces are summed.
static __global__ void k_xs_slots(uint32_t c0, uint32_t c1, uint32_t base, int zero_only, uint32_t* __restrict__ slot) {
    uint32_t j = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c1) return;
    slot[j] = zero_only ? 0u : base + (j - c0);
}
static __global__ void k_invert_slots(const uint32_t* __restrict__ slot, uint32_t c0, uint32_t c1, uint32_t* __restrict__ col_of_slot) {
    uint32_t j = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j < c1) col_of_slot[slot[j]] = j;
}
// oc[p] <- new_slot[column of oc[p]] (old: the entries hold old slots when col_of_old is set, column ids otherwise)
static __global__ void k_remap_oc(uint32_t* __restrict__ oc, uint64_t n, const uint32_t* __restrict__ col_of_old, const uint32_t* __restrict__ new_slot) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t c = oc[p];
    if (col_of_old) c = col_of_old[c];
    oc[p] = new_slot[c];
}
// first-field entries of the second shard: user_of_case[case] = column
static __global__ void k_sec_users(const uint32_t* __restrict__ case_id, const uint32_t* __restrict__ colof, uint32_t n, uint32_t* __restrict__ user_of_case,
                                   uint32_t* flags) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t c = case_id[p];
    if (c >= n) { flags[1] = 1; return; }
    if (atomicExch(&user_of_case[c], colof[p]) != 0xffffffffu) flags[2] = 1;      // a case twice in the first field
}
// second-field entries in the caller's order: first-field column of every entry (sort key) + the checks
static __global__ void k_sec_keys(const uint32_t* __restrict__ case_id, uint32_t n, const uint32_t* __restrict__ user_of_case, uint32_t* __restrict__ seen,
                                  uint32_t c0, uint32_t* __restrict__ key, uint32_t* __restrict__ val, uint32_t* flags) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    uint32_t c = case_id[q];
    key[q] = 0; val[q] = q;
    if (c >= n) { flags[1] = 1; return; }
    if (atomicExch(&seen[c], 1u) != 0u) flags[2] = 1;                               // a case twice in the second field
    uint32_t u = user_of_case[c];
    if (u == 0xffffffffu) { flags[2] = 1; u = c0; }
    key[q] = u - c0;
}
// the shard's entry order: entry p is the caller's entry perm[p]: other column, target, {user, item} pair
static __global__ void k_sec_entries(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ case_id, const uint32_t* __restrict__ colof, uint32_t n,
                                     const uint32_t* __restrict__ user_of_case, const float* __restrict__ target, uint32_t* __restrict__ oc,
                                     uint32_t* __restrict__ rcol, float* __restrict__ y) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t q = perm[p], c = case_id[q];
    const uint32_t u = user_of_case[c];
    oc[p] = u;
    rcol[2 * (size_t)p] = u; rcol[2 * (size_t)p + 1] = colof[q];
    y[p] = target[c];
}

int ingest_second(Engine* E, uint32_t n, uint32_t ncols, const uint64_t* colptr, const uint32_t* case_id, const float* x, const float* target) {
    cudaStream_t st = E->stream;
    DevSplit& S = E->tr;
    free_second(E);
    // ---- what must hold on every rank alike (no data-dependent exits before the first collective)
    if (E->world <= 1) return fail(E, SVBFM_ERR_ARG, "set_csc(TRAIN_SECOND): needs a communicator (svbfm_comm_init)");
    if (!E->d_e) return fail(E, SVBFM_ERR_ARG, "set_csc(TRAIN_SECOND): the train split must be set first");
    if (E->cfg.method == SVBFM_VB_ONLINE || E->cfg.task != 0) return fail(E, SVBFM_ERR_ARG, "set_csc(TRAIN_SECOND): vb and mcmc regression only");
    if (!E->streams || !E->excl0 || E->runs.size() != 2)
        return fail(E, SVBFM_ERR_ARG, "set_csc(TRAIN_SECOND): the train split must be two complete one-hot fields sharded by blocks of the first field");
    const Run r0 = E->runs[0], r1 = E->runs[1];
    // ---- local validation; the verdict is agreed on by all ranks before anything else is exchanged
    std::string why;
    if (ncols > E->D || ncols < r1.col_begin) why = "num_cols out of range";
    else if (colptr[0] != 0) why = "colptr[0] != 0";
    else {
        for (uint32_t j = 0; j < ncols && why.empty(); j++) if (colptr[j + 1] < colptr[j]) why = "colptr not monotone";
    }
    const uint32_t nc_ext = S.ncols_ext;
    uint64_t off1 = 0;
    if (why.empty()) {
        const uint64_t nnz = colptr[ncols];
        off1 = colptr[std::min(r1.col_begin, ncols)];
        if (nnz != 2ull * n || off1 != n) why = "every case needs exactly one entry in each of the two fields";
        else if (n && (!case_id || !target)) why = "null entry arrays";
        else if (ncols > nc_ext) why = "more columns than the train split";
    }
    uint32_t* d_flags = nullptr;      // [0] x != 1   [1] case id out of range   [2] not one entry per case and field
    uint32_t *d_case = nullptr, *d_colof = nullptr, *d_user = nullptr, *d_seen = nullptr;
    uint64_t* d_cp = nullptr;
    float *d_x = nullptr, *d_t = nullptr;
    struct Temps { std::vector<void**> v; ~Temps() { for (void** p : v) { sv_free(*p); *p = nullptr; } } } temps;
    uint32_t *d_k = nullptr, *d_v = nullptr, *d_k1 = nullptr, *d_v1 = nullptr, *d_k2 = nullptr, *d_v2 = nullptr;      // sort scratch
    uint32_t* d_new = nullptr;
    double* d_tot = nullptr;
    for (void** p : {(void**)&d_flags, (void**)&d_case, (void**)&d_colof, (void**)&d_user, (void**)&d_seen, (void**)&d_cp, (void**)&d_x, (void**)&d_t,
                     (void**)&d_k, (void**)&d_v, (void**)&d_k1, (void**)&d_v1, (void**)&d_k2, (void**)&d_v2, (void**)&d_new, (void**)&d_tot}) temps.v.push_back(p);
    SV_CUDA(E, sv_malloc((void**)&d_flags, 8 * 4));
    SV_CUDA(E, cudaMemsetAsync(d_flags, 0, 8 * 4, st));
    uint32_t h_flags[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (why.empty() && n) {
        const uint64_t nnz = 2ull * n;
        SV_CUDA(E, sv_malloc((void**)&d_case, nnz * 4));
        SV_CUDA(E, sv_malloc((void**)&d_colof, nnz * 4));
        if (x) SV_CUDA(E, sv_malloc((void**)&d_x, nnz * 4));
        SV_CUDA(E, sv_malloc((void**)&d_t, (size_t)n * 4));
        SV_CUDA(E, sv_malloc((void**)&d_cp, ((size_t)ncols + 1) * 8));
        SV_CUDA(E, sv_malloc((void**)&d_user, (size_t)n * 4));
        SV_CUDA(E, sv_malloc((void**)&d_seen, (size_t)n * 4));
        SV_CUDA(E, cudaMemcpyAsync(d_cp, colptr, ((size_t)ncols + 1) * 8, cudaMemcpyHostToDevice, st));
        SV_CUDA(E, cudaMemcpyAsync(d_case, case_id, nnz * 4, cudaMemcpyHostToDevice, st));
        if (x) SV_CUDA(E, cudaMemcpyAsync(d_x, x, nnz * 4, cudaMemcpyHostToDevice, st));
        SV_CUDA(E, cudaMemcpyAsync(d_t, target, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        SV_CUDA(E, cudaMemsetAsync(d_user, 0xff, (size_t)n * 4, st));
        SV_CUDA(E, cudaMemsetAsync(d_seen, 0, (size_t)n * 4, st));
        if (x) k_any_not_one<<<std::min<unsigned>(nblk(nnz), 148 * 16), 256, 0, st>>>(d_x, nnz, d_flags + 0);
        k_col_of_entry<<<nblk(nnz), 256, 0, st>>>(d_cp, ncols, nnz, d_colof);
        if (dev_alloc(E, &E->sec.oc, n)) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->sec.rcol, (size_t)n * 2)) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->sec.y, n)) return SVBFM_ERR_OOM;
        k_sec_users<<<nblk(n), 256, 0, st>>>(d_case, d_colof, n, d_user, d_flags);
        // The copy's entry order: by second-field column (as handed over), inside a column ascending by FIRST-field column, so that
        // neighbouring lanes of the second field's pass gather neighbouring records (the caller's case order inside a column is
        // arbitrary; the single-GPU layout gets the same from its case re-ordering). Two stable sorts: by first-field column, then
        // by second-field column.
        SV_CUDA(E, sv_malloc((void**)&d_k, (size_t)n * 4));
        SV_CUDA(E, sv_malloc((void**)&d_v, (size_t)n * 4));
        k_sec_keys<<<nblk(n), 256, 0, st>>>(d_case + off1, n, d_user, d_seen, r0.col_begin, d_k, d_v, d_flags);
        SV_CUDA(E, cudaMemcpyAsync(h_flags, d_flags, 8 * 4, cudaMemcpyDeviceToHost, st));
        SV_CUDA(E, cudaStreamSynchronize(st));
        if (h_flags[0]) why = "x != 1";
        else if (h_flags[1]) why = "case id out of range";
        else if (h_flags[2]) why = "every case needs exactly one entry in each of the two fields";
        if (why.empty()) {
            if (int rc = sort_pairs(E, d_k, d_v, n, std::max<uint32_t>(r0.col_end - r0.col_begin, 1), &d_k1, &d_v1)) return rc;
            k_gather_u32<<<nblk(n), 256, 0, st>>>(d_colof + off1, d_v1, n, d_k);          // second-field column of the entries in that order
            if (int rc = sort_pairs(E, d_k, d_v1, n, std::max<uint32_t>(ncols, 1), &d_k2, &d_v2)) return rc;
            k_sec_entries<<<nblk(n), 256, 0, st>>>(d_v2, d_case + off1, d_colof + off1, n, d_user, d_t, E->sec.oc, E->sec.rcol, E->sec.y);
            SV_CUDA(E, cudaStreamSynchronize(st));
        }
    }
    // entry pointers of the second field's columns inside the shard, indexed by global column id
    E->sec.n = why.empty() ? n : 0;
    E->sec.h_colptr.assign((size_t)nc_ext + 1, 0);
    if (why.empty())
        for (uint32_t j = r1.col_begin; j <= nc_ext; j++) E->sec.h_colptr[j] = colptr[std::min(j, ncols)] - off1;
    // ---- the ranks' verdicts, the blocks of the second field and the global number of cases: collectives from here on
    {
        uint32_t bad = why.empty() ? 0u : 1u;
        SV_CUDA(E, cudaMemcpyAsync(d_flags + 4, &bad, 4, cudaMemcpyHostToDevice, st));
        if (int rc = allreduce(E, d_flags + 4, 1, 3 /*ncclUint32*/, 2 /*ncclMax*/)) return rc;
        SV_CUDA(E, cudaMemcpyAsync(&bad, d_flags + 4, 4, cudaMemcpyDeviceToHost, st));
        SV_CUDA(E, cudaStreamSynchronize(st));
        if (bad) { free_second(E); return fail(E, SVBFM_ERR_ARG, "set_csc(TRAIN_SECOND): " + (why.empty() ? std::string("another rank rejected its shard") : why)); }
    }
    bool excl1 = false;
    if (int rc = detect_blocks(E, r1, E->sec.h_colptr, E->blk1, excl1)) return rc;
    double tot = (double)n;
    SV_CUDA(E, sv_malloc((void**)&d_tot, 8));
    SV_CUDA(E, cudaMemcpyAsync(d_tot, &tot, 8, cudaMemcpyHostToDevice, st));
    if (int rc = allreduce(E, d_tot, 1, 8 /*ncclDouble*/, 0 /*ncclSum*/)) return rc;
    SV_CUDA(E, cudaMemcpyAsync(&tot, d_tot, 8, cudaMemcpyDeviceToHost, st));
    SV_CUDA(E, cudaStreamSynchronize(st));
    if (!excl1 || (uint64_t)tot != E->n_total) {
        free_second(E);
        return fail(E, SVBFM_ERR_ARG, !excl1 ? "set_csc(TRAIN_SECOND): the ranks' shards must cover disjoint, rank-ordered blocks of the second field's columns"
                                             : "set_csc(TRAIN_SECOND): the second shards hold " + std::to_string((uint64_t)tot) + " cases, the train split " +
                                                   std::to_string(E->n_total) + " (this rank: " + std::to_string(n) + " and " + std::to_string(S.n) + ")");
    }
    if (dev_alloc(E, &E->sec.colptr, (size_t)nc_ext + 1)) return SVBFM_ERR_OOM;
    SV_CUDA(E, cudaMemcpyAsync(E->sec.colptr, E->sec.h_colptr.data(), ((size_t)nc_ext + 1) * 8, cudaMemcpyHostToDevice, st));
    if (!E->sec.oc) {      // a rank without cases
        if (dev_alloc(E, &E->sec.oc, 1)) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->sec.rcol, 2)) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->sec.y, 1)) return SVBFM_ERR_OOM;
    }

    // ---- common record slots: the fields one after the other in column order, so that a rank's block is a contiguous run of slots.
    // Second field: inside a block by popularity (descending column length, known to the block's owner)
    const uint32_t nc0 = r0.col_end - r0.col_begin, nc1 = r1.col_end - r1.col_begin;
    E->slot_base[0] = 0; E->slot_base[1] = nc0;
    E->slot_max[0] = nc0; E->slot_max[1] = nc1;      // slots of the field
    const size_t total_slots = (size_t)nc0 + nc1;
    SV_CUDA(E, sv_malloc((void**)&d_new, (size_t)E->D * 4));
    SV_CUDA(E, cudaMemsetAsync(d_new, 0, (size_t)E->D * 4, st));
    if (nc0) k_xs_slots<<<nblk(nc0), 256, 0, st>>>(r0.col_begin, r0.col_end, E->slot_base[0], 0, d_new);
    const bool ranked = E->want_rec_rank;
    if (nc1) k_xs_slots<<<nblk(nc1), 256, 0, st>>>(r1.col_begin, r1.col_end, E->slot_base[1], ranked ? 1 : 0, d_new);
    if (ranked) {
        const uint32_t c0 = E->blk1[E->rank], nb = E->blk1[E->rank + 1] - c0;
        if (nb) {
            uint32_t *d_sk = nullptr, *d_sv = nullptr, *d_ks = nullptr, *d_vs = nullptr;
            SV_CUDA(E, sv_malloc((void**)&d_sk, (size_t)nb * 4));
            SV_CUDA(E, sv_malloc((void**)&d_sv, (size_t)nb * 4));
            k_slot_keys<<<nblk(nb), 256, 0, st>>>(E->sec.colptr, c0, nb, d_sk, d_sv);
            int rc = sort_pairs(E, d_sk, d_sv, nb, 1ull << 32, &d_ks, &d_vs);
            if (!rc) { k_slot_scatter<<<nblk(nb), 256, 0, st>>>(d_vs, E->slot_base[1] + (c0 - r1.col_begin), nb, d_new); cudaStreamSynchronize(st); }
            sv_free(d_sk); sv_free(d_sv); sv_free(d_ks); sv_free(d_vs);
            if (rc) return rc;
        }
        if (int rc = allreduce(E, d_new + r1.col_begin, nc1, 3 /*ncclUint32*/, 0 /*ncclSum*/)) return rc;
    }
    // the first copy's entries gather second-field records: old slots (or column ids) -> common slots
    {
        uint32_t* d_inv = nullptr;
        if (E->rec_rank) {
            SV_CUDA(E, sv_malloc((void**)&d_inv, (size_t)E->D * 4));
            k_invert_slots<<<nblk(nc1), 256, 0, st>>>(E->d_rec_slot, r1.col_begin, r1.col_end, d_inv);
        }
        if (r0.nnz) k_remap_oc<<<nblk(r0.nnz), 256, 0, st>>>(S.cother + S.h_colptr[r0.col_begin], r0.nnz, d_inv, d_new);
        if (n) k_remap_oc<<<nblk(n), 256, 0, st>>>(E->sec.oc, n, nullptr, d_new);
        SV_CUDA(E, cudaStreamSynchronize(st));
        sv_free(d_inv);
    }
    sv_free(E->d_rec_slot);
    E->d_rec_slot = d_new; d_new = nullptr;
    E->rec_rank = true;
    sv_free(E->d_col_of_slot); E->d_col_of_slot = nullptr;
    if (dev_alloc(E, &E->d_col_of_slot, total_slots)) return SVBFM_ERR_OOM;
    SV_CUDA(E, cudaMemsetAsync(E->d_col_of_slot, 0xff, total_slots * 4, st));
    if (nc0) k_invert_slots<<<nblk(nc0), 256, 0, st>>>(E->d_rec_slot, r0.col_begin, r0.col_end, E->d_col_of_slot);
    if (nc1) k_invert_slots<<<nblk(nc1), 256, 0, st>>>(E->d_rec_slot, r1.col_begin, r1.col_end, E->d_col_of_slot);
    if (E->cpack_cap < total_slots) {
        sv_free(E->d_cpack); E->d_cpack = nullptr;
        if (sv_malloc((void**)&E->d_cpack, total_slots * 32 /* sizeof(ColPack) */) != cudaSuccess) return fail(E, SVBFM_ERR_OOM, "cudaMalloc: record slots");
        E->cpack_cap = total_slots;
    }
    {
        // a stage holds a field in slot order; without peer mappings (fallback) stage 0 holds [world] blocks padded to the largest one
        uint32_t mx = 1;
        for (int r = 0; r < E->world; r++) mx = std::max(mx, std::max(E->blk[r + 1] - E->blk[r], E->blk1[r + 1] - E->blk1[r]));
        if (int rc = setup_exchange(E, std::max<size_t>(std::max<size_t>(nc0, nc1), (size_t)E->world * mx))) return rc;
    }
    // ---- tiles of the second side on the new shard (the first side keeps its own), second residual copy, tile sums
    const uint64_t TS = 1ull << E->ts_shift;
    std::vector<uint32_t> heavy;
    for (int ri = 0; ri < 2; ri++) {
        const Run& r = E->runs[ri];
        const std::vector<uint64_t>& cp = ri ? E->sec.h_colptr : S.h_colptr;
        const uint64_t e0 = cp[r.col_begin], nn = cp[r.col_end] - e0;
        E->s_ntiles[ri] = (uint32_t)((nn + TS - 1) / TS);
        const size_t h0 = heavy.size();
        for (uint32_t j = r.col_begin; j < r.col_end; j++) {
            uint64_t b = cp[j], e = cp[j + 1];
            if (e > b && (e - 1 - e0) / TS - (b - e0) / TS > 8 /* SV_SPAN_LIGHT */) heavy.push_back(j);
        }
        E->span_heavy_n[ri] = (uint32_t)(heavy.size() - h0);
    }
    sv_free(E->d_stile_col0); sv_free(E->d_span_heavy); sv_free(E->d_e2); sv_free(E->d_partial);
    E->d_stile_col0 = nullptr; E->d_span_heavy = nullptr; E->d_e2 = nullptr; E->d_partial = nullptr;
    if (dev_alloc(E, &E->d_stile_col0, (size_t)E->s_ntiles[0] + E->s_ntiles[1])) return SVBFM_ERR_OOM;
    if (dev_alloc(E, &E->d_span_heavy, heavy.size())) return SVBFM_ERR_OOM;
    if (dev_alloc(E, &E->d_e2, n)) return SVBFM_ERR_OOM;
    if (dev_alloc(E, &E->d_partial, ((size_t)E->s_ntiles[0] + E->s_ntiles[1]) * 8)) return SVBFM_ERR_OOM;
    SV_CUDA(E, cudaMemcpyAsync(E->d_span_heavy, heavy.data(), heavy.size() * 4, cudaMemcpyHostToDevice, st));
    E->xs = true;
    set_side_views(E);
    if (int rc = stream_tile_cols(E)) return rc;
    SV_CUDA(E, cudaStreamSynchronize(st));
    SV_CUDA(E, cudaGetLastError());
    return 0;
}

// ---- vb_online on the stream schedule: batch index lists of one epoch ------------------------------------------------
// keys for the stable sort by batch: run 0 entry q is case q; run 1 entry p is case crow1[p]
static __global__ void k_vbo_keys(const uint16_t* __restrict__ rbatch, const uint32_t* __restrict__ crow1 /*null: run 0*/, uint32_t n,
                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    keys[q] = rbatch[crow1 ? crow1[q] : q];
    vals[q] = q;
}
// cnt[b * nc + (column of the case in the run - c0)] += 1; `which` selects the case's feature (0: first field, 1: second)
static __global__ void k_vbo_hist(const uint16_t* __restrict__ rbatch, const uint32_t* __restrict__ rcol, uint32_t n, int which, uint32_t c0, uint32_t nc,
                                  unsigned long long* __restrict__ cnt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t j = rcol[2 * (size_t)i + which];
    atomicAdd(&cnt[(size_t)rbatch[i] * nc + (j - c0)], 1ull);          // integer atomics: order-independent
}

// flag[b * nc + c] = column c0 + c has entries in batch b (sharded: on any rank); flag[n] = 0 closes the scan
static __global__ void k_vbo_nonempty(const unsigned long long* __restrict__ cnt, const uint32_t* __restrict__ gcnt, size_t n, uint32_t* __restrict__ flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (gcnt ? gcnt[i] != 0u : cnt[i] != 0ull) ? 1u : 0u;
    if (i == n) flag[i] = 0u;
}
// pos = exclusive scan of the flags: the flattened indices of the non-empty (batch, column) pairs in order
static __global__ void k_vbo_compact(const uint32_t* __restrict__ pos, size_t n, uint32_t* __restrict__ list) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && pos[i + 1] != pos[i]) list[pos[i]] = (uint32_t)i;
}
// flattened (batch, column) indices -> column ids; off[b] = first list entry of batch b (lower bound of b * nc)
static __global__ void k_vbo_clist_finish(uint32_t* __restrict__ list, uint32_t n_list, uint32_t nc, uint32_t c0, uint32_t num_batch, uint32_t* __restrict__ off) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= num_batch) {
        const uint64_t key = (uint64_t)t * nc;
        uint32_t lo = 0, hi = n_list;
        while (lo < hi) { uint32_t mid = lo + (hi - lo) / 2; if ((uint64_t)list[mid] < key) lo = mid + 1; else hi = mid; }
        off[t] = lo;
    }
}
static __global__ void k_vbo_clist_cols(uint32_t* __restrict__ list, uint32_t n_list, uint32_t nc, uint32_t c0) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_list) list[t] = c0 + list[t] % nc;
}
static __global__ void k_u64_to_u32(const unsigned long long* __restrict__ in, size_t n, uint32_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)in[i];
}

// For each field: idx = the run's entries stably sorted by batch (inside a batch they keep the run's column order), and
// colptr[b * nc + c] = first position in idx of column c0 + c of batch b (histogram + exclusive scan; the flattened scan
// makes colptr[(b + 1) * nc] the end of batch b). Stale after the next call.
int vbo_stream_prepare(Engine* E, uint32_t num_batch) {
    cudaStream_t st = E->stream;
    const DevSplit& S = E->tr;
    const uint32_t n = S.n;
    E->vbo_off.assign((size_t)num_batch + 1, 0);
    {
        std::vector<unsigned long long> cnt(num_batch);
        SV_CUDA(E, cudaMemcpyAsync(cnt.data(), E->d_batch_cnt, (size_t)num_batch * 8, cudaMemcpyDeviceToHost, st));
        SV_CUDA(E, cudaStreamSynchronize(st));
        for (uint32_t b = 0; b < num_batch; b++) E->vbo_off[b + 1] = E->vbo_off[b] + cnt[b];
    }
    // tiles of a batch pass: about 4 warps per scheduler on 148 SMs, between 64 and the regular tile size (or what the caller asked for)
    if (E->cfg.tile_entries || getenv("SVBFM_TILE_ENTRIES")) E->vbo_ts_shift = E->ts_shift;
    else {
        uint64_t want = (n / std::max<uint32_t>(num_batch, 1)) / (148 * 32);
        E->vbo_ts_shift = 6;
        while (E->vbo_ts_shift < E->ts_shift && (2ull << E->vbo_ts_shift) <= want) E->vbo_ts_shift++;
    }
    uint32_t max_tiles = 1;
    for (uint32_t b = 0; b < num_batch; b++)
        max_tiles = std::max<uint32_t>(max_tiles, (uint32_t)(((E->vbo_off[b + 1] - E->vbo_off[b]) >> E->vbo_ts_shift) + 1));
    if (max_tiles > E->vbo_max_tiles) {
        sv_free(E->d_vbo_tile_col0); sv_free(E->d_vbo_partial);
        E->d_vbo_tile_col0 = nullptr; E->d_vbo_partial = nullptr;
        if (dev_alloc(E, &E->d_vbo_tile_col0, (size_t)max_tiles * 2)) return SVBFM_ERR_OOM;
        if (dev_alloc(E, &E->d_vbo_partial, (size_t)max_tiles * 2 * 8)) return SVBFM_ERR_OOM;
        E->vbo_max_tiles = max_tiles;
    }
    // the batch in flight as contiguous streams (engine: k_vbo_pack)
    uint64_t max_batch = 1;
    for (uint32_t b = 0; b < num_batch; b++) max_batch = std::max<uint64_t>(max_batch, E->vbo_off[b + 1] - E->vbo_off[b]);
    const bool with_x = E->side[0].xv != nullptr;
    if (E->vbo_pack && (max_batch > E->vbo_batch_cap || (with_x && !E->d_vbo_xb[0][0]))) {
        for (int ri = 0; ri < 2; ri++) {
            sv_free(E->d_vbo_eb[ri]); sv_free(E->d_vbo_ocb[ri]); sv_free(E->d_vbo_ownb[ri]); sv_free(E->d_vbo_xb[ri][0]); sv_free(E->d_vbo_xb[ri][1]);
            E->d_vbo_eb[ri] = nullptr; E->d_vbo_ocb[ri] = nullptr; E->d_vbo_ownb[ri] = nullptr; E->d_vbo_xb[ri][0] = nullptr; E->d_vbo_xb[ri][1] = nullptr;
        }
        E->vbo_batch_cap = 0;
        for (int ri = 0; ri < 2; ri++) {
            if (dev_alloc(E, &E->d_vbo_eb[ri], max_batch)) return SVBFM_ERR_OOM;
            if (dev_alloc(E, &E->d_vbo_ocb[ri], max_batch)) return SVBFM_ERR_OOM;
            if (dev_alloc(E, &E->d_vbo_ownb[ri], max_batch)) return SVBFM_ERR_OOM;
            if (with_x && (dev_alloc(E, &E->d_vbo_xb[ri][0], max_batch) || dev_alloc(E, &E->d_vbo_xb[ri][1], max_batch))) return SVBFM_ERR_OOM;
        }
        E->vbo_batch_cap = (uint32_t)max_batch;
    }
    uint32_t *d_keys = nullptr, *d_vals = nullptr, *d_skeys = nullptr;
    SV_CUDA(E, sv_malloc((void**)&d_keys, std::max<size_t>(n, 1) * 4));
    SV_CUDA(E, sv_malloc((void**)&d_vals, std::max<size_t>(n, 1) * 4));
    for (int ri = 0; ri < 2; ri++) {
        const Run& r = E->runs[ri];
        const uint32_t nc = r.col_end - r.col_begin;
        const size_t ncp = (size_t)num_batch * nc + 1;
        sv_free(E->d_vbo_idx[ri]); sv_free(E->d_vbo_colptr[ri]);
        E->d_vbo_idx[ri] = nullptr; E->d_vbo_colptr[ri] = nullptr;
        if (dev_alloc(E, &E->d_vbo_colptr[ri], ncp)) return SVBFM_ERR_OOM;
        SV_CUDA(E, cudaMemsetAsync(E->d_vbo_colptr[ri], 0, ncp * 8, st));
        if (n) {
            k_vbo_keys<<<nblk(n), 256, 0, st>>>(E->d_rbatch, ri ? S.crow + S.h_colptr[r.col_begin] : nullptr, n, d_keys, d_vals);
            if (int rc = sort_pairs(E, d_keys, d_vals, n, std::max<uint32_t>(num_batch, 1), &d_skeys, &E->d_vbo_idx[ri])) return rc;
            sv_free(d_skeys);
            E->dev_bytes += (size_t)n * 4;
            k_vbo_hist<<<nblk(n), 256, 0, st>>>(E->d_rbatch, S.rcol, n, ri, r.col_begin, nc, E->d_vbo_colptr[ri]);
        } else if (dev_alloc(E, &E->d_vbo_idx[ri], 1)) return SVBFM_ERR_OOM;
        if (E->world > 1) {       // global number of batch entries of every column: the histogram summed over the ranks (before the scan)
            sv_free(E->d_vbo_gcnt[ri]); E->d_vbo_gcnt[ri] = nullptr;
            if (dev_alloc(E, &E->d_vbo_gcnt[ri], ncp)) return SVBFM_ERR_OOM;
            k_u64_to_u32<<<nblk(ncp - 1), 256, 0, st>>>(E->d_vbo_colptr[ri], ncp - 1, E->d_vbo_gcnt[ri]);
            if (int rc = allreduce(E, E->d_vbo_gcnt[ri], ncp - 1, 3 /*ncclUint32*/, 0 /*ncclSum*/)) return rc;
        }
        // the non-empty columns of every batch, batch after batch (the finalize of a batch walks these instead of every column of the
        // run: a 2 M-entry batch of the 200 M shape touches about a third of them)
        sv_free(E->d_vbo_clist[ri]); E->d_vbo_clist[ri] = nullptr;
        E->vbo_clist_off[ri].assign((size_t)num_batch + 1, 0);
        if (ncp > 1 && !getenv("SVBFM_VBO_NO_CLIST")) {
            uint32_t *d_flag = nullptr, *d_pos = nullptr, *d_sel = nullptr, *d_off = nullptr;
            SV_CUDA(E, sv_malloc((void**)&d_flag, ncp * 4));
            SV_CUDA(E, sv_malloc((void**)&d_pos, ncp * 4));
            SV_CUDA(E, sv_malloc((void**)&d_off, ((size_t)num_batch + 1) * 4));
            k_vbo_nonempty<<<nblk(ncp), 256, 0, st>>>(E->d_vbo_colptr[ri], E->world > 1 ? E->d_vbo_gcnt[ri] : nullptr, ncp - 1, d_flag);
            size_t scan_bytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_flag, d_pos, (int64_t)ncp, st);
            void* scan_tmp = nullptr;
            SV_CUDA(E, sv_malloc(&scan_tmp, scan_bytes ? scan_bytes : 1));
            cudaError_t se = cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, d_flag, d_pos, (int64_t)ncp, st);
            uint32_t n_sel = 0;
            if (se == cudaSuccess) se = cudaMemcpyAsync(&n_sel, d_pos + (ncp - 1), 4, cudaMemcpyDeviceToHost, st);
            if (se == cudaSuccess) se = cudaStreamSynchronize(st);
            if (se == cudaSuccess) se = sv_malloc((void**)&d_sel, std::max<size_t>(n_sel, 1) * 4);
            if (se == cudaSuccess) {
                k_vbo_compact<<<nblk(ncp - 1), 256, 0, st>>>(d_pos, ncp - 1, d_sel);
                k_vbo_clist_finish<<<nblk((uint64_t)num_batch + 1), 256, 0, st>>>(d_sel, n_sel, nc, r.col_begin, num_batch, d_off);
                se = cudaMemcpyAsync(E->vbo_clist_off[ri].data(), d_off, ((size_t)num_batch + 1) * 4, cudaMemcpyDeviceToHost, st);
                if (se == cudaSuccess) se = cudaStreamSynchronize(st);
                if (n_sel) k_vbo_clist_cols<<<nblk(n_sel), 256, 0, st>>>(d_sel, n_sel, nc, r.col_begin);
            }
            sv_free(scan_tmp); sv_free(d_flag); sv_free(d_off);
            if (se != cudaSuccess) { sv_free(d_pos); sv_free(d_sel); return fail(E, SVBFM_ERR_CUDA, std::string("vb_online column lists: ") + cudaGetErrorString(se)); }
            E->d_vbo_clist[ri] = d_sel;
            sv_free(E->d_vbo_cpos[ri]);
            E->d_vbo_cpos[ri] = d_pos;       // kept: the compact column ids of a batch (engine: k_vbo_pack)
        } else { sv_free(E->d_vbo_cpos[ri]); E->d_vbo_cpos[ri] = nullptr; }
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, E->d_vbo_colptr[ri], E->d_vbo_colptr[ri], (int64_t)ncp, st);
        void* tmp = nullptr;
        SV_CUDA(E, sv_malloc(&tmp, tmp_bytes ? tmp_bytes : 1));
        cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, E->d_vbo_colptr[ri], E->d_vbo_colptr[ri], (int64_t)ncp, st);
        sv_free(tmp);
        if (e != cudaSuccess) return fail(E, SVBFM_ERR_CUDA, std::string("vb_online scan: ") + cudaGetErrorString(e));
    }
    sv_free(d_keys); sv_free(d_vals);
    // dense per-batch column arrays (one GPU, column lists built for both fields)
    if (E->vbo_compact && E->world == 1 && E->d_vbo_cpos[0] && E->d_vbo_cpos[1]) {
        uint32_t cap = 1;
        for (uint32_t b = 0; b < num_batch; b++)
            cap = std::max(cap, (E->vbo_clist_off[0][b + 1] - E->vbo_clist_off[0][b]) + (E->vbo_clist_off[1][b + 1] - E->vbo_clist_off[1][b]));
        if (cap > E->vbo_cols_cap) {
            sv_free(E->d_vbo_cc); sv_free(E->d_vbo_ccptr); sv_free(E->d_vbo_opack_c); sv_free(E->d_vbo_colsum_c); sv_free(E->d_vbo_dT_c);
            sv_free(E->d_vbo_nextp_c); sv_free(E->d_vbo_prevm_c);
            E->d_vbo_cc = nullptr; E->d_vbo_ccptr = nullptr; E->d_vbo_opack_c = nullptr; E->d_vbo_colsum_c = nullptr; E->d_vbo_dT_c = nullptr;
            E->d_vbo_nextp_c = nullptr; E->d_vbo_prevm_c = nullptr;
            E->vbo_cols_cap = 0;
            SV_CUDA(E, sv_malloc((void**)&E->d_vbo_cc, (size_t)cap * 32));
            SV_CUDA(E, sv_malloc((void**)&E->d_vbo_opack_c, (size_t)cap * 32));
            if (dev_alloc(E, &E->d_vbo_ccptr, (size_t)cap + 2)) return SVBFM_ERR_OOM;
            if (dev_alloc(E, &E->d_vbo_colsum_c, (size_t)cap * 4)) return SVBFM_ERR_OOM;
            if (dev_alloc(E, &E->d_vbo_dT_c, cap)) return SVBFM_ERR_OOM;
            if (dev_alloc(E, &E->d_vbo_nextp_c, cap)) return SVBFM_ERR_OOM;
            if (dev_alloc(E, &E->d_vbo_prevm_c, cap)) return SVBFM_ERR_OOM;
            E->vbo_cols_cap = cap;
        }
    }
    SV_CUDA(E, cudaGetLastError());
    return 0;
}

}  // namespace svb
