// csrc/svbfm_kernels.cuh -- hand-written sm_100a kernels of the VB / MCMC coordinate sweep.
//
// Two schedules:
//   * two complete one-hot fields (user x item data, the headline shape): the TWO-COPY STREAM schedule, k_stream below:
//     one streaming pass per (step, field) over that field's own copy of the residuals + k_finalize.
//   * anything else (ragged / multi-hot rows, more fields, sharded vb_online): the general per-run schedule:
//
// One FIELD RUN (consecutive, case-disjoint columns) of one factor is swept by
//     k_sweep_reduce   warp per tile (<= tile_entries CSC entries of one column): per-entry terms -> warp
//                      shuffle reduction -> one partial {A,B,C1,C2} per tile          (gather e_i, other-field params)
//     k_combine_*      tiles -> column sums (heavy columns by a CTA; all columns when an allreduce follows)
//     k_finalize       thread per column: posterior mean/variance (VB), Gaussian draw (MCMC), natural-parameter
//                      step (vb_online); non-finite guards; delta_j; d(sum T)_j
//     k_sweep_apply    warp per tile: e_i += x*h*delta_j                               (scatter e_i)
// which is the reference's update_v / update_w / draw_v / draw_w (fm_learn_vb.h:527-644,
// fm_learn_mcmc.h:671-718,780-835, fm_learn_vb_online.h:499-627) for every column of the run at once.
//
// All sums are deterministic (fixed tile order, tree reductions); no floating-point atomics.
#pragma once
#include "svbfm_internal.h"

namespace svb {

enum { KIND_VB_W = 0, KIND_VB_V = 1, KIND_MC_W = 2, KIND_MC_V = 3, KIND_VBO_W = 4, KIND_VBO_V = 5 };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block reduction of NV values; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /*[NV*32]*/) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = warp_sum(v[k]);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; k++) smem[k * 32 + w] = v[k];
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double x = (lane < nw) ? smem[k * 32 + lane] : 0.0;
            v[k] = warp_sum(x);
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// "other fields of case i": h = sum_{c != j} mu_c x_c ; h1 = sum sigma_c x_c^2 ; h2 = sum mu_c^2 x_c^2
// (= q_i - x mu_j, S2_i - x^2 sigma_j, S3_i - x^2 mu_j^2 of fm_learn_vb.h:592-593, 628-630)
struct OtherView {            // F == 2 only: for every CSC entry, the feature id (and x) of the case's OTHER entry,
    const uint32_t* col;      // stored entry-aligned so that the sweep streams it instead of gathering the CSR row
    const float* val;
};

// two-copy stream schedule: everything an entry needs from its OTHER column in one 32-byte record (one LDG.E.256
// gather), and the constants of its OWN column for the next pass of its side (k_stream below)
struct alignas(32) ColPack { double mu, sg, delta, h4; };
struct alignas(32) OwnPack { double mu_red, h_oth, d_own, pad; };

// The 32-byte record gather of k_stream. `no_alloc`: ld.global.nc.L1::no_allocate (the line goes from L2 to the registers without
// taking an L1 line). Measured at 200 M ratings, K = 50 (gpurun call r2k, profiles/r02_k_*): the pass over the SECOND field gathers
// the first field's records in column-id order with little reuse between requests: 47.5 -> 43.5 ms per iteration without the
// allocation; the pass over the FIRST field gathers the rank-ordered records of the second (hot head dense in a few KB): 46.5 -> 61.8 ms
// without it. ld.global.nc alone and nc + L1::evict_last change nothing (108.2 / 108.6 against 107.9 ms per iteration).
__device__ __forceinline__ ColPack sv_load_record(const ColPack* p, bool no_alloc) {
#if defined(SVBFM_EMULATED)
    (void)no_alloc;
    return *p;
#else
    if (!no_alloc) return *p;
    ColPack r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.mu), "=d"(r.sg), "=d"(r.delta), "=d"(r.h4) : "l"(p));
    return r;
#endif
}

template <int FT, bool ONES, bool VAR>
__device__ __forceinline__ void others(const RowView& rv, const OtherView& ov, const double2* __restrict__ pf, uint64_t p, uint32_t i, uint32_t j,
                                       double& h, double& h1, double& h2) {
    h = 0.0; h1 = 0.0; h2 = 0.0;
    if constexpr (FT == 2) {
        uint32_t o = __ldcs(&ov.col[p]);
        double2 P = __ldg(&pf[o]);
        if constexpr (ONES) {
            h = P.x;
            if constexpr (VAR) { h1 = P.y; h2 = P.x * P.x; }
        } else {
            float x = __ldcs(&ov.val[p]);
            h = P.x * x;
            if constexpr (VAR) { h1 = P.y * x * x; h2 = P.x * P.x * x * x; }
        }
    } else {
        uint64_t b, e;
        if constexpr (FT == 1) { b = (uint64_t)i * rv.F; e = b + rv.F; }
        else { b = __ldg(&rv.rowptr[i]); e = __ldg(&rv.rowptr[i + 1]); }
        for (uint64_t k = b; k < e; k++) {
            uint32_t c = __ldg(&rv.rcol[k]);
            if (c == j) continue;
            double2 P = __ldg(&pf[c]);
            if constexpr (ONES) {
                h += P.x;
                if constexpr (VAR) { h1 += P.y; h2 += P.x * P.x; }
            } else {
                float x = __ldg(&rv.rval[k]);
                h += P.x * x;
                if constexpr (VAR) { h1 += P.y * x * x; h2 += P.x * P.x * x * x; }
            }
        }
    }
}

struct SweepArgs {
    const uint32_t* tile_col;
    const uint64_t* tile_begin;
    const uint32_t* tile_len;
    const uint32_t* exec_order;   // [tile0 + w] -> tile id executed by warp w of the launch
    const uint64_t* colptr;
    const uint32_t* crow;
    const float* cval;
    RowView rv;
    OtherView ov;
    double* e;
    const double2* pf;        // params of this factor ([D]) or the w params
    double* partial;          // [n_tiles][4]
    const double* delta;      // [D]
    uint32_t tile0, ntiles, tile_entries;
    const uint16_t* cbatch;   // vb_online: batch id of the case of every CSC entry (null otherwise)
    uint32_t batch;           // vb_online: current batch
};

// ---------------------------------------------------------------------------------------------------------
// per-entry terms of one column (pass 1 of update_v/update_w/draw_v/draw_w)
template <int KIND, int FT, bool ONES>
__global__ void __launch_bounds__(256) k_sweep_reduce(SweepArgs a) {
    uint32_t w = (blockIdx.x * (blockDim.x >> 5)) + (threadIdx.x >> 5);
    if (w >= a.ntiles) return;
    uint32_t t = __ldg(&a.exec_order[a.tile0 + w]), lane = threadIdx.x & 31;
    uint32_t j = __ldg(&a.tile_col[t]);
    uint64_t b = __ldg(&a.tile_begin[t]);
    uint64_t e_ = b + __ldg(&a.tile_len[t]);
    double2 Pj = __ldg(&a.pf[j]);
    double mu = Pj.x;
    double A = 0.0, B = 0.0, C1 = 0.0, C2 = 0.0;
    constexpr int U = 4;                     // entries per lane in flight (8 was slower: fewer resident warps): the loads of a batch are issued back to back
    for (uint64_t p0 = b + lane; p0 < e_; p0 += 32 * U) {
        bool ok[U]; uint32_t ci[U]; float xs[U]; double es[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            uint64_t p = p0 + (uint64_t)u * 32;
            ok[u] = p < e_;
            if constexpr (KIND == KIND_VBO_W || KIND == KIND_VBO_V)
                if (ok[u]) ok[u] = (__ldcs(&a.cbatch[p]) == a.batch);   // entry of another batch (vbos.h:103-109)
            ci[u] = ok[u] ? __ldcs(&a.crow[p]) : 0u;
            xs[u] = 1.0f;
            if constexpr (!ONES) if (ok[u]) xs[u] = __ldcs(&a.cval[p]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) es[u] = ok[u] ? a.e[ci[u]] : 0.0;
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!ok[u]) continue;
            uint64_t p = p0 + (uint64_t)u * 32;
            uint32_t i = ci[u];
            float xf = xs[u];
            double ei = es[u];
            double xx = (double)(xf * xf);     // the reference forms x*x in float (FM_FLOAT), then promotes
            if constexpr (KIND == KIND_VB_W || KIND == KIND_VBO_W) {
                A += xf * (ei + xf * mu);                              // vb.h:537
                B += xx;                                               // vb.h:538
                if constexpr (KIND == KIND_VBO_W) C1 += 1.0;           // |Omega_j^b|: batch entries of the column
            } else if constexpr (KIND == KIND_MC_W) {
                A += xf * (ei - mu * xf);                              // mcmc.h:677
                B += xx;
            } else if constexpr (KIND == KIND_VB_V || KIND == KIND_VBO_V) {
                double h, h1, h2;
                others<FT, ONES, true>(a.rv, a.ov, a.pf, p, i, j, h, h1, h2);
                A += xf * h * (ei + xf * mu * h);                      // vb.h:594
                B += xx * h * h + xx * h1;                             // vb.h:595
                C1 += xx * h1;                                         // sum of pass-2 h1 (vb.h:629)
                C2 += xx * h2;                                         // sum of pass-2 h2 (vb.h:630)
            } else {                                                   // KIND_MC_V
                double h, h1, h2;
                others<FT, ONES, false>(a.rv, a.ov, a.pf, p, i, j, h, h1, h2);
                double hh = xf * h;                                    // mcmc.h:789
                A += hh * ei;                                          // mcmc.h:790
                B += hh * hh;                                          // mcmc.h:791
            }
        }
    }
    A = warp_sum(A); B = warp_sum(B);
    if constexpr (KIND == KIND_VB_V || KIND == KIND_VBO_V) { C1 = warp_sum(C1); C2 = warp_sum(C2); }
    if constexpr (KIND == KIND_VBO_W) C1 = warp_sum(C1);
    if (lane == 0) {
        double2* out = reinterpret_cast<double2*>(a.partial + (size_t)t * 4);
        out[0] = make_double2(A, B);
        out[1] = make_double2(C1, C2);
    }
}

// tiles -> column sums. heavy: one CTA per heavy column (list). all: one thread per column (light only) --
// used when the column sums travel through an allreduce.
__global__ void __launch_bounds__(128) k_combine_heavy(const uint32_t* __restrict__ heavy_cols, uint32_t h0, const uint32_t* __restrict__ col_tile0,
                                                       const double* __restrict__ partial, double* __restrict__ colsum) {
    __shared__ double sm[4 * 32];
    uint32_t j = heavy_cols[h0 + blockIdx.x];
    uint32_t t0 = col_tile0[j], t1 = col_tile0[j + 1];
    double v[4] = {0, 0, 0, 0};
    for (uint32_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        const double2* p = reinterpret_cast<const double2*>(partial + (size_t)t * 4);
        double2 a = p[0], b = p[1];
        v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y;
    }
    block_sum<4>(v, sm);
    if (threadIdx.x == 0) {
        double2* o = reinterpret_cast<double2*>(colsum + (size_t)j * 4);
        o[0] = make_double2(v[0], v[1]);
        o[1] = make_double2(v[2], v[3]);
    }
}

__device__ __forceinline__ void load_colsum(uint32_t j, const uint32_t* __restrict__ col_tile0, const double* __restrict__ partial,
                                            const double* __restrict__ colsum, bool from_colsum, double& A, double& B, double& C1, double& C2) {
    uint32_t t0 = col_tile0[j], t1 = col_tile0[j + 1];
    if (from_colsum || (t1 - t0) > 8) {
        const double2* p = reinterpret_cast<const double2*>(colsum + (size_t)j * 4);
        double2 a = p[0], b = p[1];
        A = a.x; B = a.y; C1 = b.x; C2 = b.y;
    } else {
        A = B = C1 = C2 = 0.0;
        for (uint32_t t = t0; t < t1; t++) {
            const double2* p = reinterpret_cast<const double2*>(partial + (size_t)t * 4);
            double2 a = p[0], b = p[1];
            A += a.x; B += a.y; C1 += b.x; C2 += b.y;
        }
    }
}

__global__ void k_combine_light(uint32_t c0, uint32_t c1, const uint32_t* __restrict__ col_tile0, const double* __restrict__ partial,
                                double* __restrict__ colsum) {
    uint32_t j = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c1) return;
    if (col_tile0[j + 1] - col_tile0[j] > 8) return;   // heavy: written by k_combine_heavy
    double A, B, C1, C2;
    load_colsum(j, col_tile0, partial, colsum, false, A, B, C1, C2);
    double2* o = reinterpret_cast<double2*>(colsum + (size_t)j * 4);
    o[0] = make_double2(A, B);
    o[1] = make_double2(C1, C2);
}

// ---- stream schedule (k_stream): implicit tiles of TS = 2^ts_shift consecutive entries of a run, which may span several
// columns. A column that lies inside one tile has its sums written straight to colsum[j]; a column that crosses
// tile borders leaves one piece per tile: partial[T][1] in the tile where it starts, partial[T][0] in the others.
#define SV_SPAN_LIGHT 8u      // columns spanning more tiles than this are combined by a CTA (k_combine_span)
// (Summing those columns inside k_finalize instead -- the lanes of a warp over the pieces, one launch less per step and field -- was
// measured and dropped, profiles/r02_r_*: the heaviest column's ~2000 / ~3400 pieces sit on the critical path of the finalize,
// ML-10M shape 9.7 -> 10.8 ms per iteration, 200 M + 2 ms; ML-1M shape 1.53 -> 1.49 ms: a launch gap is ~1 us on a B200.)

struct SpanView {
    const uint64_t* colptr;   // null: the explicit-tile layout (col_tile0 / partial[tile][4]) is in use
    uint64_t entry0;          // first entry of the run
    uint32_t ts_shift;        // tile = 2^ts_shift entries (Engine::ts_shift)
    uint32_t light_limit;     // columns spanning more tiles than this have their sums in colsum[j] (k_combine_span); ~0u: none
};

__global__ void __launch_bounds__(128) k_combine_span(const uint32_t* __restrict__ heavy_cols, uint32_t h0, SpanView sp,
                                                      const double* __restrict__ partial, double* __restrict__ colsum) {
    __shared__ double sm[4 * 32];
    uint32_t j = heavy_cols[h0 + blockIdx.x];
    uint64_t b = sp.colptr[j], e = sp.colptr[j + 1];
    uint64_t T0 = (b - sp.entry0) >> sp.ts_shift, T1 = (e - 1 - sp.entry0) >> sp.ts_shift;
    double v[4] = {0, 0, 0, 0};
    for (uint64_t T = T0 + threadIdx.x; T <= T1; T += blockDim.x) {
        const double2* p = reinterpret_cast<const double2*>(partial + (T * 2 + (T == T0 ? 1 : 0)) * 4);
        double2 x = p[0], y = p[1];
        v[0] += x.x; v[1] += x.y; v[2] += y.x; v[3] += y.y;
    }
    block_sum<4>(v, sm);
    if (threadIdx.x == 0) {
        double2* o = reinterpret_cast<double2*>(colsum + (size_t)j * 4);
        o[0] = make_double2(v[0], v[1]);
        o[1] = make_double2(v[2], v[3]);
    }
}

// returns false when colsum[j] already holds the sums (one-tile column, heavy column, or after the allreduce)
__device__ __forceinline__ bool span_sum(uint32_t j, SpanView sp, const double* __restrict__ partial, double& A, double& B, double& C1, double& C2,
                                         bool& empty) {
    uint64_t b = sp.colptr[j], e = sp.colptr[j + 1];
    A = B = C1 = C2 = 0.0;
    empty = (e == b);
    if (empty) return true;
    uint64_t T0 = (b - sp.entry0) >> sp.ts_shift, T1 = (e - 1 - sp.entry0) >> sp.ts_shift;
    if (T0 == T1 || T1 - T0 > (uint64_t)sp.light_limit) return false;
    for (uint64_t T = T0; T <= T1; T++) {
        const double2* p = reinterpret_cast<const double2*>(partial + (T * 2 + (T == T0 ? 1 : 0)) * 4);
        double2 x = p[0], y = p[1];
        A += x.x; B += x.y; C1 += y.x; C2 += y.y;
    }
    return true;
}

// sharded stream schedule: complete colsum[j] for every column of the run and pack what travels through the
// allreduce: ab[j] = {A, B} (VB_V: B = C1 + C2). C1, C2 stay local: d(sum T) is linear in them, so every rank adds
// its own share and the shares are summed once per iteration.
template <bool VB_V>
__global__ void k_combine_light_span(uint32_t c0, uint32_t c1, SpanView sp, const double* __restrict__ partial, double* __restrict__ colsum,
                                     double2* __restrict__ ab) {
    uint32_t j = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c1) return;
    double A, B, C1, C2; bool empty;
    double2* o = reinterpret_cast<double2*>(colsum + (size_t)j * 4);
    if (span_sum(j, sp, partial, A, B, C1, C2, empty)) {
        o[0] = make_double2(A, B);
        o[1] = make_double2(C1, C2);
    } else {
        double2 x = o[0], y = o[1];
        A = x.x; B = x.y; C1 = y.x; C2 = y.y;
    }
    ab[j] = make_double2(A, VB_V ? C1 + C2 : B);
}

// ---------------------------------------------------------------------------------------------------------
// counter-based RNG (Philox4x32-10, Salmon et al. 2011) -- replaces the reference's libc rand() stream
// (src/util/random.h:150-176) for MCMC draws; matched in distribution, identical on every rank.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0; k.y += W1;
    }
    return c;
}
__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {   // (0,1), 53 bits
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double philox_normal(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint4 r = philox4x32_10(make_uint4(c0, c1, c2, c3), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    double u1 = u01(r.x, r.y), u2 = u01(r.z, r.w);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// vb_online, one GPU: what a finalize needs about a non-empty column of the batch in flight and what does not change inside the batch,
// gathered once per batch (k_vbo_cols) and read coalesced by the 2 (K + 1) finalizes of the batch (index: the column's position in
// the batch's column lists, first field first -- the same ids the packed batch carries as own-column ids, so the column sums, the
// own-column constants and d(sum T) of the batch live in dense arrays as well)
struct alignas(32) VboCol { uint32_t j, slot, group, t_v; double col_count, cnt; };

struct PeerFlags { unsigned long long* p[16]; int n, me; };      // p[r]: rank r's 16 flag words; a rank writes word `me` of every other rank

struct FinalizeArgs {
    uint32_t c0, c1;             // column range of the run
    int f;                       // factor (or -1 for w)
    int K;
    const uint32_t* col_tile0;
    const double* partial;
    const double* colsum;
    bool from_colsum;
    double2* pf;                 // params to update ([D])
    const uint32_t* group;
    const double* hyper;         // vb: sigma_w[G] or sigma_v[G][K]; mcmc: lambda
    const double* hyper_mu;      // mcmc: mu per group
    Scalars* sc;
    double* delta;               // [D]
    // stream schedule (null / 0 otherwise): records of this column for the passes that follow (see k_stream)
    SpanView span;
    const double2* ab;           // sharded: global {A, B} of every column after the allreduce (colsum keeps the local sums)
    ColPack* cpack;              // [D] gathered by the other side
    const uint32_t* rec_slot;    // record slot of every column (null: the column id; Engine::d_rec_slot)
    OwnPack* opack;              // [D] read by this side's next pass
    int rec_mode;                // 1: first field  -> cpack {new mean, new var, delta, p_prev mean}    opack {p_next mean, new mean, delta}
                                 // 2: second field -> cpack {p_next mean, p_next var, delta, old mean} opack {p_next mean, p_next mean, delta}
    const double2* p_next;       // parameters of the step that follows ([D]; null after the last one)
    const double2* p_prev;       // parameters of the step before ([D]; null at the first one)
    double* dT;                  // [D]
    double2* stage;              // cross shards: {new mean, new var} of the column at stage[slot - stage_base]: what the other ranks fetch
    uint32_t stage_base;
    uint64_t seed; int do_sample;
    // vb_online
    double2* nat;                // [D] natural params of this factor
    uint32_t* t_cnt;             // [D] t_wj (w) or t_vj (v)
    const double* col_count;     // [D]
    const uint64_t* colptr;      // batch column sizes come from colsum C-slot instead (see engine)
    const uint32_t* gcnt;        // vb_online on the sharded stream schedule: global batch entries of every column (indexed like span.colptr)
    const VboCol* cc;            // vb_online, compact columns: the batch's non-empty columns of BOTH fields (null: global column ids everywhere)
    const uint64_t* ccptr;       // ... [n_list + 1] the batch's column pointer at this field's non-empty columns
    uint32_t cid0;               // ... first compact id of this field (colsum / opack / dT are indexed by cid0 + t)
    double2* nextp_c;            // ... what the step before read as p_next[j]: this step's own parameters (valid when `carry`)
    double* prevm_c;             // ... the new mean the step before gave the column: this step's p_prev[j].x (valid when `carry`)
    int carry;                   // ... not the first step of the batch: the two arrays above replace two scattered 16-byte reads per column
    const uint32_t* col_list;    // vb_online on the stream schedule: the batch's non-empty columns of the run (null: every column c0 .. c1)
    uint32_t n_list;
    int update_t;                // vb_online: 1 when this sweep advances t_cnt (w: always; v: f == 0)
};

// stream schedule: the records of column j for the passes that follow its update (see k_stream)
// (sj: the column's record slot, oi: its index in opack)
__device__ __forceinline__ void write_records_at(const FinalizeArgs& a, uint32_t j, uint32_t sj, size_t oi, double new_mean, double new_var, double dlt,
                                                 double mu_old) {
    if (!a.rec_mode) return;
    double2 N = a.p_next ? a.p_next[j] : make_double2(0.0, 0.0);
    if (a.stage) a.stage[sj - a.stage_base] = make_double2(new_mean, new_var);
    if (a.rec_mode == 1) {
        double prev = 0.0;
        if (a.p_prev) prev = (a.cc && a.carry) ? a.prevm_c[oi] : a.p_prev[j].x;
        a.cpack[sj] = ColPack{new_mean, new_var, dlt, prev};
        a.opack[oi] = OwnPack{N.x, new_mean, dlt, 0.0};
    } else {
        a.cpack[sj] = ColPack{N.x, N.y, dlt, mu_old};
        a.opack[oi] = OwnPack{N.x, N.x, dlt, 0.0};
    }
    if (a.cc) { a.nextp_c[oi] = N; a.prevm_c[oi] = new_mean; }
}
__device__ __forceinline__ void write_records(const FinalizeArgs& a, uint32_t j, double new_mean, double new_var, double dlt, double mu_old) {
    if (!a.rec_mode) return;
    write_records_at(a, j, a.rec_slot ? a.rec_slot[j] : j, j, new_mean, new_var, dlt, mu_old);
}

template <int KIND>
__global__ void __launch_bounds__(256) k_finalize(FinalizeArgs a) {
    static_assert(KIND <= KIND_MC_V, "vb_online uses k_finalize_vbo");
    uint32_t j = a.c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.c1) return;
    double A, B, C1, C2;
    if (a.span.colptr) {
        bool empty;
        if (a.from_colsum || !span_sum(j, a.span, a.partial, A, B, C1, C2, empty)) {
            const double2* p = reinterpret_cast<const double2*>(a.colsum + (size_t)j * 4);
            double2 x = p[0], y = p[1];
            A = x.x; B = x.y; C1 = y.x; C2 = y.y;
        }
        if constexpr (KIND == KIND_VB_V) B = C1 + C2;      // vb.h:595 (k_stream leaves the B slot empty)
    } else load_colsum(j, a.col_tile0, a.partial, a.colsum, a.from_colsum, A, B, C1, C2);
    const double B_local = B;                              // this rank's share (d(sum T) of a w column, vb.h:572)
    if (a.ab) { double2 g2 = a.ab[j]; A = g2.x; B = g2.y; }
    uint32_t g = a.group[j];
    double hy = (a.f < 0) ? a.hyper[g] : a.hyper[(size_t)g * a.K + a.f];
    double alpha = a.sc->alpha;
    double2 P = a.pf[j];
    double mu_old = P.x, sg_old = P.y;
    unsigned bad = 0;
    double new_mean, new_var, dlt;
    if constexpr (KIND == KIND_VB_W || KIND == KIND_VB_V) {
        double sg = 1.0 / (hy + alpha * B);                 // vb.h:540 / :597
        double mu = sg * alpha * A;                          // vb.h:541 / :598
        if (isnan(sg) || isinf(sg)) { sg = sg_old; bad++; }  // vb.h:545-549 / :600-604
        bool skip = false;
        if (isnan(mu) || isinf(mu)) { mu = mu_old; bad++; skip = true; }   // vb.h:552-565 / :606-619 (pass 2 skipped)
        a.pf[j] = make_double2(mu, sg);
        new_mean = mu; new_var = sg; dlt = skip ? 0.0 : (mu_old - mu);
        a.delta[j] = dlt;
        if (!skip) {
            if constexpr (KIND == KIND_VB_W) a.dT[j] += B_local * (sg - sg_old);                              // vb.h:572
            else a.dT[j] += (C1 + C2) * (sg - sg_old) + C1 * (mu * mu - mu_old * mu_old);                        // vb.h:639-640
        }
    } else if constexpr (KIND == KIND_MC_W || KIND == KIND_MC_V) {
        double v_old = mu_old;
        double hm = (a.f < 0) ? a.hyper_mu[g] : a.hyper_mu[(size_t)g * a.K + a.f];
        if constexpr (KIND == KIND_MC_V) A -= v_old * B;                      // mcmc.h:793
        double s2 = 1.0 / (hy + alpha * B);                                  // mcmc.h:680 / :794
        double mean = -s2 * (alpha * A - hm * hy);                           // mcmc.h:681 / :795
        double v;
        if (isnan(s2) || isinf(s2)) v = 0.0;                                 // mcmc.h:686-687 / :800-801
        else if (a.do_sample) {
            double sd = sqrt(s2);
            if (sd == 0.0 || isnan(sd)) v = mean;                            // random.h:166-172
            else v = mean + sd * philox_normal(a.seed, j, (uint32_t)(a.f + 1), a.sc->iter, 0x5eed0001u);
        } else v = mean;
        bool skip = false;
        if (isnan(v) || isinf(v)) { v = v_old; bad++; skip = true; }         // mcmc.h:697-710 / :811-824
        a.pf[j] = make_double2(v, 0.0);
        new_mean = v; new_var = 0.0; dlt = skip ? 0.0 : (v - v_old);         // e -= h (v_old - v)  (mcmc.h:716 / :833)
        a.delta[j] = dlt;
    }
    write_records(a, j, new_mean, new_var, dlt, mu_old);
    if (bad) atomicAdd(&a.sc->nan_inf, (unsigned long long)bad);
}

// vb_online finalize: natural-parameter step. `cnt` = number of batch entries of the column (exact in fp64).
//   eta2 <- mean_n[(1-rho) eta2_old + rho (prior + alpha c_j B_n)] = (1-rho) eta2_old + rho (prior + alpha c_j B/cnt)
//   eta1 <- (1-rho) eta1_old + rho c_j alpha A/cnt ;  mu = eta1/eta2 ; sigma = 1/eta2
template <int KIND>
__global__ void __launch_bounds__(256) k_finalize_vbo(FinalizeArgs a, double* __restrict__ cnt_arr, double lamda, uint32_t t0, int update_params) {
    uint32_t j = a.c0 + blockIdx.x * blockDim.x + threadIdx.x;
    bool act = j < a.c1;
    const uint32_t tl = blockIdx.x * blockDim.x + threadIdx.x;
    VboCol cs = VboCol{0u, 0u, 0u, 0u, 0.0, 0.0};
    if (a.cc) {              // compact columns: thread tl takes the tl-th non-empty column of this field
        act = tl < a.n_list;
        if (act) cs = a.cc[a.cid0 + tl];
        j = act ? cs.j : a.c0;
    } else if (a.col_list) { // empty columns are skipped anyway (vbo.h:367, 394) and nothing reads their delta on the stream schedule
        act = tl < a.n_list;
        j = act ? a.col_list[tl] : a.c0;
        act = act && j < a.c1;
    }
    const size_t ci = a.cc ? (size_t)a.cid0 + tl : (size_t)j;       // index of the column in colsum / dT / opack
    double A = 0.0, B = 0.0, C1 = 0.0, C2 = 0.0;
    double cnt = 0.0;
    if (a.span.colptr) {
        // stream schedule: colptr is the batch's own column pointer, so the column's batch entries are counted directly.
        // Records of a skipped column are not written: no entry of the batch refers to it. No thread leaves before the
        // cooperative part below: the lanes of a warp sum a long span together.
        uint64_t cb = 0, ce = 0;
        if (act) {
            if (a.cc) { cb = a.ccptr[tl]; ce = a.ccptr[tl + 1]; }
            else { cb = a.span.colptr[j]; ce = a.span.colptr[j + 1]; }
            cnt = a.gcnt ? (double)a.gcnt[j] : (double)(ce - cb);
        }
        const bool live = act && cnt != 0.0 && update_params;
        if (act && !live && !a.cc) a.delta[j] = 0.0;
        uint64_t T0 = 0, T1 = 0;
        bool from_cs = true;                    // colsum[j] holds the sums (one-tile column, or summed before)
        if (live && !a.from_colsum) {
            const uint64_t b = cb, e = ce;
            if (e > b) {
                T0 = (b - a.span.entry0) >> a.span.ts_shift; T1 = (e - 1 - a.span.entry0) >> a.span.ts_shift;
                from_cs = (T0 == T1);
            }
        }
        // a column over 32 tiles and more (the heaviest users of a batch span hundreds of 256-entry tiles): all lanes of the warp take
        // its pieces together, fixed order; one thread alone needed ~90 us for the longest and every finalize of the batch waited for it
        const unsigned lane = threadIdx.x & 31;
        unsigned long_mask = __ballot_sync(0xffffffffu, live && !from_cs && (T1 - T0 >= 32));
        while (long_mask) {
            const int src = __ffs(long_mask) - 1;
            long_mask &= long_mask - 1;
            const uint64_t t0s = __shfl_sync(0xffffffffu, T0, src), t1s = __shfl_sync(0xffffffffu, T1, src);
            double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
            for (uint64_t T = t0s + lane; T <= t1s; T += 32) {
                const double2* p = reinterpret_cast<const double2*>(a.partial + (T * 2 + (T == t0s ? 1 : 0)) * 4);
                double2 x = p[0], y = p[1];
                v0 += x.x; v1 += x.y; v2 += y.x; v3 += y.y;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, o); v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                v2 += __shfl_xor_sync(0xffffffffu, v2, o); v3 += __shfl_xor_sync(0xffffffffu, v3, o);
            }
            if ((int)lane == src) { A = v0; B = v1; C1 = v2; C2 = v3; from_cs = false; T1 = T0 = 0; cnt = -cnt; }    // marked: sums are in registers
        }
        if (!live) return;
        if (cnt < 0.0) cnt = -cnt;              // a long span: done above
        else if (from_cs) {
            const double2* p = reinterpret_cast<const double2*>(a.colsum + ci * 4);
            double2 x = p[0], y = p[1];
            A = x.x; B = x.y; C1 = y.x; C2 = y.y;
        } else {
            for (uint64_t T = T0; T <= T1; T++) {
                const double2* p = reinterpret_cast<const double2*>(a.partial + (T * 2 + (T == T0 ? 1 : 0)) * 4);
                double2 x = p[0], y = p[1];
                A += x.x; B += x.y; C1 += y.x; C2 += y.y;
            }
        }
        if constexpr (KIND == KIND_VBO_V) B = C1 + C2;
    } else {
        if (!act) return;
        load_colsum(j, a.col_tile0, a.partial, a.colsum, a.from_colsum, A, B, C1, C2);
        if constexpr (KIND == KIND_VBO_W) { cnt = C1; cnt_arr[j] = cnt; }   // the w pass also counts the batch entries of the column
        else cnt = cnt_arr[j];
        if (cnt == 0.0 || !update_params) { a.delta[j] = 0.0; return; }     // empty columns are skipped (vbo.h:367, 394)
    }
    const double B_local = B;                                 // this rank's share (d(sum T) of a w column)
    if (a.ab) { double2 g2 = a.ab[j]; A = g2.x; B = g2.y; }   // sharded: global sums after the allreduce
    uint32_t g = a.cc ? cs.group : a.group[j];
    double hy = (a.f < 0) ? a.hyper[g] : a.hyper[(size_t)g * a.K + a.f];
    double alpha = a.sc->alpha;
    double2 P = (a.cc && a.carry) ? a.nextp_c[ci] : a.pf[j];
    double mu_dash = P.x, sg_dash = P.y;
    double2 N = a.nat[j];                                     // {eta1, eta2}
    uint32_t tc = (a.cc && KIND == KIND_VBO_V) ? cs.t_v : a.t_cnt[j];     // (t_vj does not move inside a batch: it advances with the last factor)
    double rho = pow((double)(t0 + tc), -lamda);              // vbo.h:521 / :406 (rate in force for this batch)
    double cj = a.cc ? cs.col_count : a.col_count[j];
    double eta2 = (1.0 - rho) * N.y + rho * (hy + alpha * cj * (B / cnt));      // vbo.h:515 / :579
    double eta1 = (1.0 - rho) * N.x + rho * cj * alpha * (A / cnt);             // vbo.h:516 / :580
    a.nat[j] = make_double2(eta1, eta2);
    if (a.update_t) a.t_cnt[j] = tc + (uint32_t)cnt;          // vbo.h:520 / :401
    double mu = eta1 / eta2, sg = 1.0 / eta2;                 // vbo.h:524-525 / :586-587
    unsigned bad = 0;
    if (isnan(sg) || isinf(sg)) { sg = sg_dash; bad++; }
    bool skip = false;
    if (isnan(mu) || isinf(mu)) { mu = mu_dash; bad++; skip = true; }
    a.pf[j] = make_double2(mu, sg);
    if (!a.cc) a.delta[j] = skip ? 0.0 : (mu_dash - mu);      // (nothing reads delta on the stream schedule: the passes take it from the records)
    if (!skip) {
        if constexpr (KIND == KIND_VBO_W) a.dT[ci] += B_local * (sg - sg_dash);
        else a.dT[ci] += (C1 + C2) * (sg - sg_dash) + C1 * (mu * mu - mu_dash * mu_dash);
    }
    if (a.cc) write_records_at(a, j, cs.slot, ci, mu, sg, skip ? 0.0 : (mu_dash - mu), mu_dash);
    else write_records(a, j, mu, sg, skip ? 0.0 : (mu_dash - mu), mu_dash);
    if (bad) atomicAdd(&a.sc->nan_inf, (unsigned long long)bad);
}

// per batch: the compact column table (see VboCol), the batch's column pointer at the non-empty columns, d(sum T) cleared
struct VboColsArgs {
    const uint32_t* clist[2];      // the batch's non-empty columns of run 0 / run 1
    uint32_t nl[2];
    const uint64_t* colptr[2];     // the batch's column pointers, indexed by global column id
    const uint32_t* rec_slot;      // null: the column id
    const uint32_t* group;
    const uint32_t* t_v;
    const double* col_count;
    VboCol* cc;                    // [nl[0] + nl[1]]
    uint64_t* ccptr;               // [nl[0] + 1] then [nl[1] + 1]
    double* dT_c;                  // [nl[0] + nl[1]]
};
__global__ void __launch_bounds__(256) k_vbo_cols(VboColsArgs a) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nl[0] + a.nl[1]) return;
    const int ri = t >= a.nl[0] ? 1 : 0;
    const uint32_t tt = t - (ri ? a.nl[0] : 0u);
    const uint32_t j = a.clist[ri][tt];
    const uint64_t b = a.colptr[ri][j], e = a.colptr[ri][j + 1];
    a.cc[t] = VboCol{j, a.rec_slot ? a.rec_slot[j] : j, a.group[j], a.t_v[j], a.col_count[j], (double)(e - b)};
    a.ccptr[t + ri] = b;
    if (tt + 1 == a.nl[ri]) a.ccptr[t + ri + 1] = e;
    a.dT_c[t] = 0.0;
}
// own-column constants before the first step (k_pack_init's opack, compact)
__global__ void __launch_bounds__(256) k_vbo_opack_init(const VboCol* __restrict__ cc, uint32_t nl0, uint32_t n, const double2* __restrict__ p0,
                                                        OwnPack* __restrict__ opack_c) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double m = p0[cc[t].j].x;
    opack_c[t] = t < nl0 ? OwnPack{m, 0.0, 0.0, 0.0} : OwnPack{m, m, 0.0, 0.0};
}

// pass 2: e_i += x * h * delta_j   (vb.h:571 / :638; mcmc.h:716 / :833)
template <bool IS_V, int FT, bool ONES>
__global__ void __launch_bounds__(256) k_sweep_apply(SweepArgs a) {
    uint32_t w = (blockIdx.x * (blockDim.x >> 5)) + (threadIdx.x >> 5);
    if (w >= a.ntiles) return;
    uint32_t t = __ldg(&a.exec_order[a.tile0 + w]), lane = threadIdx.x & 31;
    uint32_t j = __ldg(&a.tile_col[t]);
    double d = __ldg(&a.delta[j]);
    if (d == 0.0) return;
    uint64_t b = __ldg(&a.tile_begin[t]);
    uint64_t e_ = b + __ldg(&a.tile_len[t]);
    for (uint64_t p = b + lane; p < e_; p += 32) {
        if (a.cbatch && __ldg(&a.cbatch[p]) != a.batch) continue;
        uint32_t i = __ldg(&a.crow[p]);
        float xf = 1.0f;
        if constexpr (!ONES) xf = __ldg(&a.cval[p]);
        double hh = xf;
        if constexpr (IS_V) {
            double h, h1, h2;
            others<FT, ONES, false>(a.rv, a.ov, a.pf, p, i, j, h, h1, h2);
            hh = xf * h;
        }
        a.e[i] += hh * d;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Two-copy stream schedule for two complete one-hot fields (every case has exactly one feature in run 0 and one in
// run 1). The residuals are kept TWICE: e[i] in the entry order of run 0 (= device case order) and e2[p] in the entry
// order of run 1. Every pass over a field then streams its own copy (oc 4 B + e 8 B read + 8 B write per entry, x
// arrays when not all ones) and gathers one 32-byte record of the OTHER column from L2; there is no random access
// to the residuals at all. The copies stay bit-identical because each applies the same updates
//       e += x_j * h * delta_j                 (vb.h:571 / :638, mcmc.h:716 / :833)
// in the same order with the same operands; an update made by one side is "pending" for a copy until that copy's
// next pass:
//   steps s = w, v_0, .., v_{K-1}; per step  U(s): pass over run 0 -> finalize -> I(s): pass over run 1 -> finalize
//   pass over run 0 at step s : apply U(s-1) [own, h from the record's h4] and I(s-1) [other, h = own mean of s-1]
//   pass over run 1 at step s : apply I(s-1) [own, h4 = the user's mean of s-1] and U(s) [other, h = own mean of s]
//   then (REDUCE) the per-entry terms of pass 1 of the step (vb.h:537-538, 587-596 / mcmc.h:677, 785-792) with the
//   up-to-date residual. After the last step both copies are flushed (REDUCE = false).
// Tiles are implicit: warp t handles entries [t*TS, (t+1)*TS) of the run (TS = 2^ts_shift, 1024 by default) and walks over the columns inside
// (a 32-column window of own-column constants lives in the lanes); sums of a column that ends inside the tile are
// reduced with shuffles and written once. Fixed order everywhere: results are bit-reproducible.
struct StreamArgs {
    const uint64_t* colptr;
    uint32_t c0, c1;          // columns of the run
    uint64_t entry0;          // colptr[c0]
    uint64_t real0;           // first entry of the run in oc / xv / xo (= entry0 unless idx is set)
    const uint32_t* idx;      // IDX: entry k of this pass is the run's entry idx[entry0 + k] (vb_online: the entries of one batch,
                              // in column order; colptr then is the batch's own column pointer into idx)
    uint32_t n;               // entries of the pass
    uint32_t ntiles, ts_shift;
    const uint32_t* tile_col0;
    const uint32_t* oc;       // [nnz] other column of every entry
    const uint32_t* ownc;     // k_stream_rows: own column of every entry, indexed like oc (a packed vb_online batch carries it)
    const float* xv;          // [nnz] own x, other x (null when all ones)
    const float* xo;
    double* e;                // this side's copy of the residuals, index = entry - entry0
    const ColPack* rec;       // records of the other side's columns
    int rec_no_alloc;         // gather them without allocating L1 lines (sv_load_record)
    uint32_t rec_hot_end;     // ... or only the records from this slot on (rank layout: the slots below are the hot head that L1 should keep)
    const OwnPack* own;       // constants of this side's columns
    int has_own, own_is_w;    // a pending update of this side (own_is_w: it was a w step, h = 1)
    int has_oth, oth_is_w;    // a pending update of the other side
    double* colsum;           // [D][4]
    double* partial;          // [ntiles of the run][2][4]
};

// four sums over the warp with 6 shuffle steps instead of 20: the lanes first split the values among themselves
// (xor 16: two each, xor 8: one each), then reduce their own. Result: lane 0 -> v0, lane 8 -> v1, lane 16 -> v2,
// lane 24 -> v3 (returned in every lane's `k`; lanes with (lane & 7) == 0 hold totals). Fixed order: deterministic.
__device__ __forceinline__ double warp_sum4(double v0, double v1, double v2, double v3, uint32_t lane) {
    const bool up16 = lane & 16, up8 = lane & 8;
    double k0 = up16 ? v2 : v0, k1 = up16 ? v3 : v1;
    k0 += __shfl_xor_sync(0xffffffffu, up16 ? v0 : v2, 16);
    k1 += __shfl_xor_sync(0xffffffffu, up16 ? v1 : v3, 16);
    double k = up8 ? k1 : k0;
    k += __shfl_xor_sync(0xffffffffu, up8 ? k0 : k1, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}
// two sums: lane 0 -> v0, lane 16 -> v1
__device__ __forceinline__ double warp_sum2(double v0, double v1, uint32_t lane) {
    const bool up16 = lane & 16;
    double k = up16 ? v1 : v0;
    k += __shfl_xor_sync(0xffffffffu, up16 ? v0 : v1, 16);
    k += __shfl_xor_sync(0xffffffffu, k, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

// the pending updates of one entry. Written with explicit intrinsics: both residual copies must round identically
// whichever code path (fast / general, first / second field) applies them.
template <bool ONES>
__device__ __forceinline__ double apply_pending(double ei, float xf, float xof, const ColPack& g, double h_oth, double d_own,
                                                bool has_own, bool own_is_w, bool has_oth, bool oth_is_w) {
    if (has_own) {
        double H = own_is_w ? 1.0 : (ONES ? g.h4 : __dmul_rn(g.h4, (double)xof));
        ei = __fma_rn(ONES ? H : __dmul_rn((double)xf, H), d_own, ei);
    }
    if (has_oth) {
        double H = oth_is_w ? 1.0 : (ONES ? h_oth : __dmul_rn(h_oth, (double)xf));
        ei = __fma_rn(ONES ? H : __dmul_rn((double)xof, H), g.delta, ei);
    }
    return ei;
}

// per-entry terms of pass 1 (vb.h:537-538, 587-596 / mcmc.h:677, 785-792)
template <int KIND>
__device__ __forceinline__ void entry_terms(double ei, float xf, float xof, const ColPack& g, double mu, double& A, double& B, double& C1, double& C2) {
    double xx = (double)(xf * xf);     // the reference forms x*x in float (FM_FLOAT), then promotes
    if constexpr (KIND == KIND_VB_W) {
        A += xf * (ei + xf * mu);                              // vb.h:537
        B += xx;                                               // vb.h:538
    } else if constexpr (KIND == KIND_MC_W) {
        A += xf * (ei - mu * xf);                              // mcmc.h:677
        B += xx;
    } else if constexpr (KIND == KIND_VB_V) {
        double h = g.mu * xof, h1 = g.sg * xof * xof, h2 = g.mu * g.mu * xof * xof;
        A += xf * h * (ei + xf * mu * h);                      // vb.h:594
        C1 += xx * h1;                                         // sum of pass-2 h1 (vb.h:629)
        C2 += xx * h2;                                         // sum of pass-2 h2 (vb.h:630)
        (void)B;                                               // vb.h:595: B = sum x^2 h^2 + x^2 h1 = C2 + C1, formed in k_finalize
    } else {                                                   // KIND_MC_V
        double hh = xf * (g.mu * xof);                         // mcmc.h:789
        A += hh * ei;                                          // mcmc.h:790
        B += hh * hh;                                          // mcmc.h:791
    }
}

// ---- 1-D bulk copies global -> shared, completion on an mbarrier (TMA unit: no LSU wavefronts, no registers in flight).
// Used by the TMA variant of k_stream. Under tests/emu the copy happens at issue and the barrier keeps the same
// phase / transaction arithmetic, so that a wait on a stage that was never armed shows up as a deadlock.
#if defined(SVBFM_EMULATED)
#include <sched.h>
struct SvMbar { uint32_t phase, pending, init; int64_t tx; };
static inline void sv_mbar_init(SvMbar* b, uint32_t count) { b->phase = 0; b->pending = count; b->init = count; b->tx = 0; }
static inline void sv_mbar_flip(SvMbar* b) { if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = b->init; } }
static inline void sv_mbar_arrive_expect_tx(SvMbar* b, uint32_t bytes) { b->tx += bytes; b->pending--; sv_mbar_flip(b); }
static inline void sv_bulk_g2s(void* dst, const void* src, uint32_t bytes, SvMbar* b) {
    if (((uintptr_t)dst & 15) || ((uintptr_t)src & 15) || (bytes & 15)) { fprintf(stderr, "[emu] misaligned bulk copy\n"); abort(); }
    static long n_copies = 0;
    if (++n_copies == 1 && getenv("SVBFM_EMU_VERBOSE")) fprintf(stderr, "[emu] bulk copies in use\n");
    memcpy(dst, src, bytes); b->tx -= bytes; sv_mbar_flip(b);
}
static inline void sv_mbar_wait(SvMbar* b, uint32_t parity) {
    if (b->phase == parity) { fprintf(stderr, "[emu] mbarrier wait on a phase that will never complete\n"); abort(); }
}
static inline void sv_fence_mbar_init() {}
#else
typedef unsigned long long SvMbar;
__device__ __forceinline__ uint32_t sv_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sv_mbar_init(SvMbar* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sv_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void sv_fence_mbar_init() {      // make the initialised barriers visible to the async proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void sv_mbar_arrive_expect_tx(SvMbar* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sv_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sv_bulk_g2s(void* dst, const void* src, uint32_t bytes, SvMbar* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sv_smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(sv_smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void sv_mbar_wait(SvMbar* b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SV_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SV_DONE_%=;\n"
        "bra SV_WAIT_%=;\n"
        "SV_DONE_%=:\n"
        "}\n" ::"r"(sv_smem_u32(b)),
        "r"(parity)
        : "memory");
}
#endif

// STEADY: both sides have a pending v update (every pass of the factor loop but the first ones): no runtime flags
#ifndef SV_STREAM_U
#define SV_STREAM_U 2        // rows of 32 entries per batch (measured: 2 x 4 CTAs/SM beats 4 x 2, gpurun_out/tune_*)
#endif
// The warps of a CTA do not cooperate, so the CTA size only sets the granularity of the occupancy: 4 warps x 7 CTAs = 28 warps per
// SM at 72 registers (8 x 3 = 24 before). Measured at 200 M ratings, K = 50 (gpurun_out/r2b_*, profiles/r02_kstream_variants.txt):
// 8 x 3 109.3 ms per iteration, 4 x 7 106.2 ms. Tried and dropped in round 2: the records of batch b + 1 gathered into a second
// register set while batch b is processed (80 registers: 127 ms), prefetch.global.L1 / .L2 of the next batch's records (153 ms),
// 32-entry batches at 4 x 8 warps (119 ms): whatever adds L1TEX requests or keeps more of them in flight loses, the pass is
// bound by the 128-byte lines the record gathers touch. Also tried (profiles/r02_d_*): the ring loop
// rewritten for instruction count (shared addresses formed once, slots of two batches, no register double buffer): 586 M -> 427 M
// warp instructions per pass, but 113.9 ms instead of 107.2: issue-active fell from 61 % to 43 % while the long-scoreboard stall
// (the record gather) grew from 1.6 to 6.3 warps per issue -- the pass waits on L1TEX (data-pipe wavefronts 76 % either way),
// not on the issue slots. Ring depth 4 x 768 B instead of 6 x 768 B: 106.0 vs 107.6 ms, inside the noise of the power cap.
#ifndef SV_STREAM_WARPS
#define SV_STREAM_WARPS 4
#endif
#ifndef SV_STREAM_MINB
#define SV_STREAM_MINB 7     // resident CTAs per SM the register allocation aims at (<= 73 registers, no spills)
#endif
// per-warp ring + barriers of the TMA variant; the plain variant declares no shared memory at all
template <bool TMA, int NST, uint32_t BYTES>
struct StreamRing {
    static __device__ __forceinline__ unsigned char* ring(uint32_t) { return nullptr; }
    static __device__ __forceinline__ SvMbar* bars(uint32_t) { return nullptr; }
};
template <int NST, uint32_t BYTES>
struct StreamRing<true, NST, BYTES> {
    static __device__ __forceinline__ unsigned char* ring(uint32_t w) {
        alignas(128) __shared__ unsigned char s_ring[SV_STREAM_WARPS][NST][BYTES];
        return &s_ring[w][0][0];
    }
    static __device__ __forceinline__ SvMbar* bars(uint32_t w) {
        alignas(8) __shared__ SvMbar s_bar[SV_STREAM_WARPS][NST];
        return &s_bar[w][0];
    }
};
#ifndef SV_TMA_STAGES
#define SV_TMA_STAGES 6      // ring slots per warp (one slot = one batch: 32 U residuals + 32 U column ids); SV_TMA_STAGES - 2 batches in flight
#endif
// TMA (experiment, SVBFM_STREAM_TMA=1; all-ones data, no index list): the two streams of a warp's tile (residuals, other-column
// ids) are staged through a per-warp shared-memory ring by 1-D bulk copies that lane 0 issues SV_TMA_STAGES - 2 batches
// ahead; the lanes read their entries with LDS once the slot's mbarrier completes. Deeper prefetch than the register batch
// (bytes in flight per SM: 24 warps x 4 x 768 B instead of 24 x 768 B) at no register cost. The stores stay STG.
template <int KIND, bool ONES, bool REDUCE, bool STEADY, bool IDX = false, bool TMA = false>
__global__ void __launch_bounds__(32 * SV_STREAM_WARPS, ONES ? SV_STREAM_MINB : SV_STREAM_MINB - 1) k_stream(StreamArgs a) {      // x != 1: two more streams
    static_assert(!TMA || (ONES && !IDX), "the TMA variant covers the all-ones streams without an index list");
    constexpr bool IS_V = (KIND == KIND_VB_V || KIND == KIND_MC_V);
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int U = SV_STREAM_U;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= a.ntiles) return;
    // entry positions relative to the run (32 bit)
    const uint32_t q_begin = t << a.ts_shift;
    const uint32_t q_end = (a.n - q_begin > (1u << a.ts_shift)) ? q_begin + (1u << a.ts_shift) : a.n;
    const uint32_t* __restrict__ ocp = a.oc + a.real0;
    const float* __restrict__ xvp = ONES ? nullptr : a.xv + a.real0;
    const float* __restrict__ xop = ONES ? nullptr : a.xo + a.real0;
    const uint32_t* __restrict__ idxp = IDX ? a.idx + a.entry0 : nullptr;
    double* __restrict__ ep = a.e;
    const bool has_own = STEADY ? true : (a.has_own != 0), own_is_w = STEADY ? false : (a.own_is_w != 0);
    const bool has_oth = STEADY ? true : (a.has_oth != 0), oth_is_w = STEADY ? false : (a.oth_is_w != 0);
    const bool pend = has_own | has_oth;
    const bool need_rec = (IS_V && REDUCE) || has_oth || (has_own && !own_is_w);
    const bool rec_na = a.rec_no_alloc != 0;

    // next batch of the streams (issued one batch ahead)
    uint32_t oc_n[U]; float xs_n[U], xo_n[U]; double e_n[U];
    uint32_t r_n[IDX ? U : 1];              // IDX: the real positions of the batch (needed again for the store)
    // TMA: per-warp ring of SV_TMA_STAGES slots {32 U doubles, 32 U ids}; batch b of the tile lives in slot b % SV_TMA_STAGES
    constexpr int NST = SV_TMA_STAGES;
    constexpr uint32_t SLOT_E = 32 * U * 8, SLOT_BYTES = 32 * U * 12;
    unsigned char* const my_ring = StreamRing<TMA, NST, SLOT_BYTES>::ring(threadIdx.x >> 5);   // null unless TMA (no shared memory then)
    SvMbar* const my_bar = StreamRing<TMA, NST, SLOT_BYTES>::bars(threadIdx.x >> 5);
    auto tma_issue = [&](uint32_t bn) {     // lane 0: arm the slot of batch bn and start its two copies (whole batches only)
        const uint32_t qn = q_begin + bn * (32 * U);
        if (qn >= q_end || q_end - qn < 32 * U) return;
        SvMbar* bar = my_bar + bn % NST;
        unsigned char* dst = my_ring + (size_t)(bn % NST) * SLOT_BYTES;
        sv_mbar_arrive_expect_tx(bar, SLOT_BYTES);
        sv_bulk_g2s(dst, ep + qn, SLOT_E, bar);
        sv_bulk_g2s(dst + SLOT_E, ocp + qn, SLOT_BYTES - SLOT_E, bar);
    };
    if constexpr (TMA) {
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NST; k++) sv_mbar_init(my_bar + k, 1);
            sv_fence_mbar_init();
#pragma unroll
            for (int k = 0; k < NST; k++) tma_issue((uint32_t)k);
        }
        __syncwarp();
    }
    auto load_batch = [&](uint32_t q) {
        if (TMA && q_end - q >= 32 * U) {   // a whole batch out of the ring
            const uint32_t bi = (q - q_begin) / (32 * U), slot = bi % NST;
            sv_mbar_wait(my_bar + slot, (bi / NST) & 1u);
            const double* se = reinterpret_cast<const double*>(my_ring + (size_t)slot * SLOT_BYTES);
            const uint32_t* so = reinterpret_cast<const uint32_t*>(my_ring + (size_t)slot * SLOT_BYTES + SLOT_E);
#pragma unroll
            for (int u = 0; u < U; u++) {
                oc_n[u] = so[u * 32 + lane];
                xs_n[u] = 1.0f; xo_n[u] = 1.0f;
                e_n[u] = se[u * 32 + lane];
            }
            // the slot of batch bi - 2 is free: every lane has used its values (process() of that batch lies behind us)
            __syncwarp();
            if (bi >= 2 && lane == 0) tma_issue(bi - 2 + NST);
        } else if (q_end - q >= 32 * U) {          // a whole batch: no predicates
#pragma unroll
            for (int u = 0; u < U; u++) {
                uint32_t k = q + u * 32 + lane;
                if constexpr (IDX) { k = __ldcs(idxp + k); r_n[u] = k; }
                oc_n[u] = __ldcs(ocp + k);
                xs_n[u] = 1.0f; xo_n[u] = 1.0f;
                if constexpr (!ONES) { xs_n[u] = __ldcs(xvp + k); xo_n[u] = __ldcs(xop + k); }
                e_n[u] = __ldcs(ep + k);
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) {
                uint32_t k = q + u * 32 + lane;
                bool ok = k < q_end;
                if constexpr (IDX) { k = ok ? __ldcs(idxp + k) : 0u; r_n[u] = k; }
                oc_n[u] = ok ? __ldcs(ocp + k) : 0u;
                xs_n[u] = 1.0f; xo_n[u] = 1.0f;
                if constexpr (!ONES) if (ok) { xs_n[u] = __ldcs(xvp + k); xo_n[u] = __ldcs(xop + k); }
                e_n[u] = ok ? __ldcs(ep + k) : 0.0;
            }
        }
    };
    load_batch(q_begin);

    // window: lane l holds the constants of column jb + l
    uint32_t j = __ldg(&a.tile_col0[t]), jb = 0;
    uint32_t w_next = 0;                     // relative end of the column (clamped: ~0u beyond the run)
    double w_mu = 0.0, w_h = 0.0, w_d = 0.0;
    auto load_window = [&](uint32_t base) {
        jb = base;
        uint32_t cj = base + lane;
        if (cj < a.c1) {
            w_next = (uint32_t)(__ldg(&a.colptr[cj + 1]) - a.entry0);
            OwnPack o = a.own[cj];
            w_mu = o.mu_red; w_h = o.h_oth; w_d = o.d_own;
        } else { w_next = ~0u; w_mu = w_h = w_d = 0.0; }
    };
    uint32_t cur_b = (uint32_t)(__ldg(&a.colptr[j]) - a.entry0), next_b;
    double mu, h_oth, d_own;
    auto select = [&]() {
        int s = (int)(j - jb);
        next_b = __shfl_sync(FULL, w_next, s);
        mu = __shfl_sync(FULL, w_mu, s); h_oth = __shfl_sync(FULL, w_h, s); d_own = __shfl_sync(FULL, w_d, s);
    };
    // column that holds entry `pos` (the columns between j and it are empty); pos < n
    auto advance = [&](uint32_t pos) {
        j++;
        while (true) {
            if (j >= jb + 32) load_window(j);
            unsigned m = __ballot_sync(FULL, w_next > pos) & (FULL << (j - jb));
            if (m) { j = jb + (uint32_t)__ffs(m) - 1; break; }
            // a whole window of empty columns: binary search for the largest column with colptr[] <= pos
            uint32_t lo = jb + 31, hi = a.c1;
            while (hi - lo > 1) {
                uint32_t mid = lo + (hi - lo) / 2;
                if (__ldg(&a.colptr[mid]) - a.entry0 <= (uint64_t)pos) lo = mid; else hi = mid;
            }
            j = lo;
            load_window(j);
        }
        cur_b = pos;
        select();
    };
    load_window(j);
    select();

    double A = 0.0, B = 0.0, C1 = 0.0, C2 = 0.0;
    auto emit = [&](uint32_t cb, uint32_t nb) {    // sums of the part of column j inside this tile -> {A, B, C1, C2}
        bool whole = (cb >= q_begin) && (nb <= q_end);
        double* out = whole ? a.colsum + (size_t)j * 4 : a.partial + ((size_t)t * 2 + (cb < q_begin ? 0 : 1)) * 4;
        if constexpr (KIND == KIND_VB_V) {          // slot B stays 0: B = C1 + C2 (k_finalize)
            double k = warp_sum4(A, 0.0, C1, C2, lane);
            if ((lane & 7) == 0) out[lane >> 3] = k;
        } else {
            double k = warp_sum2(A, B, lane);
            if ((lane & 15) == 0) out[lane >> 4] = k;
        }
        A = B = C1 = C2 = 0.0;
    };
    bool done = false;

    // arithmetic of one batch (entries q0 .. q0 + 32 U) once its streams and records are in registers
    auto process = [&](uint32_t q0, const float (&xs)[U], const float (&xo)[U], const double (&es)[U], const uint32_t (&rr)[IDX ? U : 1],
                       const ColPack (&g)[U]) {
        const bool full = (q_end - q0 >= 32 * U);
        if (full && next_b >= q0 + 32 * U) {
            // the whole batch lies inside column j: no per-lane predicates
#pragma unroll
            for (int u = 0; u < U; u++) {
                double ei = apply_pending<ONES>(es[u], xs[u], xo[u], g[u], h_oth, d_own, has_own, own_is_w, has_oth, oth_is_w);
                if (pend) { uint32_t k = q0 + u * 32 + lane; if constexpr (IDX) k = rr[u]; __stcs(ep + k, ei); }
                if constexpr (REDUCE) entry_terms<KIND>(ei, xs[u], xo[u], g[u], mu, A, B, C1, C2);
            }
            if (next_b == q0 + 32 * U) {             // column j ends exactly with the batch
                if constexpr (REDUCE) emit(cur_b, next_b);
                if (next_b >= q_end) done = true;
                else advance(next_b);
            }
            return;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t row_b = q0 + u * 32;
            if (row_b >= q_end || done) break;
            const uint32_t row_e = (q_end - row_b > 32) ? row_b + 32 : q_end;
            const uint32_t k = row_b + lane;
            double ei = es[u];
            uint32_t seg_b = row_b;
            while (true) {
                const uint32_t seg_e = next_b < row_e ? next_b : row_e;
                if (k >= seg_b && k < seg_e) {
                    ei = apply_pending<ONES>(ei, xs[u], xo[u], g[u], h_oth, d_own, has_own, own_is_w, has_oth, oth_is_w);
                    if constexpr (REDUCE) entry_terms<KIND>(ei, xs[u], xo[u], g[u], mu, A, B, C1, C2);
                }
                if (next_b > row_e) break;
                // column j ends inside (or at the end of) this row
                if constexpr (REDUCE) emit(cur_b, next_b);
                seg_b = next_b;
                if (seg_b >= q_end) { done = true; break; }
                advance(seg_b);
                if (seg_b >= row_e) break;
            }
            if (pend && k < row_e) { uint32_t kr = k; if constexpr (IDX) kr = rr[u]; __stcs(ep + kr, ei); }
        }
    };
    auto gather = [&](uint32_t q0, const uint32_t (&oc)[U], ColPack (&g)[U]) {
        if ((STEADY || need_rec) && q_end - q0 >= 32 * U) {
#pragma unroll
            for (int u = 0; u < U; u++) g[u] = sv_load_record(a.rec + oc[u], rec_na || oc[u] >= a.rec_hot_end);
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (need_rec && q0 + u * 32 + lane < q_end) g[u] = sv_load_record(a.rec + oc[u], rec_na || oc[u] >= a.rec_hot_end);
                else g[u] = ColPack{0.0, 0.0, 0.0, 0.0};
            }
        }
    };
    for (uint32_t q0 = q_begin; q0 < q_end && !done; q0 += 32 * U) {
        float xs[U], xo[U]; double es[U]; ColPack g[U];
        uint32_t rr[IDX ? U : 1];
#pragma unroll
        for (int u = 0; u < U; u++) { xs[u] = xs_n[u]; xo[u] = xo_n[u]; es[u] = e_n[u]; if constexpr (IDX) rr[u] = r_n[u]; }
        gather(q0, oc_n, g);
        if (q_end - q0 > 32 * U) load_batch(q0 + 32 * U);
        process(q0, xs, xo, es, rr, g);
    }
    if constexpr (REDUCE) if (!done) emit(cur_b, next_b);   // column j continues in the next tile
}

// ---- k_stream_rows: the same pass for streams whose columns are a handful of entries (a packed vb_online batch: 2 M entries of the
// 200 M shape end ~340 k columns, six entries per column). k_stream walks the columns of a tile one after the other: every column
// end is a dependent chain of a warp reduction and a window step (~60 instructions, ~250 cycles), 750 instructions per 32 entries on
// such a batch (ncu, profiles/r02_n_*). Here every lane knows the column of its entry (a.ownc, packed with the batch), fetches that
// column's constants itself, and a row of 32 entries is reduced at once:
//   * the column that is open at the start of the row keeps per-lane sums across rows and is reduced when it ends (exactly
//     k_stream's arithmetic: the same per-lane sums, the same butterfly);
//   * the columns that begin and end inside the row take one segmented scan over the lanes (as many steps as the longest of them
//     needs, whatever their number) and their last lanes write the sums;
//   * the row's last column becomes the open one.
// Same outputs as k_stream (colsum for a column inside the tile, partial[t][0 / 1] for the piece of a column that began before the
// tile / runs past it), so the finalize kernels do not care which of the two ran. Fixed order everywhere: reproducible.
// Measured on a B200 (profiles/r02_q_*, 2 M-entry batches): tiles of 128 / 256 / 512 entries 32.0 / 34.2 / 35.8 ms per 1020 passes
// (1024: 53.4); the gathers of row r + 1 issued before row r is reduced (streams two rows ahead, 64 registers): 33.7 against 34.2,
// dropped.
#ifndef SV_ROWS_MINB
#define SV_ROWS_MINB 8
#endif
template <int KIND, bool ONES, bool REDUCE, bool STEADY>
__global__ void __launch_bounds__(128, SV_ROWS_MINB) k_stream_rows(StreamArgs a) {
    static_assert(KIND == KIND_VB_W || KIND == KIND_VB_V, "vb / vb_online passes");
    constexpr bool IS_V = (KIND == KIND_VB_V);
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= a.ntiles) return;
    const uint32_t q_begin = t << a.ts_shift;
    const uint32_t q_end = (a.n - q_begin > (1u << a.ts_shift)) ? q_begin + (1u << a.ts_shift) : a.n;
    const uint32_t* __restrict__ ocp = a.oc + a.real0;
    const uint32_t* __restrict__ cp = a.ownc + a.real0;
    const float* __restrict__ xvp = ONES ? nullptr : a.xv + a.real0;
    const float* __restrict__ xop = ONES ? nullptr : a.xo + a.real0;
    double* __restrict__ ep = a.e;
    const bool has_own = STEADY ? true : (a.has_own != 0), own_is_w = STEADY ? false : (a.own_is_w != 0);
    const bool has_oth = STEADY ? true : (a.has_oth != 0), oth_is_w = STEADY ? false : (a.oth_is_w != 0);
    const bool pend = has_own | has_oth;
    const bool need_rec = (IS_V && REDUCE) || has_oth || (has_own && !own_is_w);

    // the open column: per-lane sums across rows
    uint32_t open_col = __ldg(cp + q_begin);
    bool open_prev = q_begin > 0 && __ldg(cp + q_begin - 1) == open_col;       // it began in an earlier tile
    double A = 0.0, B = 0.0, C1 = 0.0, C2 = 0.0;
    auto emit_open = [&](bool whole) {
        double* out = whole ? a.colsum + (size_t)open_col * 4 : a.partial + ((size_t)t * 2 + (open_prev ? 0 : 1)) * 4;
        if constexpr (KIND == KIND_VB_V) {          // slot B stays 0: B = C1 + C2 (finalize)
            double k = warp_sum4(A, 0.0, C1, C2, lane);
            if ((lane & 7) == 0) out[lane >> 3] = k;
        } else {
            double k = warp_sum2(A, B, lane);
            if ((lane & 15) == 0) out[lane >> 4] = k;
        }
        A = B = C1 = C2 = 0.0;
    };

    // streams of the next row, issued one row ahead
    uint32_t c_n, oc_n; double e_n; float xs_n = 1.0f, xo_n = 1.0f;
    auto load_row = [&](uint32_t row_b) {
        const uint32_t k = row_b + lane;
        const bool ok = k < q_end;
        c_n = ok ? __ldg(cp + k) : 0xffffffffu;
        oc_n = ok ? __ldg(ocp + k) : 0u;
        e_n = ok ? ep[k] : 0.0;
        if constexpr (!ONES) { xs_n = ok ? __ldg(xvp + k) : 1.0f; xo_n = ok ? __ldg(xop + k) : 1.0f; }
    };
    load_row(q_begin);
    for (uint32_t row_b = q_begin; row_b < q_end; row_b += 32) {
        const uint32_t k = row_b + lane;
        const bool ok = k < q_end;
        const uint32_t c = c_n, oc = oc_n;
        const double e0 = e_n;
        const float xs = xs_n, xo = xo_n;
        // constants of the lane's own column, record of its other column
        double mu = 0.0, h_oth = 0.0, d_own = 0.0;
        if (ok) {
            const double2 m = *reinterpret_cast<const double2*>(&a.own[c]);
            mu = m.x; h_oth = m.y; d_own = a.own[c].d_own;
        }
        ColPack g = ColPack{0.0, 0.0, 0.0, 0.0};
        if (need_rec && ok) g = sv_load_record(a.rec + oc, a.rec_no_alloc != 0);
        if (row_b + 32 < q_end) load_row(row_b + 32);
        double ei = e0;
        if (ok) {
            ei = apply_pending<ONES>(e0, xs, xo, g, h_oth, d_own, has_own, own_is_w, has_oth, oth_is_w);
            if (pend) ep[k] = ei;
        }
        if constexpr (REDUCE) {
            double tA = 0.0, tB = 0.0, tC1 = 0.0, tC2 = 0.0;
            if (ok) entry_terms<KIND>(ei, xs, xo, g, mu, tA, tB, tC1, tC2);
            const bool in_open = (c == open_col);
            if (in_open) { A += tA; B += tB; C1 += tC1; C2 += tC2; }
            const uint32_t last_lane = (q_end - row_b >= 32) ? 31u : q_end - row_b - 1;
            const uint32_t c_last = __shfl_sync(FULL, c, (int)last_lane);
            if (c_last != open_col) {                       // the open column ends inside this row (uniform)
                emit_open(!open_prev);
                const bool in_last = ok && (c == c_last);
                const bool interior = ok && !in_open && !in_last;
                if (__any_sync(FULL, interior)) {
                    // segmented inclusive scan over the lanes (the entries are in column order). Bit l of H: lane l starts a column;
                    // pos: the lane's distance from the head of its column; the scan takes steps only up to the longest interior
                    // column of the row (uniform) -- a handful of entries on the rows that have such columns at all
                    const uint32_t c_dn = __shfl_up_sync(FULL, c, 1);
                    const unsigned H = __ballot_sync(FULL, lane == 0 || c_dn != c);
                    const uint32_t pos = lane - (31u - (uint32_t)__clz(H & (FULL >> (31u - lane))));
                    const uint32_t longest = __reduce_max_sync(FULL, interior ? pos : 0u);
                    double s0 = interior ? tA : 0.0, s1 = interior ? (IS_V ? tC1 : tB) : 0.0, s2 = interior ? tC2 : 0.0;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        if ((uint32_t)d > longest) break;
                        const bool take = pos >= (uint32_t)d;       // lane - d lies in the lane's own column
                        const double u0 = __shfl_up_sync(FULL, s0, d), u1 = __shfl_up_sync(FULL, s1, d);
                        if (take) { s0 += u0; s1 += u1; }
                        if constexpr (IS_V) { const double u2 = __shfl_up_sync(FULL, s2, d); if (take) s2 += u2; }
                    }
                    if (interior && lane < 31u && ((H >> (lane + 1)) & 1u)) {     // last lane of its column (an interior lane is never the row's last lane)
                        double2* out = reinterpret_cast<double2*>(a.colsum + (size_t)c * 4);
                        if constexpr (IS_V) { out[0] = make_double2(s0, 0.0); out[1] = make_double2(s1, s2); }
                        else out[0] = make_double2(s0, s1);
                    }
                }
                if (in_last) { A = tA; B = tB; C1 = tC1; C2 = tC2; }
                open_col = c_last; open_prev = false;
            }
        }
    }
    if constexpr (REDUCE) {
        const bool runs_on = q_end < a.n && __ldg(cp + q_end) == open_col;
        emit_open(!open_prev && !runs_on);
    }
}

// pass 2 in CASE order (streaming): the run's columns are case-disjoint, so every case has at most one feature
// inside [c0, c1); e_i += x * h * delta_j with delta and the parameters served from L2. Used when the run covers
// a large share of the cases; the CSC-order kernel above is kept for small runs.
struct RowApplyArgs {
    RowView rv;
    uint32_t n, c0, c1;
    double* e;
    const double2* pf;
    const double* delta;
    const uint16_t* rbatch;   // vb_online: batch id per case (null otherwise)
    uint32_t batch;
};

template <bool IS_V, int FT, bool ONES>
__global__ void __launch_bounds__(256) k_row_apply(RowApplyArgs a) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        if (a.rbatch && __ldg(&a.rbatch[i]) != a.batch) continue;
        if constexpr (FT == 2) {
            uint2 c = __ldg(reinterpret_cast<const uint2*>(a.rv.rcol) + i);
            bool in0 = (c.x >= a.c0) & (c.x < a.c1), in1 = (c.y >= a.c0) & (c.y < a.c1);
            if (!(in0 | in1)) continue;
            uint32_t j = in0 ? c.x : c.y, o = in0 ? c.y : c.x;
            double d = __ldg(&a.delta[j]);
            if (d == 0.0) continue;
            float xj = 1.0f, xo = 1.0f;
            if constexpr (!ONES) {
                float2 xv = __ldg(reinterpret_cast<const float2*>(a.rv.rval) + i);
                xj = in0 ? xv.x : xv.y; xo = in0 ? xv.y : xv.x;
            }
            double hh = xj;
            if constexpr (IS_V) hh = xj * (__ldg(&a.pf[o]).x * xo);
            a.e[i] += hh * d;
        } else {
            uint64_t b, e_;
            if constexpr (FT == 1) { b = (uint64_t)i * a.rv.F; e_ = b + a.rv.F; }
            else { b = __ldg(&a.rv.rowptr[i]); e_ = __ldg(&a.rv.rowptr[i + 1]); }
            uint64_t own = e_;
            for (uint64_t k = b; k < e_; k++) {
                uint32_t c = __ldg(&a.rv.rcol[k]);
                if (c >= a.c0 && c < a.c1) { own = k; break; }
            }
            if (own == e_) continue;
            uint32_t j = __ldg(&a.rv.rcol[own]);
            double d = __ldg(&a.delta[j]);
            if (d == 0.0) continue;
            float xj = 1.0f;
            if constexpr (!ONES) xj = __ldg(&a.rv.rval[own]);
            double hh = xj;
            if constexpr (IS_V) {
                double h = 0.0;
                for (uint64_t k = b; k < e_; k++) {
                    if (k == own) continue;
                    double m = __ldg(&a.pf[__ldg(&a.rv.rcol[k])]).x;
                    if constexpr (ONES) h += m; else h += m * __ldg(&a.rv.rval[k]);
                }
                hh = xj * h;
            }
            a.e[i] += hh * d;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// dense passes over the residuals
#ifndef SV_RGRID
#define SV_RGRID 1184   // 148 SMs x 8 (tests/emu builds with a smaller grid: every CUDA thread is a fiber there)
#endif

// partial[b*3 + {0,1,2}] = sum e, sum e^2, sum clamp(e)^2  (vb.h:513, :451, vbs.h:153-162)
__global__ void __launch_bounds__(256) k_reduce_e(const double* __restrict__ e, uint32_t n, const Scalars* sc, double* __restrict__ partial,
                                                  const uint16_t* __restrict__ rbatch, uint32_t batch) {
    __shared__ double sm[3 * 32];
    double lo = sc->min_target, hi = sc->max_target;
    double v[3] = {0, 0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (rbatch && rbatch[i] != batch) continue;
        double x = e[i];
        v[0] += x; v[1] += x * x;
        double p = fmax(lo, fmin(hi, x));
        v[2] += p * p;
    }
    block_sum<3>(v, sm);
    if (threadIdx.x == 0) { partial[blockIdx.x * 3 + 0] = v[0]; partial[blockIdx.x * 3 + 1] = v[1]; partial[blockIdx.x * 3 + 2] = v[2]; }
}

// out[k] (+)= sum_b partial[b*nv + k]
__global__ void __launch_bounds__(256) k_reduce_final(const double* __restrict__ partial, uint32_t nblocks, uint32_t nv, double* __restrict__ out, int accumulate) {
    __shared__ double sm[32];
    for (uint32_t k = 0; k < nv; k++) {
        double v[1] = {0.0};
        for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) v[0] += partial[(size_t)b * nv + k];
        block_sum<1>(v, sm);
        if (threadIdx.x == 0) out[k] = accumulate ? out[k] + v[0] : v[0];
    }
}

__global__ void __launch_bounds__(256) k_shift_e(double* __restrict__ e, uint32_t n, const Scalars* sc) {
    double d = sc->w0_delta;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) e[i] += d;
}

// vb_online on the stream schedule: the same two passes over the entries of one batch only (list = their positions in e)
__global__ void __launch_bounds__(256) k_reduce_e_list(const double* __restrict__ e, const uint32_t* __restrict__ list, uint32_t n, const Scalars* sc,
                                                       double* __restrict__ partial) {
    __shared__ double sm[3 * 32];
    double lo = sc->min_target, hi = sc->max_target;
    double v[3] = {0, 0, 0};
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        double x = e[__ldcs(&list[k])];
        v[0] += x; v[1] += x * x;
        double p = fmax(lo, fmin(hi, x));
        v[2] += p * p;
    }
    block_sum<3>(v, sm);
    if (threadIdx.x == 0) { partial[blockIdx.x * 3 + 0] = v[0]; partial[blockIdx.x * 3 + 1] = v[1]; partial[blockIdx.x * 3 + 2] = v[2]; }
}
__global__ void __launch_bounds__(256) k_shift_e_list(double* __restrict__ e, const uint32_t* __restrict__ list, uint32_t n, const Scalars* sc) {
    double d = sc->w0_delta;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) e[__ldcs(&list[k])] += d;
}

// sum over columns of dT[j]; zeroes dT
__global__ void __launch_bounds__(256) k_reduce_dT(double* __restrict__ dT, uint32_t n, double* __restrict__ partial) {
    __shared__ double sm[32];
    double v[1] = {0.0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { v[0] += dT[i]; dT[i] = 0.0; }
    block_sum<1>(v, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = v[0];
}

__global__ void k_add_scalar(double* dst, const double* src) { *dst += *src; }

// ---- vb_online on the stream schedule: per epoch, the entries of every batch in column order ------------------------
// second residual copy for the entries of one batch: e2[p] = e[crow1[p]] for p in idx1[lo, hi)
__global__ void k_gather_e_idx(const double* __restrict__ e, const uint32_t* __restrict__ crow1, const uint32_t* __restrict__ idx, uint32_t lo,
                               uint32_t hi, double* __restrict__ e2) {
    uint32_t k = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= hi) return;
    uint32_t p = idx[k];
    e2[p] = e[crow1[p]];
}

// The batch in flight as contiguous streams. A batch is a random 1 / num_batch of the cases: read through idx, every 8-byte
// residual and 4-byte column id of a pass costs a 32-byte sector (ncu, 200 M ratings in 100 batches: 588 MB from HBM per pass
// for 40 MB of entries, profiles/r02_n_*). Packed once per batch (this kernel: one such scattered read), the 2 (K + 1) passes,
// the reductions and the w0 shifts of the batch stream them; k_vbo_unpack puts the residuals back at the end of the batch so
// that e / e2 read as before (svbfm_get_residuals, svbfm_copies_max_diff).
struct VboPackArgs {
    const double* e;               // residuals in device case order (fresh prediction of the batch's cases)
    const uint32_t* idx[2];        // the batch's entries of run 0 / run 1 (positions inside the run), already offset to the batch
    const uint32_t* crow1;         // case of every entry of run 1
    const uint32_t* oc[2];         // other-column ids of run 0 / run 1, offset to the run's first entry
    const float* xv[2];            // own x / other x of the runs (null: all ones)
    const float* xo[2];
    const uint32_t* rcol;          // [n][2] the two columns of every case (device case order)
    uint32_t n;
    double* eb[2];
    uint32_t* ocb[2];
    uint32_t* ownb[2];             // own column of every batch entry (what k_stream_rows keys its row reductions by)
    const uint32_t* cpos[2];       // compact columns (null: global ids): rank of (batch, column) among the epoch's non-empty pairs, offset to
                                   // this batch and indexed by global column id; ownb then holds cpos[ri][j] - csub[ri]
    uint32_t csub[2];
    float* xvb[2];
    float* xob[2];
};
__global__ void __launch_bounds__(256) k_vbo_pack(VboPackArgs a) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n) return;
    const uint32_t p0 = __ldcs(a.idx[0] + k), p1 = __ldcs(a.idx[1] + k);
    const uint32_t i1 = __ldg(a.crow1 + p1);
    a.eb[0][k] = a.e[p0];                          // run 0's entry order is the device case order
    a.eb[1][k] = a.e[i1];
    uint32_t j0 = __ldg(a.rcol + 2 * (size_t)p0), j1 = __ldg(a.rcol + 2 * (size_t)i1 + 1);
    if (a.cpos[0]) { j0 = __ldg(a.cpos[0] + j0) - a.csub[0]; j1 = __ldg(a.cpos[1] + j1) - a.csub[1]; }
    a.ownb[0][k] = j0;
    a.ownb[1][k] = j1;
    a.ocb[0][k] = __ldg(a.oc[0] + p0);
    a.ocb[1][k] = __ldg(a.oc[1] + p1);
    if (a.xv[0]) {
        a.xvb[0][k] = __ldg(a.xv[0] + p0); a.xob[0][k] = __ldg(a.xo[0] + p0);
        a.xvb[1][k] = __ldg(a.xv[1] + p1); a.xob[1][k] = __ldg(a.xo[1] + p1);
    }
}
__global__ void __launch_bounds__(256) k_vbo_unpack(const double* __restrict__ eb0, const double* __restrict__ eb1, const uint32_t* __restrict__ idx0,
                                                    const uint32_t* __restrict__ idx1, uint32_t n, double* __restrict__ e, double* __restrict__ e2) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    e[__ldcs(idx0 + k)] = eb0[k];
    e2[__ldcs(idx1 + k)] = eb1[k];
}

// sharded stream schedule: block exchange of run 0's parameters. row 0 = w (when present), then one row per factor.
struct BlockXchg {
    double2* pw;              // null: no w row
    double2* pv;
    uint32_t D, rows, maxcnt;
    int world, rank;
    uint32_t blk[17];
};
__device__ __forceinline__ double2* xchg_table(const BlockXchg& b, uint32_t row) {
    if (b.pw) return row == 0 ? b.pw : b.pv + (size_t)(row - 1) * b.D;
    return b.pv + (size_t)row * b.D;
}
__global__ void k_xchg_pack(BlockXchg b, double2* __restrict__ send) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    uint32_t lo = b.blk[b.rank], cnt = b.blk[b.rank + 1] - lo;
    if (i < cnt) send[(size_t)row * b.maxcnt + i] = xchg_table(b, row)[lo + i];
}
__global__ void k_xchg_unpack(BlockXchg b, const double2* __restrict__ recv) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    int r = blockIdx.z;
    if (r == b.rank) return;
    uint32_t lo = b.blk[r], cnt = b.blk[r + 1] - lo;
    if (i < cnt) xchg_table(b, row)[lo + i] = recv[((size_t)r * b.rows + row) * b.maxcnt + i];
}

// ---------------------------------------------------------------------------------------------------------
// case-wise prediction (replaces the 2K+1 CSC scatter passes of predict_data_and_write_to_eterms,
// fm_learn_vb.h:70-203 / fm_learn_mcmc.h:117-348, and of predict_t_and_write_to_qterms, vb.h:207-312)
enum { PRED_VB_TRAIN = 0, PRED_VB_TEST = 1, PRED_MC_TRAIN = 2, PRED_MC_TEST = 3, PRED_MC_TEST_CLASS = 4 };

// ---- binary classification (-task c, mcmc / als): the reference's own erf approximation (Abramowitz-Stegun 7.1.26,
// util/random.h:47-61) and probit link (random.h:67-69), restated so that the deterministic (als) path agrees to rounding
__device__ __forceinline__ double ref_erf(double x) {
    double t = (x >= 0) ? 1.0 / (1.0 + 0.3275911 * x) : 1.0 / (1.0 - 0.3275911 * x);
    double result = 1.0 - (t * (0.254829592 + t * (-0.284496736 + t * (1.421413741 + t * (-1.453152027 + t * 1.061405429))))) * exp(-x * x);
    return (x >= 0) ? result : -result;
}
__device__ __forceinline__ double ref_cdf_gaussian(double x) { return 0.5 + 0.5 * ref_erf(0.707106781 * x); }
__device__ __forceinline__ bool class_hit(double p, double y) { return ((p >= 0.5) && (y > 0.0)) || ((p < 0.5) && (y < 0.0)); }   // mcmcs.h:193, 333

struct PredictArgs {
    RowView rv;
    const float* y;
    uint32_t n;
    const double2* pw;
    const double2* pv;
    uint32_t D; int K, k0, k1;
    const Scalars* sc;
    double* e;            // train: residual out
    double* pred;         // test: prediction out (vb: clamped)
    double* pred_sum;     // mcmc test: running sum of clamped predictions
    double* partial;      // [grid][4]
    const uint16_t* rbatch;    // vb_online: restrict to the cases of the current batch
    uint32_t batch;
    const uint32_t* list;      // LIST: the cases to predict (vb_online on the stream schedule: the batch's own case list)
    uint32_t nlist;
};

template <int MODE, int FT, bool ONES, bool LIST = false>
__global__ void __launch_bounds__(256) k_predict(PredictArgs a) {
    __shared__ double sm[4 * 32];
    double acc[4] = {0, 0, 0, 0};
    const double w0 = a.sc->w0_mean, w0v = a.sc->w0_var;
    const double lo = a.sc->min_target, hi = a.sc->max_target;
    const double inv_it = 1.0 / (double)(a.sc->iter + 1);
    const uint32_t count = LIST ? a.nlist : a.n;
    for (uint32_t k0 = blockIdx.x * blockDim.x + threadIdx.x; k0 < count; k0 += gridDim.x * blockDim.x) {
        uint32_t i = k0;
        if constexpr (LIST) i = __ldcs(&a.list[k0]);
        else if (MODE == PRED_VB_TRAIN && a.rbatch && __ldg(&a.rbatch[i]) != a.batch) continue;
        uint64_t b, e_;
        uint32_t c2[2] = {0, 0}; float x2[2] = {1.0f, 1.0f};
        if constexpr (FT == 2) {
            uint2 c = __ldg(reinterpret_cast<const uint2*>(a.rv.rcol) + i);
            c2[0] = c.x; c2[1] = c.y;
            if constexpr (!ONES) { float2 xv = __ldg(reinterpret_cast<const float2*>(a.rv.rval) + i); x2[0] = xv.x; x2[1] = xv.y; }
            b = 0; e_ = 2;
        } else if constexpr (FT == 1) { b = (uint64_t)i * a.rv.F; e_ = b + a.rv.F; }
        else { b = __ldg(&a.rv.rowptr[i]); e_ = __ldg(&a.rv.rowptr[i + 1]); }
        auto col = [&](uint64_t k) -> uint32_t { if constexpr (FT == 2) return c2[k]; else return __ldg(&a.rv.rcol[k]); };
        auto val = [&](uint64_t k) -> float { if constexpr (ONES) return 1.0f; else if constexpr (FT == 2) return x2[k]; else return __ldg(&a.rv.rval[k]); };
        double lin = 0.0, tw = 0.0;
        for (uint64_t k = b; k < e_; k++) {
            double2 P = __ldg(&a.pw[col(k)]);
            float x = val(k);
            lin += P.x * x;                                     // vb.h:184
            if constexpr (MODE == PRED_VB_TRAIN) tw += P.y * x * x;   // vb.h:298
        }
        double pair = 0.0, sq = 0.0, tt = 0.0, tr = 0.0;
        for (int f = 0; f < a.K; f++) {
            const double2* pf = a.pv + (size_t)f * a.D;
            double s = 0.0, Q = 0.0, Z = 0.0;
            for (uint64_t k = b; k < e_; k++) {
                double2 P = __ldg(&pf[col(k)]);
                float x = val(k);
                s += P.x * x;                                   // vb.h:115
                sq += 0.5 * P.x * P.x * x * x;                  // vb.h:159
                if constexpr (MODE == PRED_VB_TRAIN) {
                    Q += P.x * x * P.x * x;                     // vb.h:241
                    Z += P.y * x * x;                           // vb.h:242
                    tr += (P.x * P.x * x * x * x * x * P.y + 0.5 * x * x * x * x * P.y * P.y);   // vb.h:277-278
                }
            }
            pair += 0.5 * s * s;                                // vb.h:128
            if constexpr (MODE == PRED_VB_TRAIN) tt += (0.5 * Z * Z + Z * Q);   // vb.h:250
        }
        double q_all = -sq;
        if (a.k1) q_all += lin;
        double yhat = pair + q_all;
        if (a.k0) yhat += w0;                                   // vb.h:196-199
        double y = (double)__ldg(&a.y[i]);
        if constexpr (MODE == PRED_VB_TRAIN) {
            double T = tt - tr;
            if (a.k1) T += tw;
            if (a.k0) T += w0v;                                 // vb.h:306-309
            a.e[i] = y - yhat;                                  // vbs.h:43
            acc[0] += T;
        } else if constexpr (MODE == PRED_VB_TEST) {
            double p = fmax(lo, fmin(hi, yhat));                // vbs.h:146-147
            a.pred[i] = p;
            double err = p - y;                                 // vbs.h:270-271
            acc[0] += err * err;
        } else if constexpr (MODE == PRED_MC_TRAIN) {
            double p = fmax(lo, fmin(hi, yhat));                // mcmcs.h:167-171
            double err = p - y;
            acc[0] += err * err;
            a.e[i] = yhat - y;                                  // mcmcs.h:172
        } else if constexpr (MODE == PRED_MC_TEST_CLASS) {
            double p = ref_cdf_gaussian(yhat);                  // mcmcs.h:178-180
            a.pred[i] = p;
            double s = a.pred_sum[i] + p;                       // mcmcs.h:181
            a.pred_sum[i] = s;
            if (class_hit(p, y)) acc[0] += 1.0;                 // acc_mcmc_this (_evaluate_class, mcmcs.h:379-397)
            if (class_hit(s * inv_it, y)) acc[1] += 1.0;        // acc_mcmc_all  (_evaluate_class_map without the MAP part, mcmcs.h:331-338)
        } else {
            a.pred[i] = yhat;                                   // mcmcs.h:156
            double p = fmax(lo, fmin(hi, yhat));
            double s = a.pred_sum[i] + p;                       // mcmcs.h:159
            a.pred_sum[i] = s;
            double e1 = p - y;                                  // rmse_this (mcmcs.h:240)
            double pm = fmax(lo, fmin(hi, s * inv_it));         // mcmcs.h:241 with normalizer 1/(i+1)
            double e2 = pm - y;
            acc[0] += e1 * e1; acc[1] += e2 * e2;
        }
    }
    block_sum<4>(acc, sm);
    if (threadIdx.x == 0)
        for (int k = 0; k < 4; k++) a.partial[blockIdx.x * 4 + k] = acc[k];
}

// ---- train prediction for two complete fields (the stream schedule's data shape) ------------------------------------
// k_predict walks the K factor rows of the [K][D] parameter matrix per case: K*F sector gathers from 1.3 GB (68 ms for
// 200 M cases, K = 50). Here the parameters are first transposed to [D][K] (k_transpose_params, 1 ms); a warp then takes
// one case at a time with its lanes over the factors: two coalesced K*16 B rows per case (the user's row is kept while
// the user does not change: device case order is sorted by user), two butterfly sums, and the 32 results of a block of
// cases are written coalesced. Same formulas as k_predict (vb.h:70-312 / mcmc.h:117-348), factor sums in lane order.
__global__ void k_transpose_params(const double2* __restrict__ pv, uint32_t D, int K, double2* __restrict__ pvT) {
    __shared__ double2 tile[32][33];
    uint32_t j0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        uint32_t f = f0 + r, j = j0 + threadIdx.x;
        if (f < (uint32_t)K && j < D) tile[r][threadIdx.x] = pv[(size_t)f * D + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        uint32_t j = j0 + r, f = f0 + threadIdx.x;
        if (f < (uint32_t)K && j < D) pvT[(size_t)j * K + f] = tile[threadIdx.x][r];
    }
}

// the means alone (mcmc / als carry no variances: half the bytes per gathered row)
__global__ void k_transpose_means(const double2* __restrict__ pv, uint32_t D, int K, double* __restrict__ pvTm) {
    __shared__ double tile[32][33];
    uint32_t j0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        uint32_t f = f0 + r, j = j0 + threadIdx.x;
        if (f < (uint32_t)K && j < D) tile[r][threadIdx.x] = pv[(size_t)f * D + j].x;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        uint32_t j = j0 + r, f = f0 + threadIdx.x;
        if (f < (uint32_t)K && j < D) pvTm[(size_t)j * K + f] = tile[threadIdx.x][r];
    }
}

struct Predict2Args {
    const uint32_t* rcol;     // [2n] CSR of the two-field cases
    const float* rval;        // [2n] or null
    const float* y;
    uint32_t n;
    const double2* pw;
    const double2* pvT;       // [D][K] {mean, variance} (vb)
    const double* pvTm;       // [D][K] means (mcmc)
    int K, k0, k1;
    const Scalars* sc;
    double* e;
    double* partial;          // [warps] sum T (vb) / sum of squared clamped errors (mcmc)
};

// G = lanes per case: 32 (one case per warp step) or 16 (two cases per warp step, one per half-warp: the fixed cost of a
// step -- shuffles of the case's columns, reduction, bookkeeping -- is shared by two cases)
template <bool MCMC, bool ONES, int NS, int G>
__global__ void __launch_bounds__(256) k_predict2(Predict2Args a) {
    static_assert(G == 32 || G == 16, "lanes per case");
    const uint32_t lane = threadIdx.x & 31, gl = lane & (G - 1), gbase = lane & ~(uint32_t)(G - 1);
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
    const uint32_t blocks = (a.n + 31) / 32;                       // blocks of 32 consecutive cases, dealt out contiguously
    const uint32_t b0 = (uint32_t)((uint64_t)blocks * w / nw), b1 = (uint32_t)((uint64_t)blocks * (w + 1) / nw);
    const double w0 = a.sc->w0_mean, w0v = a.sc->w0_var, lo = a.sc->min_target, hi = a.sc->max_target;
    double acc = 0.0;
    uint32_t u_prev = 0xffffffffu;
    double2 Pu[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) Pu[s] = make_double2(0.0, 0.0);
    for (uint32_t b = b0; b < b1; b++) {
        const uint32_t i = b * 32 + lane;
        const bool ok = i < a.n;
        uint2 c = ok ? __ldcs(reinterpret_cast<const uint2*>(a.rcol) + i) : make_uint2(0u, 0u);
        float2 xv = make_float2(1.0f, 1.0f);
        if constexpr (!ONES) if (ok) xv = __ldcs(reinterpret_cast<const float2*>(a.rval) + i);
        const double yi = ok ? (double)__ldcs(a.y + i) : 0.0;
        double lin = 0.0, tw = 0.0;
        if (ok) {
            double2 A = __ldg(&a.pw[c.x]), B = __ldg(&a.pw[c.y]);
            lin = A.x * xv.x + B.x * xv.y;                                  // vb.h:184
            tw = A.y * xv.x * xv.x + B.y * xv.y * xv.y;                     // vb.h:298
        }
        double res = 0.0;
        for (uint32_t k = 0; k < (uint32_t)G; k++) {
            const uint32_t kk = gbase + k;                                  // the case this group works on: held by lane kk
            const bool valid = b * 32 + kk < a.n;
            const uint32_t u = __shfl_sync(0xffffffffu, c.x, kk), j = __shfl_sync(0xffffffffu, c.y, kk);
            float xu = 1.0f, xj = 1.0f;
            if constexpr (!ONES) { xu = __shfl_sync(0xffffffffu, xv.x, kk); xj = __shfl_sync(0xffffffffu, xv.y, kk); }
            if (valid && u != u_prev) {
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    int f = (int)gl + G * s;
                    if constexpr (MCMC) Pu[s] = make_double2(f < a.K ? __ldg(&a.pvTm[(size_t)u * a.K + f]) : 0.0, 0.0);
                    else Pu[s] = f < a.K ? __ldg(&a.pvT[(size_t)u * a.K + f]) : make_double2(0.0, 0.0);
                }
                u_prev = u;
            }
            double yv = 0.0, tv = 0.0;
#pragma unroll
            for (int s = 0; s < NS; s++) {
                int f = (int)gl + G * s;
                double2 Pj = make_double2(0.0, 0.0);
                if constexpr (MCMC) { if (valid && f < a.K) Pj.x = __ldg(&a.pvTm[(size_t)j * a.K + f]); }
                else if (valid && f < a.K) Pj = __ldg(&a.pvT[(size_t)j * a.K + f]);
                double mu_u = Pu[s].x * xu, mu_j = Pj.x * xj;
                double sm = mu_u + mu_j;                                    // vb.h:115
                double q = mu_u * mu_u + mu_j * mu_j;                       // vb.h:159 / :241
                yv += 0.5 * sm * sm - 0.5 * q;                              // vb.h:128, 159
                if constexpr (!MCMC) {
                    double xu2 = (double)xu * xu, xj2 = (double)xj * xj;
                    double z = Pu[s].y * xu2 + Pj.y * xj2;                  // vb.h:242
                    tv += 0.5 * z * z + z * q                               // vb.h:250
                          - (mu_u * mu_u * xu2 * Pu[s].y + 0.5 * xu2 * xu2 * Pu[s].y * Pu[s].y)   // vb.h:277-278
                          - (mu_j * mu_j * xj2 * Pj.y + 0.5 * xj2 * xj2 * Pj.y * Pj.y);
                }
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {                           // sum over the lanes of the group
                yv += __shfl_xor_sync(0xffffffffu, yv, o);
                if constexpr (!MCMC) tv += __shfl_xor_sync(0xffffffffu, tv, o);
            }
            if (valid && gl == k) {                                         // lane kk keeps the result of its own case
                double yhat = yv;
                if (a.k1) yhat += lin;
                if (a.k0) yhat += w0;                                       // vb.h:196-199
                if constexpr (MCMC) {
                    double p = fmax(lo, fmin(hi, yhat));                    // mcmcs.h:167-171
                    acc += (p - yi) * (p - yi);
                    res = yhat - yi;                                        // mcmcs.h:172
                } else {
                    double T = tv;
                    if (a.k1) T += tw;
                    if (a.k0) T += w0v;                                     // vb.h:306-309
                    acc += T;
                    res = yi - yhat;                                        // vbs.h:43
                }
            }
        }
        if (ok) a.e[i] = res;
    }
    acc = warp_sum(acc);
    if (lane == 0) a.partial[w] = acc;
}

// mcmc without re-prediction: Train= from the cached residuals (yhat = e + y)
__global__ void __launch_bounds__(256) k_train_sse_from_e(const double* __restrict__ e, const float* __restrict__ y, uint32_t n, const Scalars* sc,
                                                          double* __restrict__ partial) {
    __shared__ double sm[4 * 32];
    double acc[4] = {0, 0, 0, 0};
    double lo = sc->min_target, hi = sc->max_target;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double yy = (double)y[i];
        double p = fmax(lo, fmin(hi, e[i] + yy));
        acc[0] += (p - yy) * (p - yy);
    }
    block_sum<4>(acc, sm);
    if (threadIdx.x == 0)
        for (int k = 0; k < 4; k++) partial[blockIdx.x * 4 + k] = acc[k];
}

// classification, after the re-prediction of train (e = yhat - y with y = -1 / +1): train accuracy, then the new latent
// target of every case -- a draw from the normal around yhat truncated to the side of the class (or, without sampling, its
// expected value) -- and e = yhat - target (mcmcs.h:188-221). Draws: Philox keyed by (seed, iteration, caller case id, rank),
// so they do not depend on the device case order.
struct CaseRng {
    uint64_t seed; uint32_t id, iter, salt, n;
    __device__ double uniform() { uint4 r = philox4x32_10(make_uint4(n++, id, iter, salt), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))); return u01(r.x, r.y); }
    __device__ double normal() { return philox_normal(seed, n++, id, iter, salt ^ 0x00010000u); }
    // standard normal truncated to [left, inf) (util/random.h:72-102: naive rejection for left <= 0, Robert's translated exponential else)
    __device__ double left_tgaussian(double left) {
        if (left <= 0.0) {
            double r;
            do { r = normal(); } while (r < left);
            return r;
        }
        const double alpha_star = 0.5 * (left + sqrt(left * left + 4.0));
        while (true) {
            double z = -log(1.0 - uniform()) / alpha_star + left;
            double d = z - alpha_star;
            d = exp(-(d * d) / 2);
            if (uniform() < d) return z;
        }
    }
};
__global__ void __launch_bounds__(256) k_mc_class_targets(double* __restrict__ e, const float* __restrict__ y, const uint32_t* __restrict__ perm, uint32_t n,
                                                          uint64_t seed, uint32_t rank, int do_sample, const Scalars* sc, double* __restrict__ partial) {
    __shared__ double sm[32];
    double acc[1] = {0.0};
    const uint32_t iter = sc->iter;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double yy = (double)y[i];
        const double mu = e[i] + yy;                               // yhat
        if (class_hit(ref_cdf_gaussian(mu), yy)) acc[0] += 1.0;    // mcmcs.h:191-195
        double r;                                                  // e = yhat - target
        if (do_sample) {
            CaseRng rng{seed, perm ? perm[i] : i, iter, 0x5eed0004u ^ (rank << 20), 0u};
            r = (yy >= 0.0) ? -rng.left_tgaussian(-mu) : rng.left_tgaussian(mu);   // random.h:104-114 with mean mu, stdev 1
        } else {
            double phi_minus_mu = exp(-mu * mu / 2.0) / sqrt(3.141 * 2);            // 3.141: the reference's pi (mcmcs.h:203, 213)
            double Phi_minus_mu = ref_cdf_gaussian(-mu);
            r = (yy >= 0.0) ? -(phi_minus_mu / (1 - Phi_minus_mu)) : phi_minus_mu / Phi_minus_mu;
        }
        e[i] = r;
    }
    block_sum<1>(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc[0];
}

// ---------------------------------------------------------------------------------------------------------
// per-group column sums for the hyper-parameter updates. row r of the grid's y dimension: r = 0 -> w, r = f+1 -> v_f
//   vb  : S = sum (mu^2 + sigma), L = sum log(sigma)         (vb.h:477-498, 665-677)
//   mcmc: S = sum v,              L = sum v^2                (mcmc.h:938-941, 979-982, 1020-1023, 1061-1064)
#define SV_GGRID 64
template <bool MCMC>
__global__ void __launch_bounds__(256) k_group_sums(const double2* __restrict__ pw, const double2* __restrict__ pv, uint32_t D, uint32_t G,
                                                    const uint32_t* __restrict__ group, double* __restrict__ partial /*[K+1][G][2][SV_GGRID]*/) {
    __shared__ double sm[2 * 32];
    uint32_t r = blockIdx.y;
    const double2* p = (r == 0) ? pw : pv + (size_t)(r - 1) * D;
    for (uint32_t g = 0; g < G; g++) {
        double v[2] = {0.0, 0.0};
        for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < D; j += gridDim.x * blockDim.x) {
            if (G > 1 && group[j] != g) continue;
            double2 P = p[j];
            if (MCMC) { v[0] += P.x; v[1] += P.x * P.x; }
            else { v[0] += P.x * P.x + P.y; v[1] += log(P.y); }
        }
        block_sum<2>(v, sm);
        if (threadIdx.x == 0) {
            size_t base = (((size_t)r * G + g) * 2) * SV_GGRID;
            partial[base + blockIdx.x] = v[0];
            partial[base + SV_GGRID + blockIdx.x] = v[1];
        }
    }
}
__global__ void __launch_bounds__(64) k_group_sums_final(const double* __restrict__ partial, uint32_t nrows /*(K+1)*G*2*/, double* __restrict__ out) {
    uint32_t r = blockIdx.x;
    if (r >= nrows) return;
    __shared__ double sm[32];
    double v[1] = {0.0};
    for (uint32_t b = threadIdx.x; b < SV_GGRID; b += blockDim.x) v[0] += partial[(size_t)r * SV_GGRID + b];
    block_sum<1>(v, sm);
    if (threadIdx.x == 0) out[r] = v[0];
}

// ---------------------------------------------------------------------------------------------------------
// scalar (single-thread) kernels: the O(1) / O(G*K) parts of update_all / draw_all

// update_w0 (vb.h:504-525). red[0] = global sum e
__global__ void k_vb_w0(Scalars* sc) {
    double N = sc->n_total;
    double sigma_old = sc->w0_var, mu_old = sc->w0_mean;
    sc->w0_var = 1.0 / (sc->sigma_0 + N * sc->alpha);
    double w0_temp = sc->red[0] + N * mu_old;              // sum_i (e_i + mu_0')
    sc->w0_mean = sc->w0_var * sc->alpha * w0_temp;
    sc->w0_delta = mu_old - sc->w0_mean;                   // vb.h:518
    sc->sum_t += N * (sc->w0_var - sigma_old);             // vb.h:519 summed over cases
}

// alpha, sigma_0, sigma_w, sigma_v, free energy (vb.h:446-500, 646-681). grp[(r*G+g)*2 + {0,1}] = S, L
// red[1] = global sum e^2, red[2] = global sum clamp(e)^2
__global__ void k_vb_hyper(Scalars* sc, const double* __restrict__ grp, const double* __restrict__ n_per_group, uint32_t G, int K,
                           double* __restrict__ sigma_w, double* __restrict__ sigma_v, DevStats* st) {
    double N = sc->n_total;
    double temp = sc->red[1] + sc->sum_t;                  // sum_i (e_i^2 + T_i)
    st->train_stat = sqrt(sc->red[2] / N);                 // vbs.h:162
    double alpha_old = sc->alpha;
    double alpha = N / temp;
    if (isnan(alpha) || isinf(alpha)) {                    // vb.h:456-469: revert and RETURN (no sigma updates, no free energy)
        sc->nan_inf += 1; sc->alpha = alpha_old; sc->alpha_ok = 0;
        st->has_fe = 0.0; st->free_energy = 0.0; st->alpha = alpha_old;
        return;
    }
    sc->alpha = alpha; sc->alpha_ok = 1;
    sc->sigma_0 = 1.0 / (sc->w0_mean * sc->w0_mean + sc->w0_var);      // vb.h:473
    double fe = -0.5 * alpha * temp - .5 * N * log(2 * 3.14 * (1.0 / alpha));          // vb.h:662-663 (3.14 is the reference's pi)
    fe += -0.5 * sc->sigma_0 * (sc->w0_mean * sc->w0_mean + sc->w0_var) + 0.5 * log(sc->w0_var * sc->sigma_0) + .5;   // vb.h:664
    for (uint32_t g = 0; g < G; g++) {
        double S = grp[((size_t)0 * G + g) * 2 + 0], L = grp[((size_t)0 * G + g) * 2 + 1], ng = n_per_group[g];
        double sw = ng / S;                                 // vb.h:482
        sigma_w[g] = sw;
        fe += -0.5 * sw * S + 0.5 * (L + ng * log(sw)) + .5 * ng;       // vb.h:665-669 summed over the group
    }
    for (int f = 0; f < K; f++)
        for (uint32_t g = 0; g < G; g++) {
            double S = grp[((size_t)(f + 1) * G + g) * 2 + 0], L = grp[((size_t)(f + 1) * G + g) * 2 + 1], ng = n_per_group[g];
            double sv = ng / S;                             // vb.h:496
            sigma_v[(size_t)g * K + f] = sv;
            fe += -0.5 * sv * S + 0.5 * (L + ng * log(sv)) + .5 * ng;   // vb.h:670-677
        }
    st->free_energy = fe; st->has_fe = 1.0; st->alpha = alpha;
}

// scalar RNG stream for the hyper-prior draws
struct ScalarRng {
    uint64_t seed; uint32_t iter, n;
    __device__ double uniform() { uint4 r = philox4x32_10(make_uint4(n++, 0xa11ce, iter, 0x5eed0002u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))); return u01(r.x, r.y); }
    __device__ double normal() { return philox_normal(seed, n++, 0xb0b, iter, 0x5eed0003u); }
    __device__ double gamma(double alpha) {                 // Marsaglia-Tsang, as random.h:118-144
        if (alpha < 1.0) { double u = uniform(); return gamma(alpha + 1.0) * pow(u, 1.0 / alpha); }
        double d = alpha - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d), x, v, u;
        do {
            do { x = normal(); v = 1.0 + c * x; } while (v <= 0.0);
            v = v * v * v;
            u = uniform();
        } while ((u >= (1.0 - 0.0331 * (x * x) * (x * x))) && (log(u) >= (0.5 * x * x + d * (1.0 - v + log(v)))));
        return d * v;
    }
    __device__ double gaussian(double mean, double sd) { if (sd == 0.0 || isnan(sd)) return mean; return mean + sd * normal(); }
};

// draw_alpha, draw_w0, draw_w_lambda, draw_w_mu, draw_v_lambda, draw_v_mu (mcmc.h:901-929, 628-668, 931-1089).
// red[0] = sum e, red[1] = sum e^2 (global); grp[(r*G+g)*2 + {0,1}] = sum v, sum v^2.
__global__ void k_mcmc_hyper(Scalars* sc, const double* __restrict__ grp, const double* __restrict__ n_per_group, uint32_t G, int K, int k0, int k1,
                             double* __restrict__ w_lambda, double* __restrict__ w_mu, double* __restrict__ v_lambda, double* __restrict__ v_mu,
                             uint64_t seed, int do_sample, int do_multilevel) {
    const double alpha_0 = 1.0, gamma_0 = 1.0, beta_0 = 1.0, mu_0 = 0.0, w0_mean_0 = 0.0;   // mcmc.h:1100-1107
    ScalarRng rng{seed, sc->iter, 0};
    double N = sc->n_total;
    unsigned bad = 0;
    // alpha (mcmc.h:901-929)
    if (!do_multilevel) sc->alpha = alpha_0;
    else {
        double a_old = sc->alpha;
        double a = rng.gamma((alpha_0 + N) / 2.0) / ((gamma_0 + sc->red[1]) / 2.0);
        if (isnan(a) || isinf(a)) { a = a_old; bad++; }
        sc->alpha = a;
    }
    // w0 (mcmc.h:628-668)
    sc->w0_delta = 0.0;
    if (k0) {
        double reg0 = sc->sigma_0, w0 = sc->w0_mean;
        double w0_mean = sc->red[0] - N * w0;               // sum_i (e_i - w0)
        double s2 = 1.0 / (reg0 + sc->alpha * N);
        w0_mean = -s2 * (sc->alpha * w0_mean - w0_mean_0 * reg0);
        double w0n = do_sample ? rng.gaussian(w0_mean, sqrt(s2)) : w0_mean;
        if (isnan(w0n) || isinf(w0n)) { bad++; }
        else { sc->w0_mean = w0n; sc->w0_delta = w0n - w0; }   // e -= (w0_old - w0)
    }
    if (k1) {
        if (do_multilevel) {                                // draw_w_lambda (mcmc.h:970-1007)
            for (uint32_t g = 0; g < G; g++) {
                double S1 = grp[((size_t)0 * G + g) * 2 + 0], S2 = grp[((size_t)0 * G + g) * 2 + 1], ng = n_per_group[g], m = w_mu[g];
                double gam = beta_0 * (m - mu_0) * (m - mu_0) + gamma_0 + (S2 - 2.0 * m * S1 + ng * m * m);
                double a = alpha_0 + ng + 1, old = w_lambda[g];
                double l = do_sample ? rng.gamma(a / 2.0) / (gam / 2.0) : a / gam;
                if (isnan(l) || isinf(l)) { l = old; bad++; w_lambda[g] = l; break; }
                w_lambda[g] = l;
            }
            for (uint32_t g = 0; g < G; g++) {              // draw_w_mu (mcmc.h:931-968)
                double S1 = grp[((size_t)0 * G + g) * 2 + 0], ng = n_per_group[g];
                double mean = (S1 + beta_0 * mu_0) / (ng + beta_0);
                double s2 = 1.0 / ((ng + beta_0) * w_lambda[g]), old = w_mu[g];
                double m = do_sample ? rng.gaussian(mean, sqrt(s2)) : mean;
                if (isnan(m) || isinf(m)) { w_mu[g] = old; bad++; break; }
                w_mu[g] = m;
            }
        } else for (uint32_t g = 0; g < G; g++) w_mu[g] = mu_0;
    }
    if (K > 0) {
        if (do_multilevel) {
            bool stop = false;
            for (int f = 0; f < K && !stop; f++)            // draw_v_lambda (mcmc.h:1051-1089)
                for (uint32_t g = 0; g < G; g++) {
                    double S1 = grp[((size_t)(f + 1) * G + g) * 2 + 0], S2 = grp[((size_t)(f + 1) * G + g) * 2 + 1], ng = n_per_group[g];
                    double m = v_mu[(size_t)g * K + f];
                    double gam = beta_0 * (m - mu_0) * (m - mu_0) + gamma_0 + (S2 - 2.0 * m * S1 + ng * m * m);
                    double a = alpha_0 + ng + 1, old = v_lambda[(size_t)g * K + f];
                    double l = do_sample ? rng.gamma(a / 2.0) / (gam / 2.0) : a / gam;
                    if (isnan(l) || isinf(l)) { v_lambda[(size_t)g * K + f] = old; bad++; stop = true; break; }
                    v_lambda[(size_t)g * K + f] = l;
                }
            stop = false;
            for (int f = 0; f < K && !stop; f++)            // draw_v_mu (mcmc.h:1011-1049)
                for (uint32_t g = 0; g < G; g++) {
                    double S1 = grp[((size_t)(f + 1) * G + g) * 2 + 0], ng = n_per_group[g];
                    double mean = (S1 + beta_0 * mu_0) / (ng + beta_0);
                    double s2 = 1.0 / ((ng + beta_0) * v_lambda[(size_t)g * K + f]), old = v_mu[(size_t)g * K + f];
                    double m = do_sample ? rng.gaussian(mean, sqrt(s2)) : mean;
                    if (isnan(m) || isinf(m)) { v_mu[(size_t)g * K + f] = old; bad++; stop = true; break; }
                    v_mu[(size_t)g * K + f] = m;
                }
        } else for (size_t i = 0; i < (size_t)G * K; i++) v_mu[i] = mu_0;
    }
    sc->nan_inf += bad;
}

// ---- vb_online scalars (fm_learn_vb_online.h) ---------------------------------------------------------------
// start of a batch: batch size (global), fresh sum_t slot
__global__ void k_vbo_batch_begin(Scalars* sc, const double* __restrict__ batch_n, uint32_t b) {
    sc->batch_n = batch_n[b];
    sc->sum_t = 0.0;
    sc->w0_delta = 0.0;
}

// update_w0 (vbo.h:471-497). red[0] = sum over the batch of e_i. Every summand of eta2 is the same value and eta1 is
// linear in e_i, so the reference's per-case averages reduce to the expressions below.
__global__ void k_vbo_w0(Scalars* sc) {
    double bs = sc->batch_n, N = sc->n_total, rho = sc->rho_0;
    if (bs <= 0.0) { sc->w0_delta = 0.0; return; }
    double mu_dash = sc->w0_mean, sg_dash = sc->w0_var;
    double eta2 = (1.0 - rho) * sc->nat_sg_0 + rho * (sc->sigma_0 + N * sc->alpha);                     // vbo.h:483
    double eta1 = (1.0 - rho) * sc->nat_mu_0 + rho * N * sc->alpha * (sc->red[0] / bs + mu_dash);       // vbo.h:484, averaged (:488)
    sc->nat_mu_0 = eta1; sc->nat_sg_0 = eta2;
    sc->w0_mean = eta1 / eta2;                                                                          // vbo.h:490
    sc->w0_var = 1.0 / eta2;                                                                            // vbo.h:491
    sc->w0_delta = mu_dash - sc->w0_mean;                                                               // vbo.h:494
    sc->sum_t += bs * (sc->w0_var - sg_dash);                                                           // vbo.h:495 summed
}

// hyper-parameter blend + (first / last batch) free energy (vbo.h:412-467, 629-664). red[1] = batch sum e^2.
__global__ void k_vbo_hyper(Scalars* sc, const double* __restrict__ grp, const double* __restrict__ n_per_group, uint32_t G, int K,
                            double* __restrict__ sigma_w, double* __restrict__ sigma_v, DevStats* st, int want_fe, int first_batch, double lamda) {
    double bs = sc->batch_n, rho = sc->rho_0;
    double temp = sc->red[1] + sc->sum_t;
    double alpha_old = sc->alpha;
    double alpha = (1.0 - rho) * alpha_old + rho * (bs / temp);                                         // vbo.h:419
    bool ok = !(isnan(alpha) || isinf(alpha));
    if (!ok) { sc->nan_inf += 1; alpha = alpha_old; }                                                   // vbo.h:421-434: revert and return
    sc->alpha = alpha;
    if (ok) {
        sc->sigma_0 = (1.0 - rho) * sc->sigma_0 + rho * (1.0 / (sc->w0_mean * sc->w0_mean + sc->w0_var));   // vbo.h:438
        for (uint32_t g = 0; g < G; g++)
            sigma_w[g] = (1.0 - rho) * sigma_w[g] + rho * (n_per_group[g] / grp[((size_t)0 * G + g) * 2 + 0]);   // vbo.h:448
        for (int f = 0; f < K; f++)
            for (uint32_t g = 0; g < G; g++)
                sigma_v[(size_t)g * K + f] = (1.0 - rho) * sigma_v[(size_t)g * K + f] + rho * (n_per_group[g] / grp[((size_t)(f + 1) * G + g) * 2 + 0]);   // vbo.h:462
        sc->t_w0 += 1;                                                                                  // vbo.h:466-467
        sc->rho_0 = pow((double)(1u + sc->t_w0), -lamda);
    }
    if (want_fe) {                                                                                      // vbos.h:143-146 -> vbo.h:629-664
        double fe = -0.5 * sc->alpha * temp - .5 * bs * log(2 * 3.14 * (1.0 / sc->alpha));
        fe += -0.5 * sc->sigma_0 * (sc->w0_mean * sc->w0_mean + sc->w0_var) + 0.5 * log(sc->w0_var * sc->sigma_0) + .5;
        for (uint32_t g = 0; g < G; g++) {
            double S = grp[((size_t)0 * G + g) * 2 + 0], L = grp[((size_t)0 * G + g) * 2 + 1], ng = n_per_group[g];
            fe += -0.5 * sigma_w[g] * S + 0.5 * (L + ng * log(sigma_w[g])) + .5 * ng;
        }
        for (int f = 0; f < K; f++)
            for (uint32_t g = 0; g < G; g++) {
                double S = grp[((size_t)(f + 1) * G + g) * 2 + 0], L = grp[((size_t)(f + 1) * G + g) * 2 + 1], ng = n_per_group[g];
                double sv = sigma_v[(size_t)g * K + f];
                fe += -0.5 * sv * S + 0.5 * (L + ng * log(sv)) + .5 * ng;
            }
        if (first_batch) st->pad = fe;            // free energy of batch 1 (exported as free_energy_first)
        st->free_energy = fe; st->has_fe = 1.0;   // batch B overwrites: the epoch's last value
    }
    st->alpha = sc->alpha;
}

// vb_online: the caller's batch ids are checked on the device (the copy is needed anyway; a host loop over 2*10^8 ids is not)
__global__ void k_vbo_check_batch_ids(const uint32_t* __restrict__ batch_of_case, uint32_t n, uint32_t num_batch, uint32_t* __restrict__ flag) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && batch_of_case[i] >= num_batch) *flag = 1u;
}
// vb_online: batch id of every case in device order, and of the case of every CSC entry
__global__ void k_vbo_set_rbatch(const uint32_t* __restrict__ batch_of_case, const uint32_t* __restrict__ perm, uint32_t n, uint16_t* __restrict__ rbatch) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) rbatch[d] = (uint16_t)batch_of_case[perm ? perm[d] : d];
}
__global__ void k_vbo_set_cbatch(const uint32_t* __restrict__ crow, uint64_t nnz, const uint16_t* __restrict__ rbatch, uint16_t* __restrict__ cbatch) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nnz) cbatch[p] = rbatch[crow[p]];
}
__global__ void k_vbo_batch_counts(const uint16_t* __restrict__ rbatch, uint32_t n, unsigned long long* __restrict__ counts) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) atomicAdd(&counts[rbatch[d]], 1ull);     // integer atomics: order-independent
}
__global__ void k_u64_to_f64(const unsigned long long* __restrict__ in, uint32_t n, double* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}
__global__ void k_col_counts(const uint64_t* __restrict__ colptr, uint32_t ncols, uint32_t D, double* __restrict__ out) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < D) out[j] = j < ncols ? (double)(colptr[j + 1] - colptr[j]) : 0.0;
}
__global__ void k_nat_from_params(const double2* __restrict__ p, size_t n, double2* __restrict__ nat) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { double2 P = p[i]; nat[i] = make_double2(P.x / P.y, 1.0 / P.y); }      // vbo.h:750-765
}

// end of an iteration: evaluation numbers -> stats slot. red[3] = test sse (this), red[4] = test sse (running mean),
// red[5] = train sse (mcmc)
__global__ void k_finish_iter(Scalars* sc, DevStats* st, int method, int task = 0) {
    double nt = sc->nt_total, N = sc->n_total;
    if (method == SVBFM_MCMC && task == 1) {     // classification: accuracies instead of RMSEs (mcmcs.h:262-275)
        st->rmse_this = sc->red[3] / nt;         // acc_mcmc_this
        st->test_rmse = sc->red[4] / nt;         // "Test=" (accuracy of the running mean)
        st->train_stat = sc->red[5] / N;         // "Train="
        st->alpha = sc->alpha; st->has_fe = 0.0; st->free_energy = 0.0;
    } else if (method == SVBFM_MCMC) {
        st->rmse_this = sqrt(sc->red[3] / nt);
        st->test_rmse = sqrt(sc->red[4] / nt);
        st->train_stat = sqrt(sc->red[5] / N);
        st->alpha = sc->alpha; st->has_fe = 0.0; st->free_energy = 0.0;
    } else {
        st->test_rmse = sqrt(sc->red[3] / nt);
        st->rmse_this = st->test_rmse;
    }
    st->nan_inf = (double)sc->nan_inf;
    sc->nan_inf = 0;
    sc->iter += 1;
}

// stream schedule: records before the first step s0 (parameters p0): second-field columns are gathered by the first
// pass, every column needs its own constants
__global__ void k_pack_init(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, const double2* __restrict__ p0, ColPack* __restrict__ cpack,
                            OwnPack* __restrict__ opack, const uint32_t* __restrict__ rec_slot) {
    uint32_t j = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= b1) return;
    double2 P = p0[j];
    const uint32_t sj = rec_slot ? rec_slot[j] : j;
    if (j < a1) { opack[j] = OwnPack{P.x, 0.0, 0.0, 0.0}; cpack[sj] = ColPack{0.0, 0.0, 0.0, 0.0}; }
    else if (j >= b0) { opack[j] = OwnPack{P.x, P.x, 0.0, 0.0}; cpack[sj] = ColPack{P.x, P.y, 0.0, 0.0}; }
}
// before the second field's flush pass: h4 = the first field's mean of the last step
__global__ void k_pack_h4(uint32_t c0, uint32_t c1, const double2* __restrict__ p, ColPack* __restrict__ cpack) {
    uint32_t j = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j < c1) cpack[j].h4 = p[j].x;
}
// Cross shards: the columns of a field that OTHER ranks have just updated. Their new {mean, var} arrived in `stage` (slot order);
// this rank still holds the old values, so it forms what the owner's k_finalize formed -- delta from the same two operands, the
// record of the column for the passes that follow (write_records) -- and brings its parameter table up to date. The tables stay
// complete and bit-identical on every rank without a block exchange at the end of the iteration.
struct RemoteRecArgs {
    uint32_t n;                  // slots (= columns) of the field; the field's stage is [n] in slot order
    uint32_t bnd[17];            // rank q's block is the slots [bnd[q], bnd[q + 1])
    int world, me;
    const uint32_t* col_of_slot; // [slots of both fields] column of every record slot
    const double2* src[16];      // rank q's stage: its own block is current there (p2p: the peer's memory; else the local allgather buffer)
    uint32_t stage_base;         // first record slot of the field
    double2* pf;                 // parameters of this step ([D])
    const double2* p_next;
    const double2* p_prev;
    int rec_mode, mcmc;
    ColPack* cpack;
    PeerFlags flags;             // p2p: flag words of every rank (n = 0: the stage was filled by a collective)
    unsigned long long epoch;    // number of this exchange
};
__global__ void __launch_bounds__(256) k_records_remote(RemoteRecArgs a) {
    if (a.flags.n) {
        // this rank's k_finalize (earlier on the stream) has left its block in its stage: say so to every rank, then wait until every
        // other rank has said the same. Block 0 is dispatched first; the waiting blocks depend on other GPUs only.
        const int t = (int)threadIdx.x;
        if (blockIdx.x == 0 && t < a.flags.n && t != a.flags.me) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long*>(a.flags.p[t] + a.flags.me) = a.epoch;
        }
        if (t < a.flags.n && t != a.flags.me) {
            volatile unsigned long long* f = a.flags.p[a.flags.me] + t;
#if defined(SVBFM_EMULATED)
            while (*f < a.epoch) sched_yield();
#else
            // a rank that died never raises its flag: give up after ~20 s of polling (an L2 round trip each) instead of hanging the
            // GPU; word 16 of the own flag block counts the timeouts and the host turns them into SVBFM_ERR_NCCL
            unsigned long long polls = 0;
            while (*f < a.epoch) if (++polls > (1ull << 25)) { atomicAdd(a.flags.p[a.flags.me] + 16, 1ull); break; }
#endif
        }
        __syncthreads();       // the fetches below depend on the flag values through this barrier and bypass L1: no fence needed
    }
    // One thread per slot of the OTHER ranks' blocks; inside a block in slot order (the fetch from the peer's memory is coalesced).
    // The blocks are taken in the order me + 1, me + 2, ... (mod world): at any time every rank serves one reader instead of all of
    // them. Measured at 8 GPUs, 200 M ratings (ms per iteration; the exchange's share): this kernel 29.7 (10.0); every rank starting
    // at block 0 30.8 (11.1); the owner PUSHING its block into every rank's stage with coalesced 16-byte stores from a kernel of
    // its own 34.2 (14.3); in-place ncclAllGather of the same 16 B per column 31.1 (10.0) (profiles/r02_f_*, r02_h2_*, r02_h3_*,
    // r02_j_*): ~14 MB per rank and exchange move at 150 - 190 GB/s whichever way; the time follows a rank's remote slots.
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;      // index among the remote slots, rotated order
    int q = a.me;
    bool found = false;
    for (int i = 1; i < a.world; i++) {
        q = a.me + i; if (q >= a.world) q -= a.world;
        const uint32_t c = a.bnd[q + 1] - a.bnd[q];
        if (t < c) { found = true; break; }
        t -= c;
    }
    if (!found) return;
    t += a.bnd[q];                                           // slot inside the field
    const uint32_t sj = a.stage_base + t;
    const uint32_t j = a.col_of_slot[sj];
    const double2 nw = __ldcg(a.src[q] + t), old = a.pf[j];          // written by another GPU: not through L1
    const double dlt = a.mcmc ? (nw.x - old.x) : (old.x - nw.x);      // k_finalize: skip <=> the mean did not move <=> 0
    a.pf[j] = nw;
    const double2 N = a.p_next ? a.p_next[j] : make_double2(0.0, 0.0);
    if (a.rec_mode == 1) a.cpack[sj] = ColPack{nw.x, nw.y, dlt, a.p_prev ? a.p_prev[j].x : 0.0};
    else a.cpack[sj] = ColPack{N.x, N.y, dlt, old.x};
}
// the same in slot space (cross shards): after the first field's last finalize the mean of a record is that final mean
__global__ void k_pack_h4_self(uint32_t s0, uint32_t s1, ColPack* __restrict__ cpack) {
    uint32_t sl = s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (sl < s1) cpack[sl].h4 = cpack[sl].mu;
}
// first column that intersects each implicit tile of a run (largest j in [c0, c1) with colptr[j] <= first entry of the tile)
__global__ void k_tile_col0(const uint64_t* colptr, uint32_t c0, uint32_t c1, uint32_t ntiles, uint32_t ts_shift, uint32_t* __restrict__ out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    uint64_t p = colptr[c0] + ((uint64_t)t << ts_shift);
    uint32_t lo = c0, hi = c1;            // invariant: colptr[lo] <= p < colptr[hi]
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (colptr[mid] <= p) lo = mid; else hi = mid;
    }
    out[t] = lo;
}
// the second copy of the residuals, in the entry order of the second field: e2[p] = e[case of entry p]
__global__ void k_gather_e(const double* __restrict__ e, const uint32_t* __restrict__ crow, uint32_t n, double* __restrict__ e2) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) e2[p] = e[crow[p]];
}
__global__ void __launch_bounds__(256) k_copies_max_diff(const double* __restrict__ e, const uint32_t* __restrict__ crow, uint32_t n,
                                                         const double* __restrict__ e2, unsigned long long* __restrict__ out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    double d = fabs(e2[p] - e[crow[p]]);
    if (!(d == 0.0)) atomicMax(out, (unsigned long long)__double_as_longlong(d == d ? d : INFINITY));   // non-negative doubles order like integers
}

// state pack/unpack
__global__ void k_pack(const double* __restrict__ mean, const double* __restrict__ var, size_t n, double2* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double2(mean[i], var ? var[i] : 0.0);
}
__global__ void k_unpack(const double2* __restrict__ in, size_t n, double* __restrict__ mean, double* __restrict__ var) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { double2 p = in[i]; if (mean) mean[i] = p.x; if (var) var[i] = p.y; }
}
__global__ void k_fill_f64(double* a, size_t n, double v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}
__global__ void k_permute(const double* __restrict__ src, const uint32_t* __restrict__ perm, uint32_t n, double* __restrict__ dst) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) dst[d] = src[perm ? perm[d] : d];
}
__global__ void k_unpermute(const double* __restrict__ src, const uint32_t* __restrict__ perm, uint32_t n, double* __restrict__ dst) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) dst[perm ? perm[d] : d] = src[d];
}

}  // namespace svb
