// tests/emu/include/cuda_runtime.h -- TEST INFRASTRUCTURE, never part of the product.
//
// A lockstep CPU emulation of the small slice of CUDA that csrc/*.cu uses, so that the kernels' LOGIC (index arithmetic,
// warp-synchronous shuffles / ballots, block barriers, launch order of the engine) can be exercised on a machine without a
// GPU. tests/emu/build_emu.py rewrites `k<<<grid, block, smem, stream>>>(args)` into emu::launch(...) and compiles the
// unmodified sources against this header into tests/emu/_build/libsvbfm_emu.so. Every CUDA thread of a CTA is a fiber;
// a warp-level primitive is a barrier over the 32 fibers of the warp, __syncthreads() a barrier over the CTA; CTAs run
// one after the other. Nothing here says anything about speed, memory behaviour or data races between CTAs -- that is
// what the B200 runs (pytest -m gpu, bench.py, ncu) are for. The package never loads this library: libsvbfm.so (sm_100a)
// is the only product path and it fails loudly without a GPU.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <tuple>

#define SVBFM_EMULATED 1

// ---- qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#ifndef __restrict__
#define __restrict__ __restrict
#endif

// ---- vector types
struct uint2 { uint32_t x, y; } __attribute__((aligned(8)));
struct uint3 { uint32_t x, y, z; };
struct uint4 { uint32_t x, y, z, w; } __attribute__((aligned(16)));
struct float2 { float x, y; } __attribute__((aligned(8)));
struct double2 { double x, y; } __attribute__((aligned(16)));
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct dim3 {
    uint32_t x, y, z;
    dim3(uint32_t x_ = 1, uint32_t y_ = 1, uint32_t z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---- built-in variables (set by the scheduler whenever a fiber is resumed)
extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

// ---- runtime API (synchronous; "device memory" is host memory)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorStreamCaptureUnsupported = 900 };
typedef struct emuStream* cudaStream_t;
typedef struct emuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaLimit { cudaLimitMaxL2FetchGranularity = 5 };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem; };

const char* cudaGetErrorString(cudaError_t);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int*);
cudaError_t cudaGetDevice(int*);
cudaError_t cudaSetDevice(int);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp*, int);
cudaError_t cudaDeviceSetLimit(cudaLimit, size_t);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaMalloc(void**, size_t);
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
cudaError_t cudaFree(void*);
// CUDA IPC between emulator PROCESSES: allocations of 64 KB and more are backed by a memfd, a handle names (pid, fd, size) and
// the opening process maps /proc/<pid>/fd/<fd>: peer stores and flag words behave like memory shared over NVLink
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*);
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned);
cudaError_t cudaIpcCloseMemHandle(void*);
cudaError_t cudaMemcpy(void*, const void*, size_t, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void*, const void*, size_t, cudaMemcpyKind, cudaStream_t = nullptr);
cudaError_t cudaMemset(void*, int, size_t);
cudaError_t cudaMemsetAsync(void*, int, size_t, cudaStream_t = nullptr);
cudaError_t cudaStreamCreate(cudaStream_t*);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t*, unsigned);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0);
cudaError_t cudaEventCreate(cudaEvent_t*);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t*, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t);
cudaError_t cudaEventElapsedTime(float*, cudaEvent_t, cudaEvent_t);
// stream capture / graphs: while a capture is open, launches and async copies are recorded (not executed), like on the GPU
typedef struct emuGraph* cudaGraph_t;
typedef struct emuGraph* cudaGraphExec_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode);
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t*);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long = 0);
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t);
cudaError_t cudaGraphDestroy(cudaGraph_t);

// ---- the scheduler (emu_runtime.cpp)
namespace emu {
void launch(dim3 grid, dim3 block, size_t smem, cudaStream_t st, const std::function<void()>& body, const char* name = "");
uint64_t warp_exchange(unsigned mask, uint64_t mine, int src_lane);   // every lane deposits `mine`, returns the deposit of src_lane
unsigned warp_ballot(unsigned mask, bool pred);
void block_barrier();
int lane_id();
}  // namespace emu

// ---- warp / block primitives
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    int lane = emu::lane_id();
    int s = (lane & ~(width - 1)) | (src & (width - 1));
    raw = emu::warp_exchange(mask, raw, s);
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
    return __shfl_sync(mask, v, emu::lane_id() ^ lane_mask, width);
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    int lane = emu::lane_id();
    int s = ((lane & (width - 1)) + (int)d < width) ? lane + (int)d : lane;
    return __shfl_sync(mask, v, s, 32);
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    int lane = emu::lane_id();
    int s = ((lane & (width - 1)) >= (int)d) ? lane - (int)d : lane;
    return __shfl_sync(mask, v, s, 32);
}
static inline unsigned __ballot_sync(unsigned mask, bool pred) { return emu::warp_ballot(mask, pred); }
static inline int __any_sync(unsigned mask, bool pred) { return emu::warp_ballot(mask, pred) != 0u; }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    for (int o = 16; o > 0; o >>= 1) { unsigned w = __shfl_xor_sync(mask, v, o); v = w > v ? w : v; }
    return v;
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { (void)emu::warp_ballot(mask, false); }
static inline void __syncthreads() { emu::block_barrier(); }

// ---- loads / stores with cache hints: plain accesses
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { __sync_synchronize(); T v; memcpy(&v, p, sizeof(T)); return v; }
static inline void __threadfence_system() { __sync_synchronize(); }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
template <typename T> static inline void __stcg(T* p, T v) { *p = v; }

// ---- atomics (CTAs and fibers never run concurrently)
template <typename T, typename V> static inline T atomicAdd(T* p, V v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T, typename V> static inline T atomicMax(T* p, V v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <typename T, typename V> static inline T atomicMin(T* p, V v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <typename T, typename V> static inline T atomicOr(T* p, V v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <typename T, typename V> static inline T atomicExch(T* p, V v) { T o = *p; *p = (T)v; return o; }

// ---- arithmetic intrinsics
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long l) { double r; memcpy(&r, &l, 8); return r; }
static inline double cospi(double x) { return cos(M_PI * x); }
static inline double sinpi(double x) { return sin(M_PI * x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline double normcdf(double x) { return 0.5 * erfc(-x * M_SQRT1_2); }
