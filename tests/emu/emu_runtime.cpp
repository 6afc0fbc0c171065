// tests/emu/emu_runtime.cpp -- TEST INFRASTRUCTURE (see include/cuda_runtime.h).
// Fiber scheduler of the lockstep emulator + the handful of runtime calls the engine makes. One OS thread; the CUDA
// threads of a CTA are fibers that run until they return or reach a warp / block primitive; CTAs run one at a time.
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <map>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

// SVBFM_EMU_BACKTRACE=1: print a native backtrace when an emulated kernel (or the engine) faults
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void emu_fault_handler(int sig) {
    void* frames[64];
    int n = backtrace(frames, 64);
    const char msg[] = "[emu] fatal signal, native backtrace:\n";
    (void)!write(2, msg, sizeof(msg) - 1);
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
__attribute__((constructor)) static void emu_install_fault_handler() {
    if (getenv("SVBFM_EMU_BACKTRACE")) { signal(SIGSEGV, emu_fault_handler); signal(SIGBUS, emu_fault_handler); signal(SIGABRT, emu_fault_handler); }
}

namespace {

#if defined(__x86_64__)
// void emu_switch(void** save_sp, void* load_sp): callee-saved registers on the stack, swap stack pointers
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");
#else
#error "the emulator's context switch is written for x86-64"
#endif

constexpr size_t STACK_BYTES = 256 * 1024;

struct Warp {
    uint64_t slot[2][32];
    unsigned arrived = 0, exited = 0, gen = 0;
};

struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = false;
    uint3 tid;
    int lane = 0, warp = 0;
};

struct Block {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    const std::function<void()>* body = nullptr;
    void* sched_sp = nullptr;
    int cur = -1;
    unsigned n = 0, exited = 0, bar_arrived = 0, bar_gen = 0;
} B;

std::vector<char*> g_stacks;     // reused between launches

void yield_to_scheduler() {
    Fiber& f = B.fibers[B.cur];
    emu_switch(&f.sp, B.sched_sp);
}

void release_warp_if_complete(Warp& w, unsigned mask) {
    if (w.arrived && ((w.arrived | w.exited) & mask) == mask) { w.arrived = 0; w.gen++; }
}
void release_block_if_complete() {
    if (B.bar_arrived && B.bar_arrived + B.exited == B.n) { B.bar_arrived = 0; B.bar_gen++; }
}

void fiber_main() {
    (*B.body)();
    Fiber& f = B.fibers[B.cur];
    f.done = true;
    B.exited++;
    Warp& w = B.warps[f.warp];
    w.exited |= 1u << f.lane;
    release_warp_if_complete(w, 0xffffffffu);    // every mask in csrc is the full mask
    release_block_if_complete();
    emu_switch(&f.sp, B.sched_sp);
    abort();   // never resumed
}

void run_block() {
    const unsigned n = blockDim.x * blockDim.y * blockDim.z;
    B.n = n; B.exited = 0; B.bar_arrived = 0; B.bar_gen = 0;
    B.fibers.assign(n, Fiber());
    B.warps.assign((n + 31) / 32, Warp());
    while (g_stacks.size() < n) g_stacks.push_back((char*)aligned_alloc(64, STACK_BYTES));
    for (unsigned t = 0; t < n; t++) {
        Fiber& f = B.fibers[t];
        f.tid.x = t % blockDim.x; f.tid.y = (t / blockDim.x) % blockDim.y; f.tid.z = t / (blockDim.x * blockDim.y);
        f.lane = t & 31; f.warp = t >> 5;
        f.stack = g_stacks[t];
        // initial frame: six callee-saved registers, then the "return address" fiber_main; on entry rsp % 16 == 8
        uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
        uint64_t* sp = (uint64_t*)(top - 8);      // slot a caller's `call` would have used
        *--sp = (uint64_t)(uintptr_t)&fiber_main; // popped by `ret`
        for (int r = 0; r < 6; r++) *--sp = 0;
        f.sp = sp;
    }
    // lanes a partial last warp does not have count as exited
    if (n & 31) B.warps.back().exited = ~0u << (n & 31);
    unsigned live = n;
    unsigned stuck_rounds = 0;
    while (live) {
        unsigned progressed = 0;
        for (unsigned t = 0; t < n; t++) {
            Fiber& f = B.fibers[t];
            if (f.done) continue;
            B.cur = (int)t;
            threadIdx = f.tid;
            const unsigned g0 = B.warps[f.warp].gen, b0 = B.bar_gen;
            emu_switch(&B.sched_sp, f.sp);
            if (f.done) { live--; progressed++; }
            else if (B.warps[f.warp].gen != g0 || B.bar_gen != b0) progressed++;
            else progressed += 0;
        }
        // a full round in which nothing arrived anywhere new cannot happen unless the kernel deadlocks (divergent barrier)
        if (!progressed) { if (++stuck_rounds > 4) { fprintf(stderr, "[emu] deadlock: a warp or block barrier was not reached by every thread\n"); abort(); } }
        else stuck_rounds = 0;
    }
    B.cur = -1;
}

}  // namespace

// stream capture (cudaStreamBeginCapture .. EndCapture): launches and async copies are recorded into a graph instead of executed
struct emuGraph { std::vector<std::function<void()>> nodes; };
static emuGraph* g_capture = nullptr;      // non-null while a capture is open (the engine keeps one stream of work at a time)
static bool g_capture_broken = false;      // a call that is illegal during capture was made

namespace emu {

int lane_id() { return B.fibers[B.cur].lane; }

static void warp_wait(Warp& w, unsigned mask, int lane) {
    const unsigned gen = w.gen;
    w.arrived |= 1u << lane;
    release_warp_if_complete(w, mask);
    while (w.gen == gen) yield_to_scheduler();
}

uint64_t warp_exchange(unsigned mask, uint64_t mine, int src_lane) {
    Fiber& f = B.fibers[B.cur];
    Warp& w = B.warps[f.warp];
    const unsigned buf = w.gen & 1;
    w.slot[buf][f.lane] = mine;
    warp_wait(w, mask, f.lane);
    return w.slot[buf][src_lane & 31];
}

unsigned warp_ballot(unsigned mask, bool pred) {
    Fiber& f = B.fibers[B.cur];
    Warp& w = B.warps[f.warp];
    const unsigned buf = w.gen & 1;
    const unsigned exited = w.exited;
    w.slot[buf][f.lane] = pred ? 1 : 0;
    warp_wait(w, mask, f.lane);
    unsigned r = 0;
    for (int l = 0; l < 32; l++)
        if ((mask >> l & 1) && !(exited >> l & 1) && w.slot[buf][l]) r |= 1u << l;
    return r;
}

void block_barrier() {
    const unsigned gen = B.bar_gen;
    B.bar_arrived++;
    release_block_if_complete();
    while (B.bar_gen == gen) yield_to_scheduler();
}

// SVBFM_EMU_PROFILE=1: host time and launch / thread counts per kernel name, printed at exit (where does an emulated test spend its time)
struct KernelStat { double seconds = 0; uint64_t launches = 0, threads = 0; };
static std::map<std::string, KernelStat>& kernel_stats() { static auto* m = new std::map<std::string, KernelStat>(); return *m; }   // never destroyed: read at exit
static void print_kernel_stats() {
    std::vector<std::pair<double, std::string>> v;
    for (auto& kv : kernel_stats()) v.push_back({kv.second.seconds, kv.first});
    std::sort(v.rbegin(), v.rend());
    fprintf(stderr, "[emu] %-28s %10s %10s %14s\n", "kernel", "seconds", "launches", "threads");
    for (auto& p : v) { const KernelStat& k = kernel_stats()[p.second]; fprintf(stderr, "[emu] %-28s %10.3f %10llu %14llu\n", p.second.c_str(), k.seconds, (unsigned long long)k.launches, (unsigned long long)k.threads); }
}

void launch(dim3 grid, dim3 block, size_t, cudaStream_t, const std::function<void()>& body, const char* name) {
    static const bool profile = getenv("SVBFM_EMU_PROFILE") != nullptr;
    static bool registered = false;
    if (profile && !registered) { registered = true; atexit(print_kernel_stats); }
    const auto t_begin = std::chrono::steady_clock::now();
    if (B.cur != -1) { fprintf(stderr, "[emu] nested launch\n"); abort(); }
    if (grid.x == 0 || grid.y == 0 || grid.z == 0 || block.x * block.y * block.z == 0 || block.x * block.y * block.z > 1024) {
        fprintf(stderr, "[emu] invalid launch configuration grid (%u,%u,%u) block (%u,%u,%u)\n", grid.x, grid.y, grid.z, block.x, block.y, block.z);
        abort();     // the real runtime reports cudaErrorInvalidConfiguration: a bug either way
    }
    if (g_capture) {        // recorded, not executed; the body holds its arguments by value ([=] in the rewritten launch)
        std::function<void()> copy = body;
        std::string nm = name;
        g_capture->nodes.push_back([grid, block, copy, nm] { emu::launch(grid, block, 0, nullptr, copy, nm.c_str()); });
        return;
    }
    gridDim = grid; blockDim = block;
    B.body = &body;
    for (uint32_t z = 0; z < grid.z; z++)
        for (uint32_t y = 0; y < grid.y; y++)
            for (uint32_t x = 0; x < grid.x; x++) {
                blockIdx.x = x; blockIdx.y = y; blockIdx.z = z;
                run_block();
            }
    B.body = nullptr;
    if (profile) {
        KernelStat& k = kernel_stats()[name];
        k.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
        k.launches++;
        k.threads += (uint64_t)grid.x * grid.y * grid.z * block.x * block.y * block.z;
    }
}

}  // namespace emu

// ------------------------------------------------------------------------------------------------ runtime calls
struct emuStream { int id; };

struct emuEvent { std::chrono::steady_clock::time_point t; };

const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "invalid value"; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "lockstep CPU emulator of sm_100 (tests/emu)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148; p->totalGlobalMem = (size_t)8 << 30;
    return cudaSuccess;
}
cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
// allocations of 64 KB and more live in a memfd so that another emulator process can map them (cudaIpc*)
namespace {
struct SharedBlock { int fd; size_t bytes; bool opened; };
std::map<void*, SharedBlock>& shared_blocks() { static std::map<void*, SharedBlock> m; return m; }
struct IpcRecord { char magic[8]; int pid, fd; size_t bytes; };
}  // namespace
// SVBFM_EMU_GUARD=1 ("electric fence"; compute-sanitizer is not available on the GPU pool): every allocation ends 0 .. 31 bytes
// in front of an inaccessible page (32: the alignment of the record structs), and is preceded by one, so a kernel that reads or writes past the end of a buffer (or in
// front of its page) stops the process with SIGSEGV at the access instead of reading a neighbour's bytes. IPC is off in this mode.
namespace {
struct GuardBlock { void* base; size_t bytes; };
std::map<void*, GuardBlock>& guard_blocks() { static std::map<void*, GuardBlock> m; return m; }
bool guard_mode() { static const bool on = getenv("SVBFM_EMU_GUARD") != nullptr; return on; }
}  // namespace
cudaError_t cudaMalloc(void** p, size_t n) {
    if (g_capture) g_capture_broken = true;     // not allowed while capturing
    if (guard_mode()) {
        const size_t data = (std::max<size_t>(n, 1) + 31) / 32 * 32, pages = (data + 4095) / 4096;
        const size_t bytes = (pages + 2) * 4096;
        unsigned char* base = static_cast<unsigned char*>(mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
        if (base == MAP_FAILED) return cudaErrorMemoryAllocation;
        mprotect(base, 4096, PROT_NONE);
        mprotect(base + (pages + 1) * 4096, 4096, PROT_NONE);
        unsigned char* d = base + (pages + 1) * 4096 - data;          // the buffer's (32-byte rounded) end is the guard page
        memset(d, 0xcd, data);
        guard_blocks()[d] = GuardBlock{base, bytes};
        *p = d;
        return cudaSuccess;
    }
    if (n >= 65536) {
        const size_t bytes = (n + 4095) / 4096 * 4096;
        int fd = memfd_create("svbfm_emu", 0);
        if (fd >= 0 && ftruncate(fd, (off_t)bytes) == 0) {
            void* m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            if (m != MAP_FAILED) {
                memset(m, 0xcd, n);
                shared_blocks()[m] = SharedBlock{fd, bytes, false};
                *p = m;
                return cudaSuccess;
            }
        }
        if (fd >= 0) close(fd);
    }
    *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256);
    if (!*p) return cudaErrorMemoryAllocation;
    memset(*p, 0xcd, n);       // "device memory" starts out as garbage, not zeros
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    if (guard_mode()) {
        auto g = guard_blocks().find(p);
        if (g != guard_blocks().end()) { munmap(g->second.base, g->second.bytes); guard_blocks().erase(g); return cudaSuccess; }
    }
    auto it = shared_blocks().find(p);
    if (it != shared_blocks().end()) { munmap(p, it->second.bytes); close(it->second.fd); shared_blocks().erase(it); return cudaSuccess; }
    free(p);
    return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
    auto it = shared_blocks().find(p);
    if (it == shared_blocks().end() || it->second.opened) return cudaErrorInvalidValue;
    IpcRecord r;
    memset(&r, 0, sizeof(r));
    memcpy(r.magic, "EMUIPC01", 8); r.pid = (int)getpid(); r.fd = it->second.fd; r.bytes = it->second.bytes;
    memset(h, 0, sizeof(*h));
    memcpy(h->reserved, &r, sizeof(r));
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
    IpcRecord r;
    memcpy(&r, h.reserved, sizeof(r));
    if (memcmp(r.magic, "EMUIPC01", 8) != 0) return cudaErrorInvalidValue;
    char path[64];
    snprintf(path, sizeof(path), "/proc/%d/fd/%d", r.pid, r.fd);
    int fd = open(path, O_RDWR);
    if (fd < 0) return cudaErrorInvalidValue;
    void* m = mmap(nullptr, r.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) { close(fd); return cudaErrorInvalidValue; }
    shared_blocks()[m] = SharedBlock{fd, r.bytes, true};
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) {
    auto it = shared_blocks().find(p);
    if (it == shared_blocks().end() || !it->second.opened) return cudaErrorInvalidValue;
    munmap(p, it->second.bytes); close(it->second.fd); shared_blocks().erase(it);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    if (g_capture) { g_capture->nodes.push_back([d, s, n] { if (n) memmove(d, s, n); }); return cudaSuccess; }
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
    if (g_capture) { g_capture->nodes.push_back([d, v, n] { if (n) memset(d, v, n); }); return cudaSuccess; }
    if (n) memset(d, v, n);
    return cudaSuccess;
}
cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = new emuStream{1}; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new emuStream{1}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) {
    if (g_capture) { g_capture_broken = true; return cudaErrorStreamCaptureUnsupported; }     // illegal while capturing, as on the GPU
    return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent{std::chrono::steady_clock::now()}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}

cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) {
    if (g_capture) return cudaErrorInvalidValue;
    g_capture = new emuGraph();
    g_capture_broken = false;
    return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* out) {
    if (!g_capture) { *out = nullptr; return cudaErrorInvalidValue; }
    emuGraph* g = g_capture;
    g_capture = nullptr;
    if (g_capture_broken) { delete g; *out = nullptr; return cudaErrorStreamCaptureUnsupported; }
    *out = g;
    return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* ex, cudaGraph_t g, unsigned long long) { *ex = new emuGraph(*g); return cudaSuccess; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t ex, cudaStream_t) {
    if (g_capture) return cudaErrorInvalidValue;
    for (auto& n : ex->nodes) n();
    return cudaSuccess;
}
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t ex) { delete ex; return cudaSuccess; }
cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
