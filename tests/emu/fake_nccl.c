/* tests/emu/fake_nccl.c -- TEST INFRASTRUCTURE (see include/cuda_runtime.h): the nine NCCL entry points the engine binds with
 * dlsym, between PROCESSES on one host over a POSIX shared-memory segment. Built as tests/emu/_build/libnccl.so.2 and found
 * through LD_LIBRARY_PATH by the emulated engine's dlopen("libnccl.so.2"), so that the multi-rank launch order of the engine
 * (allreduce of the column sums, block exchange, agreement on the schedule at ingest) can be exercised on CPUs.
 * Every collective is synchronous: copy my piece into my slot, barrier, combine the slots in rank order (so that every rank
 * gets bit-identical results, like NCCL), barrier. */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#define ID_BYTES 128
#define MAX_RANKS 16
#define CHUNK (4u << 20) /* bytes per rank slot */

typedef struct { char b[ID_BYTES]; } nccl_id;

typedef struct {
    volatile int ready;
    int nranks;
    pthread_barrier_t bar;
} shm_header;

typedef struct {
    int rank, nranks;
    char name[64];
    size_t bytes;
    shm_header* hdr;
    unsigned char* slots; /* [nranks][CHUNK] */
} comm_t;

static size_t elem_size(int dtype) {
    switch (dtype) {
        case 0: case 1: return 1;
        case 2: case 3: case 7: return 4;
        case 4: case 5: case 8: return 8;
        default: return 0;
    }
}

const char* ncclGetErrorString(int r) { return r == 0 ? "no error" : "fake nccl: unsupported call or argument"; }

int ncclGetUniqueId(nccl_id* id) {
    memset(id, 0, sizeof(*id));
    struct timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    snprintf(id->b, ID_BYTES, "svbfm_emu_%d_%lld_%ld", (int)getpid(), (long long)ts.tv_sec, ts.tv_nsec);
    return 0;
}

int ncclCommInitRank(void** out, int nranks, nccl_id id, int rank) {
    if (nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks) return 4;
    comm_t* c = (comm_t*)calloc(1, sizeof(comm_t));
    c->rank = rank; c->nranks = nranks;
    id.b[ID_BYTES - 1] = 0;
    snprintf(c->name, sizeof(c->name), "/%.60s", id.b);
    c->bytes = 4096 + (size_t)nranks * CHUNK;
    int creator = 1;
    int fd = shm_open(c->name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 && errno == EEXIST) { creator = 0; fd = shm_open(c->name, O_RDWR, 0600); }
    if (fd < 0) { free(c); return 2; }
    if (creator && ftruncate(fd, (off_t)c->bytes) != 0) { close(fd); free(c); return 2; }
    if (!creator) {                    /* wait until the creator has sized the segment */
        struct stat st;
        for (int k = 0; k < 20000; k++) { if (fstat(fd, &st) == 0 && (size_t)st.st_size >= c->bytes) break; usleep(500); }
    }
    void* p = mmap(NULL, c->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { free(c); return 2; }
    c->hdr = (shm_header*)p;
    c->slots = (unsigned char*)p + 4096;
    if (creator) {
        pthread_barrierattr_t a;
        pthread_barrierattr_init(&a);
        pthread_barrierattr_setpshared(&a, PTHREAD_PROCESS_SHARED);
        pthread_barrier_init(&c->hdr->bar, &a, (unsigned)nranks);
        c->hdr->nranks = nranks;
        __sync_synchronize();
        c->hdr->ready = 1;
    } else {
        for (int k = 0; k < 200000 && !c->hdr->ready; k++) usleep(500);
        if (!c->hdr->ready) return 2;
    }
    pthread_barrier_wait(&c->hdr->bar);
    *out = c;
    return 0;
}

int ncclCommDestroy(void* comm) {
    comm_t* c = (comm_t*)comm;
    if (!c) return 0;
    pthread_barrier_wait(&c->hdr->bar);
    munmap((void*)c->hdr, c->bytes);
    if (c->rank == 0) shm_unlink(c->name);
    free(c);
    return 0;
}

#define REDUCE_LOOP(T)                                                                         \
    for (size_t i = 0; i < n; i++) {                                                           \
        T acc = ((const T*)(c->slots))[i];                                                     \
        for (int r = 1; r < c->nranks; r++) {                                                  \
            T v = ((const T*)(c->slots + (size_t)r * CHUNK))[i];                               \
            acc = (op == 0) ? (T)(acc + v) : (op == 1) ? (T)(acc * v) : (op == 2) ? (acc > v ? acc : v) : (acc < v ? acc : v); \
        }                                                                                      \
        ((T*)dst)[i] = acc;                                                                    \
    }

int ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int op, void* comm, void* stream) {
    (void)stream;
    comm_t* c = (comm_t*)comm;
    size_t es = elem_size(dtype);
    if (!c || !es || op < 0 || op > 3) return 4;
    size_t per = CHUNK / es;
    for (size_t off = 0; off < count; off += per) {
        size_t n = count - off < per ? count - off : per;
        memcpy(c->slots + (size_t)c->rank * CHUNK, (const unsigned char*)send + off * es, n * es);
        pthread_barrier_wait(&c->hdr->bar);
        unsigned char* dst = (unsigned char*)recv + off * es;
        if (dtype == 8) { REDUCE_LOOP(double) }
        else if (dtype == 7) { REDUCE_LOOP(float) }
        else if (dtype == 3) { REDUCE_LOOP(uint32_t) }
        else if (dtype == 2) { REDUCE_LOOP(int32_t) }
        else if (dtype == 5) { REDUCE_LOOP(uint64_t) }
        else if (dtype == 4) { REDUCE_LOOP(int64_t) }
        else return 4;
        pthread_barrier_wait(&c->hdr->bar);
    }
    return 0;
}

int ncclAllGather(const void* send, void* recv, size_t sendcount, int dtype, void* comm, void* stream) {
    (void)stream;
    comm_t* c = (comm_t*)comm;
    size_t es = elem_size(dtype);
    if (!c || !es) return 4;
    size_t per = CHUNK / es;
    for (size_t off = 0; off < sendcount; off += per) {
        size_t n = sendcount - off < per ? sendcount - off : per;
        memcpy(c->slots + (size_t)c->rank * CHUNK, (const unsigned char*)send + off * es, n * es);
        pthread_barrier_wait(&c->hdr->bar);
        for (int r = 0; r < c->nranks; r++)
            memcpy((unsigned char*)recv + ((size_t)r * sendcount + off) * es, c->slots + (size_t)r * CHUNK, n * es);
        pthread_barrier_wait(&c->hdr->bar);
    }
    return 0;
}

int ncclBroadcast(const void* send, void* recv, size_t count, int dtype, int root, void* comm, void* stream) {
    (void)stream;
    comm_t* c = (comm_t*)comm;
    size_t es = elem_size(dtype);
    if (!c || !es || root < 0 || root >= c->nranks) return 4;
    size_t per = CHUNK / es;
    for (size_t off = 0; off < count; off += per) {
        size_t n = count - off < per ? count - off : per;
        if (c->rank == root) memcpy(c->slots, (const unsigned char*)send + off * es, n * es);
        pthread_barrier_wait(&c->hdr->bar);
        memcpy((unsigned char*)recv + off * es, c->slots, n * es);
        pthread_barrier_wait(&c->hdr->bar);
    }
    return 0;
}

int ncclGroupStart(void) { return 0; }
int ncclGroupEnd(void) { return 0; }
