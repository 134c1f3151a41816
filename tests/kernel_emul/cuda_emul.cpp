// cuda_emul.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emul.h).
#include "cuda_emul.h"
#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim

#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <time.h>
#include <ucontext.h>
#include <map>
#include <vector>

namespace cuda_emul {

enum State { RUNNABLE, WAIT_BLOCK, WAIT_WARP, DONE };

struct Fiber {
    ucontext_t ctx;
    Tls tls;
    State state;
    void *stack;
    uint32_t wgen;
};

static const size_t kStack = 512 * 1024;
static std::vector<Fiber> g_fibers;
static ucontext_t g_sched;
static const std::function<void()> *g_body;
static Fiber *g_running;
Tls *g_cur;
alignas(16) unsigned char g_dyn_smem[228 * 1024];
static uint64_t g_xchg[8][2][32];        // [warp][generation parity][lane]
static uint64_t g_pred[8][2];

static void trampoline() {
    (*g_body)();
    g_running->state = DONE;
    swapcontext(&g_running->ctx, &g_sched);
}

static void yield(State s) {
    Fiber *f = g_running;
    f->state = s;
    swapcontext(&f->ctx, &g_sched);
}

void sync_block() { yield(WAIT_BLOCK); }

uint64_t warp_exchange(uint64_t v, int src) {
    Fiber *f = g_running;
    unsigned lane = f->tls.threadIdx.x & 31u, warp = f->tls.threadIdx.x >> 5, par = f->wgen & 1u;
    g_xchg[warp][par][lane] = v;
    yield(WAIT_WARP);
    uint64_t out = (src >= 0 && src < 32) ? g_xchg[warp][par][src] : v;
    f->wgen++;
    return out;
}

uint32_t warp_ballot(bool pred) {
    Fiber *f = g_running;
    unsigned lane = f->tls.threadIdx.x & 31u, warp = f->tls.threadIdx.x >> 5, par = f->wgen & 1u;
    g_xchg[warp][par][lane] = pred ? 1u : 0u;
    yield(WAIT_WARP);
    uint32_t m = 0;
    unsigned nlanes = f->tls.blockDim.x - warp * 32u < 32u ? f->tls.blockDim.x - warp * 32u : 32u;
    for (unsigned l = 0; l < nlanes; ++l) m |= (uint32_t)(g_xchg[warp][par][l] & 1u) << l;
    f->wgen++;
    return m;
}

void launch(unsigned grid, unsigned block, const std::function<void()> &body) {
    if (block == 0 || block > 256 || (block & 31u)) { fprintf(stderr, "cuda_emul: bad block size %u\n", block); abort(); }
    if (g_fibers.size() < block) {
        size_t old = g_fibers.size();
        g_fibers.resize(block);
        for (size_t i = old; i < block; ++i) g_fibers[i].stack = malloc(kStack);
    }
    g_body = &body;
    for (unsigned b = 0; b < grid; ++b) {
        for (unsigned t = 0; t < block; ++t) {
            Fiber &f = g_fibers[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = &g_sched;
            makecontext(&f.ctx, trampoline, 0);
            f.tls.threadIdx = {t, 0, 0};
            f.tls.blockIdx = {b, 0, 0};
            f.tls.blockDim = dim3(block);
            f.tls.gridDim = dim3(grid);
            f.state = RUNNABLE;
            f.wgen = 0;
        }
        unsigned live = block;
        while (live) {
            bool ran = false;
            for (unsigned t = 0; t < block; ++t) {
                Fiber &f = g_fibers[t];
                if (f.state != RUNNABLE) continue;
                g_running = &f; g_cur = &f.tls;
                swapcontext(&g_sched, &f.ctx);
                ran = true;
                if (f.state == DONE) --live;
            }
            // release warp collectives whose live lanes all arrived
            bool released = false;
            for (unsigned w = 0; w * 32u < block; ++w) {
                unsigned waiting = 0, others = 0;
                for (unsigned l = 0; l < 32 && w * 32u + l < block; ++l) {
                    State s = g_fibers[w * 32u + l].state;
                    if (s == WAIT_WARP) ++waiting; else if (s != DONE) ++others;
                }
                if (waiting && !others) {
                    for (unsigned l = 0; l < 32 && w * 32u + l < block; ++l)
                        if (g_fibers[w * 32u + l].state == WAIT_WARP) g_fibers[w * 32u + l].state = RUNNABLE;
                    released = true;
                }
            }
            if (released) continue;
            unsigned at_block = 0, other = 0;
            for (unsigned t = 0; t < block; ++t) {
                State s = g_fibers[t].state;
                if (s == WAIT_BLOCK) ++at_block; else if (s != DONE) ++other;
            }
            if (at_block && !other) {
                for (unsigned t = 0; t < block; ++t) if (g_fibers[t].state == WAIT_BLOCK) g_fibers[t].state = RUNNABLE;
                continue;
            }
            if (!ran && live) {
                fprintf(stderr, "cuda_emul: DEADLOCK in block %u (divergent barrier?) states:", b);
                for (unsigned t = 0; t < block; ++t) fprintf(stderr, " %d", (int)g_fibers[t].state);
                fprintf(stderr, "\n");
                abort();
            }
        }
    }
}

}  // namespace cuda_emul

// ---- runtime API on host memory ----------------------------------------------------------
static std::map<void *, size_t> g_big;       // mmap'ed allocations
static const size_t kBig = 64u << 20;

cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "cuda_emul (CPU fibers, tests only)");
    p->multiProcessorCount = 4; p->totalGlobalMem = (size_t)180 << 30; p->major = 10; p->minor = 0;
    return cudaSuccess;
}
cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = (size_t)170 << 30; *t = (size_t)180 << 30; return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n) {
    if (n == 0) n = 1;
    if (n >= kBig) {
        void *m = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) return cudaErrorMemoryAllocation;
        g_big[m] = n; *p = m; return cudaSuccess;
    }
    void *m = nullptr;
    if (posix_memalign(&m, 256, n)) return cudaErrorMemoryAllocation;
    memset(m, 0xcd, n);                      // poison: device memory is not zeroed for you
    *p = m; return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
    if (!p) return cudaSuccess;
    auto it = g_big.find(p);
    if (it != g_big.end()) { munmap(p, it->second); g_big.erase(it); } else free(p);
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void **p, size_t n) {
    if (n >= kBig) {
        void *m = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) return cudaErrorMemoryAllocation;
        g_big[m] = n; *p = m; return cudaSuccess;
    }
    return posix_memalign(p, 256, n ? n : 1) ? cudaErrorMemoryAllocation : cudaSuccess;
}
cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t n) {
    // zero-fill of a big mapping: drop the pages instead of touching them
    if (v == 0 && n >= kBig) {
        uintptr_t a = ((uintptr_t)p + 4095) & ~(uintptr_t)4095, e = ((uintptr_t)p + n) & ~(uintptr_t)4095;
        for (auto &kv : g_big) {
            uintptr_t b0 = (uintptr_t)kv.first, b1 = b0 + kv.second;
            if ((uintptr_t)p >= b0 && (uintptr_t)p + n <= b1 && e > a) {
                memset(p, 0, a - (uintptr_t)p);
                madvise((void *)a, e - a, MADV_DONTNEED);
                memset((void *)e, 0, (uintptr_t)p + n - e);
                return cudaSuccess;
            }
        }
    }
    memset(p, v, n); return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { return cudaMemset(p, v, n); }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(1); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
struct emulEvent { double t; };
static double now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emulEvent{0}; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cuda_emul error"; }
