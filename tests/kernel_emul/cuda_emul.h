// cuda_emul.h -- TEST INFRASTRUCTURE ONLY.
//
// A small SIMT emulator so that CPU-only CI can execute the product's kernel *sources*
// (minimod_b200/csrc/*.cuh, *.cu compiled with -DMMC_EMUL) and compare them with the oracle
// without a GPU.  One OS thread; every CUDA thread of a CTA is a ucontext fiber; barriers and
// warp collectives are cooperative yields, so execution is deterministic and a divergent
// barrier shows up as a reported deadlock instead of a hang.  CTAs run one after another.
// It also provides the handful of CUDA runtime calls mmc_api.cu makes, backed by host memory.
//
// This is not a fallback: it is never linked into libminimod_cuda.so, lives under tests/, and
// the product loader refuses to run without the real library and a CUDA device.
#ifndef CUDA_EMUL_H
#define CUDA_EMUL_H

#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
#define __grid_constant__
#define __noinline__ __attribute__((noinline))

struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(8) uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 v; v.x = x; v.y = y; return v; }
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };

namespace cuda_emul {
struct Tls { uint3 threadIdx, blockIdx; dim3 blockDim, gridDim; };
extern Tls *g_cur;                       // the fiber that is running
extern unsigned char g_dyn_smem[228 * 1024];   // the (single) dynamic shared-memory array of the running CTA
void launch(unsigned grid, unsigned block, const std::function<void()> &body);
void sync_block();
uint64_t warp_exchange(uint64_t v, int src_lane_or_neg);   // value held by src lane (own if out of range)
uint32_t warp_ballot(bool pred);
}

#define threadIdx (::cuda_emul::g_cur->threadIdx)
#define blockIdx  (::cuda_emul::g_cur->blockIdx)
#define blockDim  (::cuda_emul::g_cur->blockDim)
#define gridDim   (::cuda_emul::g_cur->gridDim)

static inline void __syncthreads() { ::cuda_emul::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { (void)::cuda_emul::warp_ballot(true); }
static inline void __threadfence() {}

static inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned d) {
    int lane = (int)(threadIdx.x & 31u);
    return (uint32_t)::cuda_emul::warp_exchange(v, lane - (int)d);
}
static inline uint32_t __shfl_down_sync(unsigned, uint32_t v, unsigned d) {
    int lane = (int)(threadIdx.x & 31u);
    return (uint32_t)::cuda_emul::warp_exchange(v, lane + (int)d > 31 ? -1 : lane + (int)d);
}
static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) { return (uint32_t)::cuda_emul::warp_exchange(v, src & 31); }
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int m) {
    int lane = (int)(threadIdx.x & 31u);
    return (uint32_t)::cuda_emul::warp_exchange(v, lane ^ m);
}
static inline unsigned __ballot_sync(unsigned, int pred) { return ::cuda_emul::warp_ballot(pred != 0); }

template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31u));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) {
    return (uint32_t)((((((uint64_t)hi) << 32) | lo) << (sh & 31u)) >> 32);
}
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }

template <class T> static inline T emul_atomic_add(T *p, T v) { T o = *p; *p = o + v; return o; }
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return emul_atomic_add(p, v); }
static inline int atomicAdd(int *p, int v) { return emul_atomic_add(p, v); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return emul_atomic_add(p, v); }
static inline uint32_t atomicOr(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
static inline uint32_t atomicAnd(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o & v; return o; }
static inline uint32_t atomicCAS(uint32_t *p, uint32_t c, uint32_t v) { uint32_t o = *p; if (o == c) *p = v; return o; }
static inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long c, unsigned long long v) { unsigned long long o = *p; if (o == c) *p = v; return o; }
static inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
static inline uint32_t atomicMax(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }

// ---- the slice of the CUDA runtime API that mmc_api.cu uses -----------------------------
typedef int cudaError_t;
typedef struct emulStream *cudaStream_t;
typedef struct emulEvent *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaHostAllocDefault = 0 };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; size_t totalGlobalMem; int major, minor; };

cudaError_t cudaSetDevice(int);
cudaError_t cudaGetDeviceCount(int *);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *, int);
cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b);
cudaError_t cudaMalloc(void **, size_t);
cudaError_t cudaFree(void *);
cudaError_t cudaMallocHost(void **, size_t);
cudaError_t cudaFreeHost(void *);
cudaError_t cudaMemcpy(void *, const void *, size_t, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void *, const void *, size_t, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaMemset(void *, int, size_t);
cudaError_t cudaMemsetAsync(void *, int, size_t, cudaStream_t);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *, unsigned);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t *);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t);
cudaError_t cudaEventSynchronize(cudaEvent_t);
cudaError_t cudaEventElapsedTime(float *, cudaEvent_t, cudaEvent_t);
cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t);
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 2; return cudaSuccess; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

#endif
