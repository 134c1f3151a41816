// simt.h -- the one place that decides what "CUDA" means for the kernel sources.
//
// Product build (nvcc, sm_100a): the real CUDA runtime.
// Test build (-DMMC_EMUL, g++): tests/kernel_emul/cuda_emul.h, a fiber-based SIMT emulator
// that lets the CPU-only CI execute the *same* kernel source under pytest.  The emulator is
// test infrastructure: it is never compiled into libminimod_cuda.so and nothing in the
// product can fall back to it.
#ifndef MMC_SIMT_H
#define MMC_SIMT_H

#ifdef MMC_EMUL
#include "cuda_emul.h"
#define MMC_LAUNCH(kernel, grid, block, stream, ...) \
    ::cuda_emul::launch((grid), (block), [=]() { kernel(__VA_ARGS__); })
#define MMC_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) \
    ::cuda_emul::launch((grid), (block), [=]() { kernel(__VA_ARGS__); })
// dynamic shared memory: one static buffer of the hardware maximum (CTAs run one after another)
#define MMC_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(::cuda_emul::g_dyn_smem)
#else
#include <cuda_runtime.h>
#define MMC_LAUNCH(kernel, grid, block, stream, ...) \
    kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define MMC_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define MMC_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif

// ---- bulk asynchronous copy global -> shared (TMA, cp.async.bulk) completing on an mbarrier.
// One lane arms the barrier with the byte count and issues the copy; every lane that reads the data waits on the
// barrier's phase parity.  Under the SIMT emulator the copy is a memcpy and the waits are no-ops.
#ifdef MMC_EMUL
static inline void mmc_mbar_init(unsigned long long *bar, uint32_t) { *bar = 0; }
static inline void mmc_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *) { memcpy(dst, src, bytes); }
static inline void mmc_mbar_wait(unsigned long long *, uint32_t) {}
#else
__device__ __forceinline__ uint32_t mmc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mmc_mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mmc_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// arms `bar` for `bytes` and starts the copy (dst/src 16-byte aligned, bytes a multiple of 16).  The shared destination
// was last touched through the generic proxy (LDS/STS): the proxy fence orders those accesses before the async write.
__device__ __forceinline__ void mmc_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mmc_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(mmc_smem_u32(dst)), "l"(src), "r"(bytes), "r"(mmc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mmc_mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mmc_smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
#endif

#endif
