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

#endif
