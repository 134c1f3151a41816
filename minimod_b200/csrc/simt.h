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
#else
#include <cuda_runtime.h>
#define MMC_LAUNCH(kernel, grid, block, stream, ...) \
    kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif

#endif
