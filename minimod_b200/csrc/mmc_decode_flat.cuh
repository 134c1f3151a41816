// mmc_decode_flat.cuh -- k_flat_setup: per-read setup as its own kernel (sm_100a).
//
// B200's instruction caches are small (L0 ~6 KB per SM sub-partition, L1.5 32 KB per SM).  A single
// fused kernel that parses MM headers, scans CIGARs, builds rank indexes and processes calls is
// ~8000 instructions and, with 24-32 warps per SM in different phases, stalls on instruction
// fetch (profiles/r01b_*).  So the once-per-read front end runs here, one warp per read, and
// leaves its results in HBM: the read's WRead (state + MM block table) and its CIGAR prefix
// sums + bucket directory (w_setup_read, "split" mode).  k_decode_warp<PRE> then copies those
// ~1.5 KB per read into its shared-memory arena and does the per-base work.
// Reads this path cannot take (> kWBlocks MM blocks, arena too small, reads >= 2^26 bases) are
// put on a list for the self-contained k_decode_warp<!PRE>, and from there k_decode.
//
// (An all-flat variant -- index, tile sums, scan, tile calls, finish as separate kernels over
// HBM-resident per-read state -- was measured at 2.2-2.3 ms per 99.6 k-read pass against 1.6 ms
// for this split: the per-call lookups into HBM/L2-resident indexes cost more than the
// instruction-cache misses they avoid.  It is in the history of this file, not in the build.)
#ifndef MMC_DECODE_FLAT_CUH
#define MMC_DECODE_FLAT_CUH

#include "mmc_decode_warp.cuh"

namespace mmc {

constexpr int kFThreads = 256;               // 8 warps per CTA

struct FlatParams {
    WRead *reads;                            // [n_reads]: state + block table of every read, in HBM
    FlatAlloc fa;                            // pool for the CIGAR arrays (dir | cq | cr) + the consumer's arena size
    uint32_t *defer_list, *defer_n;          // reads left to k_decode_warp<!PRE>
    uint32_t read_count;
};

__global__ void __launch_bounds__(kFThreads) k_flat_setup(const __grid_constant__ DecodeParams P, const __grid_constant__ FlatParams F) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < F.read_count; r += n_warps) {
        WRead *R = &F.reads[r];
        __syncwarp();
        const bool ok = w_setup_read(P, R, nullptr, 0u, &F.fa, F.defer_list, F.defer_n, r, lane);
        __syncwarp();
        if (!ok && lane == 0) R->st.n_blocks = 0;                           // not ours (or fatal): the consumer skips it
    }
}

}  // namespace mmc

#endif  // MMC_DECODE_FLAT_CUH
