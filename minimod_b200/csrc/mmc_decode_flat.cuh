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
    FlatAlloc fa;                            // pool for the CIGAR arrays (dir | cq | cr)
    uint32_t *defer_list, *defer_n;          // reads left to k_decode_warp<!PRE>
    uint32_t read_count;
    uint32_t arena_bytes;                    // per warp, this kernel: WRead + room for dir | cq | cr
    uint32_t consumer_flex_words;            // flex capacity of the consumer's arena: decides the sampling shifts
    uint32_t stream;                         // consumer is k_decode_stream: no rank index / bitmap in its arena
};

__global__ void __launch_bounds__(kFThreads) k_flat_setup(const __grid_constant__ DecodeParams P, const __grid_constant__ FlatParams F) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t aoff = kWHeadBytes + warp * F.arena_bytes;
    const WArena A = w_arena<false>(aoff);
    const uint32_t local_words = (F.arena_bytes - kWReadBytes) / 4u;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < F.read_count; r += n_warps) {
        WRead *G = &F.reads[r];
        __syncwarp();
        bool ok = w_setup_read<false>(P, aoff, F.consumer_flex_words, local_words, F.defer_list, F.defer_n, r, lane, F.stream ? &F.fa : nullptr);
        __syncwarp();
        unsigned long long base = 0;
        const uint32_t take = (ok && !F.stream) ? A.R->st.n_stage : 0u;     // (stream mode: the table was written in place)
        if (ok && !F.stream) {                                              // CIGAR arrays -> pool
            if (lane == 0) base = atomicAdd(F.fa.cursor, (unsigned long long)take);
            base = ((unsigned long long)__shfl_sync(kFull, (uint32_t)(base >> 32), 0) << 32) | __shfl_sync(kFull, (uint32_t)base, 0);
            if (base + take > F.fa.cap) { w_defer(F.defer_list, F.defer_n, r, lane); ok = false; }
        }
        if (!ok) { if (lane == 0) G->st.n_blocks = 0; continue; }           // not ours (or fatal): the consumer skips it
        uint4 *d4 = reinterpret_cast<uint4 *>(F.fa.pool + base);
        const uint4 *s4 = reinterpret_cast<const uint4 *>(A.flex);
        for (uint32_t i = lane; i < (take >> 2); i += 32u) d4[i] = s4[i];
        const uint32_t words = (uint32_t)((sizeof(WState) + sizeof(uint32_t) * (kWBlocks + 4) + sizeof(WBlock) * A.R->st.n_blocks) / 4);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A.R);
        uint32_t *dst = reinterpret_cast<uint32_t *>(G);
        for (uint32_t i = lane; i < words; i += 32u) dst[i] = src[i];
        __syncwarp();
        if (lane == 0 && !F.stream) { G->st.flex = F.fa.pool + base; G->st.flex_home = F.fa.pool + base; }
    }
}

}  // namespace mmc

#endif  // MMC_DECODE_FLAT_CUH
