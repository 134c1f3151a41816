// mmc_decode_flat.cuh -- the decode+aggregate stage as a chain of small kernels (sm_100a).
//
// Same arithmetic as k_decode_warp (mmc_decode_warp.cuh: the phase functions are shared), but every
// phase is its own kernel and the per-read working set (WRead + scratch) lives in HBM/L2 instead
// of a shared-memory arena:
//
//   k_flat_setup      warp / read   record -> WRead, MM block table, CIGAR prefix sums + directory,
//                                   scratch allocation, text-tile records           (w_setup_read)
//   k_flat_index      warp / read   rank index per distinct base class              (w_build_index)
//   k_flat_tile_sums  warp / tile   tokens and sum of (skip+1) of every 496-byte text tile
//   k_flat_scan       thread / read exclusive scan of the tile sums: every tile's first call index
//                                   and base rank, every block's first ML index (src/mod.c:1098,1200)
//   k_flat_tile_calls warp / tile   ranks -> select -> map -> update                 (w_tile_calls)
//   k_flat_finish     warp / read   implicit calls of '.' blocks, per-read error report
//
// Why: B200's instruction caches are small (L0 ~6 KB per SM sub-partition, L1.5 32 KB per SM).  The
// fused kernel keeps 32 warps per SM in different phases of an ~8000-instruction program and
// stalls on instruction fetch; here every kernel is a few hundred instructions, all warps of
// the chip run the same loop, work units (tiles) are uniform, and the shared memory that the
// arenas took is L1 again.  Intermediates are small (~1.5 KB per read) and stay L2-resident.
// Reads this path cannot take (> kWBlocks MM blocks, scratch pool exhausted, reads >= 2^26 bases)
// are handed to k_decode_warp, and from there to k_decode.
#ifndef MMC_DECODE_FLAT_CUH
#define MMC_DECODE_FLAT_CUH

#include "mmc_decode_warp.cuh"

namespace mmc {

constexpr int kFThreads = 256;               // 8 warps per CTA in every flat kernel

struct FlatTile {                            // one 496-byte text tile of one MM block
    uint32_t read, blk, tb;                  // read index in the batch, block, first text byte (multiple of 16)
    uint32_t cnt, sum;                       // tokens / saturating sum of (skip+1)    (k_flat_tile_sums)
    uint32_t carry_cnt, carry_sum;           // the same, summed over the block's earlier tiles (k_flat_scan)
    uint32_t pad;
};

struct FlatParams {
    WRead *reads;                            // [n_reads]
    FlatAlloc fa;                            // scratch pool for cq|cr|dir|idx|rd|bitmaps
    FlatTile *tiles;
    uint32_t tile_cap;
    uint32_t *n_tiles;
    uint32_t *n_dot;                         // reads with '.' blocks to process (implicit calls)
    uint32_t *defer_list, *defer_n;          // reads left to k_decode_warp
    uint32_t read_first, read_count;         // the sub-batch [read_first, read_first + read_count) these launches work on
    uint32_t stage_words;                    // shared-memory words per warp for staging a read's scratch (tile_calls / index)
};

struct WRead1 {                              // WRead with room for one block: the shared-memory copy a tile works on
    WState   st;
    uint32_t semi[kWBlocks + 4];
    WBlock   blk[1];
};
constexpr uint32_t kWRead1Bytes = (uint32_t)((sizeof(WRead1) + 15) / 16 * 16);
constexpr uint32_t kWTileBytes = (uint32_t)((sizeof(WTile) + 15) / 16 * 16);

// lane-parallel word copy (n words; both pointers 4-byte aligned)
__device__ __forceinline__ void f_copy_words(uint32_t *dst, const uint32_t *src, uint32_t n, uint32_t lane) {
    for (uint32_t i = lane; i < n; i += 32u) dst[i] = src[i];
}
// lane-parallel 16-byte copy (n words rounded up to 4; both pointers 16-byte aligned)
__device__ __forceinline__ void f_copy_vec(uint32_t *dst, const uint32_t *src, uint32_t n, uint32_t lane) {
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    for (uint32_t i = lane; i < ((n + 3u) >> 2); i += 32u) d4[i] = s4[i];
}

__device__ __forceinline__ uint32_t f_warp_id() { return (blockIdx.x * blockDim.x + threadIdx.x) >> 5; }
__device__ __forceinline__ uint32_t f_n_warps() { return (gridDim.x * blockDim.x) >> 5; }

__global__ void __launch_bounds__(kFThreads) k_flat_setup(const __grid_constant__ DecodeParams P, const __grid_constant__ FlatParams F) {
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t i = f_warp_id(); i < F.read_count; i += f_n_warps()) {
        const uint32_t r = F.read_first + i;
        WRead *R = &F.reads[r];
        __syncwarp();
        const bool ok = w_setup_read(P, R, nullptr, 0u, &F.fa, F.defer_list, F.defer_n, r, lane);
        __syncwarp();
        if (!ok) { if (lane == 0) R->st.n_blocks = 0; continue; }           // not ours (or fatal): later kernels skip it
        if (F.fa.arena_words != 0u) continue;                               // split mode: k_decode_warp<PRE> takes it from here
        const uint32_t n_blocks = R->st.n_blocks, bm_words = ((R->st.L + 31u) >> 5) + 1u;
        uint32_t any_dot = 0;
        for (uint32_t b = 0; b < n_blocks; ++b) {
            WBlock *bd = &R->blk[b];
            const uint32_t a0 = bd->hdr_end, a1 = bd->end, t0 = a0 & ~15u;
            const uint32_t nt = a1 > t0 ? (a1 - t0 + (uint32_t)kWChunks * 16u - 1u) / ((uint32_t)kWChunks * 16u) : 0u;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(F.n_tiles, nt);
            base = __shfl_sync(kFull, base, 0);
            for (uint32_t t = lane; t < nt; t += 32u) {
                FlatTile ft;
                ft.read = r; ft.blk = b; ft.tb = t0 + t * (uint32_t)kWChunks * 16u;
                ft.cnt = 0; ft.sum = 0; ft.carry_cnt = 0; ft.carry_sum = 0; ft.pad = 0;
                if (base + t < F.tile_cap) F.tiles[base + t] = ft;
            }
            if (w_needs_bitmap(bd)) {
                any_dot = 1;
                uint32_t *bm = R->st.flex + bd->o_bm;
                for (uint32_t w = lane; w < bm_words; w += 32u) bm[w] = 0;
            }
            if (lane == 0) { bd->tile0 = base; bd->n_tiles = nt; }
        }
        if (any_dot && lane == 0) atomicAdd(F.n_dot, 1u);
    }
}

__global__ void __launch_bounds__(kFThreads) k_flat_index(const __grid_constant__ DecodeParams P, const __grid_constant__ FlatParams F) {
    MMC_DYN_SMEM(uint4, f_dyn_index);
    const uint32_t lane = threadIdx.x & 31u;
    uint8_t *mine = reinterpret_cast<uint8_t *>(f_dyn_index) + (size_t)(threadIdx.x >> 5) * (kWRead1Bytes + F.stage_words * 4u);
    WRead1 *Rs = reinterpret_cast<WRead1 *>(mine);
    uint32_t *stage = reinterpret_cast<uint32_t *>(mine + kWRead1Bytes);
    for (uint32_t i = f_warp_id(); i < F.read_count; i += f_n_warps()) {
        WRead *R = &F.reads[F.read_first + i];
        const uint32_t n_blocks = R->st.n_blocks;
        for (uint32_t b = 0; b < n_blocks; ++b) {
            WBlock *bd = &R->blk[b];
            if (!w_needs_index(bd)) continue;
            uint32_t from = b;                                              // an earlier block of the same class has it already
            for (uint32_t e = 0; e < b; ++e)
                if (w_needs_index(&R->blk[e]) && R->blk[e].cls == bd->cls) { from = e; break; }
            if (from != b) {
                if (lane == 0) { bd->cnt_cls = R->blk[from].cnt_cls; bd->rshift = R->blk[from].rshift; }
            } else {
                const uint32_t words = R->st.n_ent + 2u + R->st.n_rd;
                if (words <= F.stage_words) {                               // build in shared memory, store coalesced
                    __syncwarp();
                    f_copy_words(reinterpret_cast<uint32_t *>(&Rs->st), reinterpret_cast<const uint32_t *>(&R->st), sizeof(WState) / 4, lane);
                    f_copy_words(reinterpret_cast<uint32_t *>(&Rs->blk[0]), reinterpret_cast<const uint32_t *>(bd), sizeof(WBlock) / 4, lane);
                    __syncwarp();
                    if (lane == 0) Rs->st.flex = stage - Rs->blk[0].o_idx;  // so that flex + o_idx is the staging area
                    __syncwarp();
                    w_build_index(reinterpret_cast<WRead *>(Rs), 0u, lane);
                    f_copy_vec(R->st.flex_home + bd->o_idx, stage, words, lane);
                    if (lane == 0) { bd->cnt_cls = Rs->blk[0].cnt_cls; bd->rshift = Rs->blk[0].rshift; }
                } else {
                    w_build_index(R, b, lane);
                }
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(kFThreads) k_flat_tile_sums(const __grid_constant__ DecodeParams P, const __grid_constant__ FlatParams F) {
    __shared__ WTile s_tile[kFThreads / 32];
    const uint32_t lane = threadIdx.x & 31u;
    WTile *T = &s_tile[threadIdx.x >> 5];
    const uint32_t n_tiles = *F.n_tiles < F.tile_cap ? *F.n_tiles : F.tile_cap;
    for (uint32_t t = f_warp_id(); t < n_tiles; t += f_n_warps()) {
        FlatTile *ft = &F.tiles[t];
        WRead *R = &F.reads[ft->read];
        const WBlock *bd = &R->blk[ft->blk];
        uint32_t sum = 0;
        __syncwarp();
        const uint32_t n = w_tile_ranks(R, T, ft->tb, bd->hdr_end, bd->end, 0u, &sum, lane);
        if (lane == 0) { ft->cnt = n; ft->sum = sum; }
    }
}

__global__ void __launch_bounds__(kFThreads) k_flat_scan(const __grid_constant__ DecodeParams P, const __grid_constant__ FlatParams F) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < F.read_count; i += gridDim.x * blockDim.x) {
        WRead *R = &F.reads[F.read_first + i];
        const uint32_t n_blocks = R->st.n_blocks;
        uint32_t ml_base = 0;
        for (uint32_t b = 0; b < n_blocks; ++b) {
            WBlock *bd = &R->blk[b];
            uint32_t carry_cnt = 0, carry_sum = 0;
            for (uint32_t t = bd->tile0; t < bd->tile0 + bd->n_tiles && t < F.tile_cap; ++t) {
                FlatTile *ft = &F.tiles[t];
                ft->carry_cnt = carry_cnt; ft->carry_sum = carry_sum;
                carry_cnt += ft->cnt; carry_sum = sat_add(carry_sum, ft->sum);
            }
            bd->n_calls = carry_cnt; bd->last1 = carry_sum; bd->ml_base = ml_base;
            if (carry_cnt > 0) ml_base += carry_cnt * bd->K;                // src/mod.c:1200
        }
    }
}

__global__ void __launch_bounds__(kFThreads) k_flat_tile_calls(const __grid_constant__ DecodeParams P, const __grid_constant__ FlatParams F) {
    MMC_DYN_SMEM(uint4, f_dyn_calls);
    uint8_t *s_lut = reinterpret_cast<uint8_t *>(f_dyn_calls);
    w_stage_luts(P, s_lut);
    const uint32_t lane = threadIdx.x & 31u;
    uint8_t *mine = s_lut + kWLutSlots * 256 + (size_t)(threadIdx.x >> 5) * (kWTileBytes + kWRead1Bytes + F.stage_words * 4u);
    WTile *T = reinterpret_cast<WTile *>(mine);
    WRead1 *Rs = reinterpret_cast<WRead1 *>(mine + kWTileBytes);
    uint32_t *stage = reinterpret_cast<uint32_t *>(mine + kWTileBytes + kWRead1Bytes);
    const uint32_t n_tiles = *F.n_tiles < F.tile_cap ? *F.n_tiles : F.tile_cap;
    for (uint32_t t = f_warp_id(); t < n_tiles; t += f_n_warps()) {
        const FlatTile ft = F.tiles[t];
        if (ft.cnt == 0u) continue;
        WRead *Rg = &F.reads[ft.read];
        // the read's state, this tile's block and (if it fits) the read's lookup arrays -> shared memory
        __syncwarp();
        f_copy_words(reinterpret_cast<uint32_t *>(&Rs->st), reinterpret_cast<const uint32_t *>(&Rg->st), sizeof(WState) / 4, lane);
        f_copy_words(reinterpret_cast<uint32_t *>(&Rs->blk[0]), reinterpret_cast<const uint32_t *>(&Rg->blk[ft.blk]), sizeof(WBlock) / 4, lane);
        __syncwarp();
        if (!Rs->blk[0].any_req || Rs->st.err != 0u) continue;              // nothing to do / read already fatal
        const uint32_t n_stage = Rs->st.n_stage;
        if (n_stage <= F.stage_words) {
            f_copy_vec(stage, Rs->st.flex_home, n_stage, lane);
            __syncwarp();
            if (lane == 0) Rs->st.flex = stage;
        }
        __syncwarp();
        WRead *R = reinterpret_cast<WRead *>(Rs);
        uint32_t sum = 0;
        const uint32_t n = w_tile_ranks(R, T, ft.tb, Rs->blk[0].hdr_end, Rs->blk[0].end, ft.carry_sum, &sum, lane);
        w_tile_calls(P, R, T, s_lut, 0u, ft.blk, n, ft.carry_cnt, Rs->blk[0].ml_base, lane);
        __syncwarp();
        if (lane == 0 && Rs->st.err != 0u) atomicCAS(&Rg->st.err, 0u, Rs->st.err);
    }
}

__global__ void __launch_bounds__(kFThreads) k_flat_finish(const __grid_constant__ DecodeParams P, const __grid_constant__ FlatParams F) {
    __shared__ uint8_t s_lut[kWLutSlots * 256];
    w_stage_luts(P, s_lut);
    const uint32_t lane = threadIdx.x & 31u;
    const bool dots = *F.n_dot != 0u;
    for (uint32_t i = f_warp_id(); i < F.read_count; i += f_n_warps()) {
        const uint32_t r = F.read_first + i;
        WRead *R = &F.reads[r];
        const uint32_t n_blocks = R->st.n_blocks;
        if (n_blocks == 0u) continue;
        uint32_t err = w_err(R);
        if (dots && !err) {
            for (uint32_t b = 0; b < n_blocks; ++b)
                if (w_needs_bitmap(&R->blk[b])) w_implicit_block(P, R, s_lut, b, lane);
            err = w_err(R);
        }
        if (err) w_report(P, r, err, lane);
    }
}

}  // namespace mmc

#endif  // MMC_DECODE_FLAT_CUH
