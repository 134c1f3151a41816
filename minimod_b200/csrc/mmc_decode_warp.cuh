// mmc_decode_warp.cuh -- k_decode_warp: the fast path of the decode+aggregate stage (sm_100a).
//
// One WARP per read, reads handed out by an atomic counter; no CTA barrier in the read loop,
// only shuffles and __syncwarp.  It computes exactly what k_decode (mmc_device.cuh) computes --
// freq_view_single() + get_aln() + update_freq_map() of the reference (src/mod.c:776-1370) --
// with per-read working sets small enough for a slice of shared memory ("arena") per warp:
//
//   CIGAR      prefix sums of every 2^cshift-th op (cshift 0 unless the CIGAR is long) + a
//              directory over 2^gshift-base buckets of the read, so aln[q]/ins[q] (get_aln,
//              src/mod.c:776-881) is two shared loads and a 0-2 step search.
//   base ranks one u32 prefix count per 32<<ishift bases of the block's base class instead of
//              the bases_pos[][] tables (src/mod.c:977-981); rank -> position is an
//              interpolated probe into that index + an in-word select on the 4-bit SEQ.
//   MM text    496-byte tiles: SWAR comma masks -> token descriptors -> one token per lane:
//              SWAR decimal parse, warp scan of (skip+1) gives every call its rank
//              (src/mod.c:1098), then the call is processed by the same lane.
//   context    is_context[][] (src/ref.c:204-219) for ACGT contexts of <= 8 bases is tested
//              on a 16-base window funnel-shifted out of the 2-bit reference.
//
// The hot loop (explicit calls) is inlined once; everything rare (header parsing, implicit
// calls of '.' blocks, contig-edge / N-containing contexts, sparse cells, view rows, long
// skip counts) lives in __noinline__ functions that read the per-read state from shared
// memory, to keep the instruction footprint of the hot path small.
//
// Reads that do not fit the arena (very long CIGARs / reads, > kWBlocks MM blocks) are
// appended to a deferred list and handled by the general CTA-per-read kernel k_decode in
// the same stream right after.  Both kernels add into the same dense count arrays.
#ifndef MMC_DECODE_WARP_CUH
#define MMC_DECODE_WARP_CUH

#include "mmc_device.cuh"

namespace mmc {

constexpr int kWThreads   = 256;             // 8 warps per CTA
constexpr int kWBlocks    = 8;               // MM blocks per read on this path
constexpr int kWChunks    = 31;              // 16-byte text chunks owned per tile (lane 31 is look-ahead)
constexpr int kWTokCap    = 256;             // >= kWChunks*16/2 tokens per tile
constexpr int kWMaxCShift = 5;
constexpr int kWMaxIShift = 3;
constexpr int kWLutSlots  = 8;               // -c entries whose call LUT is staged in shared memory
constexpr uint32_t kWMaxL = 1u << 26;        // longer reads go to k_decode
constexpr uint32_t kFull  = 0xffffffffu;

enum : uint32_t { kCtxNone = 0, kCtxFast = 1, kCtxSlow = 2 };

struct WCode {                               // one modification code of a block, 8 bytes
    int8_t   ri;                             // -c entry, -1: not requested (src/mod.c:1157)
    uint8_t  outc;                           // output code id
    uint8_t  ctx_mode;                       // kCtxNone / kCtxFast / kCtxSlow
    uint8_t  ctx_len;
    uint32_t pat2;                           // context in the read's orientation, 2 bits per base
};

struct WBlock {                              // see BlockDesc
    uint32_t hdr_end, end;
    uint8_t  cls, is_n, dot, K;
    uint32_t any_req;
    // rank index of the block's base class (w_build_index) and explicit-rank bitmap ('.' blocks): word offsets into flex
    uint32_t o_idx, o_rd, rshift, cnt_cls, o_bm;
    // filled once the block's skip counts are known
    uint32_t n_calls, last1, ml_base;        // tokens, last rank + 1 (0 if none), first ML index (src/mod.c:1200)
    uint32_t tile0, n_tiles;                 // flat path: the block's text tiles
    WCode    code[kMaxCodes];
};

struct WState {                              // warp-uniform state of the read being processed
    const uint32_t *cig;
    const uint8_t  *seq, *mm, *ml;
    const uint32_t *ref2, *excm;
    unsigned long long *cells;
    uint32_t *flex;                          // dir | cq | cr | idx+rd per class | bitmaps of this read (may be a staged copy)
    uint32_t *flex_home;                     // where those words really live: the bitmaps are always updated there
    uint32_t n_stage;                        // words before the bitmaps
    uint32_t r, L, n_cig, mm_len, ml_len, rev, hp, ref_len;
    int32_t  tid, pos;
    uint32_t o_cq, o_cr, o_dir;              // word offsets into flex
    uint32_t cshift, gshift, ishift, n_samp, n_ent, n_u4, n_rd, total_q;
    uint32_t cur_cls, cur_blk;               // fused path: class / block whose index currently occupies the arena
    uint32_t carry_sum, prev_last;           // fused path: carries between the text tiles of a block
    uint32_t err;
    uint32_t n_semi, n_blocks;
    uint32_t pairs;                          // stream layout: dir (32-base buckets) | uint2 {cq, cr} per op, in HBM (w_setup_read, stream mode)
    uint32_t pad_;
};

struct WRead {                               // per read: shared memory (fused path) or global scratch (flat path)
    WState   st;
    uint32_t semi[kWBlocks + 4];
    WBlock   blk[kWBlocks];
};

struct WTile {                               // per warp, shared memory: one text tile of a block
    uint4    text[32 + 1];                   // chunk i at text[i] (+ one spill-over chunk)
    uint32_t em[32];                         // per chunk: terminator mask (',' or block end) of its 32 following bytes
    uint32_t rank[kWTokCap];                 // per token: byte offset in text -> base rank -> read position q
    unsigned long long sp_state, sp_pad;     // sparse side buffer chunk of this warp: (first slot << 8) | slots used (w_sparse_flush)
};

struct WFixed {                              // fixed part of a warp's arena on the fused path
    WRead rd;
    WTile tl;
};

// TILE = false: an arena without the WTile part (k_flat_setup): flex follows WRead directly.
struct WArena { uint8_t *s_lut; WRead *R; WTile *T; uint32_t *flex; };
constexpr uint32_t kWReadBytes = (uint32_t)((sizeof(WRead) + 15) / 16 * 16);
template <bool TILE = true>
__device__ __forceinline__ WArena w_arena(uint32_t aoff) {
    MMC_DYN_SMEM(uint4, w_dyn);
    uint8_t *base = reinterpret_cast<uint8_t *>(w_dyn);
    WArena A;
    A.s_lut = base;
    WFixed *wf = reinterpret_cast<WFixed *>(base + aoff);
    A.R = &wf->rd; A.T = &wf->tl;
    A.flex = reinterpret_cast<uint32_t *>(base + aoff + (TILE ? (uint32_t)sizeof(WFixed) : kWReadBytes));
    return A;
}

constexpr uint32_t kWHeadBytes = (uint32_t)kWLutSlots * 256u;             // call LUTs, then the warps' arenas
constexpr unsigned long long kSpSentinel = ~0ull;                         // never a valid SparseRec.a; finalize skips it
constexpr uint32_t kSpChunk = 64;                                         // slots a warp reserves at a time (w_sparse_flush)

struct WarpParams {
    uint32_t arena_bytes;                    // per warp, multiple of 16, >= sizeof(WFixed) + 256
    uint32_t *defer_list;                    // reads left to k_decode
    uint32_t *defer_n;
};

struct FlatAlloc {                           // split path: bump allocator over a global scratch pool (words)
    uint32_t *pool;
    unsigned long long *cursor;
    unsigned long long cap;
};

// ---------------------------------------------------------------------------------------
// warp scans
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t x = __shfl_up_sync(kFull, v, d);
        if (lane >= (uint32_t)d) v += x;
    }
    return v;
}
__device__ __forceinline__ uint32_t warp_incl_scan_sat(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t x = __shfl_up_sync(kFull, v, d);
        if (lane >= (uint32_t)d) v = sat_add(v, x);
    }
    return v;
}

// ---------------------------------------------------------------------------------------
// 4-bit SEQ: class tests on 8 bases at a time.  flags: bit 4k+3 set <=> nibble k of the word
// belongs to the class (src/mod.c:97: A0 C1 G2 T3 N4, every other nt16 code counts as A).
// A BAM byte holds base 2i in its HIGH nibble: within a byte, bit 7 is the earlier base.
// `pat` = the class's nibble code replicated (cls 1..4); class 0 is the template flag C0.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nib_eq_flags(uint32_t u, uint32_t pat) {
    uint32_t y = u ^ pat;
    uint32_t t = (y & 0x77777777u) + 0x77777777u;
    return ~(t | y) & 0x88888888u;
}
__device__ __forceinline__ uint32_t class_pat(uint32_t cls) {
    return cls == 0u ? 0u : cls == 1u ? 0x22222222u : cls == 2u ? 0x44444444u : cls == 3u ? 0x88888888u : 0xffffffffu;
}
template <bool C0>
__device__ __forceinline__ uint32_t class_flags(uint32_t u, uint32_t pat) {
    if (C0) {
        uint32_t f = nib_eq_flags(u, 0x22222222u) | nib_eq_flags(u, 0x44444444u) |
                     nib_eq_flags(u, 0x88888888u) | nib_eq_flags(u, 0xffffffffu);
        return ~f & 0x88888888u;
    }
    return nib_eq_flags(u, pat);
}
template <bool C0>
__device__ __forceinline__ uint32_t count_u4(uint4 v, uint32_t pat) {
    return (uint32_t)(__popc(class_flags<C0>(v.x, pat)) + __popc(class_flags<C0>(v.y, pat)) +
                      __popc(class_flags<C0>(v.z, pat)) + __popc(class_flags<C0>(v.w, pat)));
}
// flags of the first nv (0..8) bases of a word: whole bytes first, then the high nibble of the next byte
__device__ __forceinline__ uint32_t valid_flags(uint32_t f, uint32_t nv) {
    if (nv >= 8u) return f;
    uint32_t keep = (1u << (8u * (nv >> 1))) - 1u;
    if (nv & 1u) keep |= 0xf0u << (8u * (nv >> 1));
    return f & keep;
}
template <bool C0>
__device__ __forceinline__ uint32_t count_u4_tail(uint4 v, uint32_t pat, uint32_t nv /* valid bases, 1..31 */) {
    return (uint32_t)(__popc(valid_flags(class_flags<C0>(v.x, pat), nv)) +
                      __popc(valid_flags(class_flags<C0>(v.y, pat), nv > 8u ? nv - 8u : 0u)) +
                      __popc(valid_flags(class_flags<C0>(v.z, pat), nv > 16u ? nv - 16u : 0u)) +
                      __popc(valid_flags(class_flags<C0>(v.w, pat), nv > 24u ? nv - 24u : 0u)));
}
// global-memory accesses spelled out: the pointers come out of structs in shared memory, where the
// compiler cannot see their address space and would emit generic LD / ATOM with run-time space checks
__device__ __forceinline__ uint4 ld16(const uint8_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ uint32_t ldg32(const uint32_t *p) { return __ldg(p); }
__device__ __forceinline__ uint32_t ldg8(const uint8_t *p) { return (uint32_t)__ldg(p); }
__device__ __forceinline__ void red_add_u64(unsigned long long *p, unsigned long long v) {
#ifdef MMC_EMUL
    *p += v;
#else
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}

// The fused kernels' dynamic shared memory: [call LUTs][arena of warp 0][arena of warp 1]...  Non-inlined
// phase functions get a warp's arena as a byte OFFSET and rebuild their pointers from the array itself,
// so that every access compiles to LDS/STS with an immediate offset.

__device__ __forceinline__ void w_raise(WRead *R, uint32_t code) { atomicCAS(&R->st.err, 0u, code); }
__device__ __forceinline__ uint32_t w_err(WRead *R) {
    __syncwarp();
    uint32_t e = *reinterpret_cast<volatile uint32_t *>(&R->st.err);
    __syncwarp();
    return e;
}

// ---------------------------------------------------------------------------------------
// aln[q], ins[q], ins_offset[q] of get_aln() in BAM orientation (src/mod.c:776-881).
// cq[s] = (read bases consumed before op s<<cshift) << 4 | that op's type; cr[s] = reference
// bases consumed before it.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ AlnHit w_cigar_lookup(const WState &S, const uint32_t *flex, uint32_t q) {
    AlnHit h; h.aln = -1; h.ins = -1; h.insoff = 0;
    const uint32_t total_q = S.total_q;
    if (q >= total_q) return h;
    if (S.pairs) {                                               // stream layout, looked up in HBM (L1-cached)
        const uint2 *pr = reinterpret_cast<const uint2 *>(flex + S.o_cq);
        const uint32_t g = S.gshift, b = q >> g;
        uint32_t lo = ldg32(flex + S.o_dir + b);
        uint32_t hi = (((b + 1u) << g) < total_q) ? ldg32(flex + S.o_dir + b + 1u) : S.n_samp - 1u;
        const uint32_t qlim = (q + 1u) << 4;
        while (lo < hi) { const uint32_t mid = (lo + hi + 1u) >> 1; if (ldg32(&pr[mid].x) < qlim) lo = mid; else hi = mid - 1u; }
        const uint2 e = __ldg(&pr[lo]);
        const uint32_t op = e.x & 15u, d = q - (e.x >> 4);
        if (op == 0u || op == 7u || op == 8u) h.aln = S.pos + (int32_t)(e.y + d);
        else if (op == 1u) { h.ins = S.pos + (int32_t)e.y - 1; h.insoff = d + 1u; }
        return h;
    }
    const uint32_t *dir = flex + S.o_dir, *cq = flex + S.o_cq;
    const uint32_t g = S.gshift, b = q >> g;
    uint32_t lo = dir[b];
    uint32_t hi = (((b + 1u) << g) < total_q) ? dir[b + 1u] : S.n_samp - 1u;
    const uint32_t qlim = (q + 1u) << 4;
    while (lo < hi) {                                            // largest sample with q0 <= q
        uint32_t mid = (lo + hi + 1u) >> 1;
        if (cq[mid] < qlim) lo = mid; else hi = mid - 1u;
    }
    const uint32_t e = cq[lo];
    uint32_t qa = e >> 4, ra = flex[S.o_cr + lo], op = e & 15u, d = q - qa;
    const uint32_t cshift = S.cshift;
    if (cshift != 0u) {                                          // walk the <= 2^cshift raw ops of the sample
        const uint32_t *cig = S.cig;
        const uint32_t n_cig = S.n_cig;
        for (uint32_t o = lo << cshift; o < n_cig; ++o) {
            const uint32_t w = ldg32(cig + o), len = w >> 4;
            op = w & 15u;
            const bool aln = op == 0u || op == 7u || op == 8u;
            const uint32_t ql = (aln || op == 1u || op == 4u) ? len : 0u;
            d = q - qa;
            if (d < ql) break;
            qa += ql;
            if (aln || op == 2u || op == 3u) ra += len;
        }
    }
    if (op == 0u || op == 7u || op == 8u) h.aln = S.pos + (int32_t)(ra + d);
    else if (op == 1u) { h.ins = S.pos + (int32_t)ra - 1; h.insoff = d + 1u; }
    return h;
}

// ---------------------------------------------------------------------------------------
// bases_pos[cls][k] (src/mod.c:977-981): BAM position of the k-th base of the block's class.
// probe: index entry + vector that hold class rank k; rem = rank of k inside the vector
// ---------------------------------------------------------------------------------------
struct SelProbe { uint32_t u, rem; uint4 v; };
template <bool C0>
__device__ __forceinline__ SelProbe w_select_probe(const WState &S, const uint32_t *flex, const WBlock *bd, uint32_t pat, uint32_t k) {
    const uint32_t *idx = flex + bd->o_idx;
    uint32_t e = flex[bd->o_rd + (k >> bd->rshift)];             // entry of rank (k >> rshift) << rshift: at or before k's
    uint32_t nxt = idx[e + 1u];
    while (nxt <= k) { ++e; nxt = idx[e + 1u]; }                 // few steps: 2^rshift ranks span few entries
    SelProbe p;
    p.rem = k - idx[e];
    const uint32_t ishift = S.ishift;
    const uint8_t *seq = S.seq;
    p.u = e << ishift;
    p.v = ld16(seq + (size_t)p.u * 16u);
    if (ishift != 0u) {
        for (;;) {                                               // at most 2^ishift vectors
            uint32_t c = count_u4<C0>(p.v, pat);
            if (p.rem < c) break;
            p.rem -= c; ++p.u;
            p.v = ld16(seq + (size_t)p.u * 16u);
        }
    }
    return p;
}
template <bool C0>
__device__ __forceinline__ uint32_t w_select_resolve(const SelProbe &p, uint32_t pat) {
    const uint32_t f0 = class_flags<C0>(p.v.x, pat), f1 = class_flags<C0>(p.v.y, pat), f2 = class_flags<C0>(p.v.z, pat), f3 = class_flags<C0>(p.v.w, pat);
    const uint32_t s0 = (uint32_t)__popc(f0), s1 = s0 + (uint32_t)__popc(f1), s2 = s1 + (uint32_t)__popc(f2);
    uint32_t rem = p.rem, f = f0, wsel = 0, sub = 0;
    if (rem >= s0) { f = f1; wsel = 1; sub = s0; }
    if (rem >= s1) { f = f2; wsel = 2; sub = s1; }
    if (rem >= s2) { f = f3; wsel = 3; sub = s2; }
    rem -= sub;
    uint32_t off = 0, c;
    c = (uint32_t)__popc(f & 0xffffu);  if (rem >= c) { rem -= c; f >>= 16; off = 4; }
    c = (uint32_t)__popc(f & 0xffu);    if (rem >= c) { rem -= c; f >>= 8; off += 2; }
    off += (rem != 0u || !(f & 0x80u)) ? 1u : 0u;                // bit 7 = the byte's first base
    return p.u * 32u + wsel * 8u + off;
}

// ---------------------------------------------------------------------------------------
// context + base test against the packed reference (src/mod.c:1162-1172, src/ref.c:204-219)
// ---------------------------------------------------------------------------------------
// 1 = in context (then refcode is the 2-bit reference base at pos, never an exception letter),
// 0 = not in context, -1 = contig edge: use the generic test
__device__ __forceinline__ int32_t ctx_fast(const uint32_t *ref2, const uint32_t *excm, uint32_t ref_len, uint32_t pos,
                                            uint32_t m, uint32_t pat2, uint32_t &refcode) {
    if (pos + 1u < m || pos + m > ref_len) return -1;
    const uint32_t w0 = pos + 1u - m;                             // first base of the window
    const uint32_t wi = w0 >> 4, ei = w0 >> 5;
    const uint32_t lo = ldg32(ref2 + wi), hi = ldg32(ref2 + wi + 1u);
    const uint32_t elo = ldg32(excm + ei), ehi = ldg32(excm + ei + 1u);
    const uint32_t W = __funnelshift_r(lo, hi, (w0 & 15u) * 2u);  // 16 bases from w0
    const uint32_t E = __funnelshift_r(elo, ehi, w0 & 31u);       // their exception bits
    const uint32_t m2 = (1u << (2u * m)) - 1u, m1 = (1u << m) - 1u;
    uint32_t hit = 0;
    for (uint32_t j = 0; j < m; ++j)                              // occurrence starting at w0 + j
        hit |= (uint32_t)((((W >> (2u * j)) & m2) == pat2) & (((E >> j) & m1) == 0u));
    refcode = (W >> (2u * (m - 1u))) & 3u;
    return (int32_t)hit;
}

// generic context + base test (contig edges, contexts with N or longer than 8): cold
__device__ __noinline__ bool w_ctx_slow(const DecodeParams &P, const WState &S, uint32_t ri, uint32_t ref_pos, uint32_t q, uint32_t is_n) {
    const ContigDev cd = P.contigs[S.tid];
    const ReqMod &rq = P.req[ri];
    if (!in_context(cd, ref_pos, S.rev ? rq.pat_rc : rq.pat, rq.ctx_len)) return false;
    if (is_n) return true;
    const uint32_t nib = (S.seq[q >> 1] >> ((~q & 1u) << 2)) & 0xfu;
    return ref_letter(cd, ref_pos) == nt16_letter(nib);
}

// a count outside the dense arrays (ins_offset > 0, exotic haplotype / code id)
// (The fast call loop does not come here: it appends from a per-warp chunk at a converged point, w_sparse_flush below.
// Letting diverged groups of a warp share that chunk state through CAS hung on real '.'-status data, so every other
// caller keeps this stateless form.)
__device__ __noinline__ void w_add_sparse(const DecodeParams &P, uint32_t tid, uint32_t rev, int32_t ref_pos, uint32_t outc,
                                          uint32_t ins16, int32_t hap, uint32_t is_mod) {
#ifdef MMC_EMUL
    unsigned long long slot = atomicAdd(P.sparse_n, 1ull);
#else
    // one atomic per group of lanes that are here together: they take consecutive slots
    const uint32_t act = __activemask(), lane_id = threadIdx.x & 31u, leader = (uint32_t)__ffs((int)act) - 1u;
    unsigned long long slot = 0;
    if (lane_id == leader) slot = atomicAdd(P.sparse_n, (unsigned long long)__popc(act));
    slot = (((unsigned long long)__shfl_sync(act, (uint32_t)(slot >> 32), (int)leader)) << 32) | __shfl_sync(act, (uint32_t)slot, (int)leader);
    slot += (unsigned long long)__popc(act & ((1u << lane_id) - 1u));
#endif
    if (slot < P.sparse_cap) {
        SparseRec s;
        s.a = ((unsigned long long)tid << 41) | ((unsigned long long)(uint32_t)ref_pos << 9) | ((unsigned long long)rev << 8) | outc;
        s.b = ins16 | ((hap < 0 ? 256u : (uint32_t)hap) << 16);
        s.w = 1u | (is_mod << 16);
        P.sparse[slot] = s;
    }
}

// The fast path's sparse appends, made at a point where the whole warp is converged (after each round of 32 calls):
// the warp owns a chunk of kSpChunk slots of the side buffer and hands them out from shared memory, so the global
// atomic -- and the wait for its result -- happens once per chunk instead of once per round.  All lanes call it;
// nrec = records per pending lane (2 with a dense-less haplotype stratum).  Unused slots of a chunk are closed with
// sentinels (here when a chunk is left, and by w_sparse_close at the end of the kernel).
__device__ __forceinline__ void w_sparse_sentinel(const DecodeParams &P, unsigned long long slot) {
    if (slot < P.sparse_cap) { SparseRec z; z.a = kSpSentinel; z.b = 0; z.w = 0; P.sparse[slot] = z; }
}
__device__ __forceinline__ void w_sparse_flush(const DecodeParams &P, WTile *T, bool pending, uint32_t nrec, uint32_t tid, uint32_t rev, uint32_t ref_pos,
                                               uint32_t outc, uint32_t ins16, uint32_t hp, uint32_t is_mod, uint32_t lane) {
    const uint32_t m = __ballot_sync(kFull, pending);
    if (!m) return;
    const uint32_t cnt = (uint32_t)__popc(m) * nrec, mine = (uint32_t)__popc(m & ((1u << lane) - 1u)) * nrec;
    const unsigned long long st = T->sp_state;
    unsigned long long base = st >> 8;
    uint32_t used = (uint32_t)(st & 0xffull);
    __syncwarp();
    if (used + cnt > kSpChunk) {                                             // warp-uniform
        for (uint32_t u = used + lane; u < kSpChunk; u += 32u) w_sparse_sentinel(P, base + u);
        unsigned long long nb = 0;
        if (lane == 0) nb = atomicAdd(P.sparse_n, (unsigned long long)kSpChunk);
        base = (((unsigned long long)__shfl_sync(kFull, (uint32_t)(nb >> 32), 0)) << 32) | __shfl_sync(kFull, (uint32_t)nb, 0);
        used = 0;
    }
    if (lane == 0) T->sp_state = (base << 8) | (used + cnt);
    __syncwarp();
    if (pending) {
        const unsigned long long slot = base + used + mine;
        SparseRec r;
        r.a = ((unsigned long long)tid << 41) | ((unsigned long long)ref_pos << 9) | ((unsigned long long)rev << 8) | outc;
        r.w = 1u | (is_mod << 16);
        if (nrec == 2u) { r.b = ins16 | (hp << 16); if (slot < P.sparse_cap) P.sparse[slot] = r; }
        r.b = ins16 | (256u << 16);
        if (slot + nrec - 1u < P.sparse_cap) P.sparse[slot + nrec - 1u] = r;
    }
}
__device__ __forceinline__ void w_sparse_open(WTile *T, uint32_t lane) {
    if (lane == 0) T->sp_state = kSpChunk;                                   // nothing reserved yet
    __syncwarp();
}
__device__ __forceinline__ void w_sparse_close(const DecodeParams &P, const WTile *T, uint32_t lane) {
    __syncwarp();
    const unsigned long long st = T->sp_state;
    for (uint32_t u = (uint32_t)(st & 0xffull) + lane; u < kSpChunk; u += 32u) w_sparse_sentinel(P, (st >> 8) + u);
}

__device__ __forceinline__ void w_add_cell(const DecodeParams &P, const WState &S, int32_t ref_pos, uint32_t outc,
                                           uint32_t ins16, int32_t hap, uint32_t is_mod) {
    int32_t hslot = -1;
    if (!P.haplotypes || hap < 0) hslot = 0;
    else if (hap + 1 < P.n_hap_slots) hslot = hap + 1;
    if (ins16 == 0u && outc < (uint32_t)P.n_code_slots && hslot >= 0) {
        const uint32_t per_pos = 2u * (uint32_t)P.n_code_slots * (uint32_t)P.n_hap_slots;
        const uint32_t within = (S.rev * (uint32_t)P.n_code_slots + outc) * (uint32_t)P.n_hap_slots + (uint32_t)hslot;
        red_add_u64(S.cells + ((unsigned long long)(uint32_t)ref_pos * per_pos + within), 1ull | ((unsigned long long)is_mod << 32));
    } else {
        w_add_sparse(P, (uint32_t)S.tid, S.rev, ref_pos, outc, ins16, hap, is_mod);
    }
}

__device__ __noinline__ void w_emit_view(const DecodeParams &P, uint32_t r, int32_t ref_pos, int32_t read_pos, uint32_t ins_off,
                                         unsigned long long order, uint32_t outc, uint32_t prob) {
    unsigned long long slot = atomicAdd(P.view_n, 1ull);
    if (slot < P.view_cap) {
        ViewDev v;
        v.read = r; v.ref_pos = ref_pos; v.read_pos = read_pos; v.ins_off = ins_off; v.order = order;
        v.code = (uint8_t)outc; v.prob = (uint8_t)prob;
        for (int z = 0; z < 6; ++z) v.pad[z] = 0;
        P.view[slot] = v;
    }
}

// Everything after "base q of the read is a call and maps to ref_pos" (SURVEY.md A.5-A.8); see process_call().
//   rd_code  2-bit code of the read base when the block's class pins it (C,G,T), else 4
__device__ __forceinline__ void w_call_at(const DecodeParams &P, WRead *R, const uint8_t *s_lut,
                                          const WBlock *bd, uint32_t blk_ord, uint32_t q, int32_t ref_pos, uint32_t ins_off,
                                          bool implicit, uint32_t cidx, uint32_t ml_base, uint32_t rd_code) {
    const WState &S = R->st;
    const uint8_t *ml = S.ml;
    const uint32_t *ref2 = S.ref2, *excm = S.excm;
    const uint32_t K = bd->K, is_n = bd->is_n;
    for (uint32_t m = 0; m < K; ++m) {
        const WCode cd = bd->code[m];
        if (cd.ri < 0) continue;                                  // src/mod.c:1157
        uint32_t prob = 0, is_mod = 0;
        if (!implicit) {                                          // issued first: independent of the context test
            const unsigned long long ml_idx = (unsigned long long)ml_base + (unsigned long long)cidx * K + m;
            if (ml_idx < S.ml_len) prob = ldg8(ml + ml_idx); else prob = 0x100u;
        }
        if (cd.ctx_mode != kCtxNone) {                            // src/mod.c:1162-1172
            uint32_t refcode = 0;
            int32_t in = cd.ctx_mode == kCtxFast ? ctx_fast(ref2, excm, S.ref_len, (uint32_t)ref_pos, cd.ctx_len, cd.pat2, refcode) : -1;
            if (in < 0) {
                if (!w_ctx_slow(P, S, (uint32_t)cd.ri, (uint32_t)ref_pos, q, is_n)) continue;
            } else {
                if (in == 0) continue;
                if (!is_n) {                                      // ref->forward[pos] == read base
                    uint32_t rc = rd_code;
                    if (rc > 3u) {
                        const uint32_t nib = (ldg8(S.seq + (q >> 1)) >> ((~q & 1u) << 2)) & 0xfu;
                        rc = nib == 1u ? 0u : nib == 2u ? 1u : nib == 4u ? 2u : nib == 8u ? 3u : 5u;
                    }
                    if (rc != refcode) continue;
                }
            }
        }
        if (prob > 0xffu) { w_raise(R, kErrMLIndex); return; }    // src/mod.c:1174
        if (P.subtool == 1) {                                     // FREQ
            if (!implicit) {
                const uint32_t f = cd.ri < kWLutSlots ? s_lut[cd.ri * 256 + prob] : P.req[cd.ri].lut[prob];   // src/mod.c:1181-1191
                if (!(f & 1u)) continue;
                is_mod = (f >> 1) & 1u;
            }                                                     // implicit: called, unmodified, no threshold (src/mod.c:1279)
            const uint32_t ins16 = ins_off & 0xffffu;             // make_key's uint16_t (src/mod.c:428)
            if (P.haplotypes) w_add_cell(P, S, ref_pos, cd.outc, ins16, (int32_t)S.hp, is_mod);
            w_add_cell(P, S, ref_pos, cd.outc, ins16, -1, is_mod);                 // src/mod.c:906-928
        } else {                                                  // VIEW
            w_emit_view(P, S.r, ref_pos, (int32_t)(S.rev ? S.L - 1u - q : q), ins_off,
                        ((unsigned long long)blk_ord << 40) | ((unsigned long long)(implicit ? 1u : 0u) << 39) | ((unsigned long long)cidx << 8) | m,
                        cd.outc, prob);
        }
    }
}

// "base q of the read is a call" (SURVEY.md A.4): read position -> reference position, then the above
__device__ __forceinline__ void w_call(const DecodeParams &P, WRead *R, const uint32_t *flex, const uint8_t *s_lut,
                                       const WBlock *bd, uint32_t blk_ord, uint32_t q, bool implicit, uint32_t cidx,
                                       uint32_t ml_base, uint32_t rd_code) {
    const WState &S = R->st;
    AlnHit h = w_cigar_lookup(S, flex, q);
    int32_t ref_pos = h.aln;
    if (P.insertions && ref_pos < 0) {
        if (implicit && S.rev) ref_pos = w_cigar_lookup(S, flex, S.L - 1u - q).ins;    // Q9 (src/mod.c:1234,1314)
        else ref_pos = h.ins;
    }
    if (ref_pos < 0) return;                                      // src/mod.c:1127,1237,1317
    w_call_at(P, R, s_lut, bd, blk_ord, q, ref_pos, P.insertions ? h.insoff : 0u, implicit, cidx, ml_base, rd_code);
}

// commas (or any byte c) of 16 text bytes as a 16-bit mask (bit i <-> byte i)
__device__ __forceinline__ uint32_t byte_mask16(uint4 v, uint32_t c) {
    const uint32_t cc = c * 0x01010101u, M = 0x00204081u;
    uint32_t a = (zero_bytes(v.x ^ cc) * M) >> 28, b = (zero_bytes(v.y ^ cc) * M) >> 28;
    uint32_t d = (zero_bytes(v.z ^ cc) * M) >> 28, e = (zero_bytes(v.w ^ cc) * M) >> 28;
    return a | (b << 4) | (d << 8) | (e << 12);
}

__device__ __forceinline__ void w_defer(uint32_t *defer_list, uint32_t *defer_n, uint32_t r, uint32_t lane) {
    if (lane == 0) defer_list[atomicAdd(defer_n, 1u)] = r;
}
__device__ __forceinline__ void w_report(const DecodeParams &P, uint32_t r, uint32_t code, uint32_t lane) {
    if (lane == 0) atomicMin(P.err, ((unsigned long long)r << 32) | code);
}

// one MM block header (src/mod.c:1003-1062); same rules as k_decode's (2b).  Returns an error code.
template <bool TILE>
__device__ __noinline__ uint32_t w_parse_header(const DecodeParams &P, uint32_t aoff, uint32_t blk) {
    WRead *R = w_arena<TILE>(aoff).R;
    const WState &S = R->st;
    WBlock &bd = R->blk[blk];
    const uint32_t start = blk == 0 ? 0u : R->semi[blk - 1u] + 1u;
    const uint32_t end = blk < S.n_semi ? R->semi[blk] : S.mm_len;
    const uint8_t *mmg = S.mm;
    // the header is a handful of characters: fetch 48 bytes around it at once, parse from the copy
    const uint32_t w0 = start & ~15u;
    uint4 win[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) win[k] = w0 + 16u * k < S.mm_len ? ld16(mmg + w0 + 16u * k) : make_uint4(0, 0, 0, 0);
    const uint8_t *wb = reinterpret_cast<const uint8_t *>(win);
    auto at = [&](uint32_t i) -> uint32_t { return i - w0 < 48u ? (uint32_t)wb[i - w0] : ldg8(mmg + i); };
    bd.end = end; bd.hdr_end = end; bd.K = 0; bd.any_req = 0; bd.cls = 0; bd.is_n = 0; bd.dot = 1;
    bd.o_idx = 0; bd.o_rd = 0; bd.rshift = 0; bd.cnt_cls = 0; bd.o_bm = 0; bd.n_calls = 0; bd.last1 = 0; bd.ml_base = 0; bd.tile0 = 0; bd.n_tiles = 0;
    for (int k = 0; k < kMaxCodes; ++k) { WCode c; c.ri = -1; c.outc = 0; c.ctx_mode = kCtxNone; c.ctx_len = 0; c.pat2 = 0; bd.code[k] = c; }
    uint32_t i = start;
    const uint32_t base_c = i < end ? at(i) : 0u;
    const bool okb = base_c == 'A' || base_c == 'C' || base_c == 'G' || base_c == 'T' || base_c == 'U' || base_c == 'N' ||
                     base_c == 'a' || base_c == 'c' || base_c == 'g' || base_c == 't' || base_c == 'u' || base_c == 'n';
    if (!okb) return kErrMMBase;
    ++i;
    const uint32_t modbase = base_c == 'U' ? (uint32_t)'T' : base_c;                 // src/mod.c:1006
    const uint32_t strand_c = i < end ? at(i) : 0u;
    if (strand_c != '+' && strand_c != '-') return kErrMMStrand;
    ++i;
    unsigned long long codes = 0;                                                    // up to 8 code characters
    uint32_t j = 0; bool has_num = false, has_alpha = false, bad = false, too_many = false;
    while (i < end) {
        const uint32_t c = at(i);
        if (c == ',' || c == '?' || c == '.') break;
        if (c >= '0' && c <= '9') has_num = true;
        else if ((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z')) has_alpha = true;
        else { bad = true; break; }
        if (j < (uint32_t)kMaxCodes) codes |= (unsigned long long)c << (8u * j); else too_many = true;
        ++j; ++i;
    }
    if (bad || j == 0 || (has_num && has_alpha)) return kErrMMCode;
    if (too_many) return kErrTooManyCodes;
    const uint32_t K = has_num ? 1u : j;                                             // src/mod.c:1048
    if (i < end && (at(i) == '?' || at(i) == '.')) { bd.dot = at(i) == '.'; ++i; }
    bd.hdr_end = i;
    bd.K = (uint8_t)K;
    uint32_t mb = modbase;                                                           // src/mod.c:1092-1093, table :98
    if (S.rev) {
        switch (modbase) {
            case 'A': mb = 'T'; break; case 'C': mb = 'G'; break; case 'G': mb = 'C'; break;
            case 'T': mb = 'A'; break; case 'N': mb = 'N'; break;
            case 'a': mb = 't'; break; case 'c': mb = 'g'; break; case 'g': mb = 'c'; break;
            case 't': mb = 'a'; break; case 'u': mb = 'a'; break; case 'n': mb = 'n'; break;
        }
    }
    const uint32_t up = mb >= 'a' ? mb - 32u : mb;
    bd.cls = up == 'A' ? 0 : up == 'C' ? 1 : up == 'G' ? 2 : (up == 'T' || up == 'U') ? 3 : 4;
    bd.is_n = modbase == 'N';
    for (uint32_t m = 0; m < K; ++m) {
        const unsigned long long key = has_num ? codes : codes >> (8u * m);          // code m = suffix string (Q4)
        int32_t ri = -1, outc = 0;
        if (P.wild_req >= 0) {
            outc = code_id(P.code_keys, key, true);
            if (outc < 0) return kErrTooManyCodes;
            ri = P.wild_req;
        } else {
            for (int32_t q = 0; q < P.n_req; ++q)
                if (P.req[q].key == key) { ri = q; outc = q; break; }
        }
        if (ri < 0) continue;
        const ReqMod &rq = P.req[ri];
        WCode c;
        c.ri = (int8_t)ri; c.outc = (uint8_t)outc; c.ctx_len = (uint8_t)rq.ctx_len;
        c.ctx_mode = (uint8_t)((P.insertions || rq.ctx_len == 0) ? kCtxNone : rq.fast_ctx ? kCtxFast : kCtxSlow);
        c.pat2 = S.rev ? rq.pat2_rc : rq.pat2;
        bd.code[m] = c;
        bd.any_req = 1;
    }
    return kErrNone;
}

// does block b need the rank index of its class / an explicit-rank bitmap?
__device__ __forceinline__ bool w_needs_index(const WBlock *bd) { return bd->any_req && (!bd->is_n || bd->dot); }
__device__ __forceinline__ bool w_needs_bitmap(const WBlock *bd) { return bd->any_req && bd->dot; }

// ---------------------------------------------------------------------------------------
// phase 1 (per read): batch record -> WState, MM block table, scratch layout, CIGAR prefix sums.
// Returns false when the read is finished already (deferred to the next kernel, or fatal and reported).
//   fused path: fa == nullptr, the scratch is the warp's arena `flex` of flex_words words; one rank index
//               and one bitmap are shared by all blocks (rebuilt when the class changes)
//   flat path:  the scratch comes from the global pool `fa`; one index per distinct class and one
//               bitmap per '.' block, so that later kernels can work on any block / tile independently
// ---------------------------------------------------------------------------------------
//   flex_words   capacity of the arena that will hold dir|cq|cr + index + bitmap (decides the sampling shifts)
//   local_words  capacity of THIS arena for dir|cq|cr (k_flat_setup's arenas are smaller than the consumer's)
// First ML index of every block (src/mod.c:1200: ml_start_idx advances by tokens x codes block after block), without decoding
// the skip counts: the tokens of a block are counted with the token-start definition of w_tile_ranks (a byte of the block's
// list that follows a ',' -- or is the list's first byte -- and is not a ',' itself).  Lets the blocks of one read be decoded
// by different warps (k_decode_stream, split mode).  Warp-wide; blocks 1.. get bd->ml_base, block 0 keeps 0.
__device__ __noinline__ void w_block_ml_bases(WRead *R, uint32_t lane) {
    const WState &S = R->st;
    const uint8_t *mm = S.mm;
    const uint32_t mm_len = S.mm_len, n_blocks = S.n_blocks;
    uint32_t acc = 0;
    for (uint32_t j = 0; j + 1u < n_blocks; ++j) {
        const uint32_t a0 = R->blk[j].hdr_end, a1 = R->blk[j].end, K = R->blk[j].K;
        uint32_t cnt = 0, carry_bit = 0;                                            // carry_bit: the byte before this step's first chunk is ','
        for (uint32_t base = a0 & ~15u; base < a1; base += 512u) {
            const uint32_t p0 = base + lane * 16u;
            uint32_t cm = 0;
            if (p0 < a1 && p0 < mm_len) cm = byte_mask16(ld16(mm + p0), ',');
            uint32_t pbit = (__shfl_up_sync(kFull, cm, 1) >> 15) & 1u;
            if (lane == 0) pbit = carry_bit;
            uint32_t st = ((cm << 1) | pbit) & ~cm & 0xffffu;
            uint32_t lo_b = a0 > p0 ? a0 - p0 : 0u, hi_b = a1 > p0 ? a1 - p0 : 0u;
            if (lo_b > 16u) lo_b = 16u;
            if (hi_b > 16u) hi_b = 16u;
            if (a0 >= p0 && a0 < p0 + 16u && !((cm >> lo_b) & 1u)) st |= 1u << lo_b;
            st &= ~((1u << lo_b) - 1u);
            st &= (1u << hi_b) - 1u;
            cnt += (uint32_t)__popc(st);
            carry_bit = (__shfl_sync(kFull, cm, 31) >> 15) & 1u;
        }
        const uint32_t tokens = __shfl_sync(kFull, warp_incl_scan(cnt, lane), 31);
        if (tokens > 0u) acc += tokens * K;
        __syncwarp();
        if (lane == 0) R->blk[j + 1u].ml_base = acc;
    }
    __syncwarp();
}

template <bool TILE>
__device__ __noinline__ bool w_setup_read(const DecodeParams &P, uint32_t aoff, uint32_t flex_words, uint32_t local_words,
                                          uint32_t *defer_list, uint32_t *defer_n, uint32_t r, uint32_t lane, const FlatAlloc *stream = nullptr) {
    const WArena A = w_arena<TILE>(aoff);
    WRead *R = A.R;
    uint32_t *flex = A.flex;                                                         // (stream mode: re-pointed at the read's slice of the pool below)
    WState &S = R->st;
    const int32_t tid = P.tid[r];
    const uint32_t L = P.l_seq[r], n_cig = P.n_cigar[r], mm_len = P.mm_len[r];
    const uint8_t *mm = P.mm + P.mm_off[r];
    if (tid < 0 || tid >= P.n_contigs || P.contigs[tid].ref2 == nullptr) { w_report(P, r, kErrNoContig, lane); return false; }
    if (L >= kWMaxL) { w_defer(defer_list, defer_n, r, lane); return false; }
    const uint32_t *cig = P.cigar + P.cigar_off[r];
    const int32_t pos = P.pos[r];
    const uint32_t ref_len = P.contigs[tid].len;
    if (lane == 0) {
        const ContigDev cd = P.contigs[tid];
        S.cig = cig; S.seq = P.seq4 + P.seq_off[r]; S.mm = mm; S.ml = P.ml + P.ml_off[r];
        S.ref2 = cd.ref2; S.excm = cd.excm; S.cells = cd.cells; S.ref_len = cd.len;
        S.r = r; S.L = L; S.n_cig = n_cig; S.mm_len = mm_len; S.ml_len = P.ml_len[r];
        S.rev = (P.flag[r] >> 4) & 1u; S.hp = P.hp[r]; S.tid = tid; S.pos = pos;
        S.err = 0; S.cur_cls = 0xffu; S.cur_blk = 0; S.n_blocks = 0; S.pairs = stream ? 1u : 0u; S.pad_ = 0;
    }

    // ---- MM blocks: positions of ';'
    uint32_t n_semi = 0;
    for (uint32_t base = 0; base < mm_len; base += 512u) {
        const uint32_t p0 = base + lane * 16u;
        uint32_t mask = 0;
        if (p0 < mm_len) {
            mask = byte_mask16(ld16(mm + p0), ';');
            const uint32_t valid = mm_len - p0;
            if (valid < 16u) mask &= (1u << valid) - 1u;
        }
        const uint32_t c = (uint32_t)__popc(mask);
        const uint32_t incl = warp_incl_scan(c, lane);
        uint32_t ord = n_semi + incl - c;
        while (mask) {
            const uint32_t bit = (uint32_t)__ffs((int)mask) - 1u;
            mask &= mask - 1u;
            if (ord < (uint32_t)kWBlocks) R->semi[ord] = p0 + bit;
            ++ord;
        }
        n_semi += __shfl_sync(kFull, incl, 31);
    }
    uint32_t n_blocks = n_semi;
    if (mm_len > 0 && mm[mm_len - 1u] != ';') n_blocks += 1;                        // unterminated last block
    if (n_blocks > (uint32_t)kWBlocks) { w_defer(defer_list, defer_n, r, lane); return false; }
    if (lane == 0) { S.n_semi = n_semi; S.n_blocks = n_blocks; }
    __syncwarp();

    // ---- block headers, one lane each
    uint32_t herr = kErrNone, my_idx = 0, my_bm = 0, my_cls = 0;
    if (lane < n_blocks) {
        herr = w_parse_header<TILE>(P, aoff, lane);
        if (herr == kErrNone) { my_idx = w_needs_index(&R->blk[lane]); my_bm = w_needs_bitmap(&R->blk[lane]); my_cls = R->blk[lane].cls; }
    }
    __syncwarp();                                                                    // block table complete before lane 0 extends it
    const uint32_t herr_mask = __ballot_sync(kFull, herr != kErrNone);
    if (herr_mask) herr = __shfl_sync(kFull, herr, __ffs((int)herr_mask) - 1);
    const uint32_t idx_mask = __ballot_sync(kFull, my_idx != 0u), bm_mask = __ballot_sync(kFull, my_bm != 0u);
    uint32_t cls_set = 0;                                                            // classes that need an index
    for (uint32_t b = 0; b < n_blocks; ++b) {
        const uint32_t c = __shfl_sync(kFull, my_cls, (int)b);
        if ((idx_mask >> b) & 1u) cls_set |= 1u << c;
    }

    // ---- scratch layout
    const uint32_t n_u4 = (L + 31u) >> 5;
    uint32_t gshift = stream ? 5u : 8u;                                              // stream: the table lives in HBM, fine buckets keep the search to a step
    while (!stream && ((L >> gshift) + 2u) > 160u) ++gshift;
    const uint32_t n_dir = (L >> gshift) + 2u;
    const uint32_t n_rd = (L >> 6) + 2u;
    const uint32_t bm_words = ((L + 31u) >> 5) + 1u;
    // one index / bitmap, reused block after block; the streaming consumer (mmc_decode_stream.cuh) needs neither
    const uint32_t n_idx = (idx_mask && !stream) ? 1u : 0u, n_bm = (bm_mask && !stream) ? 1u : 0u;
    const uint32_t cap = flex_words;
    (void)cls_set;
    uint32_t cshift = 0, ishift = 0, n_samp, n_ent, need;
    auto a4 = [](uint32_t x) { return (x + 3u) & ~3u; };
    for (;;) {
        n_samp = (n_cig + (1u << cshift) - 1u) >> cshift;
        if (n_samp == 0u) n_samp = 1u;
        n_ent = (n_u4 + (1u << ishift) - 1u) >> ishift;
        need = a4(n_dir + 2u * n_samp) + n_idx * a4(n_ent + 2u + n_rd) + n_bm * a4(bm_words);   // 16-byte aligned pieces
        if (stream || (need <= cap && a4(n_dir + 2u * n_samp) <= local_words)) break;   // stream: un-sampled, in the pool
        if (cshift >= (uint32_t)kWMaxCShift && ishift >= (uint32_t)kWMaxIShift) { w_defer(defer_list, defer_n, r, lane); return false; }
        if ((2u * n_samp >= n_ent && cshift < (uint32_t)kWMaxCShift) || ishift >= (uint32_t)kWMaxIShift) ++cshift;
        else ++ishift;
    }
    const uint32_t o_dir = 0, o_cq = stream ? a4(n_dir) : n_dir, o_cr = o_cq + n_samp, o_var = a4(o_cq + 2u * n_samp);
    if (stream) {                                                                    // dir | {cq, cr} pairs are written straight into the pool
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(stream->cursor, (unsigned long long)o_var);
        base = ((unsigned long long)__shfl_sync(kFull, (uint32_t)(base >> 32), 0) << 32) | __shfl_sync(kFull, (uint32_t)base, 0);
        if (base + o_var > stream->cap) { w_defer(defer_list, defer_n, r, lane); return false; }
        flex = stream->pool + base;
    }
    uint2 *pairs = reinterpret_cast<uint2 *>(flex + o_cq);
    if (lane == 0) {
        S.flex = flex; S.flex_home = flex;
        S.o_dir = o_dir; S.o_cq = o_cq; S.o_cr = o_cr;
        S.cshift = cshift; S.gshift = gshift; S.ishift = ishift; S.n_samp = n_samp; S.n_ent = n_ent; S.n_u4 = n_u4; S.n_rd = n_rd;
        // per block: where its class index / bitmap live
        uint32_t next = o_var;
        for (uint32_t b = 0; b < n_blocks; ++b) {
            WBlock &bd = R->blk[b];
            if ((idx_mask >> b) & 1u) {
                bd.o_idx = o_var; bd.o_rd = o_var + n_ent + 2u;
            }
        }
        next = o_var + n_idx * a4(n_ent + 2u + n_rd);
        S.n_stage = o_var;                                         // dir | cq | cr: what k_flat_setup hands over
        for (uint32_t b = 0; b < n_blocks; ++b) {
            WBlock &bd = R->blk[b];
            if ((bm_mask >> b) & 1u) bd.o_bm = next;
        }
    }

    if (stream && n_blocks > 1u && !herr_mask) { __syncwarp(); w_block_ml_bases(R, lane); }

    // ---- CIGAR prefix sums == get_aln() (src/mod.c:811-880)
    uint32_t carry_q = 0, carry_r = 0, big = 0;
    const uint32_t cmask = (1u << cshift) - 1u;
    const uint32_t rem_ref = (pos >= 0 && (uint32_t)pos < ref_len) ? ref_len - (uint32_t)pos : 0u;   // reference bases from pos on
    if (n_cig == 0u && lane == 0) { if (stream) pairs[0] = make_uint2(15u, 0u); else { flex[o_cq] = 15u; flex[o_cr] = 0; } }
    if (!stream) {
    // arena layout: 32 ops per step; the next step's words are loaded before this step's are used.  Plain warp scans are exact
    // while every length is < 2^26 (32 x 2^26 = 2^31; the running totals saturate at kSat like sat_add); a read with a longer
    // op is left to k_decode, whose block scans saturate at every step.
        uint32_t w_next = lane < n_cig ? ldg32(cig + lane) : 0u;
        for (uint32_t base = 0; base < n_cig; base += 32u) {
            const uint32_t i = base + lane, w = w_next;
            if (i + 32u < n_cig) w_next = ldg32(cig + i + 32u);
            uint32_t op = 15u, len = 0, ql = 0, rl = 0;
            if (i < n_cig) {
                op = w & 15u; len = w >> 4;
                if (op == 0u || op == 7u || op == 8u) { ql = len; rl = len; }
                else if (op == 1u || op == 4u) ql = len;
                else if (op == 2u || op == 3u) rl = len;
                else if (op == 5u) w_raise(R, kErrHardClip);
                else w_raise(R, kErrCigarOp);
            }
            big |= len >> 26;
            const uint32_t iq = warp_incl_scan(ql, lane), ir = warp_incl_scan(rl, lane);
            if (i < n_cig) {
                uint32_t q0 = carry_q + (iq - ql), r0 = carry_r + (ir - rl);
                if (q0 > kSat) q0 = kSat;
                if (r0 > kSat) r0 = kSat;
                if (len > 0u) {
                    const bool alnop = op == 0u || op == 7u || op == 8u;
                    if ((alnop || (op == 1u && P.insertions)) && q0 + ql > L) w_raise(R, kErrCigarLen);
                    if (alnop && r0 + len > rem_ref) w_raise(R, kErrRefRange);           // pos + r0 + len - 1 >= ref_len, or pos < 0
                }
                if ((i & cmask) == 0u) { flex[o_cq + (i >> cshift)] = ((q0 < kWMaxL ? q0 : kWMaxL) << 4) | op; flex[o_cr + (i >> cshift)] = r0; }
                if (ql > 0u && q0 < L) {                              // directory: sample that holds each bucket's first base
                    const uint32_t qe = q0 + ql < L ? q0 + ql : L;
                    for (uint32_t b = (q0 + (1u << gshift) - 1u) >> gshift; (b << gshift) < qe; ++b) flex[o_dir + b] = i >> cshift;
                }
            }
            carry_q += __shfl_sync(kFull, iq, 31); if (carry_q > kSat) carry_q = kSat;
            carry_r += __shfl_sync(kFull, ir, 31); if (carry_r > kSat) carry_r = kSat;
        }
    } else {
    // stream layout (long CIGARs): 128 ops per step, four consecutive ops per lane (one 16-byte load), branch-free op tables,
    // a serial prefix inside the lane and ONE pair of warp scans over the lanes' sums; the fatal conditions are collected
    // as flags and raised after the loop.  Exact while every length is < 2^24 (128 x 2^24 = 2^31).
    uint32_t bad = 0;                                                                // 1 hard clip, 2 unhandled op, 4 CIGAR past the read, 8 past the contig
    const uint32_t insq = P.insertions ? 0x183u : 0x181u;                            // ops whose query range must lie inside the read: M = X (+ I)
    uint4 w4_next = make_uint4(0, 0, 0, 0);
    if (lane * 4u < n_cig) w4_next = ld16(reinterpret_cast<const uint8_t *>(cig + lane * 4u));   // (the pool has 64 bytes of slack past the last slice)
    for (uint32_t base = 0; base < n_cig; base += 128u) {
        const uint32_t i0 = base + lane * 4u;
        const uint4 w4 = w4_next;                                                    // the next step's words are on their way while this step scans and stores
        w4_next = make_uint4(0, 0, 0, 0);
        if (i0 + 128u < n_cig) w4_next = ld16(reinterpret_cast<const uint8_t *>(cig + i0 + 128u));
        const uint32_t wv[4] = {w4.x, w4.y, w4.z, w4.w};
        uint32_t qlv[4], rlv[4], sq = 0, sr = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t valid = i0 + (uint32_t)k < n_cig ? 1u : 0u, op = wv[k] & 15u, len = valid ? wv[k] >> 4 : 0u;
            const uint32_t ql = len & (0u - ((0x193u >> op) & 1u)), rl = len & (0u - ((0x18du >> op) & 1u));   // M I S = X / M D N = X
            bad |= valid & (op == 5u ? 1u : 0u);
            bad |= (valid & (((0x19fu | 0x20u) >> op) & 1u ^ 1u)) << 1;
            big |= len >> 24;
            qlv[k] = ql; rlv[k] = rl; sq += ql; sr += rl;
        }
        if (__ballot_sync(kFull, big != 0u)) break;                                  // (deferred below)
        const uint32_t iq = warp_incl_scan(sq, lane), ir = warp_incl_scan(sr, lane);
        uint32_t q0 = carry_q + (iq - sq), r0 = carry_r + (ir - sr);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t i = i0 + (uint32_t)k, op = wv[k] & 15u, ql = qlv[k], rl = rlv[k];
            if (q0 > kSat) q0 = kSat;
            if (r0 > kSat) r0 = kSat;
            if (i < n_cig) {
                bad |= ((ql > 0u && ((insq >> op) & 1u) && q0 + ql > L) ? 4u : 0u) | ((rl > 0u && ((0x181u >> op) & 1u) && r0 + rl > rem_ref) ? 8u : 0u);
                pairs[i] = make_uint2(((q0 < kWMaxL ? q0 : kWMaxL) << 4) | op, r0);
                const uint32_t qe = q0 + ql < L ? q0 + ql : L;                    // directory: op that holds each bucket's first base
                uint32_t b = (q0 + 31u) >> 5;
                if (ql > 0u && (b << 5) < qe) {
                    flex[o_dir + b] = i;
                    for (++b; (b << 5) < qe; ++b) flex[o_dir + b] = i;
                }
            }
            q0 += ql; r0 += rl;
        }
        carry_q += __shfl_sync(kFull, iq, 31); if (carry_q > kSat) carry_q = kSat;
        carry_r += __shfl_sync(kFull, ir, 31); if (carry_r > kSat) carry_r = kSat;
    }
    const uint32_t allbad = __ballot_sync(kFull, bad & 1u) ? 1u : __ballot_sync(kFull, bad & 2u) ? 2u : __ballot_sync(kFull, bad & 4u) ? 4u : __ballot_sync(kFull, bad & 8u) ? 8u : 0u;
    if (allbad) w_raise(R, allbad == 1u ? kErrHardClip : allbad == 2u ? kErrCigarOp : allbad == 4u ? kErrCigarLen : kErrRefRange);
    }
    if (__ballot_sync(kFull, big != 0u)) { w_defer(defer_list, defer_n, r, lane); return false; }
    if (lane == 0) S.total_q = carry_q < L ? carry_q : L;
    uint32_t err = w_err(R);
    if (!err) err = herr;
    if (err) { if (lane == 0) S.n_blocks = 0; w_report(P, r, err, lane); return false; }
    if (lane == 0) {
        const int32_t lo = pos > 0 ? pos - 1 : 0;
        long long hi = (long long)pos + carry_r + 1;
        if (hi > (long long)ref_len) hi = ref_len;
        if (lo < *reinterpret_cast<volatile int32_t *>(&P.touch_lo[tid])) atomicMin(&P.touch_lo[tid], lo);
        if ((int32_t)hi > *reinterpret_cast<volatile int32_t *>(&P.touch_hi[tid])) atomicMax(&P.touch_hi[tid], (int32_t)hi);
    }
    return true;
}

// ---------------------------------------------------------------------------------------
// phase 2 (per base class used by the read's blocks): rank index over the 4-bit SEQ
//   idx[e]  = bases of the class before entry e (32 << ishift bases per entry), idx[n_ent] = total
//   rd[j]   = entry that holds the class base of rank j << rshift
// ---------------------------------------------------------------------------------------
template <bool C0>
__device__ __forceinline__ void w_count_entries(const WState &S, uint32_t *idx, uint32_t pat, uint32_t lane) {
    const uint8_t *seq = S.seq;
    const uint32_t n_ent = S.n_ent, n_u4 = S.n_u4, ishift = S.ishift, tail = S.L & 31u;
    if (ishift == 0u) {
        const uint32_t n_full = tail ? n_u4 - 1u : n_u4;          // vectors with 32 valid bases
        uint32_t e = lane;
        for (; e + 32u < n_full; e += 64u) {                      // two loads in flight
            const uint4 v0 = ld16(seq + (size_t)e * 16u), v1 = ld16(seq + (size_t)(e + 32u) * 16u);
            idx[e] = count_u4<C0>(v0, pat);
            idx[e + 32u] = count_u4<C0>(v1, pat);
        }
        for (; e < n_full; e += 32u) idx[e] = count_u4<C0>(ld16(seq + (size_t)e * 16u), pat);
        if (tail && lane == 0) idx[n_u4 - 1u] = count_u4_tail<C0>(ld16(seq + (size_t)(n_u4 - 1u) * 16u), pat, tail);
    } else {
        for (uint32_t e = lane; e < n_ent; e += 32u) {
            uint32_t u0 = e << ishift, u1 = u0 + (1u << ishift), c = 0;
            if (u1 > n_u4) u1 = n_u4;
            for (uint32_t u = u0; u < u1; ++u) {
                const uint4 v = ld16(seq + (size_t)u * 16u);
                c += (u == n_u4 - 1u && tail) ? count_u4_tail<C0>(v, pat, tail) : count_u4<C0>(v, pat);
            }
            idx[e] = c;
        }
    }
}

// builds the index of block jb's class at (bd->o_idx, bd->o_rd) and fills bd->cnt_cls / bd->rshift
__device__ __noinline__ void w_build_index(uint32_t aoff, uint32_t jb, uint32_t lane) {
    const WArena A = w_arena(aoff);
    WRead *R = A.R;
    WState &S = R->st;
    WBlock *bd = &R->blk[jb];
    uint32_t *flex = A.flex;
    uint32_t *idx = flex + bd->o_idx, *rd = flex + bd->o_rd;
    const uint32_t cls = bd->cls, n_ent = S.n_ent, pat = class_pat(cls);
    __syncwarp();
    if (cls == 0u) w_count_entries<true>(S, idx, pat, lane); else w_count_entries<false>(S, idx, pat, lane);
    __syncwarp();
    // exclusive prefix: every lane owns an odd-length run of consecutive entries (conflict-free banks)
    const uint32_t seg = ((n_ent + 31u) >> 5) | 1u;
    uint32_t e0 = lane * seg, e1 = e0 + seg, sum = 0;
    if (e0 > n_ent) e0 = n_ent;
    if (e1 > n_ent) e1 = n_ent;
    for (uint32_t e = e0; e < e1; ++e) sum += idx[e];
    const uint32_t incl = warp_incl_scan(sum, lane);
    const uint32_t total = __shfl_sync(kFull, incl, 31);
    uint32_t run = incl - sum;
    for (uint32_t e = e0; e < e1; ++e) { const uint32_t c = idx[e]; idx[e] = run; run += c; }
    uint32_t rshift = 4;
    while (((total >> rshift) + 1u) > S.n_rd) ++rshift;
    if (lane == 0) { idx[n_ent] = total; bd->cnt_cls = total; bd->rshift = rshift; }
    __syncwarp();
    // rank directory: each multiple of 2^rshift below `total` lies in exactly one entry's rank range
    for (uint32_t e = lane; e < n_ent; e += 32u) {
        const uint32_t a = idx[e], b = idx[e + 1u];
        for (uint32_t j = (a + (1u << rshift) - 1u) >> rshift; (j << rshift) < b; ++j) rd[j] = e;
    }
    __syncwarp();
}

// skip counts with 5..9 digits: cold
__device__ __noinline__ uint32_t w_parse_long(const uint8_t *tx, uint32_t nd, uint32_t *bad) {
    uint32_t val = 0;
    for (uint32_t k = 0; k < nd; ++k) {
        const uint32_t d = tx[k];
        if (d < '0' || d > '9') *bad = 1;
        val = val * 10u + (d - '0');
    }
    return val;
}

// ---------------------------------------------------------------------------------------
// phase 3 (per text tile of a block): skip counts -> base ranks (src/mod.c:1066-1098).
// The tile is 31 chunks of 16 bytes starting at tb (lane 31 holds look-ahead text only).
// Leaves carry_in + (prefix sum of skip+1) - 1 of the tile's tokens in T->rank[0..n), returns n
// and the saturating sum of the tile's (skip+1) in *sum_out.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t w_tile_ranks(WRead *R, WTile *T, uint32_t tb, uint32_t a0, uint32_t a1, uint32_t carry_in,
                                                 uint32_t *sum_out, uint32_t lane) {
    const WState &S = R->st;
    const uint32_t mm_len = S.mm_len;
    const uint8_t *mm = S.mm;
    const uint32_t p0 = tb + lane * 16u;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (p0 < mm_len) v = ld16(mm + p0);
    T->text[lane] = v;
    const uint32_t cm = byte_mask16(v, ',');
    const uint32_t up = __shfl_up_sync(kFull, cm, 1);
    const uint32_t ncm = __shfl_down_sync(kFull, cm, 1);
    uint32_t pbit = (up >> 15) & 1u;
    if (lane == 0) pbit = (tb > a0 && ldg8(mm + tb - 1u) == ',') ? 1u : 0u;
    uint32_t st = ((cm << 1) | pbit) & ~cm & 0xffffu;              // token starts: previous byte is ','
    // clip to [a0, a1) and force a start at a0 (src/mod.c:1066: the list begins right after the header)
    uint32_t lo_b = a0 > p0 ? a0 - p0 : 0u, hi_b = a1 > p0 ? a1 - p0 : 0u;
    if (lo_b > 16u) lo_b = 16u;
    if (hi_b > 16u) hi_b = 16u;
    if (a0 >= p0 && a0 < p0 + 16u && !((cm >> lo_b) & 1u)) st |= 1u << lo_b;
    st &= ~((1u << lo_b) - 1u);
    st &= (1u << hi_b) - 1u;
    if (lane >= (uint32_t)kWChunks) st = 0;
    uint32_t em = cm | (ncm << 16);                               // token terminators: ',' or the block end
    if (a1 >= p0 && a1 - p0 < 32u) em |= 1u << (a1 - p0);
    T->em[lane] = em;
    const uint32_t n_tok = (uint32_t)__popc(st);
    const uint32_t incl = warp_incl_scan(n_tok, lane);
    const uint32_t tile_cnt = __shfl_sync(kFull, incl, 31);
    uint32_t slot = incl - n_tok;
    const uint32_t base_off = lane * 16u;
    while (st) {                                                  // byte offset of every token of this chunk
        T->rank[slot++] = base_off + (uint32_t)__ffs((int)st) - 1u;
        st &= st - 1u;
    }
    __syncwarp();
    const uint32_t *tx32 = reinterpret_cast<const uint32_t *>(T->text);
    uint32_t carry = carry_in, total = 0;
    for (uint32_t c0 = 0; c0 < tile_cnt; c0 += 32u) {
        const uint32_t c = c0 + lane;
        uint32_t x = 0;
        if (c < tile_cnt) {                                       // one token per lane: SWAR decimal parse
            const uint32_t off = T->rank[c];
            const uint32_t rest = T->em[off >> 4] >> ((off & 15u) + 1u);
            uint32_t nd = rest ? (uint32_t)__ffs((int)rest) : 32u, val = 0, bad = 0;
            if (nd > 9u) { bad = 1; nd = 0; }                     // src/mod.c:1080-1085
            if (nd <= 4u) {
                const uint32_t w0 = tx32[off >> 2], w1 = tx32[(off >> 2) + 1u];
                uint32_t dg = __funnelshift_r(w0, w1, (off & 3u) * 8u) - 0x30303030u;
                const uint32_t keep = nd >= 4u ? 0xffffffffu : ((1u << (8u * nd)) - 1u);
                bad |= (((dg + 0x76767676u) | dg) & 0x80808080u & keep) != 0u;
                dg = (dg & keep) << ((8u * (4u - nd)) & 31u);
                if (nd == 0u) dg = 0;
                const uint32_t pr = (dg * 10u + (dg >> 8)) & 0x00ff00ffu;              // (10*b0+b1) | (10*b2+b3) << 16
                val = (pr & 0xffffu) * 100u + (pr >> 16);
            } else {
                val = w_parse_long(reinterpret_cast<const uint8_t *>(T->text) + off, nd, &bad);
            }
            if (bad) { w_raise(R, kErrMMSkip); val = 0; }
            x = val + 1u < kWMaxL ? val + 1u : kWMaxL;             // >= 2^26 is past any read this path takes; 32 of them fit 32 bits
        }
        const uint32_t si = warp_incl_scan(x, lane);
        if (c < tile_cnt) T->rank[c] = sat_add(carry, si) - 1u;   // base_rank (src/mod.c:1098)
        const uint32_t rt = __shfl_sync(kFull, si, 31);
        carry = sat_add(carry, rt); total = sat_add(total, rt);
    }
    *sum_out = total;
    __syncwarp();
    return tile_cnt;
}

// ---------------------------------------------------------------------------------------
// phase 4, common case in one tight loop: `freq`, one requested code per block (K == 1), canonical base
// A/C/G/T, context "*" or <= 8 ACGT bases, un-sampled CIGAR and index, haplotype (if any) with a
// dense stratum.  Same arithmetic as the general passes below.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool w_fast_ok(const DecodeParams &P, const WState &S, const WBlock *bd) {
    const WCode cd = bd->code[0];
    return bd->K == 1 && cd.ri >= 0 && cd.ri < kWLutSlots && P.subtool == 1 &&
           (!P.haplotypes || (int32_t)S.hp + 1 < P.n_hap_slots) &&
           cd.ctx_mode != kCtxSlow && !bd->is_n && bd->cls != 4u && S.ishift == 0u &&
           (uint32_t)cd.outc < (uint32_t)P.n_code_slots;
}

//   C0  class A (needs the nibble of the read base for the reference comparison)
//   EX  any of: --insertions, --haplotypes, sampled CIGAR (kept out of the lean instantiation)
template <bool C0, bool EX>
__device__ __forceinline__ void w_tile_calls_fast(const DecodeParams &P, WRead *R, WTile *T, uint32_t *flex, const uint8_t *s_lut, const WBlock *bd,
                                                  uint32_t n, uint32_t cidx0, uint32_t ml_base, uint32_t lane) {
    const WState &S = R->st;
    const uint32_t *idx = flex + bd->o_idx, *rd = flex + bd->o_rd, *dir = flex + S.o_dir, *cq = flex + S.o_cq, *cr = flex + S.o_cr;
    uint32_t *bm = flex + bd->o_bm;
    const uint32_t rshift = bd->rshift, cnt_cls = bd->cnt_cls, need_bm = bd->dot, cls = bd->cls, pat = class_pat(cls);
    const uint32_t rev = S.rev, total_q = S.total_q, g = S.gshift, last_samp = S.n_samp - 1u, ml_len = S.ml_len, ref_len = S.ref_len;
    const int32_t pos = S.pos;
    const uint8_t *seq = S.seq, *ml = S.ml;
    const uint32_t *ref2 = S.ref2, *excm = S.excm;
    unsigned long long *cells = S.cells;
    const WCode cd = bd->code[0];
    const uint32_t m = cd.ctx_mode == kCtxFast ? cd.ctx_len : 0u, pat2 = cd.pat2;
    const uint32_t m2 = (1u << (2u * m)) - 1u, m1 = (1u << m) - 1u;
    // occurrences of the context that can cover the call: the one starting j bases into the window puts pattern base
    // m-1-j on the call's position, and ref == read base (src/mod.c:1164) forces that base to be the block's class
    uint32_t jmask = 0;
    for (uint32_t j = 0; j < m; ++j) jmask |= (uint32_t)(((pat2 >> (2u * (m - 1u - j))) & 3u) == (C0 ? 0u : cls)) << j;
    const uint8_t *lut = s_lut + cd.ri * 256;
    const uint32_t per_pos = 2u * (uint32_t)P.n_code_slots * (uint32_t)P.n_hap_slots;
    const uint32_t within = (rev * (uint32_t)P.n_code_slots + cd.outc) * (uint32_t)P.n_hap_slots;
    const uint32_t hslot = EX && P.haplotypes ? S.hp + 1u : 0u, insertions = EX ? (uint32_t)P.insertions : 0u, cshift = EX ? S.cshift : 0u;
    const uint32_t ml0 = ml_base + cidx0;
    // EX: every lane runs the same number of rounds and the warp meets at the top of each, where the sparse records
    // the previous round left (ins_offset > 0: no dense cell) are appended from the warp's chunk (w_sparse_flush).
    const uint32_t nrec = EX && hslot ? 2u : 1u;
    uint32_t sp_pos = 0, sp_meta = 0;                                                // ins16 | is_mod << 16
    bool pending = false;
    for (uint32_t c = lane; EX ? c - lane < n : c < n; c += 32u) {
        if (EX) {
            __syncwarp();
            w_sparse_flush(P, T, pending, nrec, (uint32_t)S.tid, rev, sp_pos, cd.outc, sp_meta & 0xffffu, S.hp, sp_meta >> 16, lane);
            pending = false;
            if (c >= n) continue;
        }
        const uint32_t rank = T->rank[c];
        if (rank >= cnt_cls) { w_raise(R, kErrMMRank); continue; }                  // src/mod.c:1116
        const uint32_t mi = ml0 + c;
        const uint32_t prob = mi < ml_len ? ldg8(ml + mi) : 0x100u;                // issued early
        // ---- select: k-th base of the class
        const uint32_t k = rev ? cnt_cls - 1u - rank : rank;
        uint32_t e = rd[k >> rshift], nxt = idx[e + 1u];
        while (nxt <= k) { ++e; nxt = idx[e + 1u]; }
        uint32_t rem = k - idx[e];
        const uint4 v = ld16(seq + (size_t)e * 16u);
        const uint32_t f0 = class_flags<C0>(v.x, pat), f1 = class_flags<C0>(v.y, pat), f2 = class_flags<C0>(v.z, pat), f3 = class_flags<C0>(v.w, pat);
        const uint32_t s0 = (uint32_t)__popc(f0), s1 = s0 + (uint32_t)__popc(f1), s2 = s1 + (uint32_t)__popc(f2);
        uint32_t f = f0, wsel = 0, sub = 0;
        if (rem >= s0) { f = f1; wsel = 8; sub = s0; }
        if (rem >= s1) { f = f2; wsel = 16; sub = s1; }
        if (rem >= s2) { f = f3; wsel = 24; sub = s2; }
        rem -= sub;
        uint32_t cc = (uint32_t)__popc(f & 0xffffu);
        if (rem >= cc) { rem -= cc; f >>= 16; wsel += 4; }
        cc = (uint32_t)__popc(f & 0xffu);
        if (rem >= cc) { rem -= cc; f >>= 8; wsel += 2; }
        const uint32_t q = e * 32u + wsel + ((rem != 0u || !(f & 0x80u)) ? 1u : 0u);
        if (need_bm) atomicOr(&bm[rank >> 5], 1u << (rank & 31u));
        // ---- map: aln[q]
        uint32_t ref_pos, ins16 = 0;
        if (EX && cshift != 0u) {                                                   // sampled CIGAR (long reads): generic lookup
            const AlnHit h = w_cigar_lookup(S, flex, q);
            if (h.aln >= 0) ref_pos = (uint32_t)h.aln;
            else if (insertions && h.ins >= 0) { ref_pos = (uint32_t)h.ins; ins16 = h.insoff & 0xffffu; }
            else continue;
        } else {
        if (q >= total_q) continue;
        const uint32_t b = q >> g;
        uint32_t lo = dir[b], hi = (((b + 1u) << g) < total_q) ? dir[b + 1u] : last_samp;
        const uint32_t qlim = (q + 1u) << 4;
        while (lo < hi) { const uint32_t mid = (lo + hi + 1u) >> 1; if (cq[mid] < qlim) lo = mid; else hi = mid - 1u; }
        const uint32_t ce = cq[lo], op = ce & 15u;
        if (op == 0u || op == 7u || op == 8u) ref_pos = (uint32_t)(pos + (int32_t)(cr[lo] + q - (ce >> 4)));
        else if (EX && insertions && op == 1u) {                                    // ins[] / ins_offset (src/mod.c:1122-1127)
            const int32_t left = pos + (int32_t)cr[lo] - 1;
            if (left < 0) continue;                                                 // Q11
            ref_pos = (uint32_t)left; ins16 = (q - (ce >> 4) + 1u) & 0xffffu;       // make_key's uint16_t (src/mod.c:428)
        } else continue;                                                            // src/mod.c:1127
        }
        // ---- update
        if (m) {
            if (ref_pos + 1u < m || ref_pos + m > ref_len) {                        // contig edge: generic test
                if (!w_ctx_slow(P, S, (uint32_t)cd.ri, ref_pos, q, 0u)) continue;
            } else {
                const uint32_t w0 = ref_pos + 1u - m, wi = w0 >> 4, ei = w0 >> 5;
                const uint32_t W = __funnelshift_r(ldg32(ref2 + wi), ldg32(ref2 + wi + 1u), (w0 & 15u) * 2u);
                const uint32_t E = __funnelshift_r(ldg32(excm + ei), ldg32(excm + ei + 1u), w0 & 31u);
                uint32_t hit = 0;
                for (uint32_t jm = jmask; jm; jm &= jm - 1u) {
                    const uint32_t j = (uint32_t)__ffs((int)jm) - 1u;
                    hit |= (uint32_t)((((W >> (2u * j)) & m2) == pat2) & (((E >> j) & m1) == 0u));
                }
                if (!hit) continue;                                                 // src/mod.c:1162-1172; a hit implies ref base == class base
                if (C0) {                                                           // class 0 = 'A' and every other nt16 letter
                    const uint32_t nib = (ldg8(seq + (q >> 1)) >> ((~q & 1u) << 2)) & 0xfu;
                    if (nib != 1u) continue;                                        // the read base is not literally 'A'
                }
            }
        }
        if (prob > 0xffu) { w_raise(R, kErrMLIndex); continue; }                    // src/mod.c:1174
        const uint32_t fl = lut[prob];                                              // src/mod.c:1181-1191
        if (!(fl & 1u)) continue;
        const unsigned long long inc = 1ull | ((unsigned long long)((fl >> 1) & 1u) << 32);
        if (!EX || ins16 == 0u) {
            unsigned long long *cell = cells + ((unsigned long long)ref_pos * per_pos + within);
            red_add_u64(cell, inc);                                                 // the '*' stratum (or the only one)
            if (EX && hslot) red_add_u64(cell + hslot, inc);                        // src/mod.c:906-928
        } else {
            pending = true; sp_pos = ref_pos; sp_meta = ins16 | (((fl >> 1) & 1u) << 16);
        }
    }
    if (EX) {
        __syncwarp();
        w_sparse_flush(P, T, pending, nrec, (uint32_t)S.tid, rev, sp_pos, cd.outc, sp_meta & 0xffffu, S.hp, sp_meta >> 16, lane);
    }
}

// ---------------------------------------------------------------------------------------
// phase 4 (per text tile), general form: the explicit calls whose ranks are in T->rank[0..n), in two passes
//   select        base rank -> read position q                                 (bases_pos[][], src/mod.c:1102-1113)
//   map + update  q -> reference position, context test, ML threshold, count   (src/mod.c:1122-1199)
// ---------------------------------------------------------------------------------------
constexpr uint32_t kNoCall = 0xffffffffu;

template <bool C0>
__device__ __forceinline__ void w_pass_select(WRead *R, WTile *T, uint32_t *flex, const WBlock *bd, uint32_t pat, uint32_t n, uint32_t need_bm, uint32_t lane) {
    const WState &S = R->st;
    const uint32_t rev = S.rev, cnt_cls = bd->cnt_cls;
    uint32_t *bm = flex + bd->o_bm;
    for (uint32_t c = lane; c < n; c += 32u) {
        const uint32_t r0 = T->rank[c];
        if (r0 >= cnt_cls) { w_raise(R, kErrMMRank); T->rank[c] = kNoCall; continue; }     // src/mod.c:1116
        const SelProbe p0 = w_select_probe<C0>(S, flex, bd, pat, rev ? cnt_cls - 1u - r0 : r0);
        if (need_bm) atomicOr(&bm[r0 >> 5], 1u << (r0 & 31u));
        T->rank[c] = w_select_resolve<C0>(p0, pat);
    }
}

__device__ __noinline__ void w_tile_calls_fast_other(const DecodeParams &P, uint32_t aoff, uint32_t slot, uint32_t n, uint32_t cidx0,
                                                     uint32_t ml_base, uint32_t lane) {
    const WArena A = w_arena(aoff);
    const WBlock *bd = &A.R->blk[slot];
    const bool ex = P.insertions || P.haplotypes || A.R->st.cshift != 0u;
    if (bd->cls == 0u) {
        if (ex) w_tile_calls_fast<true, true>(P, A.R, A.T, A.flex, A.s_lut, bd, n, cidx0, ml_base, lane);
        else w_tile_calls_fast<true, false>(P, A.R, A.T, A.flex, A.s_lut, bd, n, cidx0, ml_base, lane);
    } else w_tile_calls_fast<false, true>(P, A.R, A.T, A.flex, A.s_lut, bd, n, cidx0, ml_base, lane);
}

__device__ __forceinline__ void w_tile_calls(const DecodeParams &P, uint32_t aoff, WRead *R, WTile *T, uint32_t *flex, const uint8_t *s_lut, uint32_t slot, uint32_t jb,
                                             uint32_t n, uint32_t cidx0, uint32_t ml_base, uint32_t lane) {
    const WState &S = R->st;
    const WBlock *bd = &R->blk[slot];                              // jb: the block's ordinal in the read (view row order)
    if (w_fast_ok(P, S, bd)) {
        const bool ex = P.insertions || P.haplotypes || S.cshift != 0u;
        if (ex || bd->cls == 0u) w_tile_calls_fast_other(P, aoff, slot, n, cidx0, ml_base, lane);   // own function: own registers
        else w_tile_calls_fast<false, false>(P, R, T, flex, s_lut, bd, n, cidx0, ml_base, lane);
        return;
    }
    const uint32_t cls = bd->cls, need_bm = bd->dot;
    const uint32_t pat = class_pat(cls), rd_code = cls >= 1u && cls <= 3u ? cls : 4u;
    // ---- select
    if (bd->is_n) {                                               // src/mod.c:1102-1107
        const uint32_t L = S.L, rev = S.rev;
        uint32_t *bm = flex + bd->o_bm;
        for (uint32_t c = lane; c < n; c += 32u) {
            const uint32_t rank = T->rank[c];
            if (rank >= L) { w_raise(R, kErrMMRank); T->rank[c] = kNoCall; continue; }
            if (need_bm) atomicOr(&bm[rank >> 5], 1u << (rank & 31u));
            T->rank[c] = rev ? L - 1u - rank : rank;
        }
    } else if (cls == 0u) w_pass_select<true>(R, T, flex, bd, pat, n, need_bm, lane);
    else w_pass_select<false>(R, T, flex, bd, pat, n, need_bm, lane);
    // ---- map + update
    for (uint32_t c = lane; c < n; c += 32u) {
        const uint32_t q = T->rank[c];
        if (q != kNoCall) w_call(P, R, flex, s_lut, bd, jb, q, false, cidx0 + c, ml_base, rd_code);
    }
}

// implicit calls of a '.' block (src/mod.c:1203-1367): every base of the class whose rank is not
// in the explicit-rank bitmap.  Needs bd->n_calls / last1 / ml_base.
__device__ __forceinline__ void w_implicit_block(const DecodeParams &P, WRead *R, const uint32_t *flex, const uint8_t *s_lut, uint32_t jb, uint32_t lane) {
    const WState &S = R->st;
    const WBlock *bd = &R->blk[jb];
    const uint32_t *bm = flex + bd->o_bm, *idx = flex + bd->o_idx;
    const uint32_t cnt_cls = bd->cnt_cls, ml_base = bd->ml_base;
    if (bd->is_n) {
        const uint32_t last1 = bd->n_calls > 0 ? bd->last1 : 0u;                    // last + 1
        const uint32_t bound = last1 > cnt_cls ? last1 : cnt_cls;                   // Q8
        for (uint32_t s = lane; s < bound; s += 32u) {
            if ((bm[s >> 5] >> (s & 31u)) & 1u) continue;
            w_call(P, R, flex, s_lut, bd, jb, S.rev ? S.L - 1u - s : s, true, s, ml_base, 4u);
        }
        return;
    }
    const uint32_t cls = bd->cls, pat = class_pat(cls), rd_code = cls >= 1u && cls <= 3u ? cls : 4u;
    for (uint32_t e = lane; e < S.n_ent; e += 32u) {
        uint32_t fr = idx[e];
        uint32_t u0 = e << S.ishift, u1 = u0 + (1u << S.ishift);
        if (u1 > S.n_u4) u1 = S.n_u4;
        for (uint32_t u = u0; u < u1; ++u) {
            const uint4 v = ld16(S.seq + (size_t)u * 16u);
            for (uint32_t j = 0; j < 4u; ++j) {
                const uint32_t b0 = u * 32u + 8u * j;
                const uint32_t wj = j == 0u ? v.x : j == 1u ? v.y : j == 2u ? v.z : v.w;
                const uint32_t f = valid_flags(cls == 0u ? class_flags<true>(wj, pat) : class_flags<false>(wj, pat), S.L > b0 ? S.L - b0 : 0u);
                for (uint32_t t = 0; t < 8u; ++t) {                                 // base order: byte by byte, high nibble first
                    if (!((f >> (8u * (t >> 1) + ((t & 1u) ? 3u : 7u))) & 1u)) continue;
                    const uint32_t s = S.rev ? cnt_cls - 1u - fr : fr;
                    ++fr;
                    if ((bm[s >> 5] >> (s & 31u)) & 1u) continue;
                    w_call(P, R, flex, s_lut, bd, jb, b0 + t, true, s, ml_base, rd_code);
                }
            }
        }
    }
}

// call LUTs of the first kWLutSlots -c entries -> shared memory of the CTA (all threads call it)
__device__ __forceinline__ void w_stage_luts(const DecodeParams &P, uint8_t *s_lut) {
    const uint32_t n_lut = P.n_req < kWLutSlots ? (uint32_t)P.n_req : (uint32_t)kWLutSlots;
    for (uint32_t i = threadIdx.x; i < n_lut * 256u; i += blockDim.x) s_lut[i] = P.req[i >> 8].lut[i & 255u];
    __syncthreads();
}

// fused path: the non-inlined pieces of the block loop
__device__ __noinline__ uint32_t w_fused_tile(const DecodeParams &P, uint32_t aoff, uint32_t jb, uint32_t tb,
                                              uint32_t carry_cnt, uint32_t ml_base, uint32_t lane) {
    const WArena A = w_arena(aoff);
    WRead *R = A.R;
    WState &S = R->st;
    const WBlock *bd = &R->blk[jb];
    uint32_t sum = 0;
    const uint32_t n = w_tile_ranks(R, A.T, tb, bd->hdr_end, bd->end, S.carry_sum, &sum, lane);
    if (bd->any_req && n) w_tile_calls(P, aoff, R, A.T, A.flex, A.s_lut, jb, jb, n, carry_cnt, ml_base, lane);
    __syncwarp();
    if (lane == 0) S.carry_sum = sat_add(S.carry_sum, sum);
    __syncwarp();
    return n;
}
__device__ __noinline__ void w_fused_implicit(const DecodeParams &P, uint32_t aoff, uint32_t jb, uint32_t lane) {
    const WArena A = w_arena(aoff);
    w_implicit_block(P, A.R, A.flex, A.s_lut, jb, lane);
}

// MINB = resident CTAs per SM the register allocation is bounded for (2: 128 regs, 3: 80, 4: 64);
// the host picks one (and the matching arena size) per context.
// PRE  = the reads were prepared by k_flat_setup in "split" mode (mmc_decode_flat.cuh): this kernel then
//        only copies WRead + dir|cq|cr into its arena instead of running w_setup_read, which keeps the
//        header parser and the CIGAR scan out of its instruction footprint.
struct PreParams { const WRead *reads; uint32_t n; };

template <int MINB, bool PRE>
__global__ void __launch_bounds__(kWThreads, MINB) k_decode_warp(const __grid_constant__ DecodeParams P, const __grid_constant__ WarpParams W,
                                                                 const __grid_constant__ PreParams Q) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t aoff = kWHeadBytes + warp * W.arena_bytes;
    const WArena A = w_arena(aoff);
    w_stage_luts(P, A.s_lut);
    w_sparse_open(A.T, lane);
    WRead *R = A.R;
    uint32_t *flex = A.flex;
    const uint32_t flex_words = (W.arena_bytes - (uint32_t)sizeof(WFixed)) / 4u;
    WState &S = R->st;
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) {
            r = atomicAdd(P.work_counter, 1u);
            if (!PRE && P.read_list) r = r < *P.read_list_n ? P.read_list[r] : 0xffffffffu;   // reads deferred by an earlier kernel
        }
        r = __shfl_sync(kFull, r, 0);
        __syncwarp();
        if (PRE) {
            if (r >= Q.n) break;
            const WRead *G = &Q.reads[r];
            const uint32_t nb = G->st.n_blocks;
            if (nb == 0u) continue;                               // deferred or fatal in k_flat_setup
            const uint32_t *src = reinterpret_cast<const uint32_t *>(G);
            uint32_t *dst = reinterpret_cast<uint32_t *>(R);
            const uint32_t words = (uint32_t)((sizeof(WState) + sizeof(uint32_t) * (kWBlocks + 4) + sizeof(WBlock) * nb) / 4);
            for (uint32_t i = lane; i < words; i += 32u) dst[i] = src[i];
            __syncwarp();
            const uint4 *s4 = reinterpret_cast<const uint4 *>(S.flex_home);
            uint4 *d4 = reinterpret_cast<uint4 *>(flex);
            for (uint32_t i = lane; i < ((S.n_stage + 3u) >> 2); i += 32u) d4[i] = s4[i];
            __syncwarp();
            if (lane == 0) { S.flex = flex; S.flex_home = flex; S.cur_cls = 0xffu; S.cur_blk = 0; }
            __syncwarp();
        } else {
            if (r >= P.n_reads) break;
            if (!w_setup_read<true>(P, aoff, flex_words, flex_words, W.defer_list, W.defer_n, r, lane)) continue;
        }

        // ---- blocks in order
        const uint32_t n_blocks = S.n_blocks;
        uint32_t ml_base = 0, err = 0;
        for (uint32_t jb = 0; jb < n_blocks; ++jb) {
            WBlock *bd = &R->blk[jb];
            const uint32_t a0 = bd->hdr_end, a1 = bd->end, cls = bd->cls;      // cls is 4 when the canonical base is N
            const bool need_bm = w_needs_bitmap(bd);
            if (w_needs_index(bd)) {
                if (S.cur_cls != cls) {
                    w_build_index(aoff, jb, lane);
                    if (lane == 0) { S.cur_cls = cls; S.cur_blk = jb; }
                } else if (lane == 0) { bd->cnt_cls = R->blk[S.cur_blk].cnt_cls; bd->rshift = R->blk[S.cur_blk].rshift; }
            }
            if (need_bm) {
                const uint32_t words = ((S.L + 31u) >> 5) + 1u;
                for (uint32_t w = lane; w < words; w += 32u) flex[bd->o_bm + w] = 0;
            }
            if (lane == 0) S.carry_sum = 0;
            __syncwarp();
            uint32_t carry_cnt = 0;
            for (uint32_t tb = a0 & ~15u; tb < a1; tb += (uint32_t)kWChunks * 16u)
                carry_cnt += w_fused_tile(P, aoff, jb, tb, carry_cnt, ml_base, lane);
            err = w_err(R);
            if (err) break;
            if (lane == 0) { bd->n_calls = carry_cnt; bd->last1 = S.carry_sum; bd->ml_base = ml_base; }
            __syncwarp();
            if (need_bm) {
                w_fused_implicit(P, aoff, jb, lane);
                err = w_err(R);
                if (err) break;
            }
            if (carry_cnt > 0) ml_base += carry_cnt * bd->K;                    // src/mod.c:1200
        }
        if (err) w_report(P, S.r, err, lane);
    }
    w_sparse_close(P, w_arena(kWHeadBytes + (threadIdx.x >> 5) * W.arena_bytes).T, lane);   // (recomputed: nothing kept live across the read loop)
}

}  // namespace mmc

#endif  // MMC_DECODE_WARP_CUH
