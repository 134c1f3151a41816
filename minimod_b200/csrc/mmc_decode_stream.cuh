// mmc_decode_stream.cuh -- k_decode_stream: the decode+aggregate stage as a streaming merge (sm_100a).
//
// One WARP per read (reads prepared by k_flat_setup, handed out by an atomic counter), but -- unlike
// k_decode_warp -- no per-read rank index, rank directory or explicit-rank bitmap: the skip counts of an
// MM block give strictly increasing base ranks (src/mod.c:1098), so the calls of a block can be MERGED
// against the read's SEQ, streamed once in rank order (from the 5' end of the original read: BAM order
// for forward reads, reversed for reverse reads, src/mod.c:1109-1113):
//
//   count step   32 SEQ vectors (1024 bases) per step: class flags + popcount per lane, one warp scan;
//                the vectors and their running class counts go into a 128-slot ring in shared memory.
//   round        32 calls: a 7-step binary search of the ring for the vector that holds the call's rank,
//                in-vector select on the staged copy (no second trip to L2), then the same
//                map -> context -> threshold -> red.global.add.u64 tail as k_decode_warp.
//   '.' blocks   explicit calls set a bit per base in the ring; a vector that leaves the ring emits its
//                remaining class bases as implicit calls (src/mod.c:1203-1367) -- no L-bit bitmap.
//
// Shared memory per warp is constant (ring 3 KB + text tile 1.7 KB + read state 1.2 KB) plus the read's CIGAR
// prefix arrays, so 28-32 warps are resident per SM whatever the read length (k_decode_warp: 24 on 15 kb reads,
// 8 on 50 kb reads), and the per-read index build (23 % of k_decode_warp's instructions) is gone.
//
// The SEQ ring is fed by cp.async.bulk + mbarrier (s_prefetch): the 32 vectors of the next count step are in flight while
// the current round of calls is decoded.  The per-read record (WRead, ~0.4 KB) is copied by the lanes; the CIGAR table stays
// where k_flat_setup wrote it and is looked up in place.
#ifndef MMC_DECODE_STREAM_CUH
#define MMC_DECODE_STREAM_CUH

#include "mmc_decode_warp.cuh"

namespace mmc {

constexpr int kSThreads = 128;               // 4 warps per CTA: fine-grained residency
constexpr uint32_t kSRing = 128u, kSMask = kSRing - 1u;
constexpr uint32_t kSWindow = 96u;           // searchable positions behind p_hi (the other 32 slots receive the next chunk)

struct SRing {
    uint4    vec[kSRing];                    // SEQ vector of stream position p at slot p & kSMask
    uint32_t pre[kSRing];                    // class bases of the read before that vector, in stream order
    uint32_t exp[kSRing];                    // '.' blocks: bit t <-> base t of the vector is an explicit call
};
// warp-uniform state of the SEQ stream of one MM block
struct SStream {
    uint32_t p_hi;                           // vectors appended so far (stream positions [0, p_hi))
    uint32_t cum;                            // class bases in them
    uint32_t p_cov;                          // lower bound of the vector that holds the first unprocessed call
    uint32_t p_ret;                          // '.' blocks: vectors before it have emitted their implicit calls
    uint32_t mode, pat;                      // 0: nibble == pat, 1: class A (everything but C,G,T,N), 2: every base (canonical base N)
    uint32_t n_u4, tail_n, rev, dot;
    uint32_t pf_on, pf_ph;                   // a bulk copy of the vectors [p_hi, p_hi + 32) is in flight / landed; parity of the warp's mbarrier
};

struct SFixed { WRead rd; WTile tl; SRing rg; SStream zs; unsigned long long bar; };
static_assert(sizeof(SFixed) % 16 == 0 && offsetof(SFixed, rg) % 16 == 0, "arena pieces are 16-byte aligned");
struct SArena { uint8_t *s_lut; WRead *R; WTile *T; SRing *G; SStream *Z; unsigned long long *bar; };
__device__ __forceinline__ SArena s_arena(uint32_t aoff) {
    MMC_DYN_SMEM(uint4, w_dyn);
    uint8_t *base = reinterpret_cast<uint8_t *>(w_dyn);
    SArena A;
    A.s_lut = base;
    SFixed *sf = reinterpret_cast<SFixed *>(base + aoff);
    A.R = &sf->rd; A.T = &sf->tl; A.G = &sf->rg; A.Z = &sf->zs; A.bar = &sf->bar;
    return A;
}

struct StreamParams {
    uint32_t arena_bytes;                    // per warp, multiple of 16, >= sizeof(SFixed) + 256
    uint32_t head_bytes;                     // call LUTs in front of the arenas
    uint32_t split;                          // work unit = (read, even / odd blocks) instead of a read: half as many reads in flight
    uint32_t pad_;
};

__device__ __forceinline__ uint32_t s_flags(uint32_t u, uint32_t mode, uint32_t pat) {
    return mode == 0u ? nib_eq_flags(u, pat) : mode == 1u ? class_flags<true>(u, 0u) : 0x88888888u;
}
struct SVecFlags { uint32_t f0, f1, f2, f3; };
// class flags of the vector at BAM vector index u (nv = its valid bases)
__device__ __forceinline__ SVecFlags s_vec_flags(uint4 v, uint32_t mode, uint32_t pat, uint32_t nv) {
    SVecFlags F;
    F.f0 = s_flags(v.x, mode, pat); F.f1 = s_flags(v.y, mode, pat); F.f2 = s_flags(v.z, mode, pat); F.f3 = s_flags(v.w, mode, pat);
    if (nv < 32u) {
        F.f0 = valid_flags(F.f0, nv); F.f1 = valid_flags(F.f1, nv > 8u ? nv - 8u : 0u);
        F.f2 = valid_flags(F.f2, nv > 16u ? nv - 16u : 0u); F.f3 = valid_flags(F.f3, nv > 24u ? nv - 24u : 0u);
    }
    return F;
}
// nibble flags (bit 7 = first base of a byte, bit 3 = second) of 8 bases -> 8-bit mask, bit t <-> base t
__device__ __forceinline__ uint32_t s_mask8(uint32_t f) {
    const uint32_t y = ((f >> 7) & 0x01010101u) | ((f >> 2) & 0x02020202u);
    return (y * 0x01041040u) >> 24;
}

// The SEQ stream is fed by bulk asynchronous copies (cp.async.bulk -> UBLKCP, completing on the warp's mbarrier): the 32
// vectors the next count step will look at are already on their way into the ring while the current round of calls is
// searched, selected and mapped.  Reverse reads stream from the end of SEQ: a chunk lands in memory order, so position p
// sits at slot (p & kSMask) ^ 31 (pre / exp stay at p & kSMask).
__device__ __forceinline__ uint32_t s_vslot(const SStream &Z, uint32_t p) { return (p & kSMask) ^ (Z.rev ? 31u : 0u); }
__device__ __forceinline__ void s_prefetch(const uint8_t *seq, SRing *G, unsigned long long *bar, SStream &Z, uint32_t lane) {
    if (Z.p_hi >= Z.n_u4) return;
    const uint32_t p0 = Z.p_hi, cnt = Z.n_u4 - p0 < 32u ? Z.n_u4 - p0 : 32u;
    __syncwarp();                                                  // every lane is done with the slots about to be overwritten
    if (lane == 0) {
        const uint32_t u_lo = Z.rev ? Z.n_u4 - p0 - cnt : p0;
        mmc_bulk_g2s(&G->vec[(p0 & kSMask) + (Z.rev ? 32u - cnt : 0u)], seq + (size_t)u_lo * 16u, cnt * 16u, bar);
    }
    Z.pf_on = 1u;
}
__device__ __forceinline__ void s_prefetch_wait(unsigned long long *bar, SStream &Z) {
    if (Z.pf_on) { mmc_mbar_wait(bar, Z.pf_ph); Z.pf_ph ^= 1u; Z.pf_on = 0u; }
}

// start of block jb: reset the stream (a copy a previous block left in flight is drained first) and start fetching its head
__device__ __noinline__ void s_open_block(uint32_t aoff, uint32_t jb, uint32_t lane) {
    const SArena A = s_arena(aoff);
    WState &S = A.R->st;
    const WBlock *bd = &A.R->blk[jb];
    SStream Z = *A.Z;
    s_prefetch_wait(A.bar, Z);
    Z.p_hi = 0; Z.cum = 0; Z.p_cov = 0; Z.p_ret = 0;
    Z.mode = bd->is_n ? 2u : bd->cls == 0u ? 1u : 0u;
    Z.pat = class_pat(bd->cls);
    Z.n_u4 = S.n_u4; Z.tail_n = (S.L & 31u) ? (S.L & 31u) : 32u; Z.rev = S.rev;
    Z.dot = (bd->dot && bd->any_req) ? 1u : 0u;
    if (bd->any_req) s_prefetch(S.seq, A.G, A.bar, Z, lane);
    __syncwarp();
    if (lane == 0) { *A.Z = Z; S.carry_sum = 0; }
    __syncwarp();
}

// implicit calls of the vectors at stream positions [p_from, p_to) of a '.' block: every class base that no
// explicit call selected (src/mod.c:1203-1367); rank s = bases of the class before it in stream order.
__device__ __noinline__ void s_retire(const DecodeParams &P, uint32_t aoff, uint32_t jb, uint32_t p_from, uint32_t p_to, uint32_t bound, uint32_t lane) {
    const SArena A = s_arena(aoff);
    WRead *R = A.R;
    const WState &S = R->st;
    const WBlock *bd = &R->blk[jb];
    const SRing *G = A.G;
    const uint32_t mode = bd->is_n ? 2u : bd->cls == 0u ? 1u : 0u, pat = class_pat(bd->cls);
    const uint32_t cls = bd->cls, rd_code = (!bd->is_n && cls >= 1u && cls <= 3u) ? cls : 4u;
    const uint32_t n_u4 = S.n_u4, rev = S.rev, tail_n = (S.L & 31u) ? (S.L & 31u) : 32u;
    __syncwarp();                                                  // the explicit bits other lanes set are visible
    for (uint32_t p = p_from + lane; p < p_to; p += 32u) {
        const uint32_t sl = p & kSMask, u = rev ? n_u4 - 1u - p : p;
        const SVecFlags F = s_vec_flags(G->vec[sl ^ (rev ? 31u : 0u)], mode, pat, u == n_u4 - 1u ? tail_n : 32u);
        const uint32_t m32 = s_mask8(F.f0) | (s_mask8(F.f1) << 8) | (s_mask8(F.f2) << 16) | (s_mask8(F.f3) << 24);
        uint32_t imp = m32 & ~G->exp[sl];
        const uint32_t pre = G->pre[sl], tot = (uint32_t)__popc(m32);
        while (imp) {
            const uint32_t t = (uint32_t)__ffs((int)imp) - 1u;
            imp &= imp - 1u;
            const uint32_t before = (uint32_t)__popc(m32 & ((1u << t) - 1u));
            const uint32_t s = pre + (rev ? tot - 1u - before : before);
            if (s >= bound) continue;
            w_call(P, R, S.flex, A.s_lut, bd, jb, u * 32u + t, true, s, 0u, rd_code);
        }
    }
    __syncwarp();
}

// number of literal 'N' letters of the read: the tail bound of the implicit loop of an N+x. block (Q8, src/mod.c:1290)
__device__ __noinline__ uint32_t s_count_n(uint32_t aoff, uint32_t lane) {
    const WState &S = s_arena(aoff).R->st;
    const uint32_t n_u4 = S.n_u4, tail = S.L & 31u;
    uint32_t c = 0;
    for (uint32_t u = lane; u < n_u4; u += 32u) {
        const uint4 v = ld16(S.seq + (size_t)u * 16u);
        c += (u == n_u4 - 1u && tail) ? count_u4_tail<false>(v, 0xffffffffu, tail) : count_u4<false>(v, 0xffffffffu);
    }
    return __shfl_sync(kFull, warp_incl_scan(c, lane), 31);
}

// one count step: the 32 vectors the last prefetch brought in are counted and join the searchable window; the next 32 are
// requested.  The copy overwrites the vec slots of positions [p_hi - 96, p_hi - 64) (new p_hi): callers keep the window
// they search inside the last 96 positions, and '.' blocks retire what falls out first.
__device__ __forceinline__ void s_append(const DecodeParams &P, uint32_t aoff, uint32_t jb, const uint8_t *seq, SRing *G, unsigned long long *bar,
                                         SStream &Z, uint32_t lane) {
    s_prefetch_wait(bar, Z);
    const uint32_t p = Z.p_hi + lane;
    uint32_t c = 0;
    if (p < Z.n_u4) {
        const uint32_t u = Z.rev ? Z.n_u4 - 1u - p : p;
        const uint4 v = G->vec[s_vslot(Z, p)];
        const SVecFlags F = s_vec_flags(v, Z.mode, Z.pat, u == Z.n_u4 - 1u ? Z.tail_n : 32u);
        c = (uint32_t)(__popc(F.f0) + __popc(F.f1) + __popc(F.f2) + __popc(F.f3));
    }
    const uint32_t incl = warp_incl_scan(c, lane);
    if (p < Z.n_u4) {
        G->pre[p & kSMask] = Z.cum + incl - c;
        if (Z.dot) G->exp[p & kSMask] = 0u;
    }
    Z.cum += __shfl_sync(kFull, incl, 31);
    Z.p_hi = Z.p_hi + 32u < Z.n_u4 ? Z.p_hi + 32u : Z.n_u4;
    __syncwarp();
    if (Z.dot && Z.p_hi + 32u > Z.p_ret + kSRing) {
        const uint32_t to = Z.p_hi + 32u - kSRing;
        s_retire(P, aoff, jb, Z.p_ret, to, 0xffffffffu, lane);
        Z.p_ret = to;
    }
    s_prefetch(seq, G, bar, Z, lane);
}

// ---------------------------------------------------------------------------------------
// phase A of a text tile: base ranks T->rank[0..n) -> read positions (bases_pos[][], src/mod.c:1102-1113), in place.
// Rounds of up to 32 calls: the stream is advanced until the round's first call is covered (and, ring permitting, its
// last), every covered call finds its vector by binary search of the ring and selects inside the staged copy.
// A class-A read base that is not literally 'A' is marked in bit 31 (it can never equal the reference, src/mod.c:1164).
// ---------------------------------------------------------------------------------------
__device__ __noinline__ void s_select_tile(const DecodeParams &P, uint32_t aoff, uint32_t jb, uint32_t n, uint32_t lane) {
    const SArena A = s_arena(aoff);
    WRead *R = A.R;
    WTile *T = A.T;
    SRing *G = A.G;
    const uint8_t *seq = R->st.seq;
    SStream Z = *A.Z;
    uint32_t c0 = 0;
    while (c0 < n) {
        const uint32_t c = c0 + lane;
        const uint32_t k = c < n ? T->rank[c] : 0xffffffffu;       // base rank == stream rank (src/mod.c:1098,1109-1113)
        const uint32_t k0 = __shfl_sync(kFull, k, 0);
        while (Z.cum <= k0 && Z.p_hi < Z.n_u4) { const uint32_t at = Z.p_hi; s_append(P, aoff, jb, seq, G, A.bar, Z, lane); Z.p_cov = at; }
        if (Z.cum <= k0) {                                         // src/mod.c:1116: more skips than bases of the class
            w_raise(R, kErrMMRank);
            for (uint32_t x = c; x < n; x += 32u) T->rank[x] = kNoCall;
            break;
        }
        const uint32_t klast = __shfl_sync(kFull, k, (int)(n - c0 < 32u ? n - c0 - 1u : 31u));
        while (Z.cum <= klast && Z.p_hi < Z.n_u4 && Z.p_hi + 32u <= Z.p_cov + kSWindow) s_append(P, aoff, jb, seq, G, A.bar, Z, lane);
        const bool act = k < Z.cum;
        const uint32_t am = __ballot_sync(kFull, act);
        c0 += (uint32_t)__popc(am);
        uint32_t p = Z.p_cov;
        if (act) {                                                 // largest p in [p_cov, p_hi) with pre[p] <= k
#pragma unroll
            for (uint32_t st = 64u; st; st >>= 1) {                 // (window <= kSWindow = 96 < 128)
                const uint32_t t = p + st;
                if (t < Z.p_hi && G->pre[t & kSMask] <= k) p = t;
            }
        }
        Z.p_cov = __shfl_sync(kFull, p, 31 - __clz((int)am));      // the next round's first call lies at or after this vector
        if (!act) continue;
        const uint32_t sl = p & kSMask, u = Z.rev ? Z.n_u4 - 1u - p : p;
        const uint4 v = G->vec[s_vslot(Z, p)];
        uint32_t rem = k - G->pre[sl];
        const SVecFlags F = s_vec_flags(v, Z.mode, Z.pat, u == Z.n_u4 - 1u ? Z.tail_n : 32u);
        const uint32_t s0 = (uint32_t)__popc(F.f0), s1 = s0 + (uint32_t)__popc(F.f1), s2 = s1 + (uint32_t)__popc(F.f2);
        if (Z.rev) rem = s2 + (uint32_t)__popc(F.f3) - 1u - rem;
        uint32_t f = F.f0, wsel = 0, sub = 0, wv = v.x;
        if (rem >= s0) { f = F.f1; wsel = 8; sub = s0; wv = v.y; }
        if (rem >= s1) { f = F.f2; wsel = 16; sub = s1; wv = v.z; }
        if (rem >= s2) { f = F.f3; wsel = 24; sub = s2; wv = v.w; }
        rem -= sub;
        uint32_t cc = (uint32_t)__popc(f & 0xffffu);
        if (rem >= cc) { rem -= cc; f >>= 16; wsel += 4; }
        cc = (uint32_t)__popc(f & 0xffu);
        if (rem >= cc) { rem -= cc; f >>= 8; wsel += 2; }
        const uint32_t off = wsel + ((rem != 0u || !(f & 0x80u)) ? 1u : 0u);
        uint32_t q = u * 32u + off;
        if (Z.dot) atomicOr(&G->exp[sl], 1u << off);
        if (Z.mode == 1u) {
            const uint32_t o8 = off & 7u, nib = (wv >> (8u * (o8 >> 1) + ((o8 & 1u) ? 0u : 4u))) & 0xfu;
            if (nib != 1u) q |= 0x80000000u;
        }
        T->rank[c] = q;
    }
    __syncwarp();
    if (lane == 0) *A.Z = Z;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------
// phase B of a text tile, common case: read positions T->rank[0..n) -> reference positions -> context -> threshold ->
// dense cells.  `freq`, one requested code per block (K == 1), canonical base A/C/G/T (w_fast_ok).
//   C0  class A (bit 31 of the position: the read base is not literally 'A')
//   EX  any of: --insertions, --haplotypes, sampled CIGAR (kept out of the lean instantiation)
// ---------------------------------------------------------------------------------------
template <bool C0, bool EX>
__device__ __forceinline__ void s_tail_fast(const DecodeParams &P, WRead *R, WTile *T, const uint8_t *s_lut, const WBlock *bd,
                                            uint32_t n, uint32_t cidx0, uint32_t ml_base, uint32_t lane) {
    const WState &S = R->st;
    const uint32_t *flex = S.flex;                                 // dir | {cq, cr}: where k_flat_setup wrote them (HBM, read through L1)
    const uint32_t *dir = flex + S.o_dir;
    const uint2 *pr = reinterpret_cast<const uint2 *>(flex + S.o_cq);
    const uint32_t rev = S.rev, total_q = S.total_q, g = S.gshift, last_samp = S.n_samp - 1u, ml_len = S.ml_len, ref_len = S.ref_len;
    const int32_t pos = S.pos;
    const uint8_t *ml = S.ml;
    const uint32_t *ref2 = S.ref2, *excm = S.excm;
    unsigned long long *cells = S.cells;
    const WCode cd = bd->code[0];
    const uint32_t cls = bd->cls;
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
    const uint32_t nrec = EX && hslot ? 2u : 1u;
    const uint32_t ml0 = ml_base + cidx0;
    uint32_t sp_pos = 0, sp_meta = 0;                              // EX: a sparse record waiting for the next converged point
    bool pending = false;
    for (uint32_t c = lane; EX ? c - lane < n : c < n; c += 32u) {
        if (EX) {
            __syncwarp();
            w_sparse_flush(P, T, pending, nrec, (uint32_t)S.tid, rev, sp_pos, cd.outc, sp_meta & 0xffffu, S.hp, sp_meta >> 16, lane);
            pending = false;
            if (c >= n) continue;
        }
        const uint32_t qq = T->rank[c];
        if (qq == kNoCall) continue;
        const uint32_t q = qq & 0x7fffffffu;
        const uint32_t mi = ml0 + c;
        const uint32_t prob = mi < ml_len ? ldg8(ml + mi) : 0x100u;                    // issued early
        // ---- map: aln[q] (get_aln, src/mod.c:776-881)
        uint32_t ref_pos, ins16 = 0;
        if (EX && cshift != 0u) {                                                       // sampled CIGAR (long reads): generic lookup
            const AlnHit h = w_cigar_lookup(S, flex, q);
            if (h.aln >= 0) ref_pos = (uint32_t)h.aln;
            else if (insertions && h.ins >= 0) { ref_pos = (uint32_t)h.ins; ins16 = h.insoff & 0xffffu; }
            else continue;
        } else {
            if (q >= total_q) continue;
            const uint32_t b = q >> g;
            uint32_t lo = ldg32(dir + b), hi = (((b + 1u) << g) < total_q) ? ldg32(dir + b + 1u) : last_samp;
            const uint32_t qlim = (q + 1u) << 4;
            while (lo < hi) { const uint32_t mid = (lo + hi + 1u) >> 1; if (ldg32(&pr[mid].x) < qlim) lo = mid; else hi = mid - 1u; }   // 32-base buckets: a step or none
            const uint2 en = __ldg(&pr[lo]);
            const uint32_t ce = en.x, op = ce & 15u;
            if (op == 0u || op == 7u || op == 8u) ref_pos = (uint32_t)(pos + (int32_t)(en.y + q - (ce >> 4)));
            else if (EX && insertions && op == 1u) {                                    // ins[] / ins_offset (src/mod.c:1122-1127)
                const int32_t left = pos + (int32_t)en.y - 1;
                if (left < 0) continue;                                                 // Q11
                ref_pos = (uint32_t)left; ins16 = (q - (ce >> 4) + 1u) & 0xffffu;       // make_key's uint16_t (src/mod.c:428)
            } else continue;                                                            // src/mod.c:1127
        }
        // ---- context + base test (src/mod.c:1162-1172)
        if (m) {
            if (ref_pos + 1u < m || ref_pos + m > ref_len) {                            // contig edge: generic test
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
                if (!hit) continue;                                                     // a hit implies ref base == class base
                if (C0 && (qq >> 31)) continue;                                         // ... and the read base must be that letter
            }
        }
        if (prob > 0xffu) { w_raise(R, kErrMLIndex); continue; }                        // src/mod.c:1174
        const uint32_t fl = lut[prob];                                                  // src/mod.c:1181-1191
        if (!(fl & 1u)) continue;
        const unsigned long long inc = 1ull | ((unsigned long long)((fl >> 1) & 1u) << 32);
        if (!EX || ins16 == 0u) {
            unsigned long long *cell = cells + ((unsigned long long)ref_pos * per_pos + within);
            red_add_u64(cell, inc);                                                     // the '*' stratum (or the only one)
            if (EX && hslot) red_add_u64(cell + hslot, inc);                            // src/mod.c:906-928
        } else {
            pending = true; sp_pos = ref_pos; sp_meta = ins16 | (((fl >> 1) & 1u) << 16);
        }
    }
    if (EX) {
        __syncwarp();
        w_sparse_flush(P, T, pending, nrec, (uint32_t)S.tid, rev, sp_pos, cd.outc, sp_meta & 0xffffu, S.hp, sp_meta >> 16, lane);
    }
}

// one text tile of block jb: skip counts -> ranks -> read positions -> counts.  Returns the tile's token count.
//   tail  0 fast C/G/T   1 fast, class A   2 fast + --insertions / --haplotypes / sampled CIGAR   3 = 1 + 2   4 general (w_call)
__device__ __noinline__ uint32_t s_tile(const DecodeParams &P, uint32_t aoff, uint32_t jb, uint32_t tb, uint32_t carry_cnt, uint32_t ml_base,
                                        uint32_t lane, uint32_t tail) {
    const SArena A = s_arena(aoff);
    WRead *R = A.R;
    WTile *T = A.T;
    WState &S = R->st;
    const WBlock *bd = &R->blk[jb];
    uint32_t sum = 0;
    const uint32_t n = w_tile_ranks(R, T, tb, bd->hdr_end, bd->end, S.carry_sum, &sum, lane);
    if (bd->any_req && n) {
        s_select_tile(P, aoff, jb, n, lane);
        if (tail == 0u) s_tail_fast<false, false>(P, R, T, A.s_lut, bd, n, carry_cnt, ml_base, lane);
        else if (tail == 1u) s_tail_fast<true, false>(P, R, T, A.s_lut, bd, n, carry_cnt, ml_base, lane);
        else if (tail == 2u) s_tail_fast<false, true>(P, R, T, A.s_lut, bd, n, carry_cnt, ml_base, lane);
        else if (tail == 3u) s_tail_fast<true, true>(P, R, T, A.s_lut, bd, n, carry_cnt, ml_base, lane);
        else {                                                     // several codes per block, `view`, canonical base N, slow contexts
            const uint32_t cls = bd->cls, rd_code = (!bd->is_n && cls >= 1u && cls <= 3u) ? cls : 4u;
            for (uint32_t c = lane; c < n; c += 32u) {
                const uint32_t qq = T->rank[c];
                if (qq != kNoCall) w_call(P, R, S.flex, A.s_lut, bd, jb, qq & 0x7fffffffu, false, carry_cnt + c, ml_base, rd_code);
            }
        }
    }
    __syncwarp();
    if (lane == 0) S.carry_sum = sat_add(S.carry_sum, sum);
    __syncwarp();
    return n;
}

// end of a '.' block: the class bases no explicit call selected are implicit calls (src/mod.c:1203-1367)
__device__ __noinline__ void s_block_finish(const DecodeParams &P, uint32_t aoff, uint32_t jb, uint32_t n_calls, uint32_t lane) {
    const SArena A = s_arena(aoff);
    WRead *R = A.R;
    const WState &S = R->st;
    const WBlock *bd = &R->blk[jb];
    SStream Z = *A.Z;
    uint32_t bound = 0xffffffffu;
    if (bd->is_n) {                                                // Q8: [0,last) U (last, #N letters)
        const uint32_t cnt_n = s_count_n(aoff, lane), last1 = n_calls > 0u ? S.carry_sum : 0u;
        bound = last1 > cnt_n ? last1 : cnt_n;
    }
    while (Z.p_hi < Z.n_u4 && Z.cum < bound) s_append(P, aoff, jb, S.seq, A.G, A.bar, Z, lane);
    s_retire(P, aoff, jb, Z.p_ret, Z.p_hi, bound, lane);
    __syncwarp();
    if (lane == 0) *A.Z = Z;                                       // (a chunk may still be in flight: the next s_open_block drains it)
    __syncwarp();
}

// MINB = resident CTAs per SM the register allocation is bounded for (8: 64 registers).
template <int MINB>
__global__ void __launch_bounds__(kSThreads, MINB) k_decode_stream(const __grid_constant__ DecodeParams P, const __grid_constant__ StreamParams W,
                                                                   const __grid_constant__ PreParams Q) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t aoff = W.head_bytes + warp * W.arena_bytes;
    const SArena A = s_arena(aoff);
    {
        const uint32_t n_lut = P.n_req < kWLutSlots ? (uint32_t)P.n_req : (uint32_t)kWLutSlots;
        for (uint32_t i = threadIdx.x; i < n_lut * 256u; i += blockDim.x) A.s_lut[i] = P.req[i >> 8].lut[i & 255u];
        __syncthreads();
    }
    w_sparse_open(A.T, lane);
    if (lane == 0) { mmc_mbar_init(A.bar, 1u); A.Z->pf_on = 0u; A.Z->pf_ph = 0u; }
    __syncwarp();
    WRead *R = A.R;
    WState &S = R->st;
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(P.work_counter, 1u);
        r = __shfl_sync(kFull, r, 0);
        __syncwarp();
        // split mode: the blocks of a read stream SEQ independently (their first ML index comes from k_flat_setup), so two
        // warps take the even and the odd blocks of the same read at the same time.  Long reads with several strata per cell
        // (config 4: 50 kb, 128 bytes of cells per position) then keep half as many reads -- half the window of the count
        // arrays -- in flight for the same number of warps, which is what the L2 can hold.
        const uint32_t half = W.split ? (r & 1u) : 0u, stride = W.split ? 2u : 1u;
        if (W.split) r >>= 1;
        if (r >= Q.n) break;
        const WRead *G = &Q.reads[r];
        const uint32_t nb = G->st.n_blocks;
        if (nb <= half) continue;                                  // deferred or fatal in k_flat_setup (0 blocks), or no block for this half
        {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(G);
            uint32_t *dst = reinterpret_cast<uint32_t *>(R);
            const uint32_t words = (uint32_t)((sizeof(WState) + sizeof(uint32_t) * (kWBlocks + 4) + sizeof(WBlock) * nb) / 4);
            for (uint32_t i = lane; i < words; i += 32u) dst[i] = src[i];
            __syncwarp();
            __syncwarp();                                          // (the CIGAR table stays where k_flat_setup wrote it: HBM, read through L1)
        }
        // ---- blocks in order
        const uint32_t n_blocks = S.n_blocks;
        uint32_t ml_base = 0, err = 0;
        const bool ex = P.insertions || P.haplotypes || S.cshift != 0u;
        for (uint32_t jb = half; jb < n_blocks; jb += stride) {
            const WBlock *bd = &R->blk[jb];
            if (W.split) ml_base = bd->ml_base;
            const uint32_t a0 = bd->hdr_end, a1 = bd->end;
            const uint32_t tail = !w_fast_ok(P, S, bd) ? 4u : (bd->cls == 0u ? 1u : 0u) + (ex ? 2u : 0u);
            s_open_block(aoff, jb, lane);
            uint32_t carry_cnt = 0;
            for (uint32_t tb = a0 & ~15u; tb < a1; tb += (uint32_t)kWChunks * 16u) {
                carry_cnt += s_tile(P, aoff, jb, tb, carry_cnt, ml_base, lane, tail);
                if (*reinterpret_cast<volatile uint32_t *>(&S.err)) break;      // (re-read converged below)
            }
            err = w_err(R);
            if (err) break;
            if (bd->dot && bd->any_req) {
                s_block_finish(P, aoff, jb, carry_cnt, lane);
                err = w_err(R);
                if (err) break;
            }
            if (carry_cnt > 0u) ml_base += carry_cnt * bd->K;      // src/mod.c:1200
        }
        if (err) w_report(P, S.r, err, lane);
    }
    {                                                              // no bulk copy may be in flight when the CTA retires
        const SArena E = s_arena(W.head_bytes + (threadIdx.x >> 5) * W.arena_bytes);
        SStream Z = *E.Z;
        s_prefetch_wait(E.bar, Z);
        w_sparse_close(P, E.T, lane);
    }
}

}  // namespace mmc

#endif  // MMC_DECODE_STREAM_CUH
