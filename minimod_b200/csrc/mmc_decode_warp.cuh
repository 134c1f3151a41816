// mmc_decode_warp.cuh -- k_decode_warp: the fast path of the decode+aggregate stage (sm_100a).
//
// One WARP per read, reads handed out by an atomic counter; no CTA barrier anywhere, only
// shuffles and __syncwarp.  It computes exactly what k_decode (mmc_device.cuh) computes --
// freq_view_single() + get_aln() + update_freq_map() of the reference (src/mod.c:776-1370) --
// with per-read working sets small enough for a slice of shared memory ("arena") per warp:
//
//   CIGAR      sampled prefix sums (every 2^cshift-th op) + a directory over 2^gshift-base
//              buckets of the read, so aln[q]/ins[q] (get_aln, src/mod.c:776-881) is two shared
//              loads, a 0-2 step search and a walk over <= 2^cshift raw CIGAR words.
//   base ranks one u32 prefix count per 32<<ishift bases of the block's base class instead of
//              the bases_pos[][] tables (src/mod.c:977-981); rank -> position is an
//              interpolated probe into that index + an in-word select on the 4-bit SEQ.
//   MM text    496-byte tiles: SWAR comma masks -> token starts -> one parse per skip count
//              -> warp scan of (skip+1) gives every call its rank (src/mod.c:1098).
//   context    is_context[][] (src/ref.c:204-219) for ACGT contexts of <= 8 bases is tested
//              on a 16-base window funnel-shifted out of the 2-bit reference.
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
constexpr int kWValCap    = 256;             // >= kWChunks*16/2 tokens per tile
constexpr int kWMaxCShift = 5;
constexpr int kWMaxIShift = 3;
constexpr uint32_t kFull  = 0xffffffffu;

struct WBlock {                              // 32 bytes, see BlockDesc
    uint32_t hdr_end, end;
    uint8_t  cls, is_n, dot, K;
    int8_t   req[kMaxCodes];
    uint8_t  outc[kMaxCodes];
    uint32_t any_req;
};

struct WFixed {                              // fixed part of a warp's arena
    uint32_t err;
    uint32_t pad0[3];
    uint32_t semi[kWBlocks + 4];
    WBlock   blk[kWBlocks];
    uint4    text[32];                       // the tile's text, chunk i at text[i]
    uint32_t val[kWValCap];                  // skip+1 of the tile's tokens, in order
};

struct WarpParams {
    uint32_t arena_bytes;                    // per warp, multiple of 16, >= sizeof(WFixed) + 256
    uint32_t *defer_list;                    // reads left to k_decode
    uint32_t *defer_n;
};

struct WRead {                               // warp-uniform per-read state (registers)
    uint32_t r, L, n_cig, mm_len, ml_len, rev, hp;
    int32_t  tid, pos;
    const uint32_t *cig;
    const uint8_t *seq, *mm, *ml;
    ContigDev cd;
    uint32_t *cq, *cr, *dir, *idx, *bm;      // arena
    uint32_t cshift, gshift, ishift, n_samp, n_ent, n_u4, total_q;
    uint32_t cnt_cls;
    float    scale;                          // n_ent / cnt_cls
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
// 4-bit SEQ: class tests on 8 bases at a time.  A BAM byte holds base 2i in its high nibble.
// flags: bit 4k+3 set <=> nibble k of `u` belongs to class `cls` (src/mod.c:97: A0 C1 G2 T3 N4,
// every other nt16 code counts as A).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nib_eq_flags(uint32_t u, uint32_t pat) {
    uint32_t y = u ^ pat;
    uint32_t t = (y & 0x77777777u) + 0x77777777u;
    return ~(t | y) & 0x88888888u;
}
__device__ __forceinline__ uint32_t class_flags(uint32_t u, uint32_t cls) {
    if (cls == 0u) {
        uint32_t f = nib_eq_flags(u, 0x22222222u) | nib_eq_flags(u, 0x44444444u) |
                     nib_eq_flags(u, 0x88888888u) | nib_eq_flags(u, 0xffffffffu);
        return ~f & 0x88888888u;
    }
    const uint32_t pat = cls == 1u ? 0x22222222u : cls == 2u ? 0x44444444u : cls == 3u ? 0x88888888u : 0xffffffffu;
    return nib_eq_flags(u, pat);
}
// nibble k of the result is base k of the word (undo BAM's high-nibble-first packing)
__device__ __forceinline__ uint32_t base_order(uint32_t x) {
    return ((x & 0x0f0f0f0fu) << 4) | ((x >> 4) & 0x0f0f0f0fu);
}
__device__ __forceinline__ uint32_t count_u4(uint4 v, uint32_t cls) {
    return (uint32_t)(__popc(class_flags(v.x, cls)) + __popc(class_flags(v.y, cls)) +
                      __popc(class_flags(v.z, cls)) + __popc(class_flags(v.w, cls)));
}
// flags of the first `nv` (0..8) bases of a base-ordered word
__device__ __forceinline__ uint32_t valid_flags(uint32_t f, uint32_t nv) {
    return nv >= 8u ? f : (f & ((1u << (4u * nv)) - 1u));
}
__device__ __forceinline__ uint32_t count_u4_tail(uint4 v, uint32_t cls, uint32_t nv /* 1..31 valid bases */) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w}, c = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t n = nv > 8u * j ? nv - 8u * j : 0u;
        c += (uint32_t)__popc(valid_flags(class_flags(base_order(w[j]), cls), n));
    }
    return c;
}
__device__ __forceinline__ uint4 ld16(const uint8_t *p) { return *reinterpret_cast<const uint4 *>(p); }

// ---------------------------------------------------------------------------------------
// errors: first one raised in the warp wins (as in k_decode)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void w_raise(WFixed *wf, uint32_t code) { atomicCAS(&wf->err, 0u, code); }
__device__ __forceinline__ uint32_t w_err(WFixed *wf) {
    __syncwarp();
    uint32_t e = *reinterpret_cast<volatile uint32_t *>(&wf->err);
    __syncwarp();
    return e;
}

// ---------------------------------------------------------------------------------------
// aln[q], ins[q], ins_offset[q] of get_aln() in BAM orientation (src/mod.c:776-881)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ AlnHit w_cigar_lookup(const WRead &R, uint32_t q) {
    AlnHit h; h.aln = -1; h.ins = -1; h.insoff = 0;
    if (q >= R.total_q) return h;
    const uint32_t b = q >> R.gshift;
    uint32_t lo = R.dir[b];
    uint32_t hi = (((b + 1u) << R.gshift) < R.total_q) ? R.dir[b + 1u] : R.n_samp - 1u;
    while (lo < hi) {                                            // largest sample with cq <= q
        uint32_t mid = (lo + hi + 1u) >> 1;
        if (R.cq[mid] <= q) lo = mid; else hi = mid - 1u;
    }
    uint32_t qa = R.cq[lo], ra = R.cr[lo], o = lo << R.cshift;
    for (; o < R.n_cig; ++o) {
        const uint32_t w = R.cig[o], op = w & 15u, len = w >> 4;
        const bool aln = op == 0u || op == 7u || op == 8u;
        const uint32_t ql = (aln || op == 1u || op == 4u) ? len : 0u;
        const uint32_t d = q - qa;
        if (d < ql) {
            if (aln) h.aln = R.pos + (int32_t)(ra + d);
            else if (op == 1u) { h.ins = R.pos + (int32_t)ra - 1; h.insoff = d + 1u; }
            return h;
        }
        qa += ql;
        if (aln || op == 2u || op == 3u) ra += len;
    }
    return h;
}

// ---------------------------------------------------------------------------------------
// bases_pos[cls][k] (src/mod.c:977-981): BAM position of the k-th base of the indexed class
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t w_select(const WRead &R, uint32_t cls, uint32_t k) {
    const uint32_t *idx = R.idx;
    uint32_t e = (uint32_t)((float)k * R.scale);
    if (e >= R.n_ent) e = R.n_ent - 1u;
    uint32_t lo, hi;                                             // idx[lo] <= k < idx[hi]
    if (idx[e] <= k) {
        lo = e; hi = e + 1u;
        uint32_t step = 1u;
        while (hi < R.n_ent && idx[hi] <= k) { lo = hi; hi = hi + step < R.n_ent ? hi + step : R.n_ent; step <<= 1; }
    } else {
        hi = e; lo = e - 1u;
        uint32_t step = 1u;
        while (idx[lo] > k) { hi = lo; lo = lo >= step ? lo - step : 0u; step <<= 1; }
    }
    while (hi - lo > 1u) {
        uint32_t mid = (lo + hi) >> 1;
        if (idx[mid] <= k) lo = mid; else hi = mid;
    }
    uint32_t rem = k - idx[lo];
    uint32_t u = lo << R.ishift;
    uint4 v;
    for (;;) {                                                   // at most 2^ishift vectors
        v = ld16(R.seq + (size_t)u * 16u);
        if (R.ishift == 0u) break;
        uint32_t c = count_u4(v, cls);
        if (rem < c) break;
        rem -= c; ++u;
    }
    uint32_t f0 = class_flags(base_order(v.x), cls), f1 = class_flags(base_order(v.y), cls);
    uint32_t f2 = class_flags(base_order(v.z), cls), f3 = class_flags(base_order(v.w), cls);
    uint32_t c0 = (uint32_t)__popc(f0), c1 = (uint32_t)__popc(f1), c2 = (uint32_t)__popc(f2);
    uint32_t f = f0, wsel = 0;
    if (rem >= c0) { rem -= c0; f = f1; wsel = 1; if (rem >= c1) { rem -= c1; f = f2; wsel = 2; if (rem >= c2) { rem -= c2; f = f3; wsel = 3; } } }
    uint32_t sh = 0, c;
    c = (uint32_t)__popc(f & 0xffffu);        if (rem >= c) { rem -= c; sh = 16; }
    c = (uint32_t)__popc((f >> sh) & 0xffu);  if (rem >= c) { rem -= c; sh += 8; }
    c = (f >> (sh + 3u)) & 1u;                if (rem >= c) { sh += 4; }
    return u * 32u + wsel * 8u + (sh >> 2);
}

// ---------------------------------------------------------------------------------------
// context + base test against the packed reference (src/mod.c:1162-1172, src/ref.c:204-219)
// ---------------------------------------------------------------------------------------
// 1 = in context (then refcode is the 2-bit reference base at pos, never an exception letter),
// 0 = not in context, -1 = use the generic test (contig edge)
__device__ __forceinline__ int32_t ctx_fast(const ContigDev &cd, uint32_t pos, uint32_t m, uint32_t pat2, uint32_t &refcode) {
    if (pos + 1u < m || (unsigned long long)pos + m > cd.len) return -1;
    const uint32_t w0 = pos + 1u - m;                             // first base of the window
    const uint32_t wi = w0 >> 4, ei = w0 >> 5;
    const uint32_t lo = cd.ref2[wi], hi = cd.ref2[wi + 1u];
    const uint32_t elo = cd.excm[ei], ehi = cd.excm[ei + 1u];
    const uint32_t W = __funnelshift_r(lo, hi, (w0 & 15u) * 2u);  // 16 bases from w0
    const uint32_t E = __funnelshift_r(elo, ehi, w0 & 31u);       // their exception bits
    const uint32_t m2 = m >= 16u ? 0xffffffffu : ((1u << (2u * m)) - 1u), m1 = (1u << m) - 1u;
    bool hit = false;
    for (uint32_t j = 0; j < m; ++j)                              // occurrence starting at w0 + j
        hit = hit || ((((W >> (2u * j)) & m2) == pat2) && (((E >> j) & m1) == 0u));
    refcode = (W >> (2u * (m - 1u))) & 3u;
    return hit ? 1 : 0;
}

__device__ __forceinline__ void w_add_cell(const DecodeParams &P, const WRead &R, int32_t ref_pos, uint32_t outc,
                                           uint32_t ins16, int32_t hap, uint32_t is_mod) {
    int32_t hslot = -1;
    if (!P.haplotypes || hap < 0) hslot = 0;
    else if (hap + 1 < P.n_hap_slots) hslot = hap + 1;
    if (ins16 == 0u && outc < (uint32_t)P.n_code_slots && hslot >= 0) {
        unsigned long long i = (((unsigned long long)(uint32_t)ref_pos * 2ull + R.rev) * (unsigned)P.n_code_slots + outc) * (unsigned)P.n_hap_slots + (unsigned)hslot;
        atomicAdd(&R.cd.cells[i], 1ull | ((unsigned long long)is_mod << 32));
    } else {
        unsigned long long slot = atomicAdd(P.sparse_n, 1ull);
        if (slot < P.sparse_cap) {
            SparseRec s;
            s.a = ((unsigned long long)(uint32_t)R.tid << 41) | ((unsigned long long)(uint32_t)ref_pos << 9) | ((unsigned long long)R.rev << 8) | outc;
            s.b = ins16 | ((hap < 0 ? 256u : (uint32_t)hap) << 16);
            s.w = 1u | (is_mod << 16);
            P.sparse[slot] = s;
        }
    }
}

// Everything after "base q of the read is a call" (SURVEY.md A.4-A.8); see process_call().
__device__ __forceinline__ void w_call(const DecodeParams &P, const WRead &R, WFixed *wf, const WBlock *bdp, uint32_t blk_ord,
                                       uint32_t q, bool implicit, uint32_t cidx, uint32_t ml_base) {
    const WBlock &bd = *bdp;                                      // stays in shared memory (dynamic indexing)
    AlnHit h = w_cigar_lookup(R, q);
    int32_t ref_pos = h.aln;
    if (ref_pos < 0 && P.insertions) {
        if (implicit && R.rev) ref_pos = w_cigar_lookup(R, R.L - 1u - q).ins;   // Q9 (src/mod.c:1234,1314)
        else ref_pos = h.ins;
    }
    if (ref_pos < 0) return;                                      // src/mod.c:1127,1237,1317
    const uint32_t ins_off = P.insertions ? h.insoff : 0u;

    int32_t nib = -1, ref_match = -1;
    for (uint32_t m = 0; m < bd.K; ++m) {
        const int32_t ri = bd.req[m];
        if (ri < 0) continue;                                     // src/mod.c:1157
        const ReqMod &rq = P.req[ri];
        if (!P.insertions && rq.ctx_len > 0) {                    // src/mod.c:1162-1172
            uint32_t refcode = 0;
            int32_t in = rq.fast_ctx ? ctx_fast(R.cd, (uint32_t)ref_pos, (uint32_t)rq.ctx_len, R.rev ? rq.pat2_rc : rq.pat2, refcode) : -1;
            if (in < 0) {
                if (!in_context(R.cd, (uint32_t)ref_pos, R.rev ? rq.pat_rc : rq.pat, rq.ctx_len)) continue;
            } else if (in == 0) continue;
            if (!bd.is_n) {
                if (ref_match < 0) {
                    if (nib < 0) nib = (int32_t)((R.seq[q >> 1] >> ((~q & 1u) << 2)) & 0xfu);
                    if (in > 0) ref_match = (uint32_t)nib == (1u << refcode) ? 1 : 0;
                    else ref_match = ref_letter(R.cd, (uint32_t)ref_pos) == nt16_letter((uint32_t)nib) ? 1 : 0;
                }
                if (!ref_match) continue;
            }
        }
        uint32_t prob = 0, is_mod = 0;
        if (!implicit) {
            unsigned long long ml_idx = (unsigned long long)ml_base + (unsigned long long)cidx * bd.K + m;
            if (ml_idx >= R.ml_len) { w_raise(wf, kErrMLIndex); return; }   // src/mod.c:1174
            prob = R.ml[ml_idx];
        }
        if (P.subtool == 1) {                                     // FREQ
            if (!implicit) {
                uint32_t f = rq.lut[prob];                        // src/mod.c:1181-1191
                if (!(f & 1u)) continue;
                is_mod = (f >> 1) & 1u;
            }
            const uint32_t ins16 = ins_off & 0xffffu;
            if (P.haplotypes) w_add_cell(P, R, ref_pos, bd.outc[m], ins16, (int32_t)R.hp, is_mod);
            w_add_cell(P, R, ref_pos, bd.outc[m], ins16, -1, is_mod);     // src/mod.c:906-928
        } else {                                                  // VIEW
            unsigned long long slot = atomicAdd(P.view_n, 1ull);
            if (slot < P.view_cap) {
                ViewDev v;
                v.read = R.r; v.ref_pos = ref_pos;
                v.read_pos = (int32_t)(R.rev ? R.L - 1u - q : q);
                v.ins_off = ins_off;
                v.order = ((unsigned long long)blk_ord << 40) | ((unsigned long long)(implicit ? 1u : 0u) << 39) |
                          ((unsigned long long)cidx << 8) | m;
                v.code = bd.outc[m]; v.prob = (uint8_t)prob;
                for (int z = 0; z < 6; ++z) v.pad[z] = 0;
                P.view[slot] = v;
            }
        }
    }
}

// commas of 16 text bytes as a 16-bit mask (bit i <-> byte i)
__device__ __forceinline__ uint32_t byte_mask16(uint4 v, uint32_t c) {
    const uint32_t cc = c * 0x01010101u, M = 0x00204081u;
    uint32_t a = (zero_bytes(v.x ^ cc) * M) >> 28, b = (zero_bytes(v.y ^ cc) * M) >> 28;
    uint32_t d = (zero_bytes(v.z ^ cc) * M) >> 28, e = (zero_bytes(v.w ^ cc) * M) >> 28;
    return a | (b << 4) | (d << 8) | (e << 12);
}

__device__ __forceinline__ void w_defer(const WarpParams &W, uint32_t r, uint32_t lane) {
    if (lane == 0) W.defer_list[atomicAdd(W.defer_n, 1u)] = r;
}
__device__ __forceinline__ void w_report(const DecodeParams &P, uint32_t r, uint32_t code, uint32_t lane) {
    if (lane == 0) atomicMin(P.err, ((unsigned long long)r << 32) | code);
}

// parse one MM block header (src/mod.c:1003-1062); same rules as k_decode's (2b).  Returns an error code.
__device__ __forceinline__ uint32_t w_parse_header(const DecodeParams &P, const WRead &R, uint32_t start, uint32_t end, WBlock &bd) {
    bd.end = end; bd.hdr_end = end; bd.K = 0; bd.any_req = 0; bd.cls = 0; bd.is_n = 0; bd.dot = 1;
    for (int k = 0; k < kMaxCodes; ++k) { bd.req[k] = -1; bd.outc[k] = 0; }
    uint32_t i = start;
    const uint32_t base_c = i < end ? R.mm[i] : 0u;
    const bool okb = base_c == 'A' || base_c == 'C' || base_c == 'G' || base_c == 'T' || base_c == 'U' || base_c == 'N' ||
                     base_c == 'a' || base_c == 'c' || base_c == 'g' || base_c == 't' || base_c == 'u' || base_c == 'n';
    if (!okb) return kErrMMBase;
    ++i;
    const uint32_t modbase = base_c == 'U' ? (uint32_t)'T' : base_c;                 // src/mod.c:1006
    const uint32_t strand_c = i < end ? R.mm[i] : 0u;
    if (strand_c != '+' && strand_c != '-') return kErrMMStrand;
    ++i;
    uint8_t codes[kMaxCodes];
    uint32_t j = 0; bool has_num = false, has_alpha = false, bad = false, too_many = false;
    while (i < end) {
        const uint32_t c = R.mm[i];
        if (c == ',' || c == '?' || c == '.') break;
        if (c >= '0' && c <= '9') has_num = true;
        else if ((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z')) has_alpha = true;
        else { bad = true; break; }
        if (j < (uint32_t)kMaxCodes) codes[j] = (uint8_t)c; else too_many = true;
        ++j; ++i;
    }
    if (bad || j == 0 || (has_num && has_alpha)) return kErrMMCode;
    if (too_many) return kErrTooManyCodes;
    const uint32_t K = has_num ? 1u : j;                                             // src/mod.c:1048
    if (i < end && (R.mm[i] == '?' || R.mm[i] == '.')) { bd.dot = R.mm[i] == '.'; ++i; }
    bd.hdr_end = i;
    bd.K = (uint8_t)K;
    uint32_t mb = modbase;                                                           // src/mod.c:1092-1093, table :98
    if (R.rev) {
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
        unsigned long long key = 0;                                                  // code m = suffix string (Q4)
        const uint32_t from = has_num ? 0u : m;
        for (uint32_t z = from; z < j; ++z) key |= (unsigned long long)codes[z] << (8u * (z - from));
        if (P.wild_req >= 0) {
            int32_t id = code_id(P.code_keys, key, true);
            if (id < 0) return kErrTooManyCodes;
            bd.req[m] = (int8_t)P.wild_req; bd.outc[m] = (uint8_t)id; bd.any_req = 1;
        } else {
            for (int32_t q = 0; q < P.n_req; ++q)
                if (P.req[q].key == key) { bd.req[m] = (int8_t)q; bd.outc[m] = (uint8_t)q; bd.any_req = 1; break; }
        }
    }
    return kErrNone;
}

__device__ __forceinline__ void w_process_read(const DecodeParams &P, const WarpParams &W, WFixed *wf, uint32_t *flex,
                                               uint32_t flex_words, uint32_t r, uint32_t lane) {
    WRead R;
    R.r = r;
    R.tid = P.tid[r]; R.pos = P.pos[r];
    R.L = P.l_seq[r]; R.n_cig = P.n_cigar[r];
    R.mm_len = P.mm_len[r]; R.ml_len = P.ml_len[r];
    R.rev = (P.flag[r] >> 4) & 1u;
    R.hp = P.hp[r];
    R.cig = P.cigar + P.cigar_off[r];
    R.seq = P.seq4 + P.seq_off[r];
    R.mm = P.mm + P.mm_off[r];
    R.ml = P.ml + P.ml_off[r];
    if (lane == 0) wf->err = 0;
    __syncwarp();
    if (R.tid < 0 || R.tid >= P.n_contigs || P.contigs[R.tid].ref2 == nullptr) { w_report(P, r, kErrNoContig, lane); return; }
    if (R.L >= (1u << 28)) { w_report(P, r, kErrSeqTooLong, lane); return; }
    R.cd = P.contigs[R.tid];

    // ---- MM blocks: positions of ';'
    uint32_t n_semi = 0;
    for (uint32_t base = 0; base < R.mm_len; base += 512u) {
        const uint32_t p0 = base + lane * 16u;
        uint32_t mask = 0;
        if (p0 < R.mm_len) {
            mask = byte_mask16(ld16(R.mm + p0), ';');
            const uint32_t valid = R.mm_len - p0;
            if (valid < 16u) mask &= (1u << valid) - 1u;
        }
        const uint32_t c = (uint32_t)__popc(mask);
        const uint32_t incl = warp_incl_scan(c, lane);
        uint32_t ord = n_semi + incl - c;
        while (mask) {
            const uint32_t bit = (uint32_t)__ffs((int)mask) - 1u;
            mask &= mask - 1u;
            if (ord < (uint32_t)kWBlocks) wf->semi[ord] = p0 + bit;
            ++ord;
        }
        n_semi += __shfl_sync(kFull, incl, 31);
    }
    uint32_t n_blocks = n_semi;
    if (R.mm_len > 0 && R.mm[R.mm_len - 1u] != ';') n_blocks += 1;                   // unterminated last block
    if (n_blocks > (uint32_t)kWBlocks) { w_defer(W, r, lane); return; }
    __syncwarp();

    // ---- block headers, one lane each
    uint32_t herr = kErrNone, my_dot_work = 0;
    if (lane < n_blocks) {
        WBlock bd;
        const uint32_t start = lane == 0 ? 0u : wf->semi[lane - 1u] + 1u;
        const uint32_t end = lane < n_semi ? wf->semi[lane] : R.mm_len;
        herr = w_parse_header(P, R, start, end, bd);
        my_dot_work = herr == kErrNone && bd.any_req && bd.dot;
        wf->blk[lane] = bd;
    }
    const uint32_t herr_mask = __ballot_sync(kFull, herr != kErrNone);
    if (herr_mask) herr = __shfl_sync(kFull, herr, __ffs((int)herr_mask) - 1);
    const bool need_bm_any = __ballot_sync(kFull, my_dot_work != 0u) != 0u;

    // ---- arena layout
    R.n_u4 = (R.L + 31u) >> 5;
    R.gshift = 8;
    while (((R.L >> R.gshift) + 2u) > 160u) ++R.gshift;
    const uint32_t n_dir = (R.L >> R.gshift) + 2u;
    const uint32_t bm_words = need_bm_any ? ((R.L + 31u) >> 5) + 1u : 0u;
    R.cshift = 0; R.ishift = 0;
    for (;;) {
        R.n_samp = (R.n_cig + (1u << R.cshift) - 1u) >> R.cshift;
        if (R.n_samp == 0u) R.n_samp = 1u;
        R.n_ent = (R.n_u4 + (1u << R.ishift) - 1u) >> R.ishift;
        const uint32_t need = n_dir + bm_words + 2u * R.n_samp + R.n_ent + 2u;
        if (need <= flex_words) break;
        if (R.cshift >= (uint32_t)kWMaxCShift && R.ishift >= (uint32_t)kWMaxIShift) { w_defer(W, r, lane); return; }
        if ((2u * R.n_samp >= R.n_ent && R.cshift < (uint32_t)kWMaxCShift) || R.ishift >= (uint32_t)kWMaxIShift) ++R.cshift;
        else ++R.ishift;
    }
    R.dir = flex;
    R.cq = R.dir + n_dir;
    R.cr = R.cq + R.n_samp;
    R.idx = R.cr + R.n_samp;
    R.bm = R.idx + R.n_ent + 2u;

    // ---- CIGAR prefix sums == get_aln() (src/mod.c:811-880)
    uint32_t carry_q = 0, carry_r = 0;
    const uint32_t cmask = (1u << R.cshift) - 1u;
    if (R.n_cig == 0u && lane == 0) { R.cq[0] = 0; R.cr[0] = 0; }
    for (uint32_t base = 0; base < R.n_cig; base += 32u) {
        const uint32_t i = base + lane;
        uint32_t op = 15u, len = 0, ql = 0, rl = 0;
        if (i < R.n_cig) {
            const uint32_t w = R.cig[i];
            op = w & 15u; len = w >> 4;
            if (op == 0u || op == 7u || op == 8u) { ql = len; rl = len; }
            else if (op == 1u || op == 4u) ql = len;
            else if (op == 2u || op == 3u) rl = len;
            else if (op == 5u) w_raise(wf, kErrHardClip);
            else w_raise(wf, kErrCigarOp);
        }
        const uint32_t iq = warp_incl_scan_sat(ql, lane), ir = warp_incl_scan_sat(rl, lane);
        if (i < R.n_cig) {
            const uint32_t q0 = sat_add(carry_q, iq >= kSat ? kSat : iq - ql), r0 = sat_add(carry_r, ir >= kSat ? kSat : ir - rl);
            const bool alnop = op == 0u || op == 7u || op == 8u;
            if ((alnop || (op == 1u && P.insertions)) && len > 0 && sat_add(q0, ql) > R.L) w_raise(wf, kErrCigarLen);
            if (alnop && len > 0) {
                long long last = (long long)R.pos + r0 + len - 1;
                if (R.pos < 0 || last >= (long long)R.cd.len) w_raise(wf, kErrRefRange);
            }
            if ((i & cmask) == 0u) { R.cq[i >> R.cshift] = q0; R.cr[i >> R.cshift] = r0; }
            if (ql > 0u && q0 < R.L) {                            // directory: sample that holds each bucket's first base
                const uint32_t qe = sat_add(q0, ql) < R.L ? q0 + ql : R.L;
                for (uint32_t b = (q0 + (1u << R.gshift) - 1u) >> R.gshift; (b << R.gshift) < qe; ++b) R.dir[b] = i >> R.cshift;
            }
        }
        carry_q = sat_add(carry_q, __shfl_sync(kFull, iq, 31));
        carry_r = sat_add(carry_r, __shfl_sync(kFull, ir, 31));
    }
    R.total_q = carry_q < R.L ? carry_q : R.L;
    uint32_t err = w_err(wf);
    if (!err) err = herr;
    if (err) { w_report(P, r, err, lane); return; }
    if (lane == 0) {
        const int32_t lo = R.pos > 0 ? R.pos - 1 : 0;
        long long hi = (long long)R.pos + carry_r + 1;
        if (hi > (long long)R.cd.len) hi = R.cd.len;
        if (lo < *reinterpret_cast<volatile int32_t *>(&P.touch_lo[R.tid])) atomicMin(&P.touch_lo[R.tid], lo);
        if ((int32_t)hi > *reinterpret_cast<volatile int32_t *>(&P.touch_hi[R.tid])) atomicMax(&P.touch_hi[R.tid], (int32_t)hi);
    }

    // ---- blocks in order
    uint32_t cur_cls = 0xffu, ml_base = 0;
    R.cnt_cls = 0; R.scale = 0.f;
    for (uint32_t jb = 0; jb < n_blocks; ++jb) {
        const WBlock *bdp = &wf->blk[jb];
        WBlock bd;                                                // scalar fields only; req[]/outc[] are read through bdp
        bd.hdr_end = bdp->hdr_end; bd.end = bdp->end; bd.cls = bdp->cls; bd.is_n = bdp->is_n; bd.dot = bdp->dot; bd.K = bdp->K; bd.any_req = bdp->any_req;
        const uint32_t a0 = bd.hdr_end, a1 = bd.end;
        const bool work = bd.any_req != 0u;
        const bool need_bm = work && bd.dot;

        uint32_t cnt_cls = 0;
        if (work && (!bd.is_n || bd.dot)) {
            const uint32_t cls = bd.cls;                          // 4 when the canonical base is N
            if (cur_cls != cls) {
                __syncwarp();
                uint32_t carry = 0;
                const uint32_t tail = R.L & 31u;
                for (uint32_t e0 = 0; e0 < R.n_ent; e0 += 32u) {
                    const uint32_t e = e0 + lane;
                    uint32_t c = 0;
                    if (e < R.n_ent) {
                        uint32_t u0 = e << R.ishift, u1 = u0 + (1u << R.ishift);
                        if (u1 > R.n_u4) u1 = R.n_u4;
                        for (uint32_t u = u0; u < u1; ++u) {
                            const uint4 v = ld16(R.seq + (size_t)u * 16u);
                            c += (u == R.n_u4 - 1u && tail) ? count_u4_tail(v, cls, tail) : count_u4(v, cls);
                        }
                    }
                    const uint32_t incl = warp_incl_scan(c, lane);
                    if (e < R.n_ent) R.idx[e] = carry + incl - c;
                    carry += __shfl_sync(kFull, incl, 31);
                }
                if (lane == 0) R.idx[R.n_ent] = carry;
                cur_cls = cls;
                R.cnt_cls = carry;
                R.scale = carry ? (float)R.n_ent / (float)carry : 0.f;
                __syncwarp();
            }
            cnt_cls = R.cnt_cls;
        }
        if (need_bm) {
            const uint32_t words = ((R.L + 31u) >> 5) + 1u;
            for (uint32_t w = lane; w < words; w += 32u) R.bm[w] = 0;
            __syncwarp();
        }

        // ---- skip counts -> ranks -> calls, one text tile (31 chunks of 16 bytes + 1 look-ahead chunk) at a time
        uint32_t carry_cnt = 0, carry_sum = 0, prev_last = 0;
        for (uint32_t tb = a0 & ~15u; tb < a1; tb += (uint32_t)kWChunks * 16u) {
            const uint32_t p0 = tb + lane * 16u;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (p0 < R.mm_len) v = ld16(R.mm + p0);
            wf->text[lane] = v;
            const uint32_t cm = byte_mask16(v, ',');
            const uint32_t up = __shfl_up_sync(kFull, cm, 1);
            const uint32_t ncm = __shfl_down_sync(kFull, cm, 1);
            const uint32_t pbit = lane == 0 ? prev_last : (up >> 15) & 1u;
            prev_last = (__shfl_sync(kFull, cm, kWChunks - 1) >> 15) & 1u;
            uint32_t st = ((cm << 1) | pbit) & ~cm & 0xffffu;      // token starts: previous byte is ','
            // clip to [a0, a1) and force a start at a0 (src/mod.c:1066: the list begins right after the header)
            uint32_t lo_b = a0 > p0 ? a0 - p0 : 0u, hi_b = a1 > p0 ? a1 - p0 : 0u;
            if (lo_b > 16u) lo_b = 16u;
            if (hi_b > 16u) hi_b = 16u;
            if (a0 >= p0 && a0 < p0 + 16u && !((cm >> lo_b) & 1u)) st |= 1u << lo_b;
            st &= ~((1u << lo_b) - 1u);
            st &= (1u << hi_b) - 1u;
            if (lane >= (uint32_t)kWChunks) st = 0;
            uint32_t em = cm | (ncm << 16);                       // token terminators: ',' or the block end
            if (a1 >= p0 && a1 - p0 < 32u) em |= 1u << (a1 - p0);
            const uint32_t n_tok = (uint32_t)__popc(st);
            const uint32_t incl = warp_incl_scan(n_tok, lane);
            const uint32_t tile_cnt = __shfl_sync(kFull, incl, 31);
            uint32_t slot = incl - n_tok;
            __syncwarp();
            const uint8_t *tx = reinterpret_cast<const uint8_t *>(wf->text) + lane * 16u;
            while (st) {
                const uint32_t b = (uint32_t)__ffs((int)st) - 1u;
                st &= st - 1u;
                const uint32_t rest = em >> (b + 1u);
                const uint32_t nd = rest ? (uint32_t)__ffs((int)rest) : 32u;
                uint32_t val = 0;
                if (nd > 9u) w_raise(wf, kErrMMSkip);              // src/mod.c:1080-1085
                else {
                    bool bad = false;
                    for (uint32_t k = 0; k < nd; ++k) {
                        const uint32_t d = tx[b + k];
                        if (d < '0' || d > '9') bad = true;
                        val = val * 10u + (d - '0');
                    }
                    if (bad) { w_raise(wf, kErrMMSkip); val = 0; }
                }
                wf->val[slot++] = val + 1u;
            }
            __syncwarp();
            for (uint32_t c0 = 0; c0 < tile_cnt; c0 += 32u) {
                const uint32_t c = c0 + lane;
                const uint32_t x = c < tile_cnt ? wf->val[c] : 0u;
                const uint32_t si = warp_incl_scan_sat(x, lane);
                if (work && c < tile_cnt) {
                    const uint32_t rank = sat_add(carry_sum, si) - 1u;          // base_rank (src/mod.c:1098)
                    bool ok = true;
                    uint32_t q = 0;
                    if (bd.is_n) {                                              // src/mod.c:1102-1107
                        if (rank >= R.L) { w_raise(wf, kErrMMRank); ok = false; }
                        else q = R.rev ? R.L - 1u - rank : rank;
                    } else {                                                    // src/mod.c:1109-1113
                        if (rank >= cnt_cls) { w_raise(wf, kErrMMRank); ok = false; }
                        else q = w_select(R, bd.cls, R.rev ? cnt_cls - 1u - rank : rank);
                    }
                    if (ok) {
                        if (need_bm) atomicOr(&R.bm[rank >> 5], 1u << (rank & 31u));
                        w_call(P, R, wf, bdp, jb, q, false, carry_cnt + c, ml_base);
                    }
                }
                carry_sum = sat_add(carry_sum, __shfl_sync(kFull, si, 31));
            }
            carry_cnt += tile_cnt;
            __syncwarp();
        }
        err = w_err(wf);
        if (err) break;

        // ---- implicit calls of a '.' block (src/mod.c:1203-1367)
        if (need_bm) {
            if (bd.is_n) {
                const uint32_t last1 = carry_cnt > 0 ? carry_sum : 0u;          // last + 1
                const uint32_t bound = last1 > cnt_cls ? last1 : cnt_cls;       // Q8
                for (uint32_t s = lane; s < bound; s += 32u) {
                    if ((R.bm[s >> 5] >> (s & 31u)) & 1u) continue;
                    const uint32_t q = R.rev ? R.L - 1u - s : s;
                    w_call(P, R, wf, bdp, jb, q, true, s, ml_base);
                }
            } else {
                for (uint32_t e = lane; e < R.n_ent; e += 32u) {
                    uint32_t fr = R.idx[e];
                    uint32_t u0 = e << R.ishift, u1 = u0 + (1u << R.ishift);
                    if (u1 > R.n_u4) u1 = R.n_u4;
                    for (uint32_t u = u0; u < u1; ++u) {
                        const uint4 v = ld16(R.seq + (size_t)u * 16u);
                        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t b0 = u * 32u + 8u * j;
                            uint32_t f = valid_flags(class_flags(base_order(w4[j]), bd.cls), R.L > b0 ? R.L - b0 : 0u);
                            while (f) {
                                const uint32_t bit = (uint32_t)__ffs((int)f) - 1u;
                                f &= f - 1u;
                                const uint32_t s = R.rev ? cnt_cls - 1u - fr : fr;
                                ++fr;
                                if ((R.bm[s >> 5] >> (s & 31u)) & 1u) continue;
                                w_call(P, R, wf, bdp, jb, b0 + (bit >> 2), true, s, ml_base);
                            }
                        }
                    }
                }
            }
            err = w_err(wf);
            if (err) break;
        }
        if (carry_cnt > 0) ml_base += carry_cnt * bd.K;                         // src/mod.c:1200
    }
    if (err) w_report(P, r, err, lane);
}

// MINB = resident CTAs per SM the register allocation is bounded for (2: 128 regs, 3: 80, 4: 64);
// the host picks one (and the matching arena size) per context.
template <int MINB>
__global__ void __launch_bounds__(kWThreads, MINB) k_decode_warp(DecodeParams P, WarpParams W) {
    MMC_DYN_SMEM(uint4, w_dyn);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint8_t *arena = reinterpret_cast<uint8_t *>(w_dyn) + (size_t)warp * W.arena_bytes;
    WFixed *wf = reinterpret_cast<WFixed *>(arena);
    uint32_t *flex = reinterpret_cast<uint32_t *>(arena + sizeof(WFixed));
    const uint32_t flex_words = (W.arena_bytes - (uint32_t)sizeof(WFixed)) / 4u;
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(P.work_counter, 1u);
        r = __shfl_sync(kFull, r, 0);
        if (r >= P.n_reads) break;
        w_process_read(P, W, wf, flex, flex_words, r, lane);
        __syncwarp();
    }
}

}  // namespace mmc

#endif  // MMC_DECODE_WARP_CUH
