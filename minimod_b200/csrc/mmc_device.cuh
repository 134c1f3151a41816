// mmc_device.cuh -- device-side data layout and kernels of libminimod_cuda.so (sm_100a).
//
// Kernel inventory (each replaces the reference CPU function named; there is no reference
// GPU code):
//   k_ref_pack        load_ref() upper-casing/U->T + the byte maps of load_ref_contexts()
//                     (src/ref.c:72-78,177-229): ASCII -> 2-bit bases + 1-bit exception mask.
//                     Contexts are evaluated on the fly from the packed reference, so no
//                     per-mod byte maps exist at all.
//   k_decode          freq_view_single() + get_aln() + update_freq_map()/add_view_entry()
//                     + merge_freq_maps() (src/mod.c:743-1370): one CTA per read.
//   k_count_nonzero / k_scan_tiles / k_emit_records
//                     the collect + (contig,pos) sort of print_freq_output()
//                     (src/mod.c:644-664): the dense arrays are already position ordered,
//                     so "sort" is a stream compaction.
//
// All arithmetic is integer; results are bit-exact by construction.
#ifndef MMC_DEVICE_CUH
#define MMC_DEVICE_CUH

#include <stdint.h>
#include "simt.h"

namespace mmc {

// ---------------------------------------------------------------------------------------
// compile-time geometry of k_decode's shared memory
// ---------------------------------------------------------------------------------------
constexpr int kMaxThreads   = 256;           // CTA size upper bound (multiple of 32)
constexpr int kChunk        = 16;            // MM text bytes parsed per thread per tile
constexpr int kMaxBlocks    = 32;            // MM blocks handled per round (more -> more rounds)
constexpr int kMaxCodes     = 8;             // modification codes in one MM block ("C+mh" = 2)
constexpr int kCigSmem      = 2048;          // CIGAR ops whose prefix sums live in shared memory
constexpr int kIdxSmem      = 1024;          // chunks of the base-rank index (64<<s bases each)
constexpr int kBitmapWords  = 2048;          // explicit-rank bitmap in shared memory (65536 ranks)
constexpr int kRankCap      = kMaxThreads * (kChunk / 2);   // tokens in one text tile
constexpr int kCodeTable    = 256;           // distinct output code strings
constexpr uint32_t kSat     = 0x7fffffffu;   // saturation cap of prefix sums

// per-read fatal conditions (the reference exit(1)s on all of them)
enum ReadError : uint32_t {
    kErrNone = 0,
    kErrHardClip,        // src/mod.c:841-844
    kErrCigarOp,         // src/mod.c:845-848
    kErrCigarLen,        // ASSERT read_pos < seq_len, src/mod.c:853,865
    kErrRefRange,        // ASSERT ref_pos in [0, ref_len), src/mod.c:860
    kErrNoContig,        // ASSERT ref != NULL, src/mod.c:793
    kErrMMBase,          // ASSERT valid_bases, src/mod.c:1005
    kErrMMStrand,        // ASSERT valid_strands, src/mod.c:1012
    kErrMMCode,          // src/mod.c:1030,1053,1054
    kErrMMSkip,          // src/mod.c:1080-1085 (non-numeric / too long / negative skip count)
    kErrMMRank,          // more skips than canonical bases, ASSERT read_pos, src/mod.c:1116
    kErrMLIndex,         // ASSERT ml_idx < ml_len, src/mod.c:1174
    kErrTooManyCodes,    // library limit: > kMaxCodes codes in a block / > kCodeTable codes
    kErrSeqTooLong,      // library limit: l_qseq >= 2^28
};

struct ReqMod {                       // one -c entry on the device (modcodem_t, src/minimod.h:60-64)
    unsigned long long key;           // code string packed little-endian into 8 bytes
    int32_t ctx_len;                  // 0 => context "*"
    uint8_t pat[32];                  // context, forward orientation (chars A,C,G,T,N)
    uint8_t pat_rc[32];               // reverse-complemented context (src/ref.c:183-194)
    uint8_t lut[256];                 // bit0 called, bit1 mod
    uint32_t fast_ctx;                // context is 1..8 bases of A/C/G/T: window test on the 2-bit reference
    uint32_t pat2, pat2_rc;           // context / its reverse complement, 2 bits per base, first base lowest
};

struct ContigDev {                    // per header contig (ref_t, src/ref.h:36-41, + its dense count array)
    const uint32_t *ref2;             // 2 bits per base, 16 per word; nullptr: contig not in the FASTA
    const uint32_t *excm;             // 1 bit per base: not A/C/G/T after upper-casing, U->T
    const uint32_t *exc_start;        // sorted run starts of exception letters (contig positions)
    const uint8_t  *exc_letter;       // upper-cased letter of each run
    unsigned long long *cells;        // [len][2 strands][n_code_slots][n_hap_slots] (n_called | n_mod<<32)
    uint32_t n_exc;
    uint32_t len;
};

struct SparseRec {                    // a count outside the dense arrays
    unsigned long long a;             // tid<<41 | pos<<9 | strand<<8 | code
    uint32_t b;                       // ins_offset(16) | hap9<<16   (hap9 256 == '*'/none)
    uint32_t w;                       // n_called | n_mod<<16
};

struct ViewDev {                      // one view row before host-side first-wins de-duplication
    uint32_t read;
    int32_t  ref_pos;
    int32_t  read_pos;
    uint32_t ins_off;
    unsigned long long order;         // processing order inside the read (block, phase, call, code)
    uint8_t  code, prob, pad[6];
};

struct DecodeParams {
    // ---- the batch (mmc_batch_t mirrored in HBM)
    uint32_t n_reads;
    const int32_t  *tid;
    const int32_t  *pos;
    const uint32_t *l_seq;
    const uint32_t *n_cigar;
    const uint32_t *mm_len;
    const uint32_t *ml_len;
    const unsigned long long *cigar_off;
    const unsigned long long *seq_off;
    const unsigned long long *mm_off;
    const unsigned long long *ml_off;
    const uint16_t *flag;
    const uint8_t  *hp;
    const uint32_t *cigar;
    const uint8_t  *seq4;
    const uint8_t  *mm;
    const uint8_t  *ml;
    // ---- options
    const ReqMod *req;
    int32_t n_req;
    int32_t wild_req;                 // index of the "*" code entry or -1 (src/mod.c:1146-1158)
    int32_t insertions;
    int32_t haplotypes;
    int32_t subtool;
    unsigned long long *code_keys;    // [kCodeTable] output code dictionary
    // ---- reference + dense counts: contigs[tid].cells[((pos*2+strand)*n_code_slots+code)*n_hap_slots+hslot]
    int32_t n_contigs;
    const ContigDev *contigs;
    int32_t n_code_slots;
    int32_t n_hap_slots;              // 1, or 1 + dense_haps with --haplotypes (slot 0 == '*')
    int32_t *touch_lo;                // per contig: lowest / highest+1 position any read covered
    int32_t *touch_hi;
    // ---- sparse side buffer
    SparseRec *sparse;
    unsigned long long sparse_cap;
    unsigned long long *sparse_n;
    // ---- view rows
    ViewDev *view;
    unsigned long long view_cap;
    unsigned long long *view_n;
    // ---- errors: min over (read<<32 | code), ~0 if none
    unsigned long long *err;
    // ---- per-CTA global scratch for reads that exceed the shared-memory capacities
    uint32_t *scratch;
    unsigned long long scratch_words_per_cta;
    uint32_t scratch_cig_words;       // words reserved for each of cq / cr
    uint32_t *work_counter;           // dynamic read scheduler
    const uint32_t *read_list;        // k_decode only: reads to process (nullptr: 0..n_reads-1)
    const uint32_t *read_list_n;      //   and how many (device memory, written by k_decode_warp)
    // test hooks: shrink the shared-memory capacities to force the scratch paths
    int32_t cig_smem_cap;
    int32_t bitmap_smem_words;
    int32_t idx_smem_cap;
};

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sat_add(uint32_t x, uint32_t y) {
    uint32_t s = x + y;               // x,y <= kSat so no wrap
    return s > kSat ? kSat : s;
}

// bit 7 of every byte of the result is set iff that byte of y is zero (exact, no carries)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t y) {
    uint32_t t = (y & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | y | 0x7f7f7f7fu);
}
// 4-bit mask: bit k set iff byte k of w equals c
__device__ __forceinline__ uint32_t byte_eq_mask(uint32_t w, uint32_t c) {
    uint32_t z = zero_bytes(w ^ (c * 0x01010101u)) >> 7;
    return (z | (z >> 7) | (z >> 14) | (z >> 21)) & 0xfu;
}

// Exclusive block scan of the pair (a,b), both saturating at kSat.  ws: >= 16 words of shared
// memory.  All threads of the CTA must call it.  Returns the CTA totals in ta,tb.
__device__ __forceinline__ void block_scan2_sat(uint32_t &a, uint32_t &b, uint32_t &ta, uint32_t &tb, uint32_t *ws) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31u) >> 5;
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t xa = __shfl_up_sync(0xffffffffu, ia, d);
        uint32_t xb = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= (uint32_t)d) { ia = sat_add(ia, xa); ib = sat_add(ib, xb); }
    }
    if (lane == 31) { ws[warp] = ia; ws[8 + warp] = ib; }
    __syncthreads();
    uint32_t pa = 0, pb = 0, sa = 0, sb = 0;
    for (uint32_t w = 0; w < nwarp; ++w) {
        uint32_t va = ws[w], vb = ws[8 + w];
        if (w < warp) { pa = sat_add(pa, va); pb = sat_add(pb, vb); }
        sa = sat_add(sa, va); sb = sat_add(sb, vb);
    }
    // inclusive-minus-own is the exclusive value unless the inclusive one already saturated
    a = sat_add(pa, ia >= kSat ? kSat : ia - a);
    b = sat_add(pb, ib >= kSat ? kSat : ib - b);
    ta = sa; tb = sb;
    __syncthreads();
}

// A CTA-uniform snapshot of the read's error flag (control flow around barriers must only
// ever branch on values every thread agrees on).
__device__ __forceinline__ uint32_t block_err(const uint32_t *err) {
    __syncthreads();
    uint32_t e = *reinterpret_cast<const volatile uint32_t *>(err);
    __syncthreads();
    return e;
}

__device__ __forceinline__ uint32_t nt16_letter(uint32_t nib) {     // seq_nt16_str = "=ACMGRSVTWYHKDBN"
    const unsigned long long L0 = ((unsigned long long)'=') | ((unsigned long long)'A' << 8) | ((unsigned long long)'C' << 16) |
                                  ((unsigned long long)'M' << 24) | ((unsigned long long)'G' << 32) | ((unsigned long long)'R' << 40) |
                                  ((unsigned long long)'S' << 48) | ((unsigned long long)'V' << 56);
    const unsigned long long L1 = ((unsigned long long)'T') | ((unsigned long long)'W' << 8) | ((unsigned long long)'Y' << 16) |
                                  ((unsigned long long)'H' << 24) | ((unsigned long long)'K' << 32) | ((unsigned long long)'D' << 40) |
                                  ((unsigned long long)'B' << 48) | ((unsigned long long)'N' << 56);
    return (uint32_t)(((nib < 8u ? L0 : L1) >> ((nib & 7u) * 8u)) & 0xffull);
}

// ---------------------------------------------------------------------------------------
// packed reference access
// ---------------------------------------------------------------------------------------
typedef ContigDev RefView;

// upper-cased (U->T) reference letter at contig position g  ==  ref->forward[pos] (src/ref.c:72-78)
__device__ __forceinline__ uint32_t ref_letter(const RefView &rv, uint32_t g) {
    uint32_t ex = (rv.excm[g >> 5] >> (uint32_t)(g & 31u)) & 1u;
    if (!ex) {
        uint32_t b = (rv.ref2[g >> 4] >> (uint32_t)((g & 15u) * 2u)) & 3u;
        return (0x54474341u >> (b * 8u)) & 0xffu;             // "ACGT"
    }
    // rare: last run start <= g
    uint32_t lo = 0, hi = rv.n_exc;                           // invariant: start[lo] <= g < start[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (rv.exc_start[mid] <= g) lo = mid; else hi = mid;
    }
    return rv.exc_letter[lo];
}

// is_context[..][pos] of the reference (src/ref.c:142-162,204-219): pos lies inside some
// occurrence of pat (length m) in the contig.
__device__ __forceinline__ bool in_context(const RefView &rv, uint32_t pos, const uint8_t *pat, int32_t m) {
    const uint32_t len = rv.len;
    if ((uint32_t)m > len) return false;
    uint32_t s_lo = pos + 1u >= (uint32_t)m ? pos + 1u - (uint32_t)m : 0u;
    uint32_t s_hi = pos <= len - (uint32_t)m ? pos : len - (uint32_t)m;
    for (uint32_t s = s_lo; s <= s_hi; ++s) {
        bool ok = true;
        for (int32_t i = 0; i < m; ++i) {
            if (ref_letter(rv, s + (uint32_t)i) != pat[i]) { ok = false; break; }
        }
        if (ok) return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------
// k_ref_pack: one thread per 32 reference positions
// ---------------------------------------------------------------------------------------
struct RefPackParams {
    const uint8_t *ascii;             // chunk of the contig
    uint32_t g_first;                 // contig position of ascii[0] (multiple of 32)
    uint32_t n;                       // letters in this chunk
    uint32_t prev_letter;             // upper-cased letter before ascii[0] in this contig, 0 at contig start
    uint32_t *ref2;
    uint32_t *excm;
    uint32_t *exc_start;              // unsorted append; host sorts
    uint8_t *exc_letter;
    uint32_t exc_cap;
    uint32_t *exc_n;
};

__device__ __forceinline__ uint32_t norm_letter(uint32_t c) {       // toupper + U->T (C locale)
    if (c >= 'a' && c <= 'z') c -= 32u;
    return c == 'U' ? (uint32_t)'T' : c;
}

__global__ void k_ref_pack(RefPackParams p) {
    unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long i0 = gid * 32ull;
    if (i0 >= p.n) return;
    uint32_t lo = 0, hi = 0, ex = 0;
    uint32_t prev = i0 == 0 ? p.prev_letter : norm_letter(p.ascii[i0 - 1]);
    for (uint32_t k = 0; k < 32; ++k) {
        unsigned long long i = i0 + k;
        if (i >= p.n) break;
        uint32_t c = norm_letter(p.ascii[i]);
        uint32_t b = c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
        if (b == 4u) {
            ex |= 1u << k;
            if (c != prev) {                                   // run start of an exception letter
                uint32_t slot = atomicAdd(p.exc_n, 1u);
                if (slot < p.exc_cap) { p.exc_start[slot] = p.g_first + (uint32_t)i; p.exc_letter[slot] = (uint8_t)c; }
            }
            b = 0;
        }
        if (k < 16) lo |= b << (2 * k); else hi |= b << (2 * (k - 16));
        prev = c;
    }
    unsigned long long g = (unsigned long long)p.g_first + i0;
    p.ref2[g >> 4] = lo;
    p.ref2[(g >> 4) + 1] = hi;
    p.excm[g >> 5] = ex;
}

// ---------------------------------------------------------------------------------------
// k_decode
// ---------------------------------------------------------------------------------------
struct BlockDesc {
    uint32_t hdr_end;                 // first byte after the status flag
    uint32_t end;                     // position of ';' (or mm_len)
    uint8_t  cls;                     // base class after strand complement: A0 C1 G2 T3 N4 (src/mod.c:97)
    uint8_t  is_n;                    // canonical base is exactly 'N' (src/mod.c:1102,1164)
    uint8_t  dot;                     // status '.' -> implicit calls (src/mod.c:1203)
    uint8_t  K;                       // mod_codes_len
    int16_t  req[kMaxCodes];          // -c entry per code, -1: not required
    uint8_t  outc[kMaxCodes];         // output code id
    uint8_t  any_req;
};

struct ReadShared {
    uint32_t total_q, total_r;
    uint32_t err;
    uint32_t cur_cls;                 // class the base index currently describes, 0xff none
    uint32_t idx_shift, idx_chunks, idx_total;
    uint32_t n_semi;
    int32_t  prev_end;                // ';' position that ends the block before this round's first
    uint32_t ml_base;
    uint32_t next_read;
};

struct ReadCtx {                      // uniform per read, lives in registers
    uint32_t r;
    int32_t tid, pos;
    uint32_t L, n_cig, mm_len, ml_len, rev, hp;
    const uint32_t *cig;
    const uint8_t *seq, *mm, *ml;
    uint32_t *cq, *cr;                // shared or scratch
    uint32_t *bm;                     // shared or scratch
    ContigDev cd;                     // reference + dense counts of the read's contig
};

// class mask of the 64 bases of word w: bit j <-> base 64w+j
__device__ __forceinline__ unsigned long long class_mask64(const uint8_t *seq, uint32_t w, uint32_t L, uint32_t cls) {
    const uint4 *p = reinterpret_cast<const uint4 *>(seq + (size_t)w * 32u);
    uint4 v0 = p[0], v1 = p[1];
    uint32_t x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    unsigned long long m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t u = ((x[k] & 0x0f0f0f0fu) << 4) | ((x[k] >> 4) & 0x0f0f0f0fu);   // nibble i <-> base 8k+i
        uint32_t z;
        if (cls == 0) {   // everything that is not C,G,T,N (src/mod.c:97 default 0)
            uint32_t yc = u ^ 0x22222222u, yg = u ^ 0x44444444u, yt = u ^ 0x88888888u, yn = u ^ 0xffffffffu;
            uint32_t nc = (yc | (yc >> 1) | (yc >> 2) | (yc >> 3));
            uint32_t ng = (yg | (yg >> 1) | (yg >> 2) | (yg >> 3));
            uint32_t nt = (yt | (yt >> 1) | (yt >> 2) | (yt >> 3));
            uint32_t nn = (yn | (yn >> 1) | (yn >> 2) | (yn >> 3));
            z = nc & ng & nt & nn & 0x11111111u;
        } else {
            uint32_t pat = cls == 1 ? 0x22222222u : cls == 2 ? 0x44444444u : cls == 3 ? 0x88888888u : 0xffffffffu;
            uint32_t y = u ^ pat;
            z = ~(y | (y >> 1) | (y >> 2) | (y >> 3)) & 0x11111111u;
        }
        z = (z | (z >> 3)) & 0x03030303u;
        z = (z | (z >> 6)) & 0x000f000fu;
        z = (z | (z >> 12)) & 0xffu;
        m |= (unsigned long long)z << (8 * k);
    }
    uint32_t base0 = w * 64u;
    if (base0 + 64u > L) {
        uint32_t valid = L > base0 ? L - base0 : 0u;
        m &= valid >= 64u ? ~0ull : ((1ull << valid) - 1ull);
    }
    return m;
}

// position of the n-th (0-based) set bit of m; n < popcount(m)
__device__ __forceinline__ uint32_t nth_set_bit64(unsigned long long m, uint32_t n) {
    uint32_t x = (uint32_t)m, base = 0;
    uint32_t c = __popc(x);
    if (n >= c) { n -= c; x = (uint32_t)(m >> 32); base = 32; }
    uint32_t sh = 0;
    c = __popc(x & 0xffffu);        if (n >= c) { n -= c; sh += 16; }
    c = __popc((x >> sh) & 0xffu);  if (n >= c) { n -= c; sh += 8; }
    c = __popc((x >> sh) & 0xfu);   if (n >= c) { n -= c; sh += 4; }
    c = __popc((x >> sh) & 0x3u);   if (n >= c) { n -= c; sh += 2; }
    c = (x >> sh) & 1u;             if (n >= c) { sh += 1; }
    return base + sh;
}

struct AlnHit { int32_t aln; int32_t ins; uint32_t insoff; };

// aln[q], ins[q], ins_offset[q] of get_aln() in BAM orientation (src/mod.c:776-881)
__device__ __forceinline__ AlnHit cigar_lookup(const ReadCtx &rc, uint32_t total_q, uint32_t q) {
    AlnHit h; h.aln = -1; h.ins = -1; h.insoff = 0;
    if (q >= total_q || rc.n_cig == 0) return h;
    uint32_t lo = 0, hi = rc.n_cig;                              // (cq[lo]>>4) <= q < (cq[hi]>>4)
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if ((rc.cq[mid] >> 4) <= q) lo = mid; else hi = mid;
    }
    uint32_t e = rc.cq[lo], op = e & 15u, d = q - (e >> 4);
    if (op == 0u || op == 7u || op == 8u) h.aln = rc.pos + (int32_t)(rc.cr[lo] + d);
    else if (op == 1u) { h.ins = rc.pos + (int32_t)rc.cr[lo] - 1; h.insoff = d + 1u; }
    return h;
}

__device__ __forceinline__ void raise(ReadShared *rs, uint32_t code) {
    atomicCAS(&rs->err, 0u, code);
}

__device__ __forceinline__ unsigned long long vload_u64(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// output-code dictionary: id of `key`, inserting when `insert`; -1 if absent/full
__device__ __forceinline__ int32_t code_id(unsigned long long *table, unsigned long long key, bool insert) {
    for (int32_t i = 0; i < kCodeTable; ++i) {
        unsigned long long cur = vload_u64(&table[i]);
        if (cur == key) return i;
        if (cur == 0ull) {
            if (!insert) return -1;
            unsigned long long old = atomicCAS(&table[i], 0ull, key);
            if (old == 0ull || old == key) return i;
        }
    }
    return -1;
}

// update_freq_map() (src/mod.c:883-929): one 64-bit add per cell, n_called low / n_mod high
__device__ __forceinline__ void add_cell(const DecodeParams &P, const ReadCtx &rc, int32_t ref_pos,
                                         uint32_t outc, uint32_t ins16, int32_t hap /* -1 => '*'/none */,
                                         uint32_t is_mod) {
    int32_t hslot = -1;
    if (!P.haplotypes || hap < 0) hslot = 0;
    else if (hap + 1 < P.n_hap_slots) hslot = hap + 1;
    if (ins16 == 0u && outc < (uint32_t)P.n_code_slots && hslot >= 0) {
        unsigned long long idx = (((unsigned long long)(uint32_t)ref_pos * 2ull + rc.rev) * (unsigned)P.n_code_slots + outc) * (unsigned)P.n_hap_slots + (unsigned)hslot;
        atomicAdd(&rc.cd.cells[idx], 1ull | ((unsigned long long)is_mod << 32));
    } else {
        unsigned long long slot = atomicAdd(P.sparse_n, 1ull);
        if (slot < P.sparse_cap) {
            SparseRec s;
            s.a = ((unsigned long long)(uint32_t)rc.tid << 41) | ((unsigned long long)(uint32_t)ref_pos << 9) | ((unsigned long long)rc.rev << 8) | outc;
            s.b = ins16 | ((hap < 0 ? 256u : (uint32_t)hap) << 16);
            s.w = 1u | (is_mod << 16);
            P.sparse[slot] = s;
        }
    }
}

// Everything after "this base of the read is a call": A.4-A.8 of SURVEY.md.
//   q        BAM-orientation read position
//   implicit call comes from the skipped-base loops (src/mod.c:1203-1367)
//   cidx     index of the call in its block (explicit) / its rank (implicit), for ML and view order
__device__ __forceinline__ void process_call(const DecodeParams &P, const ReadCtx &rc, ReadShared *rs,
                                             const BlockDesc &bd, uint32_t blk_ord, uint32_t q,
                                             bool implicit, uint32_t cidx, uint32_t ml_base) {
    const uint32_t total_q = rs->total_q;
    AlnHit h = cigar_lookup(rc, total_q, q);
    int32_t ref_pos = h.aln;
    if (ref_pos < 0 && P.insertions) {
        if (implicit && rc.rev) {
            // Q9: the reference indexes its FASTQ-oriented ins[] with the BAM-orientation
            // position (src/mod.c:1234,1314), i.e. it reads ins_bam[L-1-q].
            AlnHit h2 = cigar_lookup(rc, total_q, rc.L - 1u - q);
            ref_pos = h2.ins;
        } else {
            ref_pos = h.ins;
        }
    }
    if (ref_pos < 0) return;                                       // src/mod.c:1127,1237,1317
    const uint32_t ins_off = P.insertions ? h.insoff : 0u;
    const uint32_t nib = (rc.seq[q >> 1] >> ((~q & 1u) << 2)) & 0xfu;
    const uint32_t read_letter = nt16_letter(nib);
    const RefView &rv = rc.cd;

    int32_t ref_match = -1;                                        // lazily evaluated ref==read
    for (uint32_t m = 0; m < bd.K; ++m) {
        int32_t ri = bd.req[m];
        if (ri < 0) continue;                                      // src/mod.c:1157
        const ReqMod &rq = P.req[ri];
        if (!P.insertions && rq.ctx_len > 0) {                     // src/mod.c:1162-1172
            if (!in_context(rv, (uint32_t)ref_pos, rc.rev ? rq.pat_rc : rq.pat, rq.ctx_len)) continue;
            if (!bd.is_n) {
                if (ref_match < 0) ref_match = ref_letter(rv, (uint32_t)ref_pos) == read_letter ? 1 : 0;
                if (!ref_match) continue;
            }
        }
        uint32_t prob = 0, is_mod = 0;
        if (!implicit) {
            unsigned long long ml_idx = (unsigned long long)ml_base + (unsigned long long)cidx * bd.K + m;
            if (ml_idx >= rc.ml_len) { raise(rs, kErrMLIndex); return; }   // src/mod.c:1174
            prob = rc.ml[ml_idx];
        }
        if (P.subtool == 1) {                                      // FREQ
            if (!implicit) {
                uint32_t f = rq.lut[prob];                         // src/mod.c:1181-1191
                if (!(f & 1u)) continue;
                is_mod = (f >> 1) & 1u;
            }                                                      // implicit: called, unmodified, no threshold (src/mod.c:1279)
            const uint32_t ins16 = ins_off & 0xffffu;              // make_key's uint16_t (src/mod.c:428)
            if (P.haplotypes) {
                add_cell(P, rc, ref_pos, bd.outc[m], ins16, (int32_t)rc.hp, is_mod);
                add_cell(P, rc, ref_pos, bd.outc[m], ins16, -1, is_mod);       // src/mod.c:906-928
            } else {
                add_cell(P, rc, ref_pos, bd.outc[m], ins16, -1, is_mod);
            }
        } else {                                                   // VIEW
            unsigned long long slot = atomicAdd(P.view_n, 1ull);
            if (slot < P.view_cap) {
                ViewDev v;
                v.read = rc.r; v.ref_pos = ref_pos;
                v.read_pos = (int32_t)(rc.rev ? rc.L - 1u - q : q);
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

// bases_pos[cls][k] (src/mod.c:977-981) without materialising it: sampled rank index + in-word select
__device__ __forceinline__ uint32_t select_base(const ReadCtx &rc, const ReadShared *rs, const uint32_t *idx,
                                                uint32_t cls, uint32_t k) {
    uint32_t lo = 0, hi = rs->idx_chunks;                         // idx[lo] <= k < idx[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (idx[mid] <= k) lo = mid; else hi = mid;
    }
    uint32_t rem = k - idx[lo];
    uint32_t w = lo << rs->idx_shift;
    for (;;) {
        unsigned long long m = class_mask64(rc.seq, w, rc.L, cls);
        uint32_t pc = (uint32_t)__popcll(m);
        if (rem < pc) return w * 64u + nth_set_bit64(m, rem);
        rem -= pc; ++w;
    }
}

__global__ void __launch_bounds__(kMaxThreads) k_decode(DecodeParams P) {
    __shared__ uint32_t s_cq[kCigSmem];
    __shared__ uint32_t s_cr[kCigSmem];
    __shared__ uint32_t s_idx[kIdxSmem + 1];
    __shared__ uint32_t s_bm[kBitmapWords];
    __shared__ uint32_t s_rank[kRankCap];
    __shared__ uint32_t s_semi[kMaxBlocks];
    __shared__ BlockDesc s_blk[kMaxBlocks];
    __shared__ uint32_t s_ws[32];
    __shared__ __align__(16) uint8_t s_text[kMaxThreads * kChunk + 32];
    __shared__ ReadShared s_rs;

    const uint32_t t = threadIdx.x, NT = blockDim.x;
    ReadShared *rs = &s_rs;
    uint32_t *scratch = P.scratch ? P.scratch + (size_t)blockIdx.x * P.scratch_words_per_cta : nullptr;

    for (;;) {
        // ---- dynamic read scheduler
        __syncthreads();
        if (t == 0) {
            uint32_t nx = atomicAdd(P.work_counter, 1u);
            if (P.read_list) nx = nx < *P.read_list_n ? P.read_list[nx] : 0xffffffffu;   // reads deferred by k_decode_warp
            rs->next_read = nx;
        }
        __syncthreads();
        const uint32_t r = rs->next_read;
        if (r >= P.n_reads) break;

        ReadCtx rc;
        rc.r = r;
        rc.tid = P.tid[r]; rc.pos = P.pos[r];
        rc.L = P.l_seq[r]; rc.n_cig = P.n_cigar[r];
        rc.mm_len = P.mm_len[r]; rc.ml_len = P.ml_len[r];
        rc.rev = (P.flag[r] >> 4) & 1u;
        rc.hp = P.hp[r];
        rc.cig = P.cigar + P.cigar_off[r];
        rc.seq = P.seq4 + P.seq_off[r];
        rc.mm = P.mm + P.mm_off[r];
        rc.ml = P.ml + P.ml_off[r];
        const bool cig_in_smem = rc.n_cig <= (uint32_t)P.cig_smem_cap;
        rc.cq = cig_in_smem ? s_cq : scratch;
        rc.cr = cig_in_smem ? s_cr : scratch + P.scratch_cig_words;
        const bool bm_in_smem = ((rc.L + 31u) >> 5) <= (uint32_t)P.bitmap_smem_words;
        rc.bm = bm_in_smem ? s_bm : scratch + 2ull * P.scratch_cig_words;
        rc.cd.ref2 = nullptr; rc.cd.excm = nullptr; rc.cd.exc_start = nullptr; rc.cd.exc_letter = nullptr;
        rc.cd.cells = nullptr; rc.cd.n_exc = 0; rc.cd.len = 0;

        if (t == 0) {
            rs->err = 0; rs->cur_cls = 0xffu; rs->ml_base = 0; rs->prev_end = -1;
            rs->total_q = 0; rs->total_r = 0;
            if (rc.tid < 0 || rc.tid >= P.n_contigs || P.contigs[rc.tid].ref2 == nullptr) rs->err = kErrNoContig;
            else if (rc.L >= (1u << 28)) rs->err = kErrSeqTooLong;
        }
        uint32_t err = block_err(&rs->err);          // CTA-uniform from here on

        if (!err) {
            rc.cd = P.contigs[rc.tid];

            // ---- (1) CIGAR prefix sums == get_aln() (src/mod.c:811-880) as two scans
            uint32_t carry_q = 0, carry_r = 0;
            for (uint32_t base = 0; base < rc.n_cig; base += NT) {
                uint32_t i = base + t, op = 15u, len = 0, ql = 0, rl = 0;
                if (i < rc.n_cig) {
                    uint32_t w = rc.cig[i];
                    op = w & 15u; len = w >> 4;
                    if (op == 0u || op == 7u || op == 8u) { ql = len; rl = len; }
                    else if (op == 1u || op == 4u) ql = len;
                    else if (op == 2u || op == 3u) rl = len;
                    else if (op == 5u) raise(rs, kErrHardClip);
                    else raise(rs, kErrCigarOp);
                }
                uint32_t eq = ql, er = rl, tq, tr;
                block_scan2_sat(eq, er, tq, tr, s_ws);
                if (i < rc.n_cig) {
                    uint32_t q0 = sat_add(carry_q, eq), r0 = sat_add(carry_r, er);
                    const bool alnop = op == 0u || op == 7u || op == 8u;
                    if ((alnop || (op == 1u && P.insertions)) && len > 0 && sat_add(q0, ql) > rc.L)
                        raise(rs, kErrCigarLen);
                    if (alnop && len > 0) {
                        long long last = (long long)rc.pos + r0 + len - 1;
                        if (rc.pos < 0 || last >= (long long)rc.cd.len) raise(rs, kErrRefRange);
                    }
                    if (q0 >= (1u << 28)) q0 = (1u << 28) - 1u;     // only reachable on an error path
                    rc.cq[i] = (q0 << 4) | op;
                    rc.cr[i] = r0;
                }
                carry_q = sat_add(carry_q, tq); carry_r = sat_add(carry_r, tr);
            }
            if (t == 0) {
                rs->total_q = carry_q < rc.L ? carry_q : rc.L;
                rs->total_r = carry_r;
            }
            err = block_err(&rs->err);
            if (!err && t == 0) {
                int32_t lo = rc.pos > 0 ? rc.pos - 1 : 0;
                long long hi = (long long)rc.pos + carry_r + 1;
                if (hi > (long long)rc.cd.len) hi = rc.cd.len;
                atomicMin(&P.touch_lo[rc.tid], lo);
                atomicMax(&P.touch_hi[rc.tid], (int32_t)hi);
            }
        }

        // ---- (2) MM blocks, kMaxBlocks per round
        uint32_t n_blocks = 0, n_semi = 0;
        for (uint32_t b0 = 0; !err && (b0 == 0 || b0 < n_blocks); b0 += kMaxBlocks) {
            // (2a) ordered positions of ';' with ordinal in [b0, b0+kMaxBlocks)
            uint32_t carry = 0;
            for (uint32_t base = 0; base < rc.mm_len; base += NT * kChunk) {
                uint32_t p0 = base + t * kChunk, mask = 0;
                if (p0 < rc.mm_len) {
                    uint4 v = *reinterpret_cast<const uint4 *>(rc.mm + p0);
                    mask = byte_eq_mask(v.x, ';') | (byte_eq_mask(v.y, ';') << 4) |
                           (byte_eq_mask(v.z, ';') << 8) | (byte_eq_mask(v.w, ';') << 12);
                    uint32_t valid = rc.mm_len - p0;
                    if (valid < 16u) mask &= (1u << valid) - 1u;
                }
                uint32_t e = (uint32_t)__popc(mask), dummy = 0, tot, td;
                block_scan2_sat(e, dummy, tot, td, s_ws);
                uint32_t ord = carry + e;
                while (mask) {
                    uint32_t bit = (uint32_t)__ffs((int)mask) - 1u;
                    mask &= mask - 1u;
                    if (ord >= b0 && ord < b0 + kMaxBlocks) s_semi[ord - b0] = p0 + bit;
                    ++ord;
                }
                carry += tot;
            }
            __syncthreads();
            if (b0 == 0) {
                n_semi = carry;
                n_blocks = carry;
                if (rc.mm_len > 0 && rc.mm[rc.mm_len - 1] != ';') n_blocks += 1;   // unterminated last block
            }
            const uint32_t nb_round = n_blocks - b0 < (uint32_t)kMaxBlocks ? n_blocks - b0 : (uint32_t)kMaxBlocks;

            // (2b) block headers: base, strand, codes, status (src/mod.c:1003-1062)
            if (t < nb_round) {
                BlockDesc bd;
                uint32_t start = t == 0 ? (uint32_t)(rs->prev_end + 1) : s_semi[t - 1] + 1u;
                uint32_t end = (b0 + t) < n_semi ? s_semi[t] : rc.mm_len;
                bd.end = end; bd.hdr_end = end; bd.K = 0; bd.any_req = 0; bd.cls = 0; bd.is_n = 0; bd.dot = 1;
                for (int k = 0; k < kMaxCodes; ++k) { bd.req[k] = -1; bd.outc[k] = 0; }
                uint32_t i = start;
                uint32_t base_c = i < end ? rc.mm[i] : 0u;
                bool okb = base_c == 'A' || base_c == 'C' || base_c == 'G' || base_c == 'T' || base_c == 'U' || base_c == 'N' ||
                           base_c == 'a' || base_c == 'c' || base_c == 'g' || base_c == 't' || base_c == 'u' || base_c == 'n';
                if (!okb) raise(rs, kErrMMBase);
                else {
                    ++i;
                    uint32_t modbase = base_c == 'U' ? (uint32_t)'T' : base_c;                 // src/mod.c:1006
                    uint32_t strand_c = i < end ? rc.mm[i] : 0u;
                    if (strand_c != '+' && strand_c != '-') raise(rs, kErrMMStrand);
                    else {
                        ++i;
                        uint8_t codes[kMaxCodes];
                        uint32_t j = 0; bool has_num = false, has_alpha = false, bad = false, too_many = false;
                        while (i < end) {
                            uint32_t c = rc.mm[i];
                            if (c == ',' || c == '?' || c == '.') break;
                            if (c >= '0' && c <= '9') has_num = true;
                            else if ((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z')) has_alpha = true;
                            else { bad = true; break; }
                            if (j < (uint32_t)kMaxCodes) codes[j] = (uint8_t)c; else too_many = true;
                            ++j; ++i;
                        }
                        if (bad || j == 0 || (has_num && has_alpha)) raise(rs, kErrMMCode);
                        else if (too_many) raise(rs, kErrTooManyCodes);
                        else {
                            uint32_t K = has_num ? 1u : j;                                   // src/mod.c:1048
                            if (i < end && (rc.mm[i] == '?' || rc.mm[i] == '.')) { bd.dot = rc.mm[i] == '.'; ++i; }
                            bd.hdr_end = i;
                            bd.K = (uint8_t)K;
                            // complement when the read is reverse (src/mod.c:1092-1093, table :98)
                            uint32_t mb = modbase;
                            if (rc.rev) {
                                switch (modbase) {
                                    case 'A': mb = 'T'; break; case 'C': mb = 'G'; break; case 'G': mb = 'C'; break;
                                    case 'T': mb = 'A'; break; case 'N': mb = 'N'; break;
                                    case 'a': mb = 't'; break; case 'c': mb = 'g'; break; case 'g': mb = 'c'; break;
                                    case 't': mb = 'a'; break; case 'u': mb = 'a'; break; case 'n': mb = 'n'; break;
                                }
                            }
                            uint32_t up = mb >= 'a' ? mb - 32u : mb;
                            bd.cls = up == 'A' ? 0 : up == 'C' ? 1 : up == 'G' ? 2 : (up == 'T' || up == 'U') ? 3 : 4;
                            bd.is_n = modbase == 'N';
                            for (uint32_t m = 0; m < K; ++m) {
                                // code m is the suffix string codes[m..] (Q4, src/mod.c:1148-1152)
                                unsigned long long key = 0;
                                uint32_t from = has_num ? 0u : m;
                                for (uint32_t z = from; z < j; ++z) key |= (unsigned long long)codes[z] << (8u * (z - from));
                                if (P.wild_req >= 0) {
                                    int32_t id = code_id(P.code_keys, key, true);
                                    if (id < 0) { raise(rs, kErrTooManyCodes); break; }
                                    bd.req[m] = (int16_t)P.wild_req; bd.outc[m] = (uint8_t)id; bd.any_req = 1;
                                } else {
                                    for (int32_t q = 0; q < P.n_req; ++q)
                                        if (P.req[q].key == key) { bd.req[m] = (int16_t)q; bd.outc[m] = (uint8_t)q; bd.any_req = 1; break; }
                                }
                            }
                        }
                    }
                }
                s_blk[t] = bd;
            }
            err = block_err(&rs->err);

            // (2c) blocks in order
            for (uint32_t jb = 0; jb < nb_round && !err; ++jb) {
                const BlockDesc bd = s_blk[jb];
                const uint32_t a0 = bd.hdr_end, a1 = bd.end;
                const uint32_t ml_base = rs->ml_base;
                const bool work = bd.any_req != 0;
                const bool need_bm = work && bd.dot;

                // base-rank index for this block's class
                uint32_t cnt_cls = 0;
                if (work && (!bd.is_n || bd.dot)) {
                    const uint32_t cls = bd.cls;                   // 4 when the canonical base is N
                    if (rs->cur_cls != cls) {
                        __syncthreads();
                        const uint32_t n_words = (rc.L + 63u) >> 6;
                        uint32_t sh = 0;
                        while (((n_words + (1u << sh) - 1u) >> sh) > (uint32_t)P.idx_smem_cap) ++sh;
                        const uint32_t n_chunks = (n_words + (1u << sh) - 1u) >> sh;
                        for (uint32_t c = t; c < n_chunks; c += NT) {
                            uint32_t w0 = c << sh, w1 = w0 + (1u << sh), s = 0;
                            if (w1 > n_words) w1 = n_words;
                            for (uint32_t w = w0; w < w1; ++w) s += (uint32_t)__popcll(class_mask64(rc.seq, w, rc.L, cls));
                            s_idx[c] = s;
                        }
                        __syncthreads();
                        const uint32_t seg = (n_chunks + NT - 1u) / NT;
                        uint32_t c0 = t * seg, c1 = c0 + seg, local = 0;
                        if (c0 > n_chunks) c0 = n_chunks;
                        if (c1 > n_chunks) c1 = n_chunks;
                        for (uint32_t c = c0; c < c1; ++c) local += s_idx[c];
                        uint32_t e = local, dummy = 0, tot, td;
                        block_scan2_sat(e, dummy, tot, td, s_ws);
                        for (uint32_t c = c0; c < c1; ++c) { uint32_t v = s_idx[c]; s_idx[c] = e; e += v; }
                        if (t == 0) {
                            s_idx[n_chunks] = tot;
                            rs->cur_cls = cls; rs->idx_shift = sh; rs->idx_chunks = n_chunks; rs->idx_total = tot;
                        }
                        __syncthreads();
                    }
                    cnt_cls = rs->idx_total;
                }
                if (need_bm) {
                    const uint32_t words = (rc.L + 31u) >> 5;
                    for (uint32_t w = t; w < words; w += NT) rc.bm[w] = 0;
                    __syncthreads();
                }

                // (2d) skip counts -> ranks -> calls, one text tile at a time
                uint32_t carry_cnt = 0, carry_sum = 0;
                const uint32_t tile0 = a0 & ~15u;
                for (uint32_t tb = tile0; tb < a1; tb += NT * kChunk) {
                    // stage [tb-16, tb+NT*16+16) of the text: s_text[p - tb + 16] == mm[p]
                    for (uint32_t v = t; v < NT + 2u; v += NT) {
                        long long src = (long long)tb - 16 + (long long)v * 16;
                        uint4 val = make_uint4(0, 0, 0, 0);
                        if (src >= 0 && (uint32_t)src < rc.mm_len) val = *reinterpret_cast<const uint4 *>(rc.mm + src);
                        *reinterpret_cast<uint4 *>(s_text + v * 16u) = val;
                    }
                    __syncthreads();
#define MMC_TX(p) ((uint32_t)s_text[(p) - tb + 16u])
                    const uint32_t p0 = tb + t * kChunk;
                    const uint32_t pa = p0 > a0 ? p0 : a0, pe = p0 + kChunk < a1 ? p0 + kChunk : a1;
                    uint32_t cnt = 0, sum = 0, e_cnt = 0, e_sum = 0;
                    for (int pass = 0; pass < 2; ++pass) {
                        if (pass == 1) {
                            e_cnt = cnt; e_sum = sum;
                            uint32_t tcnt, tsum;
                            block_scan2_sat(e_cnt, e_sum, tcnt, tsum, s_ws);
                            cnt = tcnt; sum = tsum;                         // tile totals, identical in all threads
                            if (!work) break;
                        }
                        uint32_t run_cnt = 0, run_sum = 0;
                        if (pa < pe) {
                            bool prev_comma = pa == a0 || MMC_TX(pa - 1u) == ',';
                            for (uint32_t p = pa; p < pe; ++p) {
                                uint32_t c = MMC_TX(p);
                                if (c == ',') { prev_comma = true; continue; }
                                if (prev_comma) {                            // a skip count starts here
                                    uint32_t v = 0, nd = 0, q = p;
                                    bool bad = false;
                                    while (q < a1) {
                                        uint32_t d = MMC_TX(q);
                                        if (d == ',') break;
                                        if (d < '0' || d > '9' || nd >= 9u) { bad = true; break; }
                                        v = v * 10u + (d - '0'); ++nd; ++q;
                                    }
                                    if (bad) raise(rs, kErrMMSkip);
                                    run_sum = sat_add(run_sum, v + 1u);
                                    if (pass == 1) {
                                        uint32_t incl = sat_add(sat_add(carry_sum, e_sum), run_sum);
                                        s_rank[e_cnt + run_cnt] = incl - 1u;       // base_rank (src/mod.c:1098)
                                    }
                                    ++run_cnt;
                                }
                                prev_comma = false;
                            }
                        }
                        if (pass == 0) { cnt = run_cnt; sum = run_sum; }
                    }
#undef MMC_TX
                    __syncthreads();
                    if (work && rs->err == 0) {            // non-uniform read is fine: no barrier inside
                        for (uint32_t c = t; c < cnt; c += NT) {
                            const uint32_t rank = s_rank[c];
                            uint32_t q;
                            if (bd.is_n) {                                   // src/mod.c:1102-1107
                                if (rank >= rc.L) { raise(rs, kErrMMRank); continue; }
                                q = rc.rev ? rc.L - 1u - rank : rank;
                            } else {                                         // src/mod.c:1109-1113
                                if (rank >= cnt_cls) { raise(rs, kErrMMRank); continue; }
                                q = select_base(rc, rs, s_idx, bd.cls, rc.rev ? cnt_cls - 1u - rank : rank);
                            }
                            if (need_bm) atomicOr(&rc.bm[rank >> 5], 1u << (rank & 31u));
                            process_call(P, rc, rs, bd, b0 + jb, q, false, carry_cnt + c, ml_base);
                        }
                    }
                    carry_cnt += cnt; carry_sum = sat_add(carry_sum, sum);
                    __syncthreads();
                }
                err = block_err(&rs->err);

                // (2e) implicit calls of a '.' block (src/mod.c:1203-1367)
                if (need_bm && !err) {
                    if (bd.is_n) {
                        // ranks [0,last) and (last, #N letters): Q8
                        const uint32_t last1 = carry_cnt > 0 ? carry_sum : 0u;   // last+1
                        const uint32_t bound = last1 > cnt_cls ? last1 : cnt_cls;
                        for (uint32_t s = t; s < bound; s += NT) {
                            if ((rc.bm[s >> 5] >> (s & 31u)) & 1u) continue;
                            const uint32_t q = rc.rev ? rc.L - 1u - s : s;
                            process_call(P, rc, rs, bd, b0 + jb, q, true, s, ml_base);
                        }
                    } else {
                        const uint32_t n_words = (rc.L + 63u) >> 6;
                        const uint32_t sh = rs->idx_shift;
                        for (uint32_t w = t; w < n_words; w += NT) {
                            unsigned long long m = class_mask64(rc.seq, w, rc.L, bd.cls);
                            if (!m) continue;
                            // forward rank of the first class base in this word
                            uint32_t fr = s_idx[w >> sh];
                            for (uint32_t ww = (w >> sh) << sh; ww < w; ++ww)
                                fr += (uint32_t)__popcll(class_mask64(rc.seq, ww, rc.L, bd.cls));
                            while (m) {
                                uint32_t bit = (uint32_t)__ffsll((long long)m) - 1u;
                                m &= m - 1ull;
                                const uint32_t s = rc.rev ? cnt_cls - 1u - fr : fr;
                                ++fr;
                                if ((rc.bm[s >> 5] >> (s & 31u)) & 1u) continue;
                                process_call(P, rc, rs, bd, b0 + jb, w * 64u + bit, true, s, ml_base);
                            }
                        }
                    }
                    err = block_err(&rs->err);
                }
                if (t == 0 && carry_cnt > 0) rs->ml_base = ml_base + carry_cnt * bd.K;   // src/mod.c:1200
                __syncthreads();
            }
            if (t == 0 && nb_round == (uint32_t)kMaxBlocks) rs->prev_end = (int32_t)s_semi[kMaxBlocks - 1];
            __syncthreads();
        }

        if (t == 0 && err != 0)
            atomicMin(P.err, ((unsigned long long)r << 32) | err);
    }
}

// ---------------------------------------------------------------------------------------
// finalize: dense arrays -> position-ordered records
// ---------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------
// SEQ transport (mmc_batch_t.seq2, include/minimod_cuda.h): SEQ is ~87 % of the bytes of a batch and the
// end-to-end path is PCIe-bound, so the packer can send 2 bits per base plus a list of the bases that are
// not A,C,G,T; these two kernels rebuild BAM's 4-bit pool in HBM, bit for bit, before the decode kernels run.
//   k_unpack_seq2   8 bytes of seq2 (32 bases, first base in bits 7:6 of byte 0) -> 16 bytes of seq4
//   k_patch_seq4    exception e: nibble (e >> 4) of the pool := e & 15
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t seq2_expand16(uint32_t v) {              // 2 bytes of seq2 -> 4 bytes of seq4
    v &= 0xffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;                                        // 2-bit group k -> low half of nibble k
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    const uint32_t x0 = v & 0x11111111u, x1 = (v >> 1) & 0x11111111u, n0 = x0 ^ 0x11111111u, n1 = x1 ^ 0x11111111u;
    v = (n1 & n0) | ((n1 & x0) << 1) | ((x1 & n0) << 2) | ((x1 & x0) << 3);  // code c -> nt16 one-hot 1 << c
    return ((v & 0x00ff00ffu) << 8) | ((v >> 8) & 0x00ff00ffu);              // first base of a byte in its high nibble, bytes in memory order
}

__global__ void __launch_bounds__(256) k_unpack_seq2(const uint2 *in, uint4 *out, unsigned long long n8) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint2 v = in[i];
        uint4 o;
        o.x = seq2_expand16(v.x); o.y = seq2_expand16(v.x >> 16); o.z = seq2_expand16(v.y); o.w = seq2_expand16(v.y >> 16);
        out[i] = o;
    }
}

__global__ void __launch_bounds__(256) k_patch_seq4(const unsigned long long *exc, unsigned long long n, uint32_t *seq4_words) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long e = exc[i], nib = e >> 4, byte = nib >> 1;
        const uint32_t shift = (uint32_t)(byte & 3ull) * 8u + ((nib & 1ull) ? 0u : 4u);
        uint32_t *w = seq4_words + (byte >> 2);
        atomicAnd(w, ~(0xfu << shift));                                      // neighbours in the word may be patched by other threads
        atomicOr(w, (uint32_t)(e & 15ull) << shift);
    }
}

// ---------------------------------------------------------------------------------------
// CIGAR transport (mmc_batch_t.cig8, include/minimod_cuda.h): ONT alignments have an op per ~17 bases, a third of a
// whole-genome batch's bytes as 32-bit words.  The packer can send a byte per op (op | min(len,15) << 4) plus two
// escape lists per read (one byte, then 32 bits, for the lengths that do not fit); this kernel rebuilds BAM's word
// pool in HBM, word for word, before the decode kernels run.  One warp per read, a lane per op, 32 ops per step; an
// op's place in the escape lists is its rank among the escaped ops (ballot + popcount, running over the steps).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_unpack_cigar(const uint8_t *cig8, const unsigned long long *cig8_off, const uint32_t *n_cigar,
                                                      const unsigned long long *cigar_off, uint32_t n_reads, uint32_t *cigar) {
    const uint32_t lane = threadIdx.x & 31u, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_reads; r += n_warps) {
        const uint8_t *blob = cig8 + cig8_off[r];
        const uint32_t n = n_cigar[r];
        const uint32_t n1 = *reinterpret_cast<const uint32_t *>(blob);
        const uint8_t *ops = blob + 4, *l1 = ops + ((n + 3u) & ~3u);
        const uint32_t *l2 = reinterpret_cast<const uint32_t *>(l1 + ((n1 + 3u) & ~3u));
        uint32_t *out = cigar + cigar_off[r];
        uint32_t c1 = 0, c2 = 0;                                             // escaped ops before this step
        for (uint32_t base = 0; base < n; base += 32u) {
            const uint32_t i = base + lane;
            const uint32_t b = i < n ? ops[i] : 0u;
            const uint32_t op = b & 15u, l4 = b >> 4;
            const uint32_t m1 = __ballot_sync(0xffffffffu, l4 == 15u);
            const uint32_t v = l4 == 15u ? l1[c1 + (uint32_t)__popc(m1 & ((1u << lane) - 1u))] : 0u;
            const uint32_t m2 = __ballot_sync(0xffffffffu, v == 255u);
            uint32_t len = l4;
            if (l4 == 15u) len = v < 255u ? 15u + v : l2[c2 + (uint32_t)__popc(m2 & ((1u << lane) - 1u))];
            if (i < n) out[i] = (len << 4) | op;
            c1 += (uint32_t)__popc(m1); c2 += (uint32_t)__popc(m2);
        }
    }
}

// One launch covers every contig range of a finalize / drain: a job per range, tiles numbered across the jobs.
struct FinJob {
    const unsigned long long *cells;   // first cell of the scanned range
    unsigned long long n_cells;
    unsigned long long tile0;          // number of this job's first tile
    int32_t tid;
    int32_t lo;                        // contig position of cells[0]
};
struct FinalizeParams {
    const FinJob *jobs;                // sorted by tile0
    uint32_t n_jobs;
    int32_t n_code_slots, n_hap_slots, haplotypes;
    uint32_t *tile_count;              // per tile (CTA-sized range of cells)
    unsigned long long *tile_offset;   // exclusive scan of tile_count over all tiles of all jobs
    uint32_t cells_per_tile;
    void *out;                         // mmc_freq_rec_t[]
    uint32_t *mask;                    // one bit per cell of every tile: non-zero (written by k_count_nonzero, read by k_emit_records)
    uint32_t *overflow;                // set when a cell shows n_mod > n_called: n_called wrapped (src/mod.c:899-901)
};
__device__ __forceinline__ FinJob fin_job_of(const FinalizeParams &p, unsigned long long tile) {
    uint32_t lo = 0, hi = p.n_jobs - 1u;                          // last job with tile0 <= tile
    while (lo < hi) { const uint32_t mid = (lo + hi + 1u) >> 1; if (p.jobs[mid].tile0 <= tile) lo = mid; else hi = mid - 1u; }
    return p.jobs[lo];
}

// Pass 1 of the compaction: the only full read of the dense range.  One CTA (8 warps) per tile; every warp owns a contiguous
// eighth of the tile and leaves one ballot word per 32 cells, so that pass 2 never touches an empty cell again.
__global__ void __launch_bounds__(256) k_count_nonzero(const __grid_constant__ FinalizeParams p) {
    __shared__ uint32_t ws[8];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const FinJob j = fin_job_of(p, blockIdx.x);
    const unsigned long long c0 = ((unsigned long long)blockIdx.x - j.tile0) * p.cells_per_tile;
    unsigned long long c1 = c0 + p.cells_per_tile;
    if (c1 > j.n_cells) c1 = j.n_cells;
    const uint32_t per_warp = p.cells_per_tile / 8u;
    unsigned long long w0 = c0 + (unsigned long long)warp * per_warp, w1 = w0 + per_warp;
    if (w0 > c1) w0 = c1;
    if (w1 > c1) w1 = c1;
    uint32_t *mask = p.mask + (unsigned long long)blockIdx.x * (p.cells_per_tile / 32u) + warp * (per_warp / 32u);
    uint32_t cnt = 0, myword = 0, bad = 0;
    const uint32_t n_words = (uint32_t)((w1 - w0 + 31u) >> 5);
    for (uint32_t w = 0; w < n_words; w += 4u) {                  // four loads in flight per lane
        unsigned long long v[4];
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k) {
            const unsigned long long c = w0 + (unsigned long long)(w + k) * 32u + lane;
            v[k] = c < w1 ? j.cells[c] : 0ull;
        }
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k) {
            const uint32_t f = __ballot_sync(0xffffffffu, v[k] != 0ull);
            bad |= (uint32_t)((uint32_t)(v[k] >> 32) > (uint32_t)v[k]);
            if (lane == ((w + k) & 31u)) myword = f;
            cnt += (uint32_t)__popc(f);
        }
        if (((w + 4u) & 31u) == 0u || w + 4u >= n_words) {          // 32 words collected (or the slice ends): one coalesced store
            const uint32_t first = w & ~31u, mine = first + lane;
            if (mine < n_words) mask[mine] = myword;
            myword = 0;
        }
    }
    if (__ballot_sync(0xffffffffu, bad != 0u) && lane == 0) atomicOr(p.overflow, 1u);
    if (lane == 0) ws[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < 8; ++w) tot += ws[w];
        p.tile_count[blockIdx.x] = tot;
    }
}

// single-CTA exclusive scan of tile counts
__global__ void k_scan_tiles(const uint32_t *count, unsigned long long *offset, uint32_t n, unsigned long long *total) {
    __shared__ uint32_t ws[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n ? count[i] : 0u, e = v, d = 0, tot, td;
        block_scan2_sat(e, d, tot, td, ws);
        if (i < n) offset[i] = carry + e;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

struct FreqRecDev {                    // == mmc_freq_rec_t
    int32_t tid, pos;
    uint32_t n_called, n_mod;
    uint16_t ins_offset;
    int16_t hap;
    uint8_t strand, code;
    uint16_t reserved;
};

// Pass 2: one CTA (8 warps) per non-empty tile, one barrier per tile.  Every warp owns a contiguous eighth of the tile; it
// reads the ballot words pass 1 left (32 words = 1024 cells per load), the warp totals are prefix-summed through shared
// memory, and only the non-zero cells are fetched again and written as rows in cell order.
__global__ void __launch_bounds__(256) k_emit_records(const __grid_constant__ FinalizeParams p) {
    __shared__ uint32_t ws[8];
    if (p.tile_count[blockIdx.x] == 0u) return;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const FinJob j = fin_job_of(p, blockIdx.x);
    const unsigned long long c0 = ((unsigned long long)blockIdx.x - j.tile0) * p.cells_per_tile;
    unsigned long long c1 = c0 + p.cells_per_tile;
    if (c1 > j.n_cells) c1 = j.n_cells;
    const uint32_t per_warp = p.cells_per_tile / 8u;
    unsigned long long w0 = c0 + (unsigned long long)warp * per_warp, w1 = w0 + per_warp;
    if (w0 > c1) w0 = c1;
    if (w1 > c1) w1 = c1;
    const uint32_t *mask = p.mask + (unsigned long long)blockIdx.x * (p.cells_per_tile / 32u) + warp * (per_warp / 32u);
    const uint32_t n_words = (uint32_t)((w1 - w0 + 31u) >> 5);
    // sparse slice (CpG tables: less than one row per ballot word): every lane owns kWpl consecutive ballot words of the warp's
    // slice (per_warp / 32 / 32 = 4 with 32 K-cell tiles), its rows start at the exclusive prefix of the lanes' (and, through
    // shared memory, the warps') non-zero counts, and it walks the set bits of its own words -- no warp-wide loop, all lanes busy
    constexpr uint32_t kWplMax = 8;
    const uint32_t wpl = (per_warp / 32u + 31u) / 32u;
    uint32_t m[kWplMax], cnt = 0;
#pragma unroll
    for (uint32_t k = 0; k < kWplMax; ++k) {
        const uint32_t wi = lane * wpl + k;
        m[k] = (k < wpl && wi < n_words) ? mask[wi] : 0u;
        cnt += (uint32_t)__popc(m[k]);
    }
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += x; }
    if (lane == 31u) ws[warp] = incl;
    __syncthreads();
    const uint32_t wtot = ws[warp];
    if (wtot == 0u) return;
    uint32_t wbase = 0;
    for (uint32_t w = 0; w < warp; ++w) wbase += ws[w];
    FreqRecDev *out = reinterpret_cast<FreqRecDev *>(p.out) + p.tile_offset[blockIdx.x];
    const uint32_t spp = 2u * (uint32_t)p.n_code_slots * (uint32_t)p.n_hap_slots;   // slots per position
    const unsigned long long pos0 = w0 / spp;                                      // one 64-bit division per warp; the rest is 32-bit
    const uint32_t rem0 = (uint32_t)(w0 - pos0 * spp);
    if (wtot > 3u * n_words) {
        // dense slice (config 4: eight rows per ballot word): the warp takes one word at a time, a lane per cell, so that
        // the 32 rows of a step are consecutive in memory (coalesced stores); the lane-per-word walk below would scatter them
        const uint32_t lt = (1u << lane) - 1u;
        uint32_t running = wbase;
        for (uint32_t wb = 0; wb < n_words; wb += 32u) {
            const uint32_t mword = wb + lane < n_words ? mask[wb + lane] : 0u;
            uint32_t todo = __ballot_sync(0xffffffffu, mword != 0u);
            while (todo) {
                const uint32_t wi = (uint32_t)__ffs((int)todo) - 1u;
                todo &= todo - 1u;
                const uint32_t f = __shfl_sync(0xffffffffu, mword, (int)wi);
                if ((f >> lane) & 1u) {
                    const uint32_t local = (wb + wi) * 32u + lane, t = rem0 + local;
                    const unsigned long long v = j.cells[w0 + local];
                    const uint32_t dpos = t / spp;
                    uint32_t slot = t - dpos * spp;
                    const uint32_t hslot = slot % (uint32_t)p.n_hap_slots; slot /= (uint32_t)p.n_hap_slots;
                    const uint32_t code = slot % (uint32_t)p.n_code_slots; slot /= (uint32_t)p.n_code_slots;
                    FreqRecDev rec;
                    rec.tid = j.tid; rec.pos = j.lo + (int32_t)(pos0 + dpos);
                    rec.n_called = (uint32_t)v; rec.n_mod = (uint32_t)(v >> 32);
                    rec.ins_offset = 0;
                    rec.hap = p.haplotypes ? (int16_t)((int32_t)hslot - 1) : (int16_t)-1;
                    rec.strand = (uint8_t)slot; rec.code = (uint8_t)code; rec.reserved = 0;
                    out[running + (uint32_t)__popc(f & lt)] = rec;
                }
                running += (uint32_t)__popc(f);
            }
        }
        return;
    }
    if (cnt == 0u) return;
    uint32_t at = wbase + incl - cnt;
#pragma unroll
    for (uint32_t k = 0; k < kWplMax; ++k) {
        uint32_t f = m[k];
        const uint32_t wi = lane * wpl + k;
        while (f) {
            const uint32_t b = (uint32_t)__ffs((int)f) - 1u;
            f &= f - 1u;
            const uint32_t local = wi * 32u + b, t = rem0 + local;
            const unsigned long long v = j.cells[w0 + local];
            const uint32_t dpos = t / spp;
            uint32_t slot = t - dpos * spp;
            const uint32_t hslot = slot % (uint32_t)p.n_hap_slots; slot /= (uint32_t)p.n_hap_slots;
            const uint32_t code = slot % (uint32_t)p.n_code_slots; slot /= (uint32_t)p.n_code_slots;
            FreqRecDev rec;
            rec.tid = j.tid; rec.pos = j.lo + (int32_t)(pos0 + dpos);
            rec.n_called = (uint32_t)v; rec.n_mod = (uint32_t)(v >> 32);
            rec.ins_offset = 0;
            rec.hap = p.haplotypes ? (int16_t)((int32_t)hslot - 1) : (int16_t)-1;
            rec.strand = (uint8_t)slot; rec.code = (uint8_t)code; rec.reserved = 0;
            out[at++] = rec;
        }
    }
}

}  // namespace mmc

#endif  // MMC_DEVICE_CUH
