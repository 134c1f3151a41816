// mmc_api.cu -- host orchestration behind the C ABI of include/minimod_cuda.h.
//
// One context == one CUDA device.  Batches move through `n_slots` staging slots, each with
// pinned host arrays, an HBM mirror and its own stream, so the H2D copy of batch i+1
// overlaps the kernels of batch i (this is what replaces the reference's
// load || process || merge pthread pipeline, src/freq_main.c:404-474, and the per-batch
// worker pool of src/thread.c:100-158).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <chrono>
#ifndef MMC_EMUL
#include <dlfcn.h>
#include <strings.h>
#include <unistd.h>
#endif
#include <vector>

#include "../../include/minimod_cuda.h"
#include "mmc_device.cuh"
#include "mmc_decode_warp.cuh"
#include "mmc_decode_flat.cuh"
#include "mmc_decode_stream.cuh"
#include "mmc_sparse.cuh"
#ifndef MMC_EMUL
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#endif

using namespace mmc;

static_assert(sizeof(FreqRecDev) == sizeof(mmc_freq_rec_t), "record layout");
static_assert(sizeof(FreqRecDev) == 24, "record size");

namespace {

constexpr size_t kSlack = 64;                 // readable bytes after the last slice of a pool
constexpr uint32_t kTileCells = 32768;        // cells per finalize tile (256 KB: one CTA; the tile scan is a single CTA, so tiles are large)
constexpr uint32_t kExcCap = 1u << 22;        // exception-run starts per contig
static_assert(kTileCells % (8 * 32 * 32) == 0 && kTileCells / (8 * 32 * 32) <= 8, "k_emit_records: a lane owns at most 8 ballot words of its warp's slice");

std::string g_create_error;

struct Slot {
    mmc_batch_t pub;
    // host (pinned) and device arenas share one layout; offsets in bytes
    uint8_t *h_arena = nullptr, *d_arena = nullptr;
    size_t arena_bytes = 0;
    size_t o_tid, o_pos, o_lseq, o_ncig, o_mmlen, o_mllen, o_cigoff, o_seqoff, o_mmoff, o_mloff, o_flag, o_hp;
    size_t o_cigar, o_seq, o_mm, o_ml;
    size_t o_seq2 = 0, o_exc = 0, cap_exc = 0;   // seq_packing == 2: transport form of SEQ; o_seq is then device-only (last in the arena)
    size_t o_cig8 = 0, o_cig8off = 0;            // cigar_packing == 8: byte form of the CIGARs; o_cigar is then device-only
    size_t h_arena_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_h0 = nullptr, ev_h1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
    // small device state: [0] err (u64), [1] view_n (u64), then u32: [4] work counter of k_decode_warp,
    // [5] number of deferred reads, [6] work counter of k_decode
    unsigned long long *d_state = nullptr;
    uint32_t *d_defer = nullptr;               // reads k_decode_warp left to k_decode
    uint32_t *d_defer_flat = nullptr;          // reads the flat kernels left to k_decode_warp
    WRead *d_reads = nullptr;                  // flat path: per-read state
    uint32_t *d_pool = nullptr; size_t pool_words = 0;      // flat path: scratch pool
    unsigned long long *h_state = nullptr;     // pinned mirror
    ViewDev *d_view = nullptr; uint64_t view_cap = 0;    // VIEW: rows of one batch; regrown (and the batch re-run) when it overflows
    uint32_t *d_scratch = nullptr;
    size_t scratch_words = 0;
    std::vector<mmc_view_rec_t> view_out;
    bool in_flight = false, uploaded = false, acquired = false, timed = false, h2d_pending = false, view_regrown = false;
    uint32_t n_reads_submitted = 0;
    uint32_t max_cig = 0, max_l = 0; uint64_t pool_need = 0; int variant = 3;   // analyse_batch()
    uint32_t min_tid = 0xffffffffu, min_pos = 0;   // first read of the batch in coordinate order (analyse_batch; unmapped sorts last)
    bool use_stream = false;                   // this batch goes through k_decode_stream (long CIGARs / long reads) instead of k_decode_warp<PRE>
    uint32_t s_split = 0;                      // k_decode_stream: (read, even / odd blocks) work units (analyse_batch)
    int s_ctas = 8; uint32_t s_arena = 0, s_setup_flex = 4608;   // k_decode_stream: CTAs per SM, arena bytes per warp; k_flat_setup's room for dir | cq | cr
};

struct ContigHost {
    std::string name;
    uint32_t len = 0;
    ContigDev dev{};
    uint32_t *d_exc_start = nullptr;
    uint8_t *d_exc_letter = nullptr;
    bool loaded = false;
};

}  // namespace

struct mmc_ctx {
    mmc_opts_t opts{};
    std::vector<mmc_mod_t> mods;
    std::vector<ContigHost> contigs;
    std::vector<Slot> slots;
    std::string err, desc;
    int sm_count = 0, ctas_per_sm = 1, threads = 128;
    int seq_packing = 4;                                             // opts.seq_packing, or MMC_SEQ_PACKING
    int cigar_packing = 32;                                          // opts.cigar_packing, or MMC_CIGAR_PACKING
    int stream_path = 1;                       // k_flat_setup + k_decode_stream (default): streaming merge, constant shared memory per warp
    uint32_t s_head = 256;                     // k_decode_stream: bytes of call LUTs in front of the arenas
    int s_minb = 6;                            // k_decode_stream<MINB>: resident CTAs per SM the registers are bounded for (MMC_STREAM_MINB)
    int s_split_mode = 0;                      // MMC_STREAM_SPLIT: 1 = (read, even / odd blocks) work units, -1 = per batch (long reads with
                                               // several MM blocks and haplotype strata).  Off by default: measured neutral
    int split_path = 1;                        // k_flat_setup + k_decode_warp<PRE>: setup split from the fused kernel
    int warp_path = 1;                         // then k_decode_warp, then k_decode for what that defers
    // k_decode_warp<MINB>: variants bounded for MINB resident CTAs per SM; the arena of a warp shrinks as MINB grows.
    // [0] unused; chosen per batch from the reads' sizes unless MMC_WARP_OCC pins one.
    uint32_t wv_arena[5] = {0, 28672u, 14208u, 9344u, 6912u};   // (228 KB / MINB - 1 KB - kWHeadBytes) / 8 warps
    uint32_t wv_setup_arena[5] = {0, 0, 0, 0, 0};               // k_flat_setup: WRead + room for dir | cq | cr
    int wv_ctas[5] = {0, 1, 1, 1, 1};
    int w_pinned = 0;                          // MMC_WARP_OCC / MMC_WARP_ARENA given: no per-batch choice
    int w_minb = 3;                            // default variant (3 CTAs/SM, 80 registers: best on 15 kb reads)
    int n_code_slots = 1, n_hap_slots = 1, wild_req = -1;
    ReqMod *d_req = nullptr;
    unsigned long long *d_code_keys = nullptr;
    ContigDev *d_contigs = nullptr;
    int32_t *d_touch = nullptr;                // [2*n_contigs]: lo then hi
    SparseRec *d_sparse = nullptr;
    unsigned long long *d_sparse_n = nullptr;
    uint64_t sparse_cap = 0, view_cap = 0;
    uint64_t sparse_seen = 0;                  // highest fill level of the side buffer any finished batch reported
    bool committed = false;
    cudaStream_t fin_stream = nullptr;
    cudaEvent_t ev_f0 = nullptr, ev_f1 = nullptr;
    // ref-pack staging
    uint8_t *d_ascii = nullptr; size_t ascii_cap = 0;
    uint32_t *d_exc_tmp_start = nullptr; uint8_t *d_exc_tmp_letter = nullptr; uint32_t *d_exc_n = nullptr;
    // finalize scratch, grown on demand and kept: tile counts / offsets / totals, device rows, pinned host rows
    uint32_t *d_tile_count = nullptr; unsigned long long *d_tile_off = nullptr, *d_totals = nullptr;
    size_t fin_tiles_cap = 0, fin_jobs_cap = 0;
    uint32_t *d_fin_mask = nullptr;                                  // one bit per scanned cell (+ 1 word: n_called overflow flag)
    FinJob *d_fin_jobs = nullptr;                                    // the contig ranges of one finalize / drain (fin_jobs_cap entries)
    FreqRecDev *d_rows = nullptr; size_t d_rows_cap = 0;
    // device-side finalize of the sparse side buffer (mmc_sparse.cuh): grow-only scratch
    uint64_t sparse_dev_min = 1u << 16;                              // fewer records than this: host sort (a few ms at most)
    void *d_sp_scratch = nullptr; size_t sp_scratch_cap = 0;         // keys, indexes, flags, offsets, sort temp
    FreqRecDev *d_srows = nullptr; size_t d_srows_cap = 0;           // reduced sparse rows
    FreqRecDev *d_merged = nullptr; size_t d_merged_cap = 0;         // dense + sparse rows in output order
    unsigned long long *d_sn_rows = nullptr, *h_sn_rows = nullptr;   // number of reduced sparse rows (device, pinned)
    mmc_freq_rec_t *h_rows = nullptr; size_t h_rows_cap = 0;       // pinned
    // mmc_freq_drain(): rows before a coordinate watermark leave while later batches are still being copied and decoded
    std::vector<uint32_t> drained_to;                              // per contig: positions below this were returned by a drain
    uint32_t wm_tid = 0, wm_pos = 0; bool wm_set = false;          // watermark of the last drain that returned rows
    bool drain_violated = false;                                   // a batch uploaded after a drain starts before its watermark
    unsigned long long sparse_lo = 0;                              // side-buffer records with a major key below this were returned by a drain
    double host_upload_ms = 0, host_launch_ms = 0; bool trace_host = false;   // MMC_TRACE_CREATE: host time inside mmc_batch_submit
    double host_sect_ms[4] = {0, 0, 0, 0};                         // launch_decode: side-buffer reserve, state reset + general-kernel scratch, pool, launches
    cudaEvent_t ev_reset = nullptr; bool reset_pending = false;    // mmc_freq_reset() clears on fin_stream; the next decode launches wait for it on the device
    std::vector<int32_t> reset_touch;                              // (source of its asynchronous copy)
    mmc_freq_rec_t *h_drain[2] = {nullptr, nullptr}; size_t h_drain_cap[2] = {0, 0}; int drain_flip = 0;   // pinned, alternating
    unsigned long long *h_totals = nullptr;                        // pinned: [0] rows of the pass, [1] overflow flag
    cudaEvent_t ev_d0 = nullptr, ev_d1 = nullptr;
    // results
    std::vector<uint32_t> need_tmp;
    std::vector<std::string> code_names;
    mmc_timers_t tm{};
    int32_t cig_smem_cap = kCigSmem, bitmap_smem_words = kBitmapWords, idx_smem_cap = kIdxSmem;
};

namespace {

int fail(mmc_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

// several contexts (one per device) may live in one process: every entry point makes its context's device current
#define MMC_DEV(ctx) do { if (ctx) cudaSetDevice((ctx)->opts.device); } while (0)

#define CU(ctx, call)                                                                           \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(ctx, MMC_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                    \
    } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

unsigned long long pack_code(const char *s) {
    unsigned long long k = 0;
    for (int i = 0; i < 8 && s[i]; ++i) k |= (unsigned long long)(uint8_t)s[i] << (8 * i);
    return k;
}

std::string unpack_code(unsigned long long k) {
    std::string s;
    for (int i = 0; i < 8; ++i) {
        char c = (char)((k >> (8 * i)) & 0xff);
        if (!c) break;
        s.push_back(c);
    }
    return s;
}

const char *read_error_text(uint32_t code) {
    switch (code) {
    case kErrHardClip: return "Hard clipping found and they are not supported.\nTry following workarounds.\n\t01. Filter out non-primary alignments\n\t\tsamtools view -h -F 2308 reads.bam -o primary_reads.bam\n\t02. Use minimap2 with -Y to use soft clipping for suplimentary alignments.";
    case kErrCigarOp: return "Unhandled CIGAR OPT";
    case kErrCigarLen: return "Assertion failed: read_pos < seq_len (CIGAR consumes more bases than the read has)";
    case kErrRefRange: return "Assertion failed: ref_pos >= 0 && ref_pos < ref_len (alignment runs past the contig)";
    case kErrNoContig: return "Contig not found in reference provided";
    case kErrMMBase: return "Invalid base in MM tag";
    case kErrMMStrand: return "Invalid strand in MM tag";
    case kErrMMCode: return "Invalid base modification code in MM tag. Modification codes should be either numeric or alphabetic, not both, and cannot be empty.";
    case kErrMMSkip: return "Invalid skip count in MM tag";
    case kErrMMRank: return "Read pos cannot exceed seq len (MM tag skips more canonical bases than the read has)";
    case kErrMLIndex: return "mod prob index mismatch. ml_idx >= ml_len";
    case kErrTooManyCodes: return "too many modification codes (library limit: 8 per MM block, 8 characters per code, 256 distinct codes)";
    case kErrSeqTooLong: return "read longer than 2^28 bases (library limit)";
    default: return "unknown per-read error";
    }
}

// host-side reads and writes of the count state (touch ranges, side-buffer fill level, cells) come after a pending reset
int reset_settle(mmc_ctx *ctx) {
    if (!ctx->reset_pending) return MMC_OK;
    CU(ctx, cudaStreamSynchronize(ctx->fin_stream));
    ctx->reset_pending = false;
    return MMC_OK;
}

int setup_slot(mmc_ctx *ctx, Slot &s) {
    const mmc_opts_t &o = ctx->opts;
    const size_t R = o.max_reads;
    // pool capacities (bytes).  Any pool filling up ends the batch early, which never changes results.
    size_t cap_seq = align_up(o.max_bytes / 2 + 4096, 64), cap_cig = align_up(o.max_bytes / 2 + 4096, 64);
    size_t cap_mm = align_up(o.max_bytes / 2 + 4096, 64), cap_ml = align_up(o.max_bytes / 4 + 4096, 64);
    if (o.cap_cigar_words) cap_cig = align_up(o.cap_cigar_words * 4, 64);
    if (o.cap_seq_bytes) cap_seq = align_up(o.cap_seq_bytes, 64);
    if (o.cap_mm_bytes) cap_mm = align_up(o.cap_mm_bytes, 64);
    if (o.cap_ml_bytes) cap_ml = align_up(o.cap_ml_bytes, 64);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t at = off; off = align_up(off + bytes, 256); return at; };
    s.o_tid = take(R * 4); s.o_pos = take(R * 4); s.o_lseq = take(R * 4); s.o_ncig = take(R * 4);
    s.o_mmlen = take(R * 4); s.o_mllen = take(R * 4);
    s.o_cigoff = take(R * 8); s.o_seqoff = take(R * 8); s.o_mmoff = take(R * 8); s.o_mloff = take(R * 8);
    s.o_flag = take(R * 2); s.o_hp = take(R);
    const bool two_bit = ctx->seq_packing == 2, cig_bytes = ctx->cigar_packing == 8;
    if (!cig_bytes) s.o_cigar = take(cap_cig + kSlack);
    if (!two_bit) s.o_seq = take(cap_seq + kSlack);
    s.o_mm = take(cap_mm + kSlack); s.o_ml = take(cap_ml + kSlack);
    s.h_arena_bytes = off;
    if (two_bit) {                                  // 2 bits per base + exceptions cross PCIe; the 4-bit pool exists on the device only
        s.cap_exc = R + cap_seq / 64 + 1024;
        s.o_seq2 = take(cap_seq / 2 + kSlack); s.o_exc = take(8 * s.cap_exc);
        s.h_arena_bytes = off;
    }
    if (cig_bytes) {                                // a byte per op + escapes cross PCIe; the word pool exists on the device only
        s.o_cig8off = take(R * 8); s.o_cig8 = take(cap_cig + kSlack);
        s.h_arena_bytes = off;
    }
    if (two_bit) s.o_seq = take(cap_seq + kSlack);
    if (cig_bytes) s.o_cigar = take(cap_cig + kSlack);
    s.arena_bytes = off;
    CU(ctx, cudaMallocHost((void **)&s.h_arena, s.h_arena_bytes));
    CU(ctx, cudaMalloc((void **)&s.d_arena, s.arena_bytes));
    CU(ctx, cudaMemset(s.d_arena, 0, s.arena_bytes));
    CU(ctx, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CU(ctx, cudaEventCreate(&s.ev_h0)); CU(ctx, cudaEventCreate(&s.ev_h1));
    CU(ctx, cudaEventCreate(&s.ev_k0)); CU(ctx, cudaEventCreate(&s.ev_k1));
    CU(ctx, cudaMalloc((void **)&s.d_state, 128));
    CU(ctx, cudaMallocHost((void **)&s.h_state, 128));
    CU(ctx, cudaMalloc((void **)&s.d_defer_flat, sizeof(uint32_t) * std::max<size_t>(1, R)));
    if (ctx->split_path || ctx->stream_path) CU(ctx, cudaMalloc((void **)&s.d_reads, sizeof(WRead) * std::max<size_t>(1, R)));
    CU(ctx, cudaMalloc((void **)&s.d_defer, sizeof(uint32_t) * std::max<size_t>(1, R)));
    if (o.subtool == MMC_VIEW) { CU(ctx, cudaMalloc((void **)&s.d_view, ctx->view_cap * sizeof(ViewDev))); s.view_cap = ctx->view_cap; }
    mmc_batch_t &b = s.pub;
    memset(&b, 0, sizeof(b));
    b.max_reads = (uint32_t)R;
    uint8_t *h = s.h_arena;
    b.tid = (int32_t *)(h + s.o_tid); b.pos = (int32_t *)(h + s.o_pos);
    b.l_seq = (uint32_t *)(h + s.o_lseq); b.n_cigar = (uint32_t *)(h + s.o_ncig);
    b.mm_len = (uint32_t *)(h + s.o_mmlen); b.ml_len = (uint32_t *)(h + s.o_mllen);
    b.cigar_off = (uint64_t *)(h + s.o_cigoff); b.seq_off = (uint64_t *)(h + s.o_seqoff);
    b.mm_off = (uint64_t *)(h + s.o_mmoff); b.ml_off = (uint64_t *)(h + s.o_mloff);
    b.flag = (uint16_t *)(h + s.o_flag); b.hp = h + s.o_hp;
    b.cigar = cig_bytes ? nullptr : (uint32_t *)(h + s.o_cigar); b.cigar_cap = cap_cig / 4;
    b.cigar_packing = cig_bytes ? 8u : 32u;
    if (cig_bytes) { b.cig8 = h + s.o_cig8; b.cig8_cap = cap_cig; b.cig8_off = (uint64_t *)(h + s.o_cig8off); }
    b.seq4 = two_bit ? nullptr : h + s.o_seq; b.seq_cap = cap_seq;
    b.seq_packing = two_bit ? 2u : 4u;
    if (two_bit) { b.seq2 = h + s.o_seq2; b.seq_exc = (uint64_t *)(h + s.o_exc); b.seq_exc_cap = s.cap_exc; }
    b.mm = (char *)(h + s.o_mm); b.mm_cap = cap_mm;
    b.ml = h + s.o_ml; b.ml_cap = cap_ml;
    b.priv = &s;
    return MMC_OK;
}

Slot *slot_of(mmc_ctx *ctx, mmc_batch_t *b) {
    if (!b) return nullptr;
    for (Slot &s : ctx->slots) if (&s.pub == b) return &s;
    return nullptr;
}

// Shape of the batch a slot holds: sizes the scratch and picks the k_decode_warp variant.  Runs at upload time so
// that re-launching an HBM-resident batch costs no host work.
void analyse_batch(mmc_ctx *ctx, Slot &s) {
    const mmc_batch_t &b = s.pub;
    const uint32_t n = b.n_reads;
    uint32_t max_cig = 0, max_l = 0;
    uint64_t pool_need = 0;                    // split path: words of scratch for every read's dir | cq | cr
    std::vector<uint32_t> &need = ctx->need_tmp;   // per read: arena words for un-sampled CIGAR arrays + rank index
    need.resize(n);
    uint64_t first = ~0ull;                    // (tid, pos) as one key; tid < 0 sorts last
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t L = b.l_seq[i], nc = b.n_cigar[i];
        first = std::min(first, ((uint64_t)(uint32_t)b.tid[i] << 32) | (uint32_t)std::max<int32_t>(0, b.pos[i]));
        max_cig = std::max(max_cig, nc); max_l = std::max(max_l, L);
        const uint64_t n_u4 = ((uint64_t)L + 31) >> 5;
        pool_need += 164 + 2ull * nc + 8;
        need[i] = (uint32_t)std::min<uint64_t>(0xffffffffu, (L >> 8) + 2 + nc / 2 + n_u4 + 2 + (L >> 6) + 2 + 12);   // CIGAR sampled 1:4 at worst
    }
    // variant for this batch: the most CTAs per SM whose arena holds ~95% of the reads with an un-sampled index
    int mb = ctx->w_minb;
    if (!ctx->w_pinned && n > 0) {
        const size_t k = (size_t)((uint64_t)(n - 1) * 95 / 100);
        std::nth_element(need.begin(), need.begin() + k, need.end());
        const uint32_t p95 = need[k];
        mb = 3;
        while (mb > 1 && (ctx->wv_arena[mb] - (uint32_t)sizeof(WFixed)) / 4u < p95) --mb;
    }
    s.max_cig = max_cig; s.max_l = max_l; s.pool_need = pool_need; s.variant = mb;
    s.min_tid = (uint32_t)(first >> 32); s.min_pos = (uint32_t)first;
    if (n && ctx->wm_set && first < (((uint64_t)ctx->wm_tid << 32) | ctx->wm_pos)) ctx->drain_violated = true;   // see mmc_freq_drain()
    // Two implementations of the stage: k_decode_warp<MINB,PRE> keeps a read's rank index and CIGAR arrays in its arena -- fewest
    // instructions while three CTAs fit an SM (HiFi: 15 kb reads, ~30 CIGAR ops) -- and k_decode_stream needs constant shared
    // memory per warp whatever the read (ONT CIGARs, 50 kb reads: 24 warps per SM where the former drops to 16 or 8).
    s.use_stream = ctx->stream_path == 2 || (ctx->stream_path == 1 && mb < 3);
    if (s.use_stream) {
        // k_decode_stream: constant arena per warp (SFixed + room for the dir | cq | cr of short CIGARs; longer ones are
        // looked up in the pool k_flat_setup wrote).  k_flat_setup builds the arrays in its own arena: sized for ~95 % of
        // the reads un-sampled (the rest get every 2nd / 4th ... op, w_setup_read).
        uint32_t p95 = 64;
        if (n > 0) {
            for (uint32_t i = 0; i < n; ++i) need[i] = std::min<uint32_t>(160u, (b.l_seq[i] >> 8) + 2u) + 2u * b.n_cigar[i] + 8u;
            const size_t k = (size_t)((uint64_t)(n - 1) * 95 / 100);
            std::nth_element(need.begin(), need.begin() + k, need.end());
            p95 = need[k];
        }
        (void)p95;
        s.s_ctas = ctx->s_minb;
        // split the blocks of a read over two warps?  Only where the window of count cells the reads in flight cover outgrows
        // the L2: long reads with several strata per cell -- and only if the reads have more than one block (a look at the
        // MM text of three reads; a wrong guess costs a no-op work unit per read, never a result).
        s.s_split = 0;
        if (ctx->s_split_mode >= 0) s.s_split = (uint32_t)ctx->s_split_mode;
        else if (n >= 3 && max_l >= 32768u && ctx->opts.haplotypes) {
            uint32_t multi = 0;
            const uint32_t pick[3] = {0u, n / 2u, n - 1u};
            for (uint32_t k = 0; k < 3; ++k) {
                const char *t = b.mm + b.mm_off[pick[k]];
                const uint32_t len = b.mm_len[pick[k]];
                const void *first = len ? memchr(t, ';', len) : nullptr;
                if (first && (const char *)first + 1 < t + len) ++multi;            // text after the first block's ';'
            }
            s.s_split = multi >= 2 ? 1u : 0u;
        }
        s.s_arena = (uint32_t)((sizeof(SFixed) + 15) & ~(size_t)15);
        s.s_setup_flex = 256;                      // stream mode: k_flat_setup writes the CIGAR table straight into the pool
        uint64_t need_s = 0;
        for (uint32_t i = 0; i < n; ++i) need_s += (((uint64_t)(b.l_seq[i] >> 5) + 2u + 3u) & ~3ull) + 2ull * std::max<uint32_t>(1u, b.n_cigar[i]) + 4u;
        s.pool_need = std::max<uint64_t>(s.pool_need, need_s);
    }
}

int upload(mmc_ctx *ctx, Slot &s) {
    const mmc_batch_t &b = s.pub;
    const size_t n = b.n_reads;
    if (n > b.max_reads || b.cigar_used > b.cigar_cap || b.seq_used > b.seq_cap || b.mm_used > b.mm_cap || b.ml_used > b.ml_cap ||
        b.seq_exc_used > b.seq_exc_cap || b.cig8_used > b.cig8_cap)
        return fail(ctx, MMC_EINVAL, "batch exceeds its capacities");
    const bool two_bit = ctx->seq_packing == 2, cig_bytes = ctx->cigar_packing == 8;
    uint64_t bytes = 0;
    CU(ctx, cudaEventRecord(s.ev_h0, s.stream));
    auto cp = [&](size_t off, size_t len) -> cudaError_t {
        if (!len) return cudaSuccess;
        bytes += len;
        return cudaMemcpyAsync(s.d_arena + off, s.h_arena + off, len, cudaMemcpyHostToDevice, s.stream);
    };
    CU(ctx, cp(s.o_tid, n * 4)); CU(ctx, cp(s.o_pos, n * 4)); CU(ctx, cp(s.o_lseq, n * 4)); CU(ctx, cp(s.o_ncig, n * 4));
    CU(ctx, cp(s.o_mmlen, n * 4)); CU(ctx, cp(s.o_mllen, n * 4));
    CU(ctx, cp(s.o_cigoff, n * 8)); CU(ctx, cp(s.o_seqoff, n * 8)); CU(ctx, cp(s.o_mmoff, n * 8)); CU(ctx, cp(s.o_mloff, n * 8));
    CU(ctx, cp(s.o_flag, n * 2)); CU(ctx, cp(s.o_hp, n));
    if (!cig_bytes) {
        CU(ctx, cp(s.o_cigar, b.cigar_used * 4));
    } else if (n) {                                             // transport form -> the word pool the kernels read
        CU(ctx, cp(s.o_cig8off, n * 8)); CU(ctx, cp(s.o_cig8, b.cig8_used));
        const unsigned grid = (unsigned)std::min<uint64_t>((n + 7) / 8, (uint64_t)ctx->sm_count * 8);
        MMC_LAUNCH(k_unpack_cigar, grid, 256u, s.stream, (const uint8_t *)(s.d_arena + s.o_cig8), (const unsigned long long *)(s.d_arena + s.o_cig8off),
                   (const uint32_t *)(s.d_arena + s.o_ncig), (const unsigned long long *)(s.d_arena + s.o_cigoff), (uint32_t)n, (uint32_t *)(s.d_arena + s.o_cigar));
        CU(ctx, cudaGetLastError());
        ctx->tm.kernel_launches += 1;
    }
    CU(ctx, cp(s.o_mm, b.mm_used)); CU(ctx, cp(s.o_ml, b.ml_used));
    if (!two_bit) {
        CU(ctx, cp(s.o_seq, b.seq_used));
    } else if (b.seq_used) {                                    // transport form -> the 4-bit pool the kernels read
        CU(ctx, cp(s.o_seq2, (b.seq_used + 1) / 2)); CU(ctx, cp(s.o_exc, 8 * b.seq_exc_used));
        const uint64_t n8 = (b.seq_used + 15) / 16;
        const unsigned grid = (unsigned)std::min<uint64_t>((n8 + 255) / 256, (uint64_t)ctx->sm_count * 8);
        MMC_LAUNCH(k_unpack_seq2, grid, 256u, s.stream, (const uint2 *)(s.d_arena + s.o_seq2), (uint4 *)(s.d_arena + s.o_seq), (unsigned long long)n8);
        CU(ctx, cudaGetLastError());
        ctx->tm.kernel_launches += 1;
        if (b.seq_exc_used) {
            const unsigned g2 = (unsigned)std::min<uint64_t>((b.seq_exc_used + 255) / 256, (uint64_t)ctx->sm_count * 8);
            MMC_LAUNCH(k_patch_seq4, g2, 256u, s.stream, (const unsigned long long *)(s.d_arena + s.o_exc), (unsigned long long)b.seq_exc_used,
                       (uint32_t *)(s.d_arena + s.o_seq));
            CU(ctx, cudaGetLastError());
            ctx->tm.kernel_launches += 1;
        }
    }
    CU(ctx, cudaEventRecord(s.ev_h1, s.stream));
    ctx->tm.h2d_bytes += bytes;
    s.uploaded = true; s.h2d_pending = true;
    s.n_reads_submitted = (uint32_t)n;
    analyse_batch(ctx, s);
    return MMC_OK;
}

int wait_slot(mmc_ctx *ctx, Slot &s);

// The side buffer holds the whole run's sparse records (the reference's hash map has no limit either): before a batch is
// launched, make sure every batch that can be in flight still fits behind the highest fill level seen so far, and grow the
// buffer (all slots drained, device-to-device copy) when it does not.  A single batch that appends more than its reserve
// is caught in wait_slot().
int reserve_sparse(mmc_ctx *ctx, const Slot &s) {
    if (ctx->opts.subtool != MMC_FREQ) return MMC_OK;
    if (!ctx->opts.insertions && !ctx->opts.haplotypes && ctx->wild_req < 0) return MMC_OK;   // every cell is dense
    const mmc_batch_t &b = s.pub;
    // appends of one batch: <= 2 records per explicit call (haplotype stratum + '*') + implicit calls inside insertions
    const uint64_t per_batch = 4 * b.ml_used + b.seq_used / 4 + (1u << 16);
    const uint64_t need = ctx->sparse_seen + per_batch * (uint64_t)ctx->slots.size();
    if (need <= ctx->sparse_cap) return MMC_OK;
    for (Slot &o : ctx->slots) { int rc = wait_slot(ctx, o); if (rc != MMC_OK) return rc; }
    const uint64_t again = ctx->sparse_seen + per_batch * (uint64_t)ctx->slots.size();
    if (again <= ctx->sparse_cap) return MMC_OK;
    uint64_t cap = ctx->sparse_cap;
    while (cap < again) cap += cap / 2 + (1u << 20);
    SparseRec *nb = nullptr;
    if (cudaMalloc((void **)&nb, sizeof(SparseRec) * cap) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, MMC_ENOMEM, "cannot grow the sparse count buffer to %llu records (%.1f GB)", (unsigned long long)cap, cap * 16 / 1e9);
    }
    unsigned long long sn = 0;
    { int rc = reset_settle(ctx); if (rc != MMC_OK) return rc; }
    CU(ctx, cudaMemcpy(&sn, ctx->d_sparse_n, 8, cudaMemcpyDeviceToHost));
    if (sn > ctx->sparse_cap) sn = ctx->sparse_cap;
    if (sn) CU(ctx, cudaMemcpy(nb, ctx->d_sparse, sizeof(SparseRec) * sn, cudaMemcpyDeviceToDevice));
    CU(ctx, cudaMemset(nb + sn, 0xff, sizeof(SparseRec) * (cap - sn)));
    CU(ctx, cudaFree(ctx->d_sparse));
    ctx->d_sparse = nb; ctx->sparse_cap = cap;
    return MMC_OK;
}

int launch_decode(mmc_ctx *ctx, Slot &s) {
    const mmc_batch_t &b = s.pub;
    const uint32_t n = s.n_reads_submitted;
    auto tick = [&](int k, std::chrono::steady_clock::time_point &t) {            // MMC_TRACE_CREATE: host time per section
        const auto now = std::chrono::steady_clock::now();
        ctx->host_sect_ms[k] += std::chrono::duration<double, std::milli>(now - t).count();
        t = now;
    };
    auto tsec = std::chrono::steady_clock::now();
    { int rc = reserve_sparse(ctx, s); if (rc != MMC_OK) return rc; }
    tick(0, tsec);
    // reset the slot's device state: err = ~0, view_n = 0, work counters and deferred count = 0
    // layout: u64 [0] err, [1] view_n, [4] pool cursor; u32 [4] work counter of k_decode_warp, [5] reads it defers,
    // [6] work counter of k_decode, [7] reads the flat path defers, [10] tiles, [11] reads with '.' blocks
    if (ctx->reset_pending) CU(ctx, cudaStreamWaitEvent(s.stream, ctx->ev_reset, 0));   // the counts are being cleared: copies may overlap that, kernels may not
    memset(s.h_state, 0, 64);
    s.h_state[0] = ~0ull;
    CU(ctx, cudaMemcpyAsync(s.d_state, s.h_state, 64, cudaMemcpyHostToDevice, s.stream));
    if (n == 0) { s.in_flight = true; s.timed = false; return MMC_OK; }

    // shape of the batch, computed once per upload (analyse_batch)
    const uint32_t max_cig = s.max_cig, max_l = s.max_l;
    const uint64_t pool_need = s.pool_need;
    const int mb = s.variant;
    const uint32_t w_arena_bytes = ctx->wv_arena[mb], setup_arena_bytes = ctx->wv_setup_arena[mb];
    unsigned grid = (unsigned)std::min<uint64_t>(n, (uint64_t)ctx->sm_count * ctx->ctas_per_sm);
    if (grid == 0) grid = 1;
    uint32_t cig_words = max_cig > (uint32_t)ctx->cig_smem_cap ? (uint32_t)align_up(max_cig, 32) : 0;
    uint32_t bm_words = ((max_l + 31u) >> 5) > (uint32_t)ctx->bitmap_smem_words ? (uint32_t)align_up((max_l + 31u) >> 5, 32) : 0;
    size_t per_cta = 2 * (size_t)cig_words + bm_words;
    if (per_cta) {
        size_t need = per_cta * grid;
        if (need > s.scratch_words) {
            CU(ctx, cudaStreamSynchronize(s.stream));
            if (s.d_scratch) CU(ctx, cudaFree(s.d_scratch));
            s.d_scratch = nullptr; s.scratch_words = 0;
            CU(ctx, cudaMalloc((void **)&s.d_scratch, need * 4));
            s.scratch_words = need;
        }
    }

    tick(1, tsec);
    DecodeParams P;
    memset(&P, 0, sizeof(P));
    uint8_t *d = s.d_arena;
    P.n_reads = n;
    P.tid = (const int32_t *)(d + s.o_tid); P.pos = (const int32_t *)(d + s.o_pos);
    P.l_seq = (const uint32_t *)(d + s.o_lseq); P.n_cigar = (const uint32_t *)(d + s.o_ncig);
    P.mm_len = (const uint32_t *)(d + s.o_mmlen); P.ml_len = (const uint32_t *)(d + s.o_mllen);
    P.cigar_off = (const unsigned long long *)(d + s.o_cigoff); P.seq_off = (const unsigned long long *)(d + s.o_seqoff);
    P.mm_off = (const unsigned long long *)(d + s.o_mmoff); P.ml_off = (const unsigned long long *)(d + s.o_mloff);
    P.flag = (const uint16_t *)(d + s.o_flag); P.hp = d + s.o_hp;
    P.cigar = (const uint32_t *)(d + s.o_cigar); P.seq4 = d + s.o_seq; P.mm = d + s.o_mm; P.ml = d + s.o_ml;
    P.req = ctx->d_req; P.n_req = ctx->opts.n_mods; P.wild_req = ctx->wild_req;
    P.insertions = ctx->opts.insertions; P.haplotypes = ctx->opts.haplotypes; P.subtool = ctx->opts.subtool;
    P.code_keys = ctx->d_code_keys;
    P.n_contigs = (int32_t)ctx->contigs.size(); P.contigs = ctx->d_contigs;
    P.n_code_slots = ctx->n_code_slots; P.n_hap_slots = ctx->n_hap_slots;
    P.touch_lo = ctx->d_touch; P.touch_hi = ctx->d_touch + ctx->contigs.size();
    P.sparse = ctx->d_sparse; P.sparse_cap = ctx->sparse_cap; P.sparse_n = ctx->d_sparse_n;
    P.view = s.d_view; P.view_cap = s.view_cap; P.view_n = s.d_state + 1;
    P.err = s.d_state;
    P.scratch = per_cta ? s.d_scratch : nullptr;
    P.scratch_words_per_cta = per_cta; P.scratch_cig_words = cig_words;
    uint32_t *st32 = (uint32_t *)s.d_state;
    P.work_counter = st32 + 4;
    P.cig_smem_cap = ctx->cig_smem_cap; P.bitmap_smem_words = ctx->bitmap_smem_words; P.idx_smem_cap = ctx->idx_smem_cap;

    FlatParams F;
    memset(&F, 0, sizeof(F));
    if (ctx->split_path || ctx->stream_path) {
        if (pool_need > s.pool_words) {
            CU(ctx, cudaStreamSynchronize(s.stream));
            if (s.d_pool) CU(ctx, cudaFree(s.d_pool));
            s.d_pool = nullptr; s.pool_words = 0;
            const uint64_t want = pool_need + pool_need / 4;   // headroom: batches differ by a few per cent, every regrowth is a device-wide synchronisation
            CU(ctx, cudaMalloc((void **)&s.d_pool, want * 4));
            s.pool_words = want;
        }
        F.reads = s.d_reads;
        F.fa.pool = s.d_pool; F.fa.cursor = s.d_state + 4; F.fa.cap = s.pool_words;
        F.defer_list = s.d_defer_flat; F.defer_n = st32 + 7;
    }

    tick(2, tsec);
    CU(ctx, cudaEventRecord(s.ev_k0, s.stream));
    if (s.use_stream) {
        // k_flat_setup prepares every read (state + CIGAR table in HBM), k_decode_stream merges the calls
        // against the SEQ stream; what k_flat_setup cannot prepare goes down the chain below
        const uint32_t setup_arena = kWReadBytes + s.s_setup_flex;
        F.arena_bytes = setup_arena; F.consumer_flex_words = 1u << 24; F.read_count = n; F.stream = 1;
        const unsigned rgrid = (unsigned)std::min<uint64_t>(((uint64_t)n + kFThreads / 32 - 1) / (kFThreads / 32), (uint64_t)ctx->sm_count * 16);
        MMC_LAUNCH_SMEM(k_flat_setup, rgrid, (unsigned)kFThreads, (size_t)kWHeadBytes + (size_t)setup_arena * (kFThreads / 32), s.stream, P, F);
        CU(ctx, cudaGetLastError());
        StreamParams SP; SP.arena_bytes = s.s_arena; SP.head_bytes = ctx->s_head; SP.split = s.s_split; SP.pad_ = 0;
        PreParams Q; Q.reads = s.d_reads; Q.n = n;
        const uint64_t units = (uint64_t)n * (s.s_split ? 2u : 1u);
        const unsigned sgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((units + kSThreads / 32 - 1) / (kSThreads / 32), (uint64_t)ctx->sm_count * s.s_ctas));
        const size_t ssmem = (size_t)ctx->s_head + (size_t)s.s_arena * (kSThreads / 32);
        if (s.s_ctas == 8) MMC_LAUNCH_SMEM((k_decode_stream<8>), sgrid, (unsigned)kSThreads, ssmem, s.stream, P, SP, Q);
        else if (s.s_ctas == 6) MMC_LAUNCH_SMEM((k_decode_stream<6>), sgrid, (unsigned)kSThreads, ssmem, s.stream, P, SP, Q);
        else if (s.s_ctas == 5) MMC_LAUNCH_SMEM((k_decode_stream<5>), sgrid, (unsigned)kSThreads, ssmem, s.stream, P, SP, Q);
        else MMC_LAUNCH_SMEM((k_decode_stream<4>), sgrid, (unsigned)kSThreads, ssmem, s.stream, P, SP, Q);
        CU(ctx, cudaGetLastError());
        ctx->tm.kernel_launches += 2;
        P.read_list = s.d_defer_flat; P.read_list_n = st32 + 7; P.work_counter = st32 + 12;
    }
    if (ctx->warp_path) {
        // fast path: one warp per read; reads that do not fit a warp's shared-memory arena go to the list
        WarpParams W;
        W.arena_bytes = w_arena_bytes; W.defer_list = s.d_defer; W.defer_n = st32 + 5;
        const uint64_t warps = n;                             // one warp per read, persistent above the resident limit
        unsigned wgrid = (unsigned)std::min<uint64_t>((warps + kWThreads / 32 - 1) / (kWThreads / 32), (uint64_t)ctx->sm_count * ctx->wv_ctas[mb]);
        if (wgrid == 0) wgrid = 1;
        const size_t wsmem = (size_t)kWHeadBytes + (size_t)w_arena_bytes * (kWThreads / 32);
        PreParams Q; Q.reads = nullptr; Q.n = 0;
        if (ctx->split_path && !s.use_stream) {
            // split path: k_flat_setup prepares every read (state + CIGAR arrays in HBM), the fused kernel does the rest
            F.arena_bytes = setup_arena_bytes;
            F.consumer_flex_words = (w_arena_bytes - (uint32_t)sizeof(WFixed)) / 4u;
            F.read_count = n;
            const unsigned rgrid = (unsigned)std::min<uint64_t>(((uint64_t)n + kFThreads / 32 - 1) / (kFThreads / 32), (uint64_t)ctx->sm_count * 16);
            MMC_LAUNCH_SMEM(k_flat_setup, rgrid, (unsigned)kFThreads, (size_t)kWHeadBytes + (size_t)setup_arena_bytes * (kFThreads / 32), s.stream, P, F);
            CU(ctx, cudaGetLastError());
            ctx->tm.kernel_launches += 1;
            Q.reads = s.d_reads; Q.n = n;
            if (mb == 1) MMC_LAUNCH_SMEM((k_decode_warp<1, true>), wgrid, (unsigned)kWThreads, wsmem, s.stream, P, W, Q);
            else if (mb == 2) MMC_LAUNCH_SMEM((k_decode_warp<2, true>), wgrid, (unsigned)kWThreads, wsmem, s.stream, P, W, Q);
            else if (mb == 3) MMC_LAUNCH_SMEM((k_decode_warp<3, true>), wgrid, (unsigned)kWThreads, wsmem, s.stream, P, W, Q);
            else MMC_LAUNCH_SMEM((k_decode_warp<4, true>), wgrid, (unsigned)kWThreads, wsmem, s.stream, P, W, Q);
            CU(ctx, cudaGetLastError());
            ctx->tm.kernel_launches += 1;
            // reads k_flat_setup could not prepare go through the self-contained kernel below
            P.read_list = s.d_defer_flat; P.read_list_n = st32 + 7; P.work_counter = st32 + 12;
            Q.reads = nullptr; Q.n = 0;
        }
        if (mb == 1) MMC_LAUNCH_SMEM((k_decode_warp<1, false>), wgrid, (unsigned)kWThreads, wsmem, s.stream, P, W, Q);
        else if (mb == 2) MMC_LAUNCH_SMEM((k_decode_warp<2, false>), wgrid, (unsigned)kWThreads, wsmem, s.stream, P, W, Q);
        else if (mb == 3) MMC_LAUNCH_SMEM((k_decode_warp<3, false>), wgrid, (unsigned)kWThreads, wsmem, s.stream, P, W, Q);
        else MMC_LAUNCH_SMEM((k_decode_warp<4, false>), wgrid, (unsigned)kWThreads, wsmem, s.stream, P, W, Q);
        CU(ctx, cudaGetLastError());
        ctx->tm.kernel_launches += 1;
        P.read_list = s.d_defer; P.read_list_n = st32 + 5; P.work_counter = st32 + 6;
    }
    MMC_LAUNCH(k_decode, grid, (unsigned)ctx->threads, s.stream, P);
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaEventRecord(s.ev_k1, s.stream));
    CU(ctx, cudaMemcpyAsync(s.h_state, s.d_state, 32, cudaMemcpyDeviceToHost, s.stream));
    CU(ctx, cudaMemcpyAsync(s.h_state + 8, ctx->d_sparse_n, 8, cudaMemcpyDeviceToHost, s.stream));   // fill level of the side buffer
    ctx->tm.kernel_launches += 1;
    ctx->tm.batches += 1;
    ctx->tm.reads += n;
    s.in_flight = true; s.timed = true;
    tick(3, tsec);
    return MMC_OK;
}

int wait_slot(mmc_ctx *ctx, Slot &s) {
    if (!s.in_flight) return MMC_OK;
    CU(ctx, cudaStreamSynchronize(s.stream));
    s.in_flight = false;
    if (s.h2d_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, s.ev_h0, s.ev_h1) == cudaSuccess) ctx->tm.h2d_ms += ms;
        s.h2d_pending = false;
    }
    if (s.timed) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1) == cudaSuccess) ctx->tm.decode_ms += ms;
        ctx->tm.deferred_reads += ((const uint32_t *)s.h_state)[5];
        ctx->tm.flat_deferred_reads += ((const uint32_t *)s.h_state)[7];
        s.timed = false;
    }
    if (s.n_reads_submitted) {
        ctx->sparse_seen = std::max<uint64_t>(ctx->sparse_seen, s.h_state[8]);
        if (s.h_state[8] > ctx->sparse_cap)
            return fail(ctx, MMC_ENOMEM, "sparse count buffer overflow (%llu records > capacity %llu) inside one batch; raise sparse_capacity (minimod: --sparse-cap) or lower -K/-B",
                        (unsigned long long)s.h_state[8], (unsigned long long)ctx->sparse_cap);
    }
    if (s.n_reads_submitted && s.h_state[0] != ~0ull) {
        uint32_t read = (uint32_t)(s.h_state[0] >> 32), code = (uint32_t)(s.h_state[0] & 0xffffffffu);
        fail(ctx, MMC_EREAD, "read #%u of the batch: %s", read, read_error_text(code));
        ctx->err += "\x1f" + std::to_string(read);           // machine-readable suffix: \x1f<read index>
        return MMC_EREAD;
    }
    if (ctx->opts.subtool == MMC_VIEW && s.n_reads_submitted && s.h_state[1] > s.view_cap) {
        // more rows than the slot's buffer holds ('.' blocks with a `*` context can emit a row per base): view rows have no
        // side effect outside the slot, so grow the buffer to the count the kernels reported and run the batch again
        const uint64_t want = s.h_state[1] + s.h_state[1] / 16 + 1024;
        if (s.view_regrown) return fail(ctx, MMC_ENOMEM, "view record buffer overflow (%llu rows > %llu) after regrowing it",
                                        (unsigned long long)s.h_state[1], (unsigned long long)s.view_cap);
        if (s.d_view) CU(ctx, cudaFree(s.d_view));
        s.d_view = nullptr; s.view_cap = 0;
        if (cudaMalloc((void **)&s.d_view, want * sizeof(ViewDev)) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, MMC_ENOMEM, "cannot grow the view record buffer to %llu rows", (unsigned long long)want);
        }
        s.view_cap = want; s.view_regrown = true;
        int rc = launch_decode(ctx, s);
        if (rc == MMC_OK) rc = wait_slot(ctx, s);
        s.view_regrown = false;
        return rc;
    }
    return MMC_OK;
}

}  // namespace

// Every environment variable the library reads, in one place.  None of them changes a result: they pick between
// implementations that the test-suite proves equivalent (so that every path can be exercised on every fixture and A/B-timed
// on the GPU), override a transport form, or turn on a trace.
//   MMC_DECODE_PATH = general | warp | split | stream     which decode kernels run (default: per batch, by the reads' shape)
//   MMC_STREAM_MINB = 4|5|6|8, MMC_STREAM_SPLIT = -1|0|1, MMC_WARP_OCC = 1..4, MMC_WARP_ARENA = bytes, MMC_DECODE_THREADS   tuning
//   MMC_TEST_SMALL_SMEM = 1                                tiny shared-memory caps: forces the global-scratch fallbacks
//   MMC_SEQ_PACKING = 2|4, MMC_CIGAR_PACKING = 8|32        transport forms (override mmc_opts_t)
//   MMC_SPARSE_DEVICE_MIN = n                              side-buffer passes on the device from n records (0: always)
//   MMC_TRACE_CREATE, MMC_TRACE_FINALIZE                   timing traces on stderr (read where they are used)
static void apply_env_overrides(mmc_ctx *ctx) {
    if (const char *e = getenv("MMC_DECODE_THREADS")) { int v = atoi(e); if (v >= 32 && v <= kMaxThreads && v % 32 == 0) ctx->threads = v; }
    if (const char *e = getenv("MMC_TEST_SMALL_SMEM")) {
        if (atoi(e)) { ctx->cig_smem_cap = 16; ctx->bitmap_smem_words = 8; ctx->idx_smem_cap = 8; }
    }
    if (const char *e = getenv("MMC_DECODE_PATH")) {
        if (!strcmp(e, "general")) { ctx->warp_path = 0; ctx->split_path = 0; ctx->stream_path = 0; }
        else if (!strcmp(e, "warp")) { ctx->split_path = 0; ctx->stream_path = 0; }
        else if (!strcmp(e, "split")) { ctx->split_path = 1; ctx->stream_path = 0; }
        else if (!strcmp(e, "stream")) ctx->stream_path = 2;          // always (default 1: per batch, by the reads' shape)
    }
    if (const char *e = getenv("MMC_STREAM_SPLIT")) { const int v = atoi(e); if (v >= -1 && v <= 1) ctx->s_split_mode = v; }
    if (const char *e = getenv("MMC_STREAM_MINB")) { int v = atoi(e); if (v == 8 || v == 6 || v == 5 || v == 4) ctx->s_minb = v; }
    if (const char *e = getenv("MMC_WARP_OCC")) { int v = atoi(e); if (v >= 1 && v <= 4) { ctx->w_minb = v; ctx->w_pinned = 1; } }
    if (const char *e = getenv("MMC_SEQ_PACKING")) { int v = atoi(e); if (v == 2 || v == 4) ctx->seq_packing = v; }
    if (const char *e = getenv("MMC_CIGAR_PACKING")) { int v = atoi(e); if (v == 8 || v == 32) ctx->cigar_packing = v; }
    if (const char *e = getenv("MMC_SPARSE_DEVICE_MIN")) ctx->sparse_dev_min = strtoull(e, nullptr, 10);
    if (const char *e = getenv("MMC_WARP_ARENA")) {          // bytes of shared memory per warp
        long v = atol(e);
        if (v >= (long)sizeof(WFixed) + 256 && v <= 28 * 1024) { ctx->wv_arena[ctx->w_minb] = (uint32_t)(v & ~15l); ctx->w_pinned = 1; }
    }
}

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int mmc_abi_version(void) { return MMC_ABI_VERSION; }

const char *mmc_strerror(const mmc_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int mmc_create(mmc_ctx **out, const mmc_opts_t *opts, int32_t n_contigs, const char *const *names, const uint32_t *lens) {
    if (!out || !opts || opts->struct_size != sizeof(mmc_opts_t)) return fail(nullptr, MMC_EINVAL, "mmc_create: bad opts (ABI mismatch?)");
    if (opts->subtool != MMC_FREQ && opts->subtool != MMC_VIEW) return fail(nullptr, MMC_EINVAL, "mmc_create: subtool must be MMC_FREQ or MMC_VIEW");
    if (opts->n_mods < 1 || opts->n_mods > MMC_MAX_MODS || !opts->mods) return fail(nullptr, MMC_EINVAL, "mmc_create: need 1..%d modification codes", MMC_MAX_MODS);
    if (n_contigs < 0 || (n_contigs > 0 && (!names || !lens))) return fail(nullptr, MMC_EINVAL, "mmc_create: bad contig table");
    // MMC_TRACE_CREATE=1: where the start-up time goes (stderr)
    const bool tr = getenv("MMC_TRACE_CREATE") != nullptr;
    const auto tr0 = std::chrono::steady_clock::now();
    auto stamp = [&](const char *what) {
        if (tr) fprintf(stderr, "[mmc_create] %8.3f ms  %s\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count(), what);
    };
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(nullptr, MMC_ECUDA, "mmc_create: no CUDA device available (libminimod_cuda has no CPU fallback)");
    if (opts->device < 0 || opts->device >= ndev) return fail(nullptr, MMC_EINVAL, "mmc_create: device %d out of range (0..%d)", opts->device, ndev - 1);

    mmc_ctx *ctx = new mmc_ctx();
    ctx->opts = *opts;
    mmc_opts_t &o = ctx->opts;
    if (o.n_slots <= 0) o.n_slots = 3;
    if (o.max_reads == 0) o.max_reads = 512;                 // init_opt(), src/minimod.c:487
    if (o.max_bytes == 0) o.max_bytes = 20 * 1000 * 1000;    // src/minimod.c:488
    // '*' + HP 0..2 = four 8-byte strata = exactly one 32-byte sector per (position, strand, code): the two reductions of a
    // call (src/mod.c:906-928) land in ONE sector and a position's hot cells take one L2 sector instead of 1.75 on average
    // (config 4: 84 % of the reductions' sectors missed L2 with five strata, profiles/r02_n1_l2_atomics.txt)
    if (o.dense_haps <= 0) o.dense_haps = 3;
    if (o.dense_codes <= 0) o.dense_codes = 8;
    if (o.dense_haps > 255) o.dense_haps = 255;
    ctx->mods.assign(opts->mods, opts->mods + opts->n_mods);
    o.mods = ctx->mods.data();
    ctx->s_head = (uint32_t)std::min<int>(opts->n_mods, kWLutSlots) * 256u;
    ctx->seq_packing = opts->seq_packing == 2 ? 2 : 4;
    ctx->cigar_packing = opts->cigar_packing == 8 ? 8 : 32;
    apply_env_overrides(ctx);
    for (int mb = 1; mb <= 4; ++mb) {                        // k_flat_setup holds WRead + the un-sampled CIGAR arrays of most reads
        const uint32_t flex = ctx->wv_arena[mb] - (uint32_t)sizeof(WFixed);
        ctx->wv_setup_arena[mb] = kWReadBytes + std::min<uint32_t>(std::max<uint32_t>(4608u, (flex / 2u) & ~15u), 24576u);
    }

#define CUC(call)                                                                                            \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            fail(nullptr, MMC_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            mmc_destroy(ctx);                                                                                \
            return MMC_ECUDA;                                                                                \
        }                                                                                                    \
    } while (0)

    stamp("driver initialised (cudaGetDeviceCount)");
    CUC(cudaSetDevice(o.device));
    CUC(cudaFree(nullptr));
    stamp("device context created");
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, o.device));
    ctx->sm_count = prop.multiProcessorCount;
    int occ = 1;
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_decode, ctx->threads, 0));
    ctx->ctas_per_sm = occ < 1 ? 1 : occ;
    {
        size_t setup_max = 0;
        for (int mb = 1; mb <= 4; ++mb) setup_max = std::max(setup_max, (size_t)kWHeadBytes + (size_t)ctx->wv_setup_arena[mb] * (kFThreads / 32));
        setup_max = std::max(setup_max, (size_t)kWHeadBytes + (size_t)(kWReadBytes + 24576u) * (kFThreads / 32));
        CUC(cudaFuncSetAttribute(k_flat_setup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)setup_max));
        CUC(cudaFuncSetAttribute((k_decode_stream<8>), cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CUC(cudaFuncSetAttribute((k_decode_stream<6>), cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CUC(cudaFuncSetAttribute((k_decode_stream<5>), cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CUC(cudaFuncSetAttribute((k_decode_stream<4>), cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
#define MMC_WARP_ATTR(MB)                                                                                                        \
        do {                                                                                                                     \
            const size_t smem = (size_t)kWHeadBytes + (size_t)ctx->wv_arena[MB] * (kWThreads / 32);                         \
            int wocc = 1;                                                                                                        \
            CUC(cudaFuncSetAttribute((k_decode_warp<MB, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
            CUC(cudaFuncSetAttribute((k_decode_warp<MB, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
            CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wocc, (k_decode_warp<MB, true>), kWThreads, smem));               \
            ctx->wv_ctas[MB] = wocc < 1 ? 1 : wocc;                                                                              \
        } while (0)
        MMC_WARP_ATTR(1); MMC_WARP_ATTR(2); MMC_WARP_ATTR(3); MMC_WARP_ATTR(4);
#undef MMC_WARP_ATTR
    }

    stamp("kernel attributes set (modules loaded)");
    // ---- -c entries -> device tables
    std::vector<ReqMod> req(o.n_mods);
    ctx->wild_req = -1;
    for (int i = 0; i < o.n_mods; ++i) {
        const mmc_mod_t &m = ctx->mods[i];
        ReqMod &r = req[i];
        memset(&r, 0, sizeof(r));
        size_t cl = strnlen(m.code, MMC_MAX_CODE_LEN + 1), xl = strnlen(m.context, MMC_MAX_CONTEXT + 1);
        if (cl == 0 || cl > MMC_MAX_CODE_LEN || xl == 0 || xl > MMC_MAX_CONTEXT) {
            fail(nullptr, MMC_EINVAL, "mmc_create: modification code/context %d empty or too long", i);
            mmc_destroy(ctx); return MMC_EINVAL;
        }
        r.key = pack_code(m.code);
        if (!strcmp(m.code, "*")) ctx->wild_req = i;
        if (!strcmp(m.context, "*")) r.ctx_len = 0;
        else {
            r.ctx_len = (int32_t)xl;
            for (size_t k = 0; k < xl; ++k) {
                char c = m.context[k];
                if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') {
                    fail(nullptr, MMC_EINVAL, "mmc_create: context '%s' must be upper-case A/C/G/T/N or '*'", m.context);
                    mmc_destroy(ctx); return MMC_EINVAL;
                }
                r.pat[k] = (uint8_t)c;
                char b = m.context[xl - 1 - k];                 // reverse complement, src/ref.c:183-194
                r.pat_rc[k] = (uint8_t)(b == 'A' ? 'T' : b == 'C' ? 'G' : b == 'G' ? 'C' : b == 'T' ? 'A' : 'N');
            }
        }
        memcpy(r.lut, m.call_lut, 256);
        r.fast_ctx = r.ctx_len >= 1 && r.ctx_len <= 8;
        for (int k = 0; k < r.ctx_len && r.fast_ctx; ++k) {
            auto code2 = [](uint8_t c) -> int { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; };
            int a = code2(r.pat[k]), b = code2(r.pat_rc[k]);
            if (a < 0 || b < 0) { r.fast_ctx = 0; break; }
            r.pat2 |= (uint32_t)a << (2 * k); r.pat2_rc |= (uint32_t)b << (2 * k);
        }
    }
    CUC(cudaMalloc((void **)&ctx->d_req, sizeof(ReqMod) * req.size()));
    CUC(cudaMemcpy(ctx->d_req, req.data(), sizeof(ReqMod) * req.size(), cudaMemcpyHostToDevice));
    std::vector<unsigned long long> keys(kCodeTable, 0ull);
    if (ctx->wild_req < 0) for (int i = 0; i < o.n_mods; ++i) keys[i] = req[i].key;
    CUC(cudaMalloc((void **)&ctx->d_code_keys, sizeof(unsigned long long) * kCodeTable));
    CUC(cudaMemcpy(ctx->d_code_keys, keys.data(), sizeof(unsigned long long) * kCodeTable, cudaMemcpyHostToDevice));

    ctx->n_code_slots = ctx->wild_req >= 0 ? o.dense_codes : o.n_mods;
    ctx->n_hap_slots = o.haplotypes ? 1 + o.dense_haps : 1;

    // ---- contigs
    ctx->contigs.resize(n_contigs);
    for (int i = 0; i < n_contigs; ++i) { ctx->contigs[i].name = names[i] ? names[i] : ""; ctx->contigs[i].len = lens[i]; }
    size_t nc = std::max<size_t>(1, (size_t)n_contigs);
    CUC(cudaMalloc((void **)&ctx->d_contigs, sizeof(ContigDev) * nc));
    CUC(cudaMemset(ctx->d_contigs, 0, sizeof(ContigDev) * nc));
    CUC(cudaMalloc((void **)&ctx->d_touch, sizeof(int32_t) * 2 * nc));

    // ---- side buffers
    ctx->sparse_cap = o.sparse_capacity ? o.sparse_capacity : (o.subtool == MMC_FREQ ? std::max<uint64_t>(1u << 20, o.max_bytes / 8) : 16);   // starting size: grows (reserve_sparse)
    CUC(cudaMalloc((void **)&ctx->d_sparse, sizeof(SparseRec) * ctx->sparse_cap));
    CUC(cudaMemset(ctx->d_sparse, 0xff, sizeof(SparseRec) * ctx->sparse_cap));   // sentinels: a slot nobody has written is skipped
    CUC(cudaMalloc((void **)&ctx->d_sparse_n, 8));
    CUC(cudaMemset(ctx->d_sparse_n, 0, 8));
    ctx->view_cap = o.view_capacity ? o.view_capacity : std::max<uint64_t>(1u << 16, o.max_bytes);
    CUC(cudaStreamCreateWithFlags(&ctx->fin_stream, cudaStreamNonBlocking));
    CUC(cudaEventCreate(&ctx->ev_f0)); CUC(cudaEventCreate(&ctx->ev_f1));
    CUC(cudaEventCreate(&ctx->ev_d0)); CUC(cudaEventCreate(&ctx->ev_d1));
    CUC(cudaEventCreate(&ctx->ev_reset));

    stamp("tables and side buffers allocated");
    ctx->slots.resize(o.n_slots);
    for (Slot &s : ctx->slots) {
        int rc = setup_slot(ctx, s);
        if (rc != MMC_OK) { g_create_error = ctx->err; mmc_destroy(ctx); return rc; }
    }
    stamp("batch slots allocated (pinned host + device arenas)");
    // touch ranges start empty: [lo x n_contigs][hi x n_contigs]
    {
        const size_t m = (size_t)n_contigs;
        std::vector<int32_t> t2(2 * std::max<size_t>(1, m));
        for (size_t i = 0; i < m; ++i) { t2[i] = INT32_MAX; t2[m + i] = 0; }
        if (m) CUC(cudaMemcpy(ctx->d_touch, t2.data(), sizeof(int32_t) * 2 * m, cudaMemcpyHostToDevice));
    }
#undef CUC
    *out = ctx;
    return MMC_OK;
}

void mmc_destroy(mmc_ctx *ctx) {
    MMC_DEV(ctx);
    if (!ctx) return;
    if (getenv("MMC_TRACE_CREATE"))
        fprintf(stderr, "[mmc_destroy] host time inside mmc_batch_submit: %.1f ms enqueuing copies / unpack kernels + batch analysis, %.1f ms launching the decode stage (scratch growth included)\n",
                ctx->host_upload_ms, ctx->host_launch_ms);
    if (getenv("MMC_TRACE_CREATE"))
        fprintf(stderr, "[mmc_destroy]   of the latter: side-buffer reserve %.1f ms, state reset + general-kernel scratch %.1f ms, scratch pool %.1f ms, kernel launches %.1f ms\n",
                ctx->host_sect_ms[0], ctx->host_sect_ms[1], ctx->host_sect_ms[2], ctx->host_sect_ms[3]);
    cudaDeviceSynchronize();
    for (Slot &s : ctx->slots) {
        if (s.h_arena) cudaFreeHost(s.h_arena);
        if (s.d_arena) cudaFree(s.d_arena);
        if (s.d_state) cudaFree(s.d_state);
        if (s.h_state) cudaFreeHost(s.h_state);
        if (s.d_view) cudaFree(s.d_view);
        if (s.d_defer) cudaFree(s.d_defer);
        if (s.d_defer_flat) cudaFree(s.d_defer_flat);
        if (s.d_reads) cudaFree(s.d_reads);
        if (s.d_pool) cudaFree(s.d_pool);
        if (s.d_scratch) cudaFree(s.d_scratch);
        if (s.ev_h0) cudaEventDestroy(s.ev_h0);
        if (s.ev_h1) cudaEventDestroy(s.ev_h1);
        if (s.ev_k0) cudaEventDestroy(s.ev_k0);
        if (s.ev_k1) cudaEventDestroy(s.ev_k1);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    for (ContigHost &c : ctx->contigs) {
        if (c.dev.ref2) cudaFree((void *)c.dev.ref2);
        if (c.dev.excm) cudaFree((void *)c.dev.excm);
        if (c.dev.cells) cudaFree(c.dev.cells);
        if (c.d_exc_start) cudaFree(c.d_exc_start);
        if (c.d_exc_letter) cudaFree(c.d_exc_letter);
    }
    if (ctx->d_req) cudaFree(ctx->d_req);
    if (ctx->d_code_keys) cudaFree(ctx->d_code_keys);
    if (ctx->d_contigs) cudaFree(ctx->d_contigs);
    if (ctx->d_touch) cudaFree(ctx->d_touch);
    if (ctx->d_sparse) cudaFree(ctx->d_sparse);
    if (ctx->d_sparse_n) cudaFree(ctx->d_sparse_n);
    if (ctx->d_ascii) cudaFree(ctx->d_ascii);
    if (ctx->d_exc_tmp_start) cudaFree(ctx->d_exc_tmp_start);
    if (ctx->d_exc_tmp_letter) cudaFree(ctx->d_exc_tmp_letter);
    if (ctx->d_exc_n) cudaFree(ctx->d_exc_n);
    if (ctx->d_tile_count) cudaFree(ctx->d_tile_count);
    if (ctx->d_tile_off) cudaFree(ctx->d_tile_off);
    if (ctx->d_fin_mask) cudaFree(ctx->d_fin_mask);
    if (ctx->d_totals) cudaFree(ctx->d_totals);
    if (ctx->d_rows) cudaFree(ctx->d_rows);
    if (ctx->d_sp_scratch) cudaFree(ctx->d_sp_scratch);
    if (ctx->d_srows) cudaFree(ctx->d_srows);
    if (ctx->d_merged) cudaFree(ctx->d_merged);
    if (ctx->d_sn_rows) cudaFree(ctx->d_sn_rows);
    if (ctx->h_sn_rows) cudaFreeHost(ctx->h_sn_rows);
    if (ctx->h_rows) cudaFreeHost(ctx->h_rows);
    for (int k = 0; k < 2; ++k) if (ctx->h_drain[k]) cudaFreeHost(ctx->h_drain[k]);
    if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
    if (ctx->ev_d0) cudaEventDestroy(ctx->ev_d0);
    if (ctx->ev_d1) cudaEventDestroy(ctx->ev_d1);
    if (ctx->ev_reset) cudaEventDestroy(ctx->ev_reset);
    if (ctx->d_fin_jobs) cudaFree(ctx->d_fin_jobs);
    if (ctx->ev_f0) cudaEventDestroy(ctx->ev_f0);
    if (ctx->ev_f1) cudaEventDestroy(ctx->ev_f1);
    if (ctx->fin_stream) cudaStreamDestroy(ctx->fin_stream);
    delete ctx;
}

int mmc_ref_add(mmc_ctx *ctx, int32_t tid, const char *seq, uint32_t len) {
    MMC_DEV(ctx);
    if (!ctx) return MMC_EINVAL;
    if (tid < 0 || (size_t)tid >= ctx->contigs.size()) return fail(ctx, MMC_EINVAL, "mmc_ref_add: tid %d out of range", tid);
    ContigHost &c = ctx->contigs[tid];
    if (c.loaded) return fail(ctx, MMC_ESTATE, "mmc_ref_add: contig %s added twice", c.name.c_str());
    if (len != c.len)                                        // the reference asserts this per read, src/mod.c:861
        return fail(ctx, MMC_EINVAL, "ref_len:%u target_len:%u (contig %s: reference and BAM header lengths differ)", len, c.len, c.name.c_str());
    if (!seq && len) return fail(ctx, MMC_EINVAL, "mmc_ref_add: null sequence");
    const size_t spp = 2 * (size_t)ctx->n_code_slots * ctx->n_hap_slots;
    const size_t n32 = ((size_t)len + 31) / 32;
    size_t free_b = 0, total_b = 0;
    CU(ctx, cudaMemGetInfo(&free_b, &total_b));
    size_t need = (ctx->opts.subtool == MMC_FREQ ? (size_t)len * spp * 8 : 0) + n32 * 12 + (64u << 20);
    if (need > free_b)
        return fail(ctx, MMC_ENOMEM, "contig %s needs %.1f GB of HBM for its packed reference and dense count array (%zu cells per position) but %.1f GB are free; shard contigs across GPUs or lower dense_haps/dense_codes",
                    c.name.c_str(), need / 1e9, spp, free_b / 1e9);
    uint32_t *ref2 = nullptr, *excm = nullptr;
    CU(ctx, cudaMalloc((void **)&ref2, n32 * 8 + 16));       // + slack: windows are read as two adjacent words
    CU(ctx, cudaMalloc((void **)&excm, n32 * 4 + 16));
    CU(ctx, cudaMemsetAsync(ref2 + n32 * 2, 0, 16, ctx->fin_stream));
    CU(ctx, cudaMemsetAsync(excm + n32, 0, 16, ctx->fin_stream));
    c.dev.ref2 = ref2; c.dev.excm = excm; c.dev.len = len;
    if (ctx->opts.subtool == MMC_FREQ) {
        CU(ctx, cudaMalloc((void **)&c.dev.cells, std::max<size_t>(8, (size_t)len * spp * 8)));
        CU(ctx, cudaMemsetAsync(c.dev.cells, 0, std::max<size_t>(8, (size_t)len * spp * 8), ctx->fin_stream));
    }
    if (!ctx->d_ascii) {
        ctx->ascii_cap = 64u << 20;
        CU(ctx, cudaMalloc((void **)&ctx->d_ascii, ctx->ascii_cap));
        CU(ctx, cudaMalloc((void **)&ctx->d_exc_tmp_start, sizeof(uint32_t) * kExcCap));
        CU(ctx, cudaMalloc((void **)&ctx->d_exc_tmp_letter, kExcCap));
        CU(ctx, cudaMalloc((void **)&ctx->d_exc_n, 4));
    }
    CU(ctx, cudaMemsetAsync(ctx->d_exc_n, 0, 4, ctx->fin_stream));
    uint32_t prev = 0;
    for (size_t off = 0; off < len; off += ctx->ascii_cap) {
        size_t n = std::min<size_t>(ctx->ascii_cap, len - off);
        CU(ctx, cudaMemcpyAsync(ctx->d_ascii, seq + off, n, cudaMemcpyHostToDevice, ctx->fin_stream));
        RefPackParams rp;
        rp.ascii = ctx->d_ascii; rp.g_first = (uint32_t)off; rp.n = (uint32_t)n; rp.prev_letter = prev;
        rp.ref2 = ref2; rp.excm = excm;
        rp.exc_start = ctx->d_exc_tmp_start; rp.exc_letter = ctx->d_exc_tmp_letter; rp.exc_cap = kExcCap; rp.exc_n = ctx->d_exc_n;
        unsigned threads = 256, grid = (unsigned)((n + 32 * (size_t)threads - 1) / (32 * (size_t)threads));
        MMC_LAUNCH(k_ref_pack, grid, threads, ctx->fin_stream, rp);
        CU(ctx, cudaGetLastError());
        ctx->tm.kernel_launches += 1;
        CU(ctx, cudaStreamSynchronize(ctx->fin_stream));    // seq+off may be pageable; keep it simple and ordered
        unsigned char last = (unsigned char)seq[off + n - 1];
        if (last >= 'a' && last <= 'z') last -= 32;
        prev = last == 'U' ? 'T' : last;
    }
    uint32_t n_exc = 0;
    CU(ctx, cudaMemcpy(&n_exc, ctx->d_exc_n, 4, cudaMemcpyDeviceToHost));
    if (n_exc > kExcCap)
        return fail(ctx, MMC_ENOMEM, "contig %s has more than %u runs of non-ACGT letters (library limit)", c.name.c_str(), kExcCap);
    if (n_exc) {
        std::vector<uint32_t> st(n_exc);
        std::vector<uint8_t> le(n_exc);
        CU(ctx, cudaMemcpy(st.data(), ctx->d_exc_tmp_start, 4 * (size_t)n_exc, cudaMemcpyDeviceToHost));
        CU(ctx, cudaMemcpy(le.data(), ctx->d_exc_tmp_letter, n_exc, cudaMemcpyDeviceToHost));
        std::vector<uint32_t> order(n_exc);
        for (uint32_t i = 0; i < n_exc; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return st[a] < st[b]; });
        std::vector<uint32_t> st2(n_exc);
        std::vector<uint8_t> le2(n_exc);
        for (uint32_t i = 0; i < n_exc; ++i) { st2[i] = st[order[i]]; le2[i] = le[order[i]]; }
        CU(ctx, cudaMalloc((void **)&c.d_exc_start, 4 * (size_t)n_exc));
        CU(ctx, cudaMalloc((void **)&c.d_exc_letter, n_exc));
        CU(ctx, cudaMemcpy(c.d_exc_start, st2.data(), 4 * (size_t)n_exc, cudaMemcpyHostToDevice));
        CU(ctx, cudaMemcpy(c.d_exc_letter, le2.data(), n_exc, cudaMemcpyHostToDevice));
    }
    c.dev.exc_start = c.d_exc_start; c.dev.exc_letter = c.d_exc_letter; c.dev.n_exc = n_exc;
    c.loaded = true;
    ctx->committed = false;
    return MMC_OK;
}

int mmc_ref_commit(mmc_ctx *ctx) {
    MMC_DEV(ctx);
    if (!ctx) return MMC_EINVAL;
    std::vector<ContigDev> tab(ctx->contigs.size());
    for (size_t i = 0; i < tab.size(); ++i) {
        if (ctx->contigs[i].loaded) tab[i] = ctx->contigs[i].dev; else memset(&tab[i], 0, sizeof(ContigDev));
    }
    if (!tab.empty()) CU(ctx, cudaMemcpy(ctx->d_contigs, tab.data(), sizeof(ContigDev) * tab.size(), cudaMemcpyHostToDevice));
    CU(ctx, cudaStreamSynchronize(ctx->fin_stream));
    ctx->committed = true;
    return MMC_OK;
}

int mmc_batch_acquire(mmc_ctx *ctx, mmc_batch_t **batch) {
    MMC_DEV(ctx);
    if (!ctx || !batch) return MMC_EINVAL;
    Slot *pick = nullptr;
    for (Slot &s : ctx->slots) if (!s.acquired) { pick = &s; break; }
    if (!pick) return fail(ctx, MMC_ESTATE, "mmc_batch_acquire: all %d slots are in use; release one first", (int)ctx->slots.size());
    int rc = wait_slot(ctx, *pick);
    if (rc != MMC_OK) return rc;
    pick->acquired = true; pick->uploaded = false;
    mmc_batch_t &b = pick->pub;
    b.n_reads = 0; b.cigar_used = b.seq_used = b.mm_used = b.ml_used = 0; b.seq_exc_used = 0; b.cig8_used = 0;
    *batch = &b;
    return MMC_OK;
}

int mmc_batch_upload(mmc_ctx *ctx, mmc_batch_t *batch) {
    MMC_DEV(ctx);
    Slot *s = ctx ? slot_of(ctx, batch) : nullptr;
    if (!s || !s->acquired) return fail(ctx, MMC_ESTATE, "mmc_batch_upload: not an acquired batch");
    if (!ctx->committed) return fail(ctx, MMC_ESTATE, "mmc_batch_upload: call mmc_ref_commit() first");
    int rc = wait_slot(ctx, *s);
    if (rc != MMC_OK) return rc;
    rc = upload(ctx, *s);
    if (rc != MMC_OK) return rc;
    CU(ctx, cudaStreamSynchronize(s->stream));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, s->ev_h0, s->ev_h1) == cudaSuccess) ctx->tm.h2d_ms += ms;
    s->h2d_pending = false;
    return MMC_OK;
}

int mmc_batch_launch(mmc_ctx *ctx, mmc_batch_t *batch) {
    MMC_DEV(ctx);
    Slot *s = ctx ? slot_of(ctx, batch) : nullptr;
    if (!s || !s->acquired || !s->uploaded) return fail(ctx, MMC_ESTATE, "mmc_batch_launch: batch was not uploaded");
    int rc = wait_slot(ctx, *s);
    if (rc != MMC_OK) return rc;
    return launch_decode(ctx, *s);
}

int mmc_batch_submit(mmc_ctx *ctx, mmc_batch_t *batch) {
    MMC_DEV(ctx);
    Slot *s = ctx ? slot_of(ctx, batch) : nullptr;
    if (!s || !s->acquired) return fail(ctx, MMC_ESTATE, "mmc_batch_submit: not an acquired batch");
    if (!ctx->committed) return fail(ctx, MMC_ESTATE, "mmc_batch_submit: call mmc_ref_commit() first");
    int rc = wait_slot(ctx, *s);
    if (rc != MMC_OK) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    rc = upload(ctx, *s);
    if (rc != MMC_OK) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    rc = launch_decode(ctx, *s);
    const auto t2 = std::chrono::steady_clock::now();
    ctx->host_upload_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
    ctx->host_launch_ms += std::chrono::duration<double, std::milli>(t2 - t1).count();
    return rc;
}

int mmc_batch_wait(mmc_ctx *ctx, mmc_batch_t *batch) {
    MMC_DEV(ctx);
    Slot *s = ctx ? slot_of(ctx, batch) : nullptr;
    if (!s) return fail(ctx, MMC_ESTATE, "mmc_batch_wait: unknown batch");
    return wait_slot(ctx, *s);
}

int mmc_batch_release(mmc_ctx *ctx, mmc_batch_t *batch) {
    MMC_DEV(ctx);
    Slot *s = ctx ? slot_of(ctx, batch) : nullptr;
    if (!s) return fail(ctx, MMC_ESTATE, "mmc_batch_release: unknown batch");
    int rc = wait_slot(ctx, *s);
    s->acquired = false; s->uploaded = false;
    return rc;
}

int mmc_sync(mmc_ctx *ctx) {
    MMC_DEV(ctx);
    if (!ctx) return MMC_EINVAL;
    int first = reset_settle(ctx);
    for (Slot &s : ctx->slots) { int rc = wait_slot(ctx, s); if (rc != MMC_OK && first == MMC_OK) first = rc; }
    return first;
}

int mmc_last_decode_ms(mmc_ctx *ctx, mmc_batch_t *batch, double *ms) {
    MMC_DEV(ctx);
    Slot *s = ctx ? slot_of(ctx, batch) : nullptr;
    if (!s || !ms) return MMC_EINVAL;
    int rc = wait_slot(ctx, *s);
    if (rc != MMC_OK) return rc;
    float f = 0;
    CU(ctx, cudaEventElapsedTime(&f, s->ev_k0, s->ev_k1));
    *ms = f;
    return MMC_OK;
}

const char *mmc_code_name(const mmc_ctx *ctx, int32_t code) {
    if (!ctx || code < 0) return "";
    if (ctx->wild_req < 0) return code < ctx->opts.n_mods ? ctx->mods[code].code : "";
    return (size_t)code < ctx->code_names.size() ? ctx->code_names[code].c_str() : "";
}

static int refresh_code_names(mmc_ctx *ctx) {
    if (ctx->wild_req < 0) return MMC_OK;
    std::vector<unsigned long long> keys(kCodeTable);
    CU(ctx, cudaMemcpy(keys.data(), ctx->d_code_keys, sizeof(unsigned long long) * kCodeTable, cudaMemcpyDeviceToHost));
    ctx->code_names.assign(kCodeTable, std::string());
    for (int i = 0; i < kCodeTable; ++i) ctx->code_names[i] = unpack_code(keys[i]);
    return MMC_OK;
}

}  // extern "C"

// ---- the two library primitives of the sparse finalize: stable LSD radix sort of (key, value) pairs, exclusive sum.
// Under the SIMT emulator (CPU CI) device memory is host memory and the same contracts are met with the C++ library.
template <typename K>
static int sort_pairs(mmc_ctx *ctx, const K *kin, K *kout, const uint32_t *vin, uint32_t *vout, uint32_t n, int end_bit, void *tmp, size_t tmp_bytes) {
#ifdef MMC_EMUL
    (void)tmp; (void)tmp_bytes; (void)ctx;
    std::vector<uint32_t> ord(n);
    for (uint32_t i = 0; i < n; ++i) ord[i] = i;
    const K mask = end_bit >= (int)(8 * sizeof(K)) ? ~(K)0 : (((K)1 << end_bit) - 1);
    std::stable_sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) { return (kin[x] & mask) < (kin[y] & mask); });
    for (uint32_t i = 0; i < n; ++i) { kout[i] = kin[ord[i]]; vout[i] = vin[ord[i]]; }
#else
    CU(ctx, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, (int)n, 0, end_bit, ctx->fin_stream));
#endif
    return MMC_OK;
}
static int exclusive_sum(mmc_ctx *ctx, const uint32_t *in, uint32_t *out, uint32_t n, void *tmp, size_t tmp_bytes) {
#ifdef MMC_EMUL
    (void)tmp; (void)tmp_bytes; (void)ctx;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n; ++i) { const uint32_t v = in[i]; out[i] = acc; acc += v; }
#else
    CU(ctx, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int)n, ctx->fin_stream));
#endif
    return MMC_OK;
}

// Sort + reduce the sn records of the sparse side buffer into ctx->d_srows (row order), count into ctx->h_sn_rows
// (valid after the next synchronize of fin_stream).  Everything is queued on fin_stream.
static int sparse_rows_on_device(mmc_ctx *ctx, uint64_t sn, unsigned long long key_lo, unsigned long long key_hi) {
    const uint32_t n = (uint32_t)sn;
    size_t tmp_bytes = 0;
#ifndef MMC_EMUL
    {
        size_t t1 = 0, t2 = 0, t3 = 0;
        CU(ctx, cub::DeviceRadixSort::SortPairs(nullptr, t1, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, 25, ctx->fin_stream));
        CU(ctx, cub::DeviceRadixSort::SortPairs(nullptr, t2, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, 64, ctx->fin_stream));
        CU(ctx, cub::DeviceScan::ExclusiveSum(nullptr, t3, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, ctx->fin_stream));
        tmp_bytes = std::max(t1, std::max(t2, t3));
    }
#endif
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b32 = up(4 * (size_t)n), b64 = up(8 * (size_t)n);
    const size_t need = 4 * b32 + 2 * b64 + up(tmp_bytes);
    if (need > ctx->sp_scratch_cap) {
        if (ctx->d_sp_scratch) cudaFree(ctx->d_sp_scratch);
        ctx->d_sp_scratch = nullptr; ctx->sp_scratch_cap = 0;
        CU(ctx, cudaMalloc(&ctx->d_sp_scratch, need + need / 8));
        ctx->sp_scratch_cap = need + need / 8;
    }
    if (sn > ctx->d_srows_cap) {
        if (ctx->d_srows) cudaFree(ctx->d_srows);
        ctx->d_srows = nullptr; ctx->d_srows_cap = 0;
        CU(ctx, cudaMalloc((void **)&ctx->d_srows, sizeof(FreqRecDev) * (sn + sn / 8)));
        ctx->d_srows_cap = sn + sn / 8;
    }
    if (!ctx->d_sn_rows) {
        CU(ctx, cudaMalloc((void **)&ctx->d_sn_rows, 8));
        CU(ctx, cudaMallocHost((void **)&ctx->h_sn_rows, 8));
    }
    uint8_t *base = reinterpret_cast<uint8_t *>(ctx->d_sp_scratch);
    uint32_t *k32a = reinterpret_cast<uint32_t *>(base), *k32b = reinterpret_cast<uint32_t *>(base + b32);
    uint32_t *ia = reinterpret_cast<uint32_t *>(base + 2 * b32), *ib = reinterpret_cast<uint32_t *>(base + 3 * b32);
    unsigned long long *k64a = reinterpret_cast<unsigned long long *>(base + 4 * b32), *k64b = reinterpret_cast<unsigned long long *>(base + 4 * b32 + b64);
    void *tmp = base + 4 * b32 + 2 * b64;
    const unsigned grid = (unsigned)std::min<uint64_t>((n + kSpThreads - 1) / kSpThreads, (uint64_t)ctx->sm_count * 8);
    int tid_bits = 1;
    while (((size_t)1 << tid_bits) <= ctx->contigs.size()) ++tid_bits;      // the all-ones sentinel stays the largest key
    int rc;
    MMC_LAUNCH(k_sparse_keys, grid, (unsigned)kSpThreads, ctx->fin_stream, (const SparseRec *)ctx->d_sparse, n, k32a, ia);
    CU(ctx, cudaGetLastError());
    if ((rc = sort_pairs<uint32_t>(ctx, k32a, k32b, ia, ib, n, 25, tmp, tmp_bytes)) != MMC_OK) return rc;
    MMC_LAUNCH(k_sparse_gather, grid, (unsigned)kSpThreads, ctx->fin_stream, (const SparseRec *)ctx->d_sparse, (const uint32_t *)ib, n, k64a, key_lo, key_hi);
    CU(ctx, cudaGetLastError());
    if ((rc = sort_pairs<unsigned long long>(ctx, k64a, k64b, ib, ia, n, std::min(64, 41 + tid_bits), tmp, tmp_bytes)) != MMC_OK) return rc;
    uint32_t *flag = k32a, *off = k32b;                                      // the minor keys are no longer needed
    MMC_LAUNCH(k_sparse_heads, grid, (unsigned)kSpThreads, ctx->fin_stream, (const SparseRec *)ctx->d_sparse, (const uint32_t *)ia, n, flag, key_lo, key_hi);
    CU(ctx, cudaGetLastError());
    if ((rc = exclusive_sum(ctx, flag, off, n, tmp, tmp_bytes)) != MMC_OK) return rc;
    MMC_LAUNCH(k_sparse_emit, grid, (unsigned)kSpThreads, ctx->fin_stream, (const SparseRec *)ctx->d_sparse, (const uint32_t *)ia, (const uint32_t *)flag,
               (const uint32_t *)off, n, ctx->d_srows, ctx->d_sn_rows, key_lo, key_hi);
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(ctx->h_sn_rows, ctx->d_sn_rows, 8, cudaMemcpyDeviceToHost, ctx->fin_stream));
    ctx->tm.kernel_launches += 4;
    return MMC_OK;
}

// The rows of the count cells in coordinate order.  drain == false: everything no drain has returned yet (after waiting
// for every batch).  drain == true: what lies before the watermark (wm_tid, wm_pos), after waiting only for the batches
// that hold a read starting before it -- later batches keep copying and decoding while these rows are compacted and read
// back (they only touch cells at or after the watermark).
static int finalize_rows(mmc_ctx *ctx, bool drain, uint32_t wm_tid, uint32_t wm_pos_given, const mmc_freq_rec_t **recs, uint64_t *n_recs) {
    int rc = MMC_OK;
    const uint64_t wm_key = ((uint64_t)wm_tid << 32) | wm_pos_given;
    // Rows leave up to ONE POSITION BELOW the watermark: a read that starts exactly at the watermark may still add a side-buffer
    // record at the position before it (the left flank of a leading insertion, src/mod.c:1122-1127), and all rows of a position
    // -- dense cells and side-buffer records -- have to leave in the same call to come out in order.
    const uint32_t wm_pos = wm_pos_given ? wm_pos_given - 1u : 0u;
    if (!drain) rc = mmc_sync(ctx);
    else for (Slot &s : ctx->slots) {
        if (!s.in_flight || ((((uint64_t)s.min_tid << 32) | s.min_pos) >= wm_key && s.n_reads_submitted)) continue;
        const int r1 = wait_slot(ctx, s);
        if (r1 != MMC_OK && rc == MMC_OK) rc = r1;
    }
    if (rc != MMC_OK) return rc;
    *n_recs = 0; *recs = nullptr;
    if ((rc = reset_settle(ctx)) != MMC_OK) return rc;
    rc = refresh_code_names(ctx);
    if (rc != MMC_OK) return rc;
    const size_t nc = ctx->contigs.size();
    if (ctx->drained_to.size() != nc) ctx->drained_to.assign(nc, 0u);
    std::vector<int32_t> touch(2 * std::max<size_t>(1, nc));
    if (nc) CU(ctx, cudaMemcpy(touch.data(), ctx->d_touch, sizeof(int32_t) * 2 * nc, cudaMemcpyDeviceToHost));
    const uint32_t spp = 2u * (uint32_t)ctx->n_code_slots * (uint32_t)ctx->n_hap_slots;
    mmc_freq_rec_t *&h_rows = drain ? ctx->h_drain[ctx->drain_flip] : ctx->h_rows;
    size_t &h_rows_cap = drain ? ctx->h_drain_cap[ctx->drain_flip] : ctx->h_rows_cap;

    struct Job { int32_t tid; int32_t lo; uint64_t n_cells; uint64_t n_tiles; uint64_t tile0; };
    std::vector<Job> jobs;
    uint64_t tiles = 0;
    for (size_t i = 0; i < nc; ++i) {
        if (!ctx->contigs[i].loaded) continue;
        int32_t lo = touch[i], hi = touch[nc + i];
        if ((int64_t)ctx->drained_to[i] > lo) lo = (int32_t)std::min<uint32_t>(ctx->drained_to[i], (uint32_t)INT32_MAX);
        if (drain) {
            if ((uint32_t)i > wm_tid) break;
            if ((uint32_t)i == wm_tid && (int64_t)wm_pos < hi) hi = (int32_t)wm_pos;
        }
        if (lo >= hi) continue;
        Job j; j.tid = (int32_t)i; j.lo = lo; j.n_cells = (uint64_t)(hi - lo) * spp;
        j.n_tiles = (j.n_cells + kTileCells - 1) / kTileCells; j.tile0 = tiles;
        tiles += j.n_tiles;
        jobs.push_back(j);
    }
    // the sparse side buffer is complete once the decodes are: its size bounds the rows it can add
    unsigned long long sn = 0;
    CU(ctx, cudaMemcpy(&sn, ctx->d_sparse_n, 8, cudaMemcpyDeviceToHost));
    if (sn > ctx->sparse_cap)
        return fail(ctx, MMC_ENOMEM, "sparse count buffer overflow (%llu records > capacity %llu); raise sparse_capacity",
                    sn, (unsigned long long)ctx->sparse_cap);
    // Which records of the side buffer this call owes: major keys in [sparse_lo, sparse_hi).  Batches in flight may be
    // appending while a drain reads the buffer: their keys are never below sparse_hi, and what they have not written yet
    // reads as sentinels (the buffer is pre-filled with them).
    unsigned long long sparse_hi = ~0ull;
    if (drain) {
        sparse_hi = ((unsigned long long)wm_tid << 41) | ((unsigned long long)wm_pos << 9);
        if (sparse_hi < ctx->sparse_lo) sparse_hi = ctx->sparse_lo;
    }
    if (drain && sn >= 0x7fffffffull) return MMC_OK;         // (too many records for one device pass: left to mmc_freq_finalize())
    const bool dev_sparse = sn > 0 && (drain || sn >= ctx->sparse_dev_min) && sn < 0x7fffffffull;   // many records (--insertions), or a drain: sorted on the device
    std::vector<SparseRec> raw(dev_sparse ? 0 : sn);
    if (sn && !dev_sparse) {
        CU(ctx, cudaMemcpyAsync(raw.data(), ctx->d_sparse, sizeof(SparseRec) * sn, cudaMemcpyDeviceToHost, ctx->fin_stream));
        ctx->tm.d2h_bytes += sizeof(SparseRec) * sn;
    }
    auto ensure_rows = [&](uint64_t rows) -> int {           // pinned result buffer: dense rows + room to merge the sparse ones in
        if (rows <= h_rows_cap) return MMC_OK;
        if (h_rows) cudaFreeHost(h_rows);
        h_rows = nullptr; h_rows_cap = 0;
        const size_t cap = rows + rows / 8 + 1024;
        CU(ctx, cudaMallocHost((void **)&h_rows, sizeof(mmc_freq_rec_t) * cap));
        h_rows_cap = cap;
        return MMC_OK;
    };
    uint64_t n_dense = 0;
    CU(ctx, cudaEventRecord(ctx->ev_f0, ctx->fin_stream));
    if (dev_sparse) {
        rc = sparse_rows_on_device(ctx, sn, ctx->sparse_lo, sparse_hi);
        if (rc != MMC_OK) return rc;
        if (!tiles) CU(ctx, cudaStreamSynchronize(ctx->fin_stream));       // (else the wait for the tile totals covers it)
    }
    if (tiles) {
        if (tiles > ctx->fin_tiles_cap) {
            if (ctx->d_tile_count) cudaFree(ctx->d_tile_count);
            if (ctx->d_tile_off) cudaFree(ctx->d_tile_off);
            ctx->d_tile_count = nullptr; ctx->d_tile_off = nullptr; ctx->fin_tiles_cap = 0;
            if (ctx->d_fin_mask) cudaFree(ctx->d_fin_mask);
            ctx->d_fin_mask = nullptr;
            CU(ctx, cudaMalloc((void **)&ctx->d_tile_count, 4 * tiles));
            CU(ctx, cudaMalloc((void **)&ctx->d_tile_off, 8 * tiles));
            CU(ctx, cudaMalloc((void **)&ctx->d_fin_mask, 4 * (tiles * (kTileCells / 32) + 1)));
            ctx->fin_tiles_cap = tiles;
        }
        if (jobs.size() > ctx->fin_jobs_cap) {
            if (ctx->d_fin_jobs) cudaFree(ctx->d_fin_jobs);
            ctx->d_fin_jobs = nullptr; ctx->fin_jobs_cap = 0;
            CU(ctx, cudaMalloc((void **)&ctx->d_fin_jobs, sizeof(FinJob) * jobs.size()));
            ctx->fin_jobs_cap = jobs.size();
        }
        if (!ctx->d_totals) {
            CU(ctx, cudaMalloc((void **)&ctx->d_totals, 8));
            CU(ctx, cudaMallocHost((void **)&ctx->h_totals, 16));
        }
        // ONE launch per pass over every contig range (a whole-genome finalize used to be 3 launches x 195 contigs)
        std::vector<FinJob> fj(jobs.size());
        for (size_t k = 0; k < jobs.size(); ++k) {
            const Job &j = jobs[k];
            fj[k].cells = ctx->contigs[j.tid].dev.cells + (uint64_t)j.lo * spp;
            fj[k].n_cells = j.n_cells; fj[k].tile0 = j.tile0; fj[k].tid = j.tid; fj[k].lo = j.lo;
        }
        CU(ctx, cudaMemcpyAsync(ctx->d_fin_jobs, fj.data(), sizeof(FinJob) * fj.size(), cudaMemcpyHostToDevice, ctx->fin_stream));
        uint32_t *d_overflow = ctx->d_fin_mask + ctx->fin_tiles_cap * (kTileCells / 32);
        CU(ctx, cudaMemsetAsync(d_overflow, 0, 4, ctx->fin_stream));
        FinalizeParams fp;
        memset(&fp, 0, sizeof(fp));
        fp.jobs = ctx->d_fin_jobs; fp.n_jobs = (uint32_t)jobs.size();
        fp.n_code_slots = ctx->n_code_slots; fp.n_hap_slots = ctx->n_hap_slots; fp.haplotypes = ctx->opts.haplotypes;
        fp.tile_count = ctx->d_tile_count; fp.tile_offset = ctx->d_tile_off; fp.cells_per_tile = kTileCells;
        fp.mask = ctx->d_fin_mask; fp.overflow = d_overflow;
        MMC_LAUNCH(k_count_nonzero, (unsigned)tiles, 256u, ctx->fin_stream, fp);
        CU(ctx, cudaGetLastError());
        MMC_LAUNCH(k_scan_tiles, 1u, 256u, ctx->fin_stream, fp.tile_count, fp.tile_offset, (uint32_t)tiles, ctx->d_totals);
        CU(ctx, cudaGetLastError());
        ctx->tm.kernel_launches += 2;
        CU(ctx, cudaMemcpyAsync(ctx->h_totals, ctx->d_totals, 8, cudaMemcpyDeviceToHost, ctx->fin_stream));
        CU(ctx, cudaMemcpyAsync(ctx->h_totals + 1, d_overflow, 4, cudaMemcpyDeviceToHost, ctx->fin_stream));
        CU(ctx, cudaStreamSynchronize(ctx->fin_stream));     // (also: fj has been copied)
        if ((uint32_t)ctx->h_totals[1])                      // src/mod.c:899-901,922-924: the reference aborts too
            return fail(ctx, MMC_ENOMEM, "n_called overflowed for a position (more than 4294967295 calls on one cell). Please report this issue.");
        const uint64_t total = ctx->h_totals[0];
        if (total) {
            if (total > ctx->d_rows_cap) {
                if (ctx->d_rows) cudaFree(ctx->d_rows);
                ctx->d_rows = nullptr; ctx->d_rows_cap = 0;
                const size_t cap = total + total / 8 + 1024;
                CU(ctx, cudaMalloc((void **)&ctx->d_rows, sizeof(FreqRecDev) * cap));
                ctx->d_rows_cap = cap;
            }
            fp.out = ctx->d_rows;
            MMC_LAUNCH(k_emit_records, (unsigned)tiles, 256u, ctx->fin_stream, fp);
            CU(ctx, cudaGetLastError());
            ctx->tm.kernel_launches += 1;
            n_dense = total;
        }
    }
    // rows that go back in one copy: the dense rows, with the device-sorted sparse rows merged in when there are any
    uint64_t n_dev_rows = n_dense;
    {
        const FreqRecDev *src = ctx->d_rows;
        const uint64_t n_srows = dev_sparse ? (uint64_t)*ctx->h_sn_rows : 0;
        if (n_srows && n_dense) {
            n_dev_rows = n_dense + n_srows;
            if (n_dev_rows > ctx->d_merged_cap) {
                if (ctx->d_merged) cudaFree(ctx->d_merged);
                ctx->d_merged = nullptr; ctx->d_merged_cap = 0;
                const size_t cap = n_dev_rows + n_dev_rows / 8 + 1024;
                CU(ctx, cudaMalloc((void **)&ctx->d_merged, sizeof(FreqRecDev) * cap));
                ctx->d_merged_cap = cap;
            }
            const unsigned grid = (unsigned)std::min<uint64_t>((n_dev_rows + kSpThreads - 1) / kSpThreads, (uint64_t)ctx->sm_count * 16);
            MMC_LAUNCH(k_merge_rows, grid, (unsigned)kSpThreads, ctx->fin_stream, (const FreqRecDev *)ctx->d_rows, (unsigned long long)n_dense,
                       (const FreqRecDev *)ctx->d_srows, (unsigned long long)n_srows, ctx->d_merged);
            CU(ctx, cudaGetLastError());
            ctx->tm.kernel_launches += 1;
            src = ctx->d_merged;
        } else if (n_srows) {
            n_dev_rows = n_srows; src = ctx->d_srows;
        }
        CU(ctx, cudaEventRecord(ctx->ev_f1, ctx->fin_stream));
        if (n_dev_rows) {
            rc = ensure_rows(n_dev_rows + (dev_sparse ? 0 : sn));
            if (rc != MMC_OK) return rc;
            CU(ctx, cudaEventRecord(ctx->ev_d0, ctx->fin_stream));
            CU(ctx, cudaMemcpyAsync(h_rows, src, sizeof(FreqRecDev) * n_dev_rows, cudaMemcpyDeviceToHost, ctx->fin_stream));
            CU(ctx, cudaEventRecord(ctx->ev_d1, ctx->fin_stream));
            ctx->tm.d2h_bytes += sizeof(FreqRecDev) * n_dev_rows;
        }
    }
    n_dense = n_dev_rows;                                     // what the host merge below (few sparse records) starts from
    // ---- sparse side buffer: sort + reduce on the host while the dense rows are on their way back ...
    std::vector<mmc_freq_rec_t> sparse;
    const bool fin_trace = getenv("MMC_TRACE_FINALIZE") != nullptr;
    const auto t_s0 = std::chrono::steady_clock::now();
    if (!raw.empty()) {
        CU(ctx, cudaStreamSynchronize(ctx->fin_stream));   // (the copy of the records)
        for (SparseRec &r : raw) if (r.a < ctx->sparse_lo || r.a >= sparse_hi) r.a = kSpSentinel;   // not this call's (see above)
        std::sort(raw.begin(), raw.end(), [](const SparseRec &x, const SparseRec &y) {
            if (x.a != y.a) return x.a < y.a;                 // tid, pos, strand, code == numeric order of a
            uint32_t xi = x.b & 0xffffu, yi = y.b & 0xffffu;
            if (xi != yi) return xi < yi;
            // hap: '*' (256) first, then 0,1,...
            int32_t xh = (int32_t)(x.b >> 16) == 256 ? -1 : (int32_t)(x.b >> 16), yh = (int32_t)(y.b >> 16) == 256 ? -1 : (int32_t)(y.b >> 16);
            return xh < yh;
        });
        sparse.reserve(raw.size());
        for (size_t i = 0; i < raw.size();) {
            if (raw[i].a == kSpSentinel) break;              // unused slots of the warps' chunks sort last
            size_t j = i;
            uint64_t called = 0, mod = 0;
            while (j < raw.size() && raw[j].a == raw[i].a && raw[j].b == raw[i].b) { called += raw[j].w & 0xffffu; mod += raw[j].w >> 16; ++j; }
            mmc_freq_rec_t r;
            r.tid = (int32_t)(raw[i].a >> 41); r.pos = (int32_t)((raw[i].a >> 9) & 0xffffffffull);
            r.strand = (uint8_t)((raw[i].a >> 8) & 1u); r.code = (uint8_t)(raw[i].a & 0xffu);
            r.ins_offset = (uint16_t)(raw[i].b & 0xffffu);
            uint32_t h9 = raw[i].b >> 16;
            r.hap = h9 == 256 ? (int16_t)-1 : (int16_t)h9;
            r.n_called = (uint32_t)called; r.n_mod = (uint32_t)mod; r.reserved = 0;
            sparse.push_back(r);
            i = j;
        }
    }
    const auto t_s1 = std::chrono::steady_clock::now();
    CU(ctx, cudaStreamSynchronize(ctx->fin_stream));
    const auto t_s2 = std::chrono::steady_clock::now();
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev_f0, ctx->ev_f1) == cudaSuccess) ctx->tm.finalize_ms += ms;
        if (n_dense && cudaEventElapsedTime(&ms, ctx->ev_d0, ctx->ev_d1) == cudaSuccess) ctx->tm.d2h_ms += ms;
    }
    // ---- ... then merged into the pinned dense rows in place, from the back: the dense rows after each sparse row
    // move up as one block (there are few sparse rows, so this is a handful of large memmoves, no second copy)
    const size_t ns = sparse.size();
    if (ns) {
        rc = ensure_rows(n_dense + ns);                      // only allocates when there were no dense rows at all
        if (rc != MMC_OK) return rc;
        auto less = [](const mmc_freq_rec_t &x, const mmc_freq_rec_t &y) {
            if (x.tid != y.tid) return x.tid < y.tid;
            if (x.pos != y.pos) return x.pos < y.pos;
            if (x.strand != y.strand) return x.strand < y.strand;
            if (x.code != y.code) return x.code < y.code;
            if (x.ins_offset != y.ins_offset) return x.ins_offset < y.ins_offset;
            return x.hap < y.hap;
        };
        mmc_freq_rec_t *rows = h_rows;
        size_t end = n_dense;                                // dense rows [0, end) are still where the copy put them
        for (size_t j = ns; j-- > 0;) {
            const mmc_freq_rec_t &sr = sparse[j];
            size_t hi = end, lo = 0, step = 1;               // gallop back from end: consecutive sparse rows are close
            while (step <= hi) {
                if (!less(sr, rows[hi - step])) { lo = hi - step + 1; break; }
                hi -= step; step <<= 1;
            }
            const size_t p = (size_t)(std::upper_bound(rows + lo, rows + hi, sr, less) - rows);   // first dense row > sr
            if (end > p) memmove(rows + p + j + 1, rows + p, (end - p) * sizeof(mmc_freq_rec_t));
            rows[p + j] = sr;
            end = p;
        }
    }
    if (fin_trace) {
        const auto t_s3 = std::chrono::steady_clock::now();
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[mmc finalize] %llu rows from the device (%s), %llu sparse records -> %zu rows merged on the host; host sort+reduce %.2f ms, wait for device %.2f ms, merge %.2f ms\n",
                (unsigned long long)n_dense, dev_sparse ? "dense + sparse, merged there" : "dense", sn, ns, ms(t_s0, t_s1), ms(t_s1, t_s2), ms(t_s2, t_s3));
    }
    *n_recs = n_dense + ns;
    *recs = *n_recs ? h_rows : nullptr;
    if (drain) {                                              // what this call returned is never returned again
        ctx->sparse_lo = sparse_hi;
        for (size_t i = 0; i < nc && (uint32_t)i <= wm_tid; ++i)
            ctx->drained_to[i] = std::max(ctx->drained_to[i], (uint32_t)i < wm_tid ? ctx->contigs[i].len : std::min(wm_pos, ctx->contigs[i].len));
        if (!ctx->wm_set || wm_key > (((uint64_t)ctx->wm_tid << 32) | ctx->wm_pos)) { ctx->wm_tid = wm_tid; ctx->wm_pos = wm_pos_given; }
        ctx->wm_set = true;
        ctx->drain_flip ^= 1;
    }
    return MMC_OK;
}

extern "C" {

int mmc_freq_finalize(mmc_ctx *ctx, const mmc_freq_rec_t **recs, uint64_t *n_recs) {
    MMC_DEV(ctx);
    if (!ctx || !recs || !n_recs) return MMC_EINVAL;
    if (ctx->opts.subtool != MMC_FREQ) return fail(ctx, MMC_ESTATE, "mmc_freq_finalize: context was created for view");
    if (ctx->drain_violated)
        return fail(ctx, MMC_EORDER, "a batch submitted after mmc_freq_drain() holds a read that starts before the drained watermark "
                                     "(input not coordinate-sorted): call mmc_freq_undrain(), drop the drained rows and finalize again");
    return finalize_rows(ctx, false, 0, 0, recs, n_recs);
}

int mmc_freq_drain(mmc_ctx *ctx, int32_t tid, uint32_t pos, const mmc_freq_rec_t **recs, uint64_t *n_recs) {
    MMC_DEV(ctx);
    if (!ctx || !recs || !n_recs || tid < 0) return MMC_EINVAL;
    if (ctx->opts.subtool != MMC_FREQ) return fail(ctx, MMC_ESTATE, "mmc_freq_drain: context was created for view");
    *recs = nullptr; *n_recs = 0;
    if (ctx->drain_violated) return MMC_OK;                   // nothing more leaves early; mmc_freq_finalize() reports it
    if (ctx->wm_set && (((uint64_t)(uint32_t)tid << 32) | pos) <= (((uint64_t)ctx->wm_tid << 32) | ctx->wm_pos)) return MMC_OK;
    return finalize_rows(ctx, true, (uint32_t)tid, pos, recs, n_recs);
}

int mmc_freq_undrain(mmc_ctx *ctx) {
    MMC_DEV(ctx);
    if (!ctx) return MMC_EINVAL;
    ctx->drained_to.assign(ctx->contigs.size(), 0u);
    ctx->wm_set = false; ctx->wm_tid = 0; ctx->wm_pos = 0; ctx->drain_violated = false; ctx->sparse_lo = 0;
    return MMC_OK;
}

int mmc_freq_reset(mmc_ctx *ctx) {
    MMC_DEV(ctx);
    if (!ctx) return MMC_EINVAL;
    int rc = mmc_sync(ctx);
    if (rc != MMC_OK) return rc;
    const size_t nc = ctx->contigs.size();
    CU(ctx, cudaStreamSynchronize(ctx->fin_stream));         // (a previous reset's copy out of reset_touch)
    std::vector<int32_t> &touch = ctx->reset_touch;
    touch.assign(2 * std::max<size_t>(1, nc), 0);
    if (nc) CU(ctx, cudaMemcpy(touch.data(), ctx->d_touch, sizeof(int32_t) * 2 * nc, cudaMemcpyDeviceToHost));
    const size_t spp = 2 * (size_t)ctx->n_code_slots * ctx->n_hap_slots;
    for (size_t i = 0; i < nc; ++i) {
        if (!ctx->contigs[i].loaded || !ctx->contigs[i].dev.cells) continue;
        int32_t lo = touch[i], hi = touch[nc + i];
        if (lo >= hi) continue;
        CU(ctx, cudaMemsetAsync(ctx->contigs[i].dev.cells + (size_t)lo * spp, 0, (size_t)(hi - lo) * spp * 8, ctx->fin_stream));
        touch[i] = INT32_MAX; touch[nc + i] = 0;
    }
    if (nc) CU(ctx, cudaMemcpyAsync(ctx->d_touch, touch.data(), sizeof(int32_t) * 2 * nc, cudaMemcpyHostToDevice, ctx->fin_stream));
    CU(ctx, cudaMemsetAsync(ctx->d_sparse_n, 0, 8, ctx->fin_stream));
    if (ctx->sparse_seen)                                  // slots that are not written read as sentinels (finalize_rows)
        CU(ctx, cudaMemsetAsync(ctx->d_sparse, 0xff, sizeof(SparseRec) * std::min<uint64_t>(ctx->sparse_seen, ctx->sparse_cap), ctx->fin_stream));
    // asynchronous: the next batches' H2D copies overlap the clearing; their kernels (launch_decode) and every later
    // finalize / drain (same stream) are ordered behind it on the device
    CU(ctx, cudaEventRecord(ctx->ev_reset, ctx->fin_stream));
    ctx->reset_pending = true;
    ctx->sparse_seen = 0;
    ctx->drained_to.assign(nc, 0u);
    ctx->wm_set = false; ctx->wm_tid = 0; ctx->wm_pos = 0; ctx->drain_violated = false; ctx->sparse_lo = 0;
    return MMC_OK;
}

int mmc_view_fetch(mmc_ctx *ctx, mmc_batch_t *batch, const mmc_view_rec_t **recs, uint64_t *n_recs) {
    MMC_DEV(ctx);
    Slot *s = ctx ? slot_of(ctx, batch) : nullptr;
    if (!s || !recs || !n_recs) return MMC_EINVAL;
    if (ctx->opts.subtool != MMC_VIEW) return fail(ctx, MMC_ESTATE, "mmc_view_fetch: context was created for freq");
    int rc = wait_slot(ctx, *s);
    if (rc != MMC_OK) return rc;
    rc = refresh_code_names(ctx);
    if (rc != MMC_OK) return rc;
    uint64_t n = s->n_reads_submitted ? s->h_state[1] : 0;
    std::vector<ViewDev> raw(n);
    if (n) {
        CU(ctx, cudaMemcpy(raw.data(), s->d_view, sizeof(ViewDev) * n, cudaMemcpyDeviceToHost));
        ctx->tm.d2h_bytes += sizeof(ViewDev) * n;
    }
    // first-wins de-duplication on (read, ref_pos, code string, uint16 ins_offset): add_view_entry(), src/mod.c:931-946
    std::sort(raw.begin(), raw.end(), [](const ViewDev &x, const ViewDev &y) {
        if (x.read != y.read) return x.read < y.read;
        if (x.ref_pos != y.ref_pos) return x.ref_pos < y.ref_pos;
        if (x.code != y.code) return x.code < y.code;
        uint32_t xi = x.ins_off & 0xffffu, yi = y.ins_off & 0xffffu;
        if (xi != yi) return xi < yi;
        return x.order < y.order;
    });
    s->view_out.clear();
    for (size_t i = 0; i < raw.size(); ++i) {
        if (i && raw[i].read == raw[i - 1].read && raw[i].ref_pos == raw[i - 1].ref_pos && raw[i].code == raw[i - 1].code &&
            (raw[i].ins_off & 0xffffu) == (raw[i - 1].ins_off & 0xffffu))
            continue;
        mmc_view_rec_t v;
        v.read = raw[i].read; v.ref_pos = raw[i].ref_pos; v.read_pos = raw[i].read_pos; v.ins_offset = raw[i].ins_off;
        v.code = raw[i].code; v.mod_prob = raw[i].prob;
        v.strand = (uint8_t)((batch->flag[v.read] >> 4) & 1u); v.hp = batch->hp[v.read];
        s->view_out.push_back(v);
    }
    *recs = s->view_out.data();
    *n_recs = s->view_out.size();
    return MMC_OK;
}

int mmc_dense_slice(mmc_ctx *ctx, int32_t tid, uint32_t start, uint32_t end, void **dev_ptr, uint64_t *n_cells) {
    if (!ctx || !dev_ptr || !n_cells) return MMC_EINVAL;
    if (tid < 0 || (size_t)tid >= ctx->contigs.size() || !ctx->contigs[tid].loaded || !ctx->contigs[tid].dev.cells)
        return fail(ctx, MMC_EINVAL, "mmc_dense_slice: contig %d has no dense counts", tid);
    if (start > end || end > ctx->contigs[tid].len) return fail(ctx, MMC_EINVAL, "mmc_dense_slice: bad range");
    { MMC_DEV(ctx); int rc = reset_settle(ctx); if (rc != MMC_OK) return rc; }
    const size_t spp = 2 * (size_t)ctx->n_code_slots * ctx->n_hap_slots;
    *dev_ptr = ctx->contigs[tid].dev.cells + (size_t)start * spp;
    *n_cells = (uint64_t)(end - start) * spp;
    return MMC_OK;
}

int mmc_dense_touch(mmc_ctx *ctx, int32_t tid, uint32_t start, uint32_t end) {
    MMC_DEV(ctx);
    if (!ctx) return MMC_EINVAL;
    if (tid < 0 || (size_t)tid >= ctx->contigs.size() || !ctx->contigs[tid].loaded) return fail(ctx, MMC_EINVAL, "mmc_dense_touch: bad contig %d", tid);
    if (start >= end || end > ctx->contigs[tid].len) return fail(ctx, MMC_EINVAL, "mmc_dense_touch: bad range");
    int rc = mmc_sync(ctx);
    if (rc != MMC_OK) return rc;
    const size_t nc = ctx->contigs.size();
    int32_t lo = 0, hi = 0;
    CU(ctx, cudaMemcpy(&lo, ctx->d_touch + tid, 4, cudaMemcpyDeviceToHost));
    CU(ctx, cudaMemcpy(&hi, ctx->d_touch + nc + tid, 4, cudaMemcpyDeviceToHost));
    lo = std::min<int32_t>(lo, (int32_t)start); hi = std::max<int32_t>(hi, (int32_t)end);
    CU(ctx, cudaMemcpy(ctx->d_touch + tid, &lo, 4, cudaMemcpyHostToDevice));
    CU(ctx, cudaMemcpy(ctx->d_touch + nc + tid, &hi, 4, cudaMemcpyHostToDevice));
    return MMC_OK;
}

// cells of a boundary region: dst += src (contexts that share a device), and the clearing of the non-owners' copies
__global__ void k_cells_add(unsigned long long *dst, const unsigned long long *src, unsigned long long n) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) dst[i] += src[i];
}

int mmc_region_reduce(mmc_ctx *const *ctxs, int32_t n_ctx, int32_t tid, double *ms_out, uint64_t *bytes_out) {
    if (ms_out) *ms_out = 0;
    if (bytes_out) *bytes_out = 0;
    if (!ctxs || n_ctx < 1 || !ctxs[0]) return MMC_EINVAL;
    mmc_ctx *c0 = ctxs[0];
    if (n_ctx == 1) return MMC_OK;
    if (tid < 0 || (size_t)tid >= c0->contigs.size()) return fail(c0, MMC_EINVAL, "mmc_region_reduce: bad contig %d", tid);
    const uint64_t len = c0->contigs[tid].len;
    std::vector<uint32_t> lo(n_ctx), hi(n_ctx);
    bool same_device = true;
    for (int k = 0; k < n_ctx; ++k) {
        mmc_ctx *c = ctxs[k];
        if (!c || c->opts.subtool != MMC_FREQ || c->contigs.size() != c0->contigs.size() || !c->contigs[tid].loaded || !c->contigs[tid].dev.cells ||
            c->n_code_slots != c0->n_code_slots || c->n_hap_slots != c0->n_hap_slots)
            return fail(c0, MMC_EINVAL, "mmc_region_reduce: context %d does not hold dense counts of contig %d with the same layout", k, tid);
        CU(c0, cudaSetDevice(c->opts.device));
        int rc = mmc_touched_range(c, tid, &lo[k], &hi[k]);
        if (rc != MMC_OK) { c0->err = c->err; return rc; }
        if (c->opts.device != c0->opts.device) same_device = false;
    }
    const size_t spp = 2 * (size_t)c0->n_code_slots * c0->n_hap_slots;
    auto slice_start = [&](int k) -> uint64_t { return ((uint64_t)k * len + (uint64_t)n_ctx - 1) / (uint64_t)n_ctx; };   // first p with p*n/len == k
    struct Region { int owner; uint32_t s, e; };
    std::vector<Region> regions;
    uint32_t reach = 0;                                       // how far the reads of the contexts left of the boundary ran
    for (int j = 1; j < n_ctx; ++j) {
        if (hi[j - 1] > lo[j - 1]) reach = std::max(reach, hi[j - 1]);
        const uint64_t s = slice_start(j), e = std::min<uint64_t>(std::min<uint64_t>(j + 1 < n_ctx ? slice_start(j + 1) : len, reach), len);
        if (e > s) regions.push_back({j, (uint32_t)s, (uint32_t)e});
    }
    if (regions.empty()) return MMC_OK;
#ifndef MMC_EMUL
    typedef struct ncclComm *comm_t;
    typedef int (*init_all_t)(comm_t *, int, const int *);
    typedef int (*allreduce_t)(const void *, void *, size_t, int, int, comm_t, cudaStream_t);
    typedef int (*void_t)(void);
    typedef int (*destroy_t)(comm_t);
    typedef const char *(*errstr_t)(int);
    void *h = nullptr;
    init_all_t p_init = nullptr; allreduce_t p_ar = nullptr; void_t p_gs = nullptr, p_ge = nullptr; destroy_t p_destroy = nullptr; errstr_t p_err = nullptr;
    std::vector<comm_t> comms(n_ctx, nullptr);
    if (!same_device) {
        // stdout is the tool's data channel.  NCCL's debug lines follow NCCL_DEBUG_FILE -- except at NCCL_DEBUG=VERSION (what
        // this image exports), where the "NCCL version ..." line is printed to stdout regardless: that level is raised to WARN
        // (same line, now through the debug file), and stdout itself points at stderr while the communicators are created.
        setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
        { const char *lv = getenv("NCCL_DEBUG"); if (lv && !strcasecmp(lv, "VERSION")) setenv("NCCL_DEBUG", "WARN", 1); }
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) return fail(c0, MMC_ECUDA, "mmc_region_reduce: contexts on several devices need NCCL, but libnccl.so.2 cannot be loaded (%s)", dlerror());
        p_init = (init_all_t)dlsym(h, "ncclCommInitAll"); p_ar = (allreduce_t)dlsym(h, "ncclAllReduce");
        p_gs = (void_t)dlsym(h, "ncclGroupStart"); p_ge = (void_t)dlsym(h, "ncclGroupEnd");
        p_destroy = (destroy_t)dlsym(h, "ncclCommDestroy"); p_err = (errstr_t)dlsym(h, "ncclGetErrorString");
        if (!p_init || !p_ar || !p_gs || !p_ge || !p_destroy) return fail(c0, MMC_ECUDA, "mmc_region_reduce: libnccl.so.2 lacks the expected symbols");
        std::vector<int> devs(n_ctx);
        for (int k = 0; k < n_ctx; ++k) devs[k] = ctxs[k]->opts.device;
        fflush(stdout);
        const int saved_out = dup(1);
        if (saved_out >= 0) dup2(2, 1);
        const int rc = p_init(comms.data(), n_ctx, devs.data());
        if (saved_out >= 0) { fflush(stdout); dup2(saved_out, 1); close(saved_out); }
        if (rc != 0) return fail(c0, MMC_ECUDA, "ncclCommInitAll failed: %s", p_err ? p_err(rc) : "?");
    }
#endif
    const auto t0 = std::chrono::steady_clock::now();
    uint64_t bytes = 0;
    for (const Region &R : regions) {
        const size_t n_cells = (size_t)(R.e - R.s) * spp;
        bytes += n_cells * 8;
        if (same_device) {
            mmc_ctx *own = ctxs[R.owner];
            for (int k = 0; k < n_ctx; ++k) {
                if (k == R.owner) continue;
                const unsigned grid = (unsigned)std::min<size_t>((n_cells + 255) / 256, (size_t)own->sm_count * 16);
                MMC_LAUNCH(k_cells_add, grid, 256u, own->fin_stream, own->contigs[tid].dev.cells + (size_t)R.s * spp,
                           (const unsigned long long *)(ctxs[k]->contigs[tid].dev.cells + (size_t)R.s * spp), (unsigned long long)n_cells);
                CU(c0, cudaGetLastError());
            }
            CU(c0, cudaStreamSynchronize(own->fin_stream));
        }
#ifndef MMC_EMUL
        else {
            int rc = p_gs();
            for (int k = 0; k < n_ctx && rc == 0; ++k) {
                unsigned long long *cells = ctxs[k]->contigs[tid].dev.cells + (size_t)R.s * spp;
                rc = p_ar(cells, cells, n_cells, 5 /* ncclUint64: n_called and n_mod never carry into each other */, 0 /* ncclSum */, comms[k], ctxs[k]->fin_stream);
            }
            const int rc2 = p_ge();
            if (rc != 0 || rc2 != 0) return fail(c0, MMC_ECUDA, "ncclAllReduce failed: %s", p_err ? p_err(rc ? rc : rc2) : "?");
            for (int k = 0; k < n_ctx; ++k) { CU(c0, cudaSetDevice(ctxs[k]->opts.device)); CU(c0, cudaStreamSynchronize(ctxs[k]->fin_stream)); }
        }
#endif
        for (int k = 0; k < n_ctx; ++k) {                    // the owner keeps the sums, every other copy is cleared
            mmc_ctx *c = ctxs[k];
            CU(c0, cudaSetDevice(c->opts.device));
            if (k == R.owner) { int rc = mmc_dense_touch(c, tid, R.s, R.e); if (rc != MMC_OK) { c0->err = c->err; return rc; } }
            else CU(c0, cudaMemsetAsync(c->contigs[tid].dev.cells + (size_t)R.s * spp, 0, n_cells * 8, c->fin_stream));
        }
        for (int k = 0; k < n_ctx; ++k) { CU(c0, cudaSetDevice(ctxs[k]->opts.device)); CU(c0, cudaStreamSynchronize(ctxs[k]->fin_stream)); }
    }
    const auto t1 = std::chrono::steady_clock::now();
#ifndef MMC_EMUL
    if (!same_device) { for (comm_t cm : comms) if (cm) p_destroy(cm); }
#endif
    if (ms_out) *ms_out = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (bytes_out) *bytes_out = bytes;
    CU(c0, cudaSetDevice(c0->opts.device));
    return MMC_OK;
}

int mmc_touched_range(mmc_ctx *ctx, int32_t tid, uint32_t *lo, uint32_t *hi) {
    MMC_DEV(ctx);
    if (!ctx || !lo || !hi) return MMC_EINVAL;
    if (tid < 0 || (size_t)tid >= ctx->contigs.size()) return fail(ctx, MMC_EINVAL, "mmc_touched_range: bad contig %d", tid);
    int rc = mmc_sync(ctx);
    if (rc != MMC_OK) return rc;
    const size_t nc = ctx->contigs.size();
    int32_t l = 0, h = 0;
    CU(ctx, cudaMemcpy(&l, ctx->d_touch + tid, 4, cudaMemcpyDeviceToHost));
    CU(ctx, cudaMemcpy(&h, ctx->d_touch + nc + tid, 4, cudaMemcpyDeviceToHost));
    if (l >= h) { *lo = 0; *hi = 0; } else { *lo = (uint32_t)l; *hi = (uint32_t)h; }
    return MMC_OK;
}

const char *mmc_describe(mmc_ctx *ctx) {
    if (!ctx) return "";
    char buf[256];
    bool any_stream = false, any_split = false;
    for (const Slot &sl : ctx->slots) { if (sl.uploaded || sl.in_flight || sl.n_reads_submitted) { if (sl.use_stream) any_stream = true; else any_split = true; } }
    if (ctx->stream_path && any_stream && !any_split)
        snprintf(buf, sizeof(buf), "k_flat_setup + k_decode_stream<%d> (dominant) + the fallback kernels k_decode_warp<%d,0> / k_decode for deferred reads",
                 ctx->s_minb, ctx->w_minb);
    else if (ctx->stream_path && any_stream)
        snprintf(buf, sizeof(buf), "k_flat_setup + k_decode_stream<%d> or k_decode_warp<MINB,PRE> per batch (by read shape) + fallback kernels", ctx->s_minb);
    else if (ctx->warp_path && ctx->split_path)
        snprintf(buf, sizeof(buf), "k_flat_setup + k_decode_warp<MINB,PRE> (dominant; MINB per batch, default %d) + k_decode_warp<MINB,0> / k_decode for deferred reads", ctx->w_minb);
    else if (ctx->warp_path) snprintf(buf, sizeof(buf), "k_decode_warp<MINB,0> + k_decode for deferred reads");
    else snprintf(buf, sizeof(buf), "k_decode (CTA per read)");
    ctx->desc = buf;
    return ctx->desc.c_str();
}

int mmc_get_timers(mmc_ctx *ctx, mmc_timers_t *out) {
    if (!ctx || !out) return MMC_EINVAL;
    *out = ctx->tm;
    return MMC_OK;
}

int mmc_reset_timers(mmc_ctx *ctx) {
    if (!ctx) return MMC_EINVAL;
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    return MMC_OK;
}

}  // extern "C"
