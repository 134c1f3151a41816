/*
 * minimod_cuda.h -- C ABI of libminimod_cuda.so
 *
 * The B200-native drop-in for minimod's per-read modification-decode and
 * frequency-aggregation hot path.  Every entry point below names the reference
 * interface it replaces (paths are into warp9seq/minimod v0.5.0).
 *
 * What stays on the host (reference behaviour, unchanged): BAM/BGZF decoding,
 * the load_db() read filters and -K/-B batching (src/minimod.c:235-333), FASTA
 * parsing (src/ref.c:46-89, kseq), -c / -m parsing (src/mod.c:204-398) and
 * text formatting (src/mod.c:560-728).
 *
 * What moves behind this ABI: freq_view_single() + get_aln() (src/mod.c:776-881,
 * 948-1370), update_freq_map()/add_view_entry() (src/mod.c:883-946),
 * merge_freq_maps() (src/mod.c:743-774), the (contig,pos) sort of
 * print_freq_output() (src/mod.c:644-664), load_ref()'s upper-casing and
 * load_ref_contexts() (src/ref.c:72-78,177-229), and the work_db() thread pool
 * (src/thread.c:100-158).
 *
 * Conventions: plain C, no exceptions, no exit().  Every function returns
 * MMC_OK (0) or a negative MMC_E* code; mmc_strerror(ctx) returns the message
 * for the most recent failure on that context, worded like the reference's own
 * stderr text where the reference has one, so the host can print it in
 * "[func::ERROR]" style and exit(EXIT_FAILURE) as the reference does
 * (src/error.h:94-152).  There is no CPU fallback: if no CUDA device is usable,
 * mmc_create() fails.
 *
 * Threading: a context is used from one host thread at a time (its error string,
 * timers and slot states are not synchronised); the pinned arrays of an acquired
 * batch may be FILLED by any thread, which is where the host's time goes.  Several
 * contexts (one per device) may be driven from different threads.
 */
#ifndef MINIMOD_CUDA_H
#define MINIMOD_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMC_ABI_VERSION 2

/* status codes */
#define MMC_OK            0
#define MMC_EINVAL       -1   /* bad argument / unsupported option combination        */
#define MMC_ECUDA        -2   /* CUDA runtime failure (message has the CUDA error)    */
#define MMC_ENOMEM       -3   /* host or device allocation failed / capacity exceeded */
#define MMC_EREAD        -4   /* a read is malformed in a way the reference treats as fatal */
#define MMC_ESTATE       -5   /* call sequence error                                  */
#define MMC_EORDER       -6   /* rows were drained early and a later batch was not in coordinate order */

/* subtool: enum subtool, src/minimod.h:88 */
#define MMC_VIEW 0
#define MMC_FREQ 1

/* bits of mmc_mod_t.call_lut[] */
#define MMC_LUT_CALLED 1u
#define MMC_LUT_MOD    2u

#define MMC_MAX_CODE_LEN 8    /* bytes of a modification code string (letters or ChEBI digits) */
#define MMC_MAX_CONTEXT  32   /* bytes of a context string */
#define MMC_MAX_MODS     64   /* -c entries */
#define MMC_ALIGN        16   /* every per-read slice in the flat pools starts 16-byte aligned */

typedef struct mmc_ctx mmc_ctx;

/* One "-c code[context]" entry with its "-m" threshold already applied:
 * replaces modcodem_t (src/minimod.h:60-64) and the khash it lives in.
 *   code      letters, ChEBI digits, or "*"             (parse_mod_codes, src/mod.c:204-326)
 *   context   upper-case, U->T, or "*"
 *   call_lut  FREQ only: for ML byte p, bit0 = counted in n_called, bit1 = counted in n_mod.
 *             Built by the host from the double threshold t with the reference's exact
 *             expressions (src/mod.c:56,1181-1191): x=(p+0.5)/256.0; x>=t -> called|mod;
 *             else x<=1-t -> called; else 0.
 */
typedef struct {
    char    code[MMC_MAX_CODE_LEN + 1];
    char    context[MMC_MAX_CONTEXT + 1];
    uint8_t call_lut[256];
} mmc_mod_t;

/* Replaces the parts of opt_t (src/minimod.h:92-121) the hot path reads. */
typedef struct {
    uint32_t struct_size;        /* sizeof(mmc_opts_t), for ABI checking */
    int32_t  subtool;            /* MMC_FREQ or MMC_VIEW */
    int32_t  n_mods;             /* opt_t.n_mods */
    const mmc_mod_t *mods;       /* index order == modcodem_t.index */
    int32_t  insertions;         /* --insertions */
    int32_t  haplotypes;         /* --haplotypes */
    int32_t  device;             /* CUDA device ordinal */
    int32_t  n_slots;            /* batches in flight (>=1); 0 -> 3 */
    uint64_t max_reads;          /* per-batch capacity in reads  (-K) */
    uint64_t max_bytes;          /* per-batch capacity in payload bytes (-B): cigar+seq+mm+ml pools */
    uint64_t sparse_capacity;    /* records in the side buffer for cells outside the dense arrays
                                    (ins_offset>0, HP>=dense_haps, code id>=dense_codes); 0 -> default */
    int32_t  dense_haps;         /* haplotype values 0..dense_haps-1 get dense strata; 0 -> 3 ('*' + HP 0,1,2: one 32-byte sector) */
    int32_t  dense_codes;        /* with a wildcard code: dense code slots; 0 -> 8 */
    uint64_t view_capacity;      /* VIEW: records per batch; 0 -> derived from max_bytes */
    /* optional explicit pool capacities per batch (0 -> derived from max_bytes) */
    uint64_t cap_cigar_words, cap_seq_bytes, cap_mm_bytes, cap_ml_bytes;
    int32_t  seq_packing;        /* how SEQ crosses PCIe: 0 or 4 -> BAM's 4-bit nibbles (batch->seq4);
                                    2 -> 2 bits per base + an exception list (batch->seq2 / seq_exc), expanded
                                    to the 4-bit form on the device.  SEQ is ~87 % of a batch's bytes. */
    int32_t  cigar_packing;      /* how CIGARs cross PCIe: 0 or 32 -> BAM's 32-bit words (batch->cigar); 8 -> the byte form
                                    (batch->cig8), expanded to the words on the device.  ONT alignments have an op per
                                    ~17 bases (a third of a whole-genome batch's bytes as words, ~1.2 bytes per op here). */
} mmc_opts_t;

/* A batch in flight.  Replaces the per-read fields of db_t (src/minimod.h:125-160) that
 * load_db() fills: the arrays are the library's pinned staging memory, written by the host
 * packer, read by the device.  Per-read i:
 *   tid,pos,flag,l_seq,n_cigar  bam1_core_t fields
 *   hp                          (uint8_t)bam_aux2i(HP) or 0       (get_hp_tag, src/mod.c:188-202)
 *   cigar[cigar_off[i] ..+n_cigar[i]]         BAM CIGAR words (len<<4|op)
 *   seq4[seq_off[i] ..+(l_seq[i]+1)/2]        BAM 4-bit packed SEQ
 *   mm[mm_off[i] ..+mm_len[i]]                MM:Z text without the NUL  (get_mm_tag_ptr, src/mod.c:123-140)
 *   ml[ml_off[i] ..+ml_len[i]]                ML:B:C bytes; ml_len 0 if absent/not B,C (get_ml_tag, src/mod.c:142-185)
 * With seq_packing == 2 the packer writes, instead of seq4,
 *   seq2[seq_off[i]/2 ..+(l_seq[i]+3)/4]      4 bases per byte, first base in bits 7:6; A 0, C 1, G 2, T 3
 *   seq_exc[0 .. seq_exc_used)                every base that is not A,C,G,T (stored as 0 in seq2) and the pad
 *                                             nibble of an odd-length read: (nibble index into the seq4 pool,
 *                                             i.e. 2*seq_off[i]+base) << 4 | BAM nt16 code (0 for the pad)
 * and still advances seq_used in seq4 bytes; the device rebuilds the identical 4-bit pool.
 * With cigar_packing == 8 the packer writes, instead of cigar, one blob per read at cig8[cig8_off[i]] (16-byte aligned):
 *   u32 n1                                    ops whose length does not fit the nibble
 *   n_cigar[i] bytes, padded to 4             op | min(len,15) << 4             (15: the length is in the next list)
 *   n1 bytes, padded to 4                     len - 15 if < 255, else 255       (255: the length is in the next list)
 *   n2 x u32                                  len, for the ops marked 255, in op order
 * and still sets cigar_off[i] / advances cigar_used in words; the device rebuilds the identical word pool.
 * Offsets are element indices into their pool (cigar: words; others: bytes); each must be a
 * multiple of MMC_ALIGN bytes, and every pool needs MMC_ALIGN bytes of slack after the last
 * slice (mmc_batch_acquire() sizes them so).  The host appends reads while
 * n_reads<max_reads and the *_used counters stay within the *_cap capacities.
 */
typedef struct {
    uint32_t  n_reads;
    uint32_t  max_reads;
    int32_t  *tid;
    int32_t  *pos;
    uint32_t *l_seq;
    uint32_t *n_cigar;
    uint32_t *mm_len;
    uint32_t *ml_len;
    uint64_t *cigar_off;
    uint64_t *seq_off;
    uint64_t *mm_off;
    uint64_t *ml_off;
    uint16_t *flag;
    uint8_t  *hp;
    uint32_t *cigar;  uint64_t cigar_cap, cigar_used;   /* in words */
    uint8_t  *seq4;   uint64_t seq_cap,   seq_used;     /* in bytes */
    char     *mm;     uint64_t mm_cap,    mm_used;
    uint8_t  *ml;     uint64_t ml_cap,    ml_used;
    void     *priv;   /* library private */
    uint8_t  *seq2;                                     /* seq_packing == 2 only */
    uint64_t *seq_exc; uint64_t seq_exc_cap, seq_exc_used;
    uint32_t  seq_packing;                              /* 4 or 2: which of seq4 / seq2 the packer fills */
    uint32_t  cigar_packing;                            /* 32 or 8: which of cigar / cig8 the packer fills */
    uint8_t  *cig8;   uint64_t cig8_cap,  cig8_used;    /* cigar_packing == 8 only; in bytes */
    uint64_t *cig8_off;
} mmc_batch_t;

/* One output row of `freq`: the decoded key + value of core->freq_map
 * (make_key/decode_key, src/mod.c:428-457; freq_t, src/minimod.h:70-73). */
typedef struct {
    int32_t  tid;
    int32_t  pos;
    uint32_t n_called;
    uint32_t n_mod;
    uint16_t ins_offset;     /* already truncated to uint16 as in make_key() */
    int16_t  hap;            /* -1: haplotypes off, or the '*' aggregate row */
    uint8_t  strand;         /* 0 '+', 1 '-' */
    uint8_t  code;           /* mmc_code_name(ctx, code) */
    uint16_t reserved;
} mmc_freq_rec_t;

/* One output row of `view`: key + view_t (src/minimod.h:75-78) of db->view_maps[read]. */
typedef struct {
    uint32_t read;           /* index of the read in its batch */
    int32_t  ref_pos;
    int32_t  read_pos;       /* FASTQ-orientation position, as printed */
    uint32_t ins_offset;     /* untruncated, as print_view_output() prints it (src/mod.c:608) */
    uint8_t  code;
    uint8_t  mod_prob;       /* ML byte; 0 for implicit ('.') calls */
    uint8_t  strand;
    uint8_t  hp;
} mmc_view_rec_t;

/* Device-side stage timers (milliseconds, CUDA events on the batch's stream), summed since
 * create or the last mmc_reset_timers().  Mirrors core_t's process/merge timers
 * (src/minimod.h:183-187). */
typedef struct {
    double   h2d_ms;         /* host->device copies of batch payloads */
    double   decode_ms;      /* decode+aggregate kernels (process_db + merge_db equivalent) */
    double   finalize_ms;    /* dense scan/compaction kernels (merge + sort equivalent) */
    double   d2h_ms;         /* result read-back */
    uint64_t batches;
    uint64_t reads;
    uint64_t kernel_launches; /* launches of this library's own kernels */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t deferred_reads; /* reads the warp-per-read kernel handed to the general CTA-per-read kernel */
    uint64_t flat_deferred_reads; /* reads the flat kernel chain handed to the warp-per-read kernel */
} mmc_timers_t;

/* ---- lifecycle: replaces init_core()/free_core() (src/minimod.c:51-161) -------------- */
int  mmc_create(mmc_ctx **out, const mmc_opts_t *opts,
                int32_t n_contigs, const char *const *names, const uint32_t *lens);
void mmc_destroy(mmc_ctx *ctx);
const char *mmc_strerror(const mmc_ctx *ctx);     /* NULL ctx -> message of a failed mmc_create() */
int  mmc_abi_version(void);

/* ---- reference: replaces load_ref()'s per-base normalisation, load_ref_contexts() and
 *      get_ref() (src/ref.c:72-78,169-229).  `seq` is the raw FASTA sequence of header
 *      contig `tid` (any case, U allowed); its length must equal lens[tid] (the reference
 *      asserts this per read, src/mod.c:861).  Contigs never added are "not found in
 *      reference" for any read that maps to them (src/mod.c:792-793). ---------------- */
int  mmc_ref_add(mmc_ctx *ctx, int32_t tid, const char *seq, uint32_t len);
int  mmc_ref_commit(mmc_ctx *ctx);

/* ---- batches: replaces init_db()/load_db() storage, process_db(), work_db() and
 *      merge_db() (src/minimod.c:164-350,373-386; src/thread.c:145-158) ------------- */
int  mmc_batch_acquire(mmc_ctx *ctx, mmc_batch_t **batch);   /* waits for a free slot */
int  mmc_batch_submit(mmc_ctx *ctx, mmc_batch_t *batch);     /* async: H2D + kernels   */
int  mmc_batch_wait(mmc_ctx *ctx, mmc_batch_t *batch);       /* MMC_EREAD if a read was fatal */
int  mmc_batch_release(mmc_ctx *ctx, mmc_batch_t *batch);    /* slot reusable          */
/* measurement helpers: upload once, then re-run the kernels on HBM-resident inputs */
int  mmc_batch_upload(mmc_ctx *ctx, mmc_batch_t *batch);
int  mmc_batch_launch(mmc_ctx *ctx, mmc_batch_t *batch);
int  mmc_sync(mmc_ctx *ctx);

/* ---- freq results: replaces output_core()'s collect+sort (src/mod.c:644-664).  Rows come
 *      back ordered by (tid, pos, strand, code, ins_offset, hap) with hap -1 first; the
 *      host orders contigs by strcmp as cmp_key_fast() does.  The array is owned by the
 *      context and valid until the next finalize/reset/destroy. ----------------------- */
int  mmc_freq_finalize(mmc_ctx *ctx, const mmc_freq_rec_t **recs, uint64_t *n_recs);
int  mmc_freq_reset(mmc_ctx *ctx);                            /* zero all counts */
const char *mmc_code_name(const mmc_ctx *ctx, int32_t code);  /* mod_code string of a row */

/* ---- streaming results for coordinate-sorted input.  The reference prints after the last read
 *      (output_core, src/minimod.c:388; its BAMs are coordinate-sorted, test/test_ext.sh:63): here the rows of
 *      finished positions can leave while later batches are still being copied and decoded, so the read-back
 *      and the text formatting overlap the rest of the job and D2H shares no time with H2D.
 *      mmc_freq_drain(tid,pos): the caller states that every read it submits FROM NOW ON starts at or after
 *      (tid,pos) in coordinate order (tid, then pos).  Returns the rows before that watermark that no earlier
 *      drain returned, in the order mmc_freq_finalize() uses; waits only for the batches that hold a read
 *      starting before the watermark.  The rows stay valid until the second-next drain/finalize/reset (two
 *      pinned buffers alternate).  Rows leave up to one position below the watermark: a read that starts exactly
 *      at it may still add a side-buffer record at the position before (a leading insertion), and the rows of
 *      a position leave together.  Side-buffer records (--insertions, haplotypes >= dense_haps, code ids >=
 *      dense_codes) below the watermark are sorted, reduced and merged into the drained rows like at the end.
 *      After drains, mmc_freq_finalize() returns the REMAINING rows: concatenated, the drains and the
 *      remainder are the single-call table.  Counts are never cleared by a drain, so a broken promise loses
 *      nothing: the library notices a later batch that starts before the watermark, mmc_freq_finalize() then
 *      fails with MMC_EORDER, and after mmc_freq_undrain() it returns the complete table (the caller drops the
 *      rows it drained). ------------------------------------------------------------------ */
int  mmc_freq_drain(mmc_ctx *ctx, int32_t tid, uint32_t pos, const mmc_freq_rec_t **recs, uint64_t *n_recs);
int  mmc_freq_undrain(mmc_ctx *ctx);

/* ---- view results: replaces output_db()'s per-read collect (src/mod.c:560-593).  Rows of
 *      the batch, ordered by (read, ref_pos, code, ins_offset) after first-wins
 *      de-duplication (add_view_entry, src/mod.c:931-946).  Valid until release. -------- */
int  mmc_view_fetch(mmc_ctx *ctx, mmc_batch_t *batch, const mmc_view_rec_t **recs, uint64_t *n_recs);

/* ---- multi-GPU: raw dense count cells of positions [start,end) of contig tid, for the
 *      boundary/halo reduce across region-sharded devices (no reference equivalent).
 *      Each cell is a uint64: n_called in the low 32 bits, n_mod in the high 32. ------- */
int  mmc_dense_slice(mmc_ctx *ctx, int32_t tid, uint32_t start, uint32_t end,
                     void **dev_ptr, uint64_t *n_cells);
/* tell the library that [start,end) of contig tid now holds counts written from outside
 * (the result of such a reduce), so that finalize scans it */
int  mmc_dense_touch(mmc_ctx *ctx, int32_t tid, uint32_t start, uint32_t end);
/* [lo,hi) of contig tid that any submitted read (or mmc_dense_touch) has covered so far; lo>=hi: nothing */
int  mmc_touched_range(mmc_ctx *ctx, int32_t tid, uint32_t *lo, uint32_t *hi);

/* ---- multi-GPU in ONE process (the reference is one process, src/freq_main.c:404-474): ctxs[0..n) -- one context per
 *      device, created over the same contig table -- hold the reads that START in consecutive slices of contig tid
 *      (context k: positions p with p*n/len == k).  A read may run past its slice, so its counts sit in its own context's
 *      copy of the next slices' cells.  This call sums those boundary cells across the contexts -- ncclAllReduce(sum,
 *      uint64) over NVLink when the contexts are on different devices (libnccl.so.2, loaded on demand), a device-local
 *      add when they share one -- leaves every position's total with the context that owns it and zeroes the other
 *      copies, so that the union of the contexts' mmc_freq_finalize() rows is the single-device table (sparse rows of a
 *      boundary stay where they were counted: add rows with equal keys).  ms / bytes (optional): device time and volume. */
int  mmc_region_reduce(mmc_ctx *const *ctxs, int32_t n_ctx, int32_t tid, double *ms, uint64_t *bytes);

int  mmc_get_timers(mmc_ctx *ctx, mmc_timers_t *out);
int  mmc_reset_timers(mmc_ctx *ctx);
/* one line naming the kernels the decode stage of this context launches (measurement reports) */
const char *mmc_describe(mmc_ctx *ctx);
/* per-launch device time (ms) of the most recent decode kernel on `batch` */
int  mmc_last_decode_ms(mmc_ctx *ctx, mmc_batch_t *batch, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* MINIMOD_CUDA_H */
