/* minimod_cuda_glue.c -- the INTEGRATION.md binding, made real.
 *
 * Linked with the UNMODIFIED reference sources (warp9seq/minimod v0.5.0, src/ *.c) it replaces the bodies of the batch
 * operators the reference's drivers call -- process_db(), merge_db(), output_db(), output_core()
 * (src/minimod.c:344,354,373,388) -- and load_ref_contexts() (src/ref.c:177) with calls into libminimod_cuda.so.
 * Everything else stays the reference's own code: option parsing (freq_main.c / view_main.c), load_ref()'s kseq loop,
 * init_core(), load_db()'s htslib loop and read filters, the load || process || merge pthread pipeline, the statistics.
 * The reference's own definitions of the five functions are renamed at compile time (-Dprocess_db=ref__process_db ...,
 * see oracle/Makefile: target _ref/minimod_ref_cuda); no reference source file is edited or copied.
 *
 * The reference calls process_db() and merge_db()/output_db() from different pthreads (src/freq_main.c:93-164); the
 * library is not re-entrant per context, so every call is taken under one mutex here.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "minimod.h"
#include "mod.h"
#include "ref.h"
#include "error.h"
#include "misc.h"
#include "khash.h"
#include "minimod_cuda.h"

KHASH_MAP_INIT_STR(refm, ref_t *)
extern khash_t(refm) *ref_map;                              /* src/ref.c:42 */
void ref__process_db(core_t *core, db_t *db);               /* the reference's own versions (summary subtool) */
void ref__output_db(core_t *core, db_t *db);

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static mmc_ctx *g_ctx = NULL;
static mmc_mod_t g_mods[MMC_MAX_MODS];
#define GLUE_SLOTS 8
static struct { db_t *db; mmc_batch_t *b; } g_slot[GLUE_SLOTS];

#define DIE(...) do { ERROR(__VA_ARGS__); exit(EXIT_FAILURE); } while (0)

/* load_ref_contexts(): contexts are evaluated on the device from the packed reference (mmc_ref_add), so no byte maps are
 * built; destroy_ref() still frees the per-mod pointer arrays, so they exist (empty). */
void load_ref_contexts(int n_mod_codes, char **mod_contexts) {
    (void)mod_contexts;
    for (khiter_t k = kh_begin(ref_map); k != kh_end(ref_map); ++k) {
        if (!kh_exist(ref_map, k)) continue;
        ref_t *ref = kh_value(ref_map, k);
        ref->is_context = (uint8_t **)calloc(n_mod_codes, sizeof(uint8_t *));
        ref->is_context_rev = (uint8_t **)calloc(n_mod_codes, sizeof(uint8_t *));
        MALLOC_CHK(ref->is_context); MALLOC_CHK(ref->is_context_rev);
    }
}

/* -c / -m tables -> flat device options; context + reference -> device (once, on the first batch) */
static void glue_init(core_t *core) {
    opt_t *opt = &core->opt;
    if (opt->n_mods > MMC_MAX_MODS) DIE("too many modification codes for libminimod_cuda (%d)", opt->n_mods);
    memset(g_mods, 0, sizeof g_mods);
    for (khint_t k = kh_begin(opt->modcodes_map); k < kh_end(opt->modcodes_map); ++k) {
        if (!kh_exist(opt->modcodes_map, k)) continue;
        modcodem_t *m = kh_value(opt->modcodes_map, k);
        mmc_mod_t *o = &g_mods[m->index];
        snprintf(o->code, sizeof o->code, "%s", kh_key(opt->modcodes_map, k));
        snprintf(o->context, sizeof o->context, "%s", m->context);
        for (int p = 0; p < 256; p++) {                      /* same doubles as src/mod.c:56,1181-1191 */
            double x = (double)((p + 0.5) / 256.0);
            o->call_lut[p] = x >= m->thresh ? (MMC_LUT_CALLED | MMC_LUT_MOD) : x <= 1 - m->thresh ? MMC_LUT_CALLED : 0;
        }
    }
    mmc_opts_t mo;
    memset(&mo, 0, sizeof mo);
    mo.struct_size = sizeof mo;
    mo.subtool = opt->subtool == FREQ ? MMC_FREQ : MMC_VIEW;
    mo.n_mods = opt->n_mods; mo.mods = g_mods;
    mo.insertions = opt->insertions; mo.haplotypes = opt->haplotypes;
    mo.device = getenv("MMC_DEVICE") ? atoi(getenv("MMC_DEVICE")) : 0;
    mo.n_slots = 4;
    mo.max_reads = (uint64_t)opt->batch_size;
    mo.max_bytes = (uint64_t)opt->batch_size_bases;
    bam_hdr_t *h = core->bam_hdr;
    if (mmc_create(&g_ctx, &mo, h->n_targets, (const char *const *)h->target_name, h->target_len) != MMC_OK) DIE("%s", mmc_strerror(NULL));
    for (int tid = 0; tid < h->n_targets; ++tid) {
        ref_t *ref = get_ref(h->target_name[tid]);
        if (!ref) continue;                                  /* "Contig not found" only matters if a read maps there (src/mod.c:793) */
        if (mmc_ref_add(g_ctx, tid, ref->forward, (uint32_t)ref->ref_seq_length) != MMC_OK)
            WARNING("%s", mmc_strerror(g_ctx));              /* length mismatch: fatal only when a read maps there (src/mod.c:861) */
    }
    if (mmc_ref_commit(g_ctx) != MMC_OK) DIE("%s", mmc_strerror(g_ctx));
}

static uint8_t hp_of(bam1_t *rec) {                          /* get_hp_tag(), src/mod.c:188-202 */
    uint8_t *s = bam_aux_get(rec, "HP");
    return s ? (uint8_t)bam_aux2i(s) : 0;
}

#define UP16(x) (((x) + 15) & ~(uint64_t)15)
static int pack_read(mmc_batch_t *b, bam1_t *rec, const char *mm, const uint8_t *ml, uint32_t ml_len) {
    uint64_t c0 = UP16(b->cigar_used * 4) / 4, s0 = UP16(b->seq_used), m0 = UP16(b->mm_used), l0 = UP16(b->ml_used);
    uint32_t i = b->n_reads, mm_len = (uint32_t)strlen(mm), sb = ((uint32_t)rec->core.l_qseq + 1) / 2;
    if (i >= b->max_reads || c0 + rec->core.n_cigar > b->cigar_cap || s0 + sb > b->seq_cap || m0 + mm_len > b->mm_cap ||
        l0 + ml_len > b->ml_cap)
        return 0;
    b->tid[i] = rec->core.tid; b->pos[i] = rec->core.pos; b->flag[i] = rec->core.flag;
    b->l_seq[i] = (uint32_t)rec->core.l_qseq; b->n_cigar[i] = rec->core.n_cigar; b->mm_len[i] = mm_len; b->ml_len[i] = ml_len;
    b->hp[i] = hp_of(rec);
    b->cigar_off[i] = c0; b->seq_off[i] = s0; b->mm_off[i] = m0; b->ml_off[i] = l0;
    memcpy(b->cigar + c0, bam_get_cigar(rec), 4 * (size_t)rec->core.n_cigar);
    memcpy(b->seq4 + s0, bam_get_seq(rec), sb);
    memcpy(b->mm + m0, mm, mm_len);
    if (ml_len) memcpy(b->ml + l0, ml, ml_len);
    b->cigar_used = c0 + rec->core.n_cigar; b->seq_used = s0 + sb; b->mm_used = m0 + mm_len; b->ml_used = l0 + ml_len;
    b->n_reads = i + 1;
    return 1;
}

static mmc_batch_t *slot_take(db_t *db) {
    for (int i = 0; i < GLUE_SLOTS; ++i)
        if (g_slot[i].db == db) { mmc_batch_t *b = g_slot[i].b; g_slot[i].db = NULL; g_slot[i].b = NULL; return b; }
    return NULL;
}
static void slot_put(db_t *db, mmc_batch_t *b) {
    for (int i = 0; i < GLUE_SLOTS; ++i)
        if (!g_slot[i].db) { g_slot[i].db = db; g_slot[i].b = b; return; }
    DIE("%s", "glue: more batches in flight than slots");
}

/* a per-read fatal condition: the library names the read by its index in the batch; print it the reference's way */
static void die_read(db_t *db) {
    const char *msg = mmc_strerror(g_ctx);
    const char *sep = strchr(msg, '\x1f');
    if (sep) {
        long idx = atol(sep + 1);
        const char *colon = strstr(msg, ": ");
        if (idx >= 0 && idx < db->n_bam_recs)
            DIE("read %s: %.*s", bam_get_qname(db->bam_recs[idx]), (int)(sep - (colon ? colon + 2 : msg)), colon ? colon + 2 : msg);
    }
    DIE("%s", msg);
}

/* process_db(), src/minimod.c:344: pack the accepted reads of the batch into the pinned SoA and submit (asynchronous) */
void process_db(core_t *core, db_t *db) {
    if (core->opt.subtool == SUMMARY) { ref__process_db(core, db); return; }
    double t0 = realtime();
    pthread_mutex_lock(&g_mu);
    if (!g_ctx) glue_init(core);
    mmc_batch_t *b = NULL;
    if (mmc_batch_acquire(g_ctx, &b) != MMC_OK) DIE("%s", mmc_strerror(g_ctx));
    for (int i = 0; i < db->n_bam_recs; ++i) {
        bam1_t *rec = db->bam_recs[i];
        const uint8_t *ml = db->ml[i];
        if (!pack_read(b, rec, db->mm[i], ml, ml ? db->ml_lens[i] : 0))
            DIE("read %s does not fit the device batch buffers; raise -B", bam_get_qname(rec));
    }
    if (mmc_batch_submit(g_ctx, b) != MMC_OK) DIE("%s", mmc_strerror(g_ctx));
    slot_put(db, b);
    pthread_mutex_unlock(&g_mu);
    core->process_db_time += realtime() - t0;
}

/* merge_db(), src/minimod.c:373: counts are aggregated on the device; only the statistics remain */
void merge_db(core_t *core, db_t *db) {
    double t0 = realtime();
    pthread_mutex_lock(&g_mu);
    mmc_batch_t *b = slot_take(db);
    if (b && mmc_batch_release(g_ctx, b) != MMC_OK) die_read(db);
    pthread_mutex_unlock(&g_mu);
    core->total_reads += db->total_reads;
    core->total_bytes += db->total_bytes;
    core->processed_reads += db->n_bam_recs;
    core->processed_bytes += db->processed_bytes;
    core->merge_db_time += realtime() - t0;
}

/* output_db(), src/minimod.c:354 (view): rows of the batch with print_view_output()'s format (src/mod.c:606-614) */
void output_db(core_t *core, db_t *db) {
    if (core->opt.subtool != VIEW) { ref__output_db(core, db); return; }
    double t0 = realtime();
    pthread_mutex_lock(&g_mu);
    mmc_batch_t *b = slot_take(db);
    if (b) {
        const mmc_view_rec_t *r = NULL; uint64_t n = 0;
        if (mmc_view_fetch(g_ctx, b, &r, &n) != MMC_OK) die_read(db);
        FILE *out = core->opt.output_fp;
        for (uint64_t j = 0; j < n; ++j) {
            bam1_t *rec = db->bam_recs[r[j].read];
            fprintf(out, "%s\t%d\t%c\t%s\t%d\t%s\t%f", core->bam_hdr->target_name[rec->core.tid], r[j].ref_pos, r[j].strand ? '-' : '+',
                    bam_get_qname(rec), r[j].read_pos, mmc_code_name(g_ctx, r[j].code), (double)((r[j].mod_prob + 0.5) / 256.0));
            if (core->opt.insertions) fprintf(out, "\t%d", (int)r[j].ins_offset);
            if (core->opt.haplotypes) fprintf(out, "\t%d", (int)r[j].hp);
            fputc('\n', out);
        }
        if (mmc_batch_release(g_ctx, b) != MMC_OK) die_read(db);
    }
    pthread_mutex_unlock(&g_mu);
    core->total_reads += db->total_reads;
    core->total_bytes += db->total_bytes;
    core->processed_reads += db->n_bam_recs;
    core->processed_bytes += db->processed_bytes;
    core->output_time += realtime() - t0;
}

static bam_hdr_t *g_hdr_for_sort;
static int cmp_tid_by_name(const void *a, const void *b) {
    return strcmp(g_hdr_for_sort->target_name[*(const int *)a], g_hdr_for_sort->target_name[*(const int *)b]);
}

/* output_core(), src/minimod.c:388 -> print_freq_output(), src/mod.c:644-728: rows arrive ordered by
 * (tid, pos, strand, code, ins_offset, hap); contigs are printed in strcmp order like cmp_key_fast() (src/mod.c:59-87) */
void output_core(core_t *core) {
    if (core->opt.subtool != FREQ) return;
    pthread_mutex_lock(&g_mu);
    if (!g_ctx) { pthread_mutex_unlock(&g_mu); return; }    /* no batch at all: nothing to print (src/mod.c:648) */
    const mmc_freq_rec_t *r = NULL; uint64_t n = 0;
    double s0 = realtime();
    if (mmc_freq_finalize(g_ctx, &r, &n) != MMC_OK) DIE("%s", mmc_strerror(g_ctx));
    core->sort_time = realtime() - s0;
    double o0 = realtime();
    bam_hdr_t *h = core->bam_hdr;
    int nt = h->n_targets;
    int *order = (int *)malloc(sizeof(int) * (nt > 0 ? nt : 1));
    uint64_t *first = (uint64_t *)calloc((size_t)nt + 1, sizeof(uint64_t)), *count = (uint64_t *)calloc((size_t)nt + 1, sizeof(uint64_t));
    MALLOC_CHK(order); MALLOC_CHK(first); MALLOC_CHK(count);
    for (int i = 0; i < nt; ++i) order[i] = i;
    for (uint64_t i = 0; i < n; ++i) { if (count[r[i].tid]++ == 0) first[r[i].tid] = i; }
    g_hdr_for_sort = h;
    qsort(order, (size_t)nt, sizeof(int), cmp_tid_by_name);
    FILE *out = core->opt.output_fp;
    for (int oi = 0; oi < nt; ++oi) {
        int tid = order[oi];
        const char *contig = h->target_name[tid];
        for (uint64_t i = first[tid]; i < first[tid] + count[tid]; ++i) {
            const mmc_freq_rec_t *f = &r[i];
            const char *code = mmc_code_name(g_ctx, f->code);
            char strand = f->strand ? '-' : '+';
            if (core->opt.bedmethyl_out) {
                double v = (double)f->n_mod * 100 / f->n_called;
                fprintf(out, "%s\t%d\t%d\t%s\t%d\t%c\t%d\t%d\t255,0,0\t%d\t%f\n", contig, f->pos, f->pos + 1, code, f->n_called, strand, f->pos,
                        f->pos + 1, f->n_called, v);
            } else {
                double v = (double)f->n_mod / f->n_called;
                fprintf(out, "%s\t%d\t%d\t%c\t%d\t%d\t%f\t%s", contig, f->pos, f->pos, strand, f->n_called, f->n_mod, v, code);
                if (core->opt.insertions) fprintf(out, "\t%d", f->ins_offset);
                if (core->opt.haplotypes) { if (f->hap == -1) fputs("\t*", out); else fprintf(out, "\t%d", f->hap); }
                fputc('\n', out);
            }
        }
    }
    if (n && out != stdout) fclose(out);
    free(order); free(first); free(count);
    core->output_time += realtime() - o0;
    mmc_destroy(g_ctx);
    g_ctx = NULL;
    pthread_mutex_unlock(&g_mu);
}
