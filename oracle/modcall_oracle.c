/*
 * modcall_oracle.c -- TEST INFRASTRUCTURE ONLY: the CPU oracle of the CUDA hot path.
 *
 * A plain-C, single-threaded restatement of the reference's per-read modification decode and
 * frequency aggregation (warp9seq/minimod v0.5.0), written from the reference's behaviour and
 * following its control flow step by step so that it can be checked against it line by line.
 * Every function cites the reference lines it restates.  Nothing in the product (libminimod_cuda,
 * the `minimod` binary, minimod_b200 python package) may link, import or call this file: it is used by tests/,
 * by __graft_entry__.smoke() and by bench.py's cpu_baseline "port" leg only, as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_port.py checks this restatement against (a) all 22
 * golden files of the reference's own test-suite (test/test.sh:66-250) and (b) the unmodified
 * reference binary oracle/_ref/minimod_ref on the fixture BAMs and on synthetic inputs.
 *
 * Deliberate differences from the reference's *mechanics* (never its results):
 *   - input is the flat structure-of-arrays batch (mmc_batch_t) instead of bam1_t records;
 *   - the string-keyed khash maps are replaced by an append-only update log that is sorted and
 *     reduced at the end (results are order independent: uint32 sums);
 *   - fatal conditions return an error string instead of exit(1).
 * Everything else -- the reversed CIGAR walk for reverse reads, the FASTQ-oriented aln/ins arrays,
 * the bases_pos tables, the character-by-character MM parser, the double-precision threshold
 * test, the KMP context byte maps -- is done the way the reference does it.
 */
#include "modcall_oracle.h"

#include <ctype.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define N_BASES 6

typedef struct {
    int32_t tid, pos;
    uint8_t strand, code;
    uint16_t ins16;
    int16_t hap;
    uint8_t called, mod;
} update_t;

typedef struct {
    uint32_t len;
    char *forward;              /* ref_t.forward, src/ref.h:38 */
    uint8_t **is_context;       /* [n_mods][len], src/ref.h:39 */
    uint8_t **is_context_rev;
} oref_t;

struct oracle_ctx {
    int subtool, n_mods, insertions, haplotypes, wildcard; /* wildcard: index of the "*" code or -1 */
    mmc_mod_t *mods;
    double *thresh;
    int n_contigs;
    uint32_t *lens;
    oref_t *refs;
    /* output code dictionary (strings) */
    char codes[256][16];
    int n_codes;
    /* freq: update log */
    update_t *upd; size_t n_upd, cap_upd;
    mmc_freq_rec_t *freq; size_t n_freq;
    /* view: rows of the last batch */
    mmc_view_rec_t *view; size_t n_view, cap_view;
    char err[512];
};

static int fail(oracle_ctx *c, const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(c->err, sizeof(c->err), fmt, ap); va_end(ap);
    return -1;
}

/* seq_nt16_str of htslib, used at src/mod.c:978,1118 */
static const char NT16[] = "=ACMGRSVTWYHKDBN";

/* base_idx_lookup, src/mod.c:97: everything not listed is 0 */
static int base_idx(int c) {
    switch (c) {
    case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3; case 'N': case 'n': return 4; default: return 0;
    }
}
/* base_complement_lookup, src/mod.c:98 (0 for anything else) */
static int base_comp(int c) {
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'U': return 'A'; case 'N': return 'N';
    case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; case 'u': return 'a'; case 'n': return 'n';
    default: return 0;
    }
}
/* valid_bases, src/mod.c:95 */
static int valid_base(int c) { return c && strchr("ACGTUNacgtun", c) != NULL; }

oracle_ctx *oracle_create(int subtool, int n_mods, const mmc_mod_t *mods, const double *thresh, int insertions, int haplotypes,
                          int n_contigs, const uint32_t *lens) {
    oracle_ctx *c = (oracle_ctx *)calloc(1, sizeof(*c));
    c->subtool = subtool; c->n_mods = n_mods; c->insertions = insertions; c->haplotypes = haplotypes;
    c->mods = (mmc_mod_t *)malloc(sizeof(mmc_mod_t) * (size_t)n_mods);
    memcpy(c->mods, mods, sizeof(mmc_mod_t) * (size_t)n_mods);
    c->thresh = (double *)calloc((size_t)n_mods, sizeof(double));
    if (thresh) memcpy(c->thresh, thresh, sizeof(double) * (size_t)n_mods);
    c->wildcard = -1;
    for (int i = 0; i < n_mods; i++) if (!strcmp(mods[i].code, "*")) c->wildcard = i;
    c->n_contigs = n_contigs;
    c->lens = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_contigs > 0 ? n_contigs : 1));
    memcpy(c->lens, lens, sizeof(uint32_t) * (size_t)n_contigs);
    c->refs = (oref_t *)calloc((size_t)(n_contigs > 0 ? n_contigs : 1), sizeof(oref_t));
    return c;
}

void oracle_destroy(oracle_ctx *c) {
    if (!c) return;
    for (int t = 0; t < c->n_contigs; t++) {
        oref_t *r = &c->refs[t];
        if (!r->forward) continue;
        for (int i = 0; i < c->n_mods; i++) { free(r->is_context[i]); free(r->is_context_rev[i]); }
        free(r->is_context); free(r->is_context_rev); free(r->forward);
    }
    free(c->refs); free(c->lens); free(c->mods); free(c->thresh); free(c->upd); free(c->freq); free(c->view); free(c);
}

const char *oracle_strerror(const oracle_ctx *c) { return c->err; }
const char *oracle_code_name(const oracle_ctx *c, int code) { return code >= 0 && code < c->n_codes ? c->codes[code] : ""; }

/* KMP search marking match starts: search_context_kmp(), src/ref.c:92-139 */
static void kmp_mark_starts(const char *pat, const char *txt, size_t N, uint8_t *starts) {
    size_t M = strlen(pat);
    if (M == 0 || M > N) return;
    size_t *lps = (size_t *)malloc(M * sizeof(size_t));
    size_t len = 0, i = 1;
    lps[0] = 0;
    while (i < M) {
        if (pat[i] == pat[len]) { len++; lps[i] = len; i++; }
        else if (len != 0) len = lps[len - 1];
        else { lps[i] = 0; i++; }
    }
    size_t j = 0;
    i = 0;
    while ((N - i) >= (M - j)) {
        int matched = pat[j] == txt[i];
        if (matched) { j++; i++; }
        if (j == M) { starts[i - j] = 1; j = lps[j - 1]; }
        else if (i < N && !matched) { if (j != 0) j = lps[j - 1]; else i++; }
    }
    free(lps);
}

/* search_context_kmp_mark_window(), src/ref.c:142-162: every base of every match window */
static void mark_windows(const char *pat, const char *txt, size_t n, uint8_t *out) {
    size_t m = strlen(pat);
    uint8_t *starts = (uint8_t *)calloc(n ? n : 1, 1);
    kmp_mark_starts(pat, txt, n, starts);
    for (size_t i = 0; i < n; i++)
        if (starts[i]) for (size_t j = i; j < i + m && j < n; j++) out[j] = 1;
    free(starts);
}

/* load_ref() normalisation (src/ref.c:72-78) + load_ref_contexts() (src/ref.c:177-229) for one contig */
int oracle_ref_add(oracle_ctx *c, int tid, const char *seq, uint32_t len) {
    if (tid < 0 || tid >= c->n_contigs) return fail(c, "bad tid %d", tid);
    oref_t *r = &c->refs[tid];
    r->len = len;
    r->forward = (char *)malloc((size_t)len + 1);
    for (uint32_t i = 0; i < len; i++) {
        int ch = toupper((unsigned char)seq[i]);
        r->forward[i] = (char)(ch == 'U' ? 'T' : ch);
    }
    r->forward[len] = 0;
    r->is_context = (uint8_t **)calloc((size_t)c->n_mods, sizeof(uint8_t *));
    r->is_context_rev = (uint8_t **)calloc((size_t)c->n_mods, sizeof(uint8_t *));
    for (int i = 0; i < c->n_mods; i++) {
        const char *ctx = c->mods[i].context;
        r->is_context[i] = (uint8_t *)calloc(len ? len : 1, 1);
        r->is_context_rev[i] = (uint8_t *)calloc(len ? len : 1, 1);
        if (!strcmp(ctx, "*")) { memset(r->is_context[i], 1, len); memset(r->is_context_rev[i], 1, len); continue; }
        char rc[MMC_MAX_CONTEXT + 1];
        size_t l = strlen(ctx);
        for (size_t j = 0; j < l; j++) rc[j] = (char)base_comp(ctx[l - j - 1]);      /* src/ref.c:183-194 */
        rc[l] = 0;
        mark_windows(ctx, r->forward, len, r->is_context[i]);
        mark_windows(rc, r->forward, len, r->is_context_rev[i]);
    }
    return 0;
}

static int code_index(oracle_ctx *c, const char *s) {
    for (int i = 0; i < c->n_codes; i++) if (!strcmp(c->codes[i], s)) return i;
    if (c->n_codes >= 256) return -1;
    snprintf(c->codes[c->n_codes], sizeof(c->codes[0]), "%s", s);
    return c->n_codes++;
}

/* update_freq_map(), src/mod.c:883-929: the keyed entry and, with haplotypes, the hap=-1 aggregate */
static void log_update(oracle_ctx *c, int32_t tid, int32_t pos, int ins_offset, int code, int rev, int hap, int called, int mod) {
    for (int k = 0; k < (hap != -1 ? 2 : 1); k++) {
        if (c->n_upd == c->cap_upd) { c->cap_upd = c->cap_upd ? c->cap_upd * 2 : 1 << 16; c->upd = (update_t *)realloc(c->upd, c->cap_upd * sizeof(update_t)); }
        update_t *u = &c->upd[c->n_upd++];
        u->tid = tid; u->pos = pos; u->strand = (uint8_t)rev; u->code = (uint8_t)code;
        u->ins16 = (uint16_t)ins_offset;                    /* make_key(... uint16_t ins_offset ...), src/mod.c:428 */
        u->hap = (int16_t)(k == 0 ? hap : -1);
        u->called = (uint8_t)called; u->mod = (uint8_t)mod;
    }
}

typedef struct { int32_t ref_pos; uint16_t ins16; uint8_t code; mmc_view_rec_t rec; } vrow_t;

typedef struct {
    vrow_t *rows; size_t n, cap;
} vlist_t;

/* add_view_entry(), src/mod.c:931-946: first entry for a key wins */
static void view_add(vlist_t *vl, int32_t ref_pos, int ins_offset, int code, const mmc_view_rec_t *rec) {
    uint16_t i16 = (uint16_t)ins_offset;
    for (size_t i = 0; i < vl->n; i++)
        if (vl->rows[i].ref_pos == ref_pos && vl->rows[i].ins16 == i16 && vl->rows[i].code == code) return;
    if (vl->n == vl->cap) { vl->cap = vl->cap ? vl->cap * 2 : 1024; vl->rows = (vrow_t *)realloc(vl->rows, vl->cap * sizeof(vrow_t)); }
    vrow_t *r = &vl->rows[vl->n++];
    r->ref_pos = ref_pos; r->ins16 = i16; r->code = (uint8_t)code; r->rec = *rec;
}

static int cmp_vrow(const void *a, const void *b) {
    const vrow_t *x = (const vrow_t *)a, *y = (const vrow_t *)b;
    if (x->ref_pos != y->ref_pos) return x->ref_pos < y->ref_pos ? -1 : 1;
    if (x->code != y->code) return x->code < y->code ? -1 : 1;
    if (x->ins16 != y->ins16) return x->ins16 < y->ins16 ? -1 : 1;
    return 0;
}

static inline int seqi(const uint8_t *s, uint32_t i) { return s[i >> 1] >> ((~i & 1) << 2) & 0xf; }   /* bam_seqi */

/* One read: get_aln() (src/mod.c:776-881) then freq_view_single() (src/mod.c:948-1370). */
static int process_read(oracle_ctx *c, const mmc_batch_t *b, uint32_t ri, vlist_t *vl) {
    const int32_t tid = b->tid[ri], pos = b->pos[ri];
    const uint32_t seq_len = b->l_seq[ri], n_cigar = b->n_cigar[ri], ml_len = b->ml_len[ri];
    const int rev = (b->flag[ri] & 16) != 0;
    const uint32_t *cigar = b->cigar + b->cigar_off[ri];
    const uint8_t *seq = b->seq4 + b->seq_off[ri];
    const char *mm_string = b->mm + b->mm_off[ri];
    const int mm_str_len = (int)b->mm_len[ri];
    const uint8_t *ml = b->ml + b->ml_off[ri];
    const int haplotype = c->haplotypes ? (int)b->hp[ri] : -1;                 /* src/mod.c:963 */
    int rc = 0;

    if (tid < 0 || tid >= c->n_contigs || !c->refs[tid].forward) return fail(c, "read %u: Contig not found in reference provided", ri);
    const oref_t *ref = &c->refs[tid];

    /* bam_endpos(): pos + reference length of the CIGAR, or pos+1 */
    int32_t end = pos;
    for (uint32_t ci = 0; ci < n_cigar; ci++) {
        int op = cigar[ci] & 15;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += (int32_t)(cigar[ci] >> 4);
    }
    if (end == pos) end = pos + 1;

    int *aln = (int *)malloc(sizeof(int) * seq_len);
    int *ins = (int *)malloc(sizeof(int) * seq_len);
    int *ins_offset = (int *)malloc(sizeof(int) * seq_len);
    int *bases_pos[N_BASES];
    int bases_pos_lens[N_BASES] = {0};
    for (int j = 0; j < N_BASES; j++) bases_pos[j] = (int *)malloc(sizeof(int) * seq_len);
    int *skip_counts = (int *)malloc(sizeof(int) * (size_t)(mm_str_len + 1));

    /* ---- get_aln(): arrays are indexed by FASTQ-orientation read position */
    for (uint32_t i = 0; i < seq_len; i++) { aln[i] = -1; ins[i] = -1; ins_offset[i] = 0; }
    {
        int read_pos = 0, ref_pos = pos;
        for (uint32_t ci = 0; ci < n_cigar; ++ci) {
            uint32_t cg = rev ? cigar[n_cigar - ci - 1] : cigar[ci];            /* src/mod.c:812-815 */
            int cigar_len = (int)(cg >> 4), cigar_op = (int)(cg & 15);
            int read_inc = 0, ref_inc = 0, is_aligned = 0, is_inserted = 0;
            if (cigar_op == 0 || cigar_op == 7 || cigar_op == 8) { is_aligned = 1; read_inc = 1; ref_inc = 1; }
            else if (cigar_op == 2) ref_inc = 1;
            else if (cigar_op == 3) ref_inc = 1;
            else if (cigar_op == 1) { read_inc = 1; is_inserted = 1; }
            else if (cigar_op == 4) read_inc = 1;
            else if (cigar_op == 5) { rc = fail(c, "read %u: Hard clipping found and they are not supported", ri); goto done; }
            else { rc = fail(c, "read %u: Unhandled CIGAR OPT Cigar: %d", ri, cigar_op); goto done; }
            for (int j = 0; j < cigar_len; ++j) {
                if (is_aligned) {
                    if (!(read_pos < (int)seq_len)) { rc = fail(c, "read %u: read_pos:%d seq_len:%d", ri, read_pos, (int)seq_len); goto done; }
                    int start = ref_pos;
                    if (rev) start = pos + end - ref_pos - 1;                      /* src/mod.c:855-857 */
                    aln[read_pos] = start;
                    if (!(ref_pos >= 0 && (uint32_t)ref_pos < ref->len)) { rc = fail(c, "read %u: ref_pos:%d ref_len:%u", ri, ref_pos, ref->len); goto done; }
                    if (ref->len != c->lens[tid]) { rc = fail(c, "read %u: ref_len:%u target_len:%u", ri, ref->len, c->lens[tid]); goto done; }
                }
                if (c->insertions && is_inserted) {
                    if (!(read_pos < (int)seq_len)) { rc = fail(c, "read %u: read_pos:%d seq_len:%d", ri, read_pos, (int)seq_len); goto done; }
                    int start = ref_pos - 1, offset = j + 1;
                    if (rev) { start = pos + end - ref_pos - 1; offset = cigar_len - j; }   /* src/mod.c:868-871 */
                    ins[read_pos] = start;
                    ins_offset[read_pos] = offset;
                }
                read_pos += read_inc;
                ref_pos += ref_inc;
            }
        }
    }

    /* ---- bases_pos tables, src/mod.c:977-981 */
    for (uint32_t i = 0; i < seq_len; i++) {
        int idx = base_idx(NT16[seqi(seq, i)]);
        bases_pos[idx][bases_pos_lens[idx]++] = (int)i;
    }

    /* ---- the MM string, block by block, src/mod.c:995-1369 */
    int i = 0, ml_start_idx = 0, blk_ord = 0;
    while (i < mm_str_len) {
        int skip_counts_len = 0, mod_codes_len = 0;
        char modbase = 0, status_flag = '.';
        char mod_codes[64];
        if (!valid_base(mm_string[i])) { rc = fail(c, "read %u: Invalid base:%c", ri, mm_string[i]); goto done; }
        modbase = mm_string[i] == 'U' ? 'T' : mm_string[i];                        /* src/mod.c:1006 */
        i++;
        if (i < mm_str_len) {
            if (mm_string[i] != '+' && mm_string[i] != '-') { rc = fail(c, "read %u: Invalid strand:%c", ri, mm_string[i]); goto done; }
            i++;
        }
        int j = 0, has_nums = 0, has_alpha = 0;
        while (i < mm_str_len && mm_string[i] != ',' && mm_string[i] != ';' && mm_string[i] != '?' && mm_string[i] != '.') {
            char ch = mm_string[i];
            if (ch >= '0' && ch <= '9') has_nums = 1;
            else if ((ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z')) has_alpha = 1;
            else { rc = fail(c, "read %u: Invalid base modification code:%c", ri, ch); goto done; }
            if (j >= 62) { rc = fail(c, "read %u: modification code string too long", ri); goto done; }
            mod_codes[j++] = ch; i++;
        }
        mod_codes[j] = 0;
        mod_codes_len = j;
        if (has_nums) mod_codes_len = 1;                                            /* src/mod.c:1048-1050 */
        if (!(mod_codes_len > 0)) { rc = fail(c, "read %u: Invalid modification codes. Modification codes cannot be empty.", ri); goto done; }
        if (has_nums && has_alpha) { rc = fail(c, "read %u: Modification codes should be either numeric or alphabetic, not both.", ri); goto done; }
        if (i < mm_str_len && (mm_string[i] == '?' || mm_string[i] == '.')) { status_flag = mm_string[i]; i++; }
        /* skip counts, src/mod.c:1064-1090 */
        int k = 0;
        while (i < mm_str_len && mm_string[i] != ';') {
            if (mm_string[i] == ',') { i++; continue; }
            char tok[10]; int l = 0;
            while (i < mm_str_len && mm_string[i] != ',' && mm_string[i] != ';') {
                if (l >= 9) { rc = fail(c, "read %u: skip count longer than 9 characters", ri); goto done; }   /* assert(l < 10) */
                tok[l++] = mm_string[i++];
            }
            tok[l] = 0;
            /* the reference uses sscanf("%d"); this restatement only accepts plain digits and flags the rest */
            for (int z = 0; z < l; z++) if (tok[z] < '0' || tok[z] > '9') { rc = fail(c, "read %u: Invalid skip count '%s'", ri, tok); goto done; }
            skip_counts[k++] = atoi(tok);
        }
        skip_counts_len = k;
        i++;

        const char mb = rev ? (char)base_comp(modbase) : modbase;                   /* src/mod.c:1092 */
        const int idx = base_idx(mb);
        int ml_idx = ml_start_idx;
        long base_rank = -1;
        for (int cc = 0; cc < skip_counts_len; cc++) {                              /* called bases, src/mod.c:1097-1199 */
            base_rank += (long)skip_counts[cc] + 1;
            long read_pos;
            if (modbase == 'N') read_pos = rev ? (long)seq_len - base_rank - 1 : base_rank;
            else {
                if (base_rank >= bases_pos_lens[idx]) { rc = fail(c, "read %u: Read pos cannot exceed seq len (rank %ld of %d)", ri, base_rank, bases_pos_lens[idx]); goto done; }
                read_pos = rev ? bases_pos[idx][bases_pos_lens[idx] - base_rank - 1] : bases_pos[idx][base_rank];
            }
            if (!(read_pos >= 0 && read_pos < (long)seq_len)) { rc = fail(c, "read %u: Read pos cannot exceed seq len. read_pos: %ld seq_len: %u", ri, read_pos, seq_len); goto done; }
            char read_base = NT16[seqi(seq, (uint32_t)read_pos)];
            int fastq_read_pos = rev ? (int)(seq_len - read_pos - 1) : (int)read_pos;
            int ref_pos = aln[fastq_read_pos];
            if (c->insertions) ref_pos = ref_pos == -1 ? ins[fastq_read_pos] : ref_pos;
            if (ref_pos == -1) { if (mod_codes_len > 0) ml_idx = ml_start_idx + cc * mod_codes_len + mod_codes_len - 1; continue; }
            for (int m = 0; m < mod_codes_len; m++) {
                ml_idx = ml_start_idx + cc * mod_codes_len + m;
                const char *mod_code = has_nums ? mod_codes : &mod_codes[m];      /* suffix string, src/mod.c:1148-1152 */
                int req = c->wildcard;
                if (req < 0) {
                    for (int q = 0; q < c->n_mods; q++) if (!strcmp(c->mods[q].code, mod_code)) { req = q; break; }
                    if (req < 0) continue;
                }
                int req_all_contexts = strcmp(c->mods[req].context, "*") == 0;
                int is_in_context = (rev && ref->is_context_rev[req][ref_pos]) || (!rev && ref->is_context[req][ref_pos]);
                int matches_reference = req_all_contexts || mb == 'N' || ref->forward[ref_pos] == read_base;
                if (c->insertions) { /* no context check with --insertions, src/mod.c:1167 */ }
                else if (is_in_context && matches_reference) { }
                else continue;
                if (!(ml_idx < (int)ml_len)) { rc = fail(c, "read %u: mod prob index mismatch. ml_idx:%d ml_len:%u", ri, ml_idx, ml_len); goto done; }
                uint8_t mod_prob = ml[ml_idx];
                int ins_off = c->insertions ? ins_offset[fastq_read_pos] : 0;
                int code = code_index(c, mod_code);
                if (code < 0) { rc = fail(c, "read %u: too many distinct modification codes", ri); goto done; }
                if (c->subtool == MMC_FREQ) {
                    int is_mod = 0, is_called = 0;
                    double thresh = c->thresh[req];
                    double mod_prob_dbl = (double)((mod_prob + 0.5) / 256.0);         /* THRESH_UINT8_TO_DBL, src/mod.c:56 */
                    if (mod_prob_dbl >= thresh) { is_called = 1; is_mod = 1; }
                    else if (mod_prob_dbl <= 1 - thresh) is_called = 1;
                    else continue;
                    log_update(c, tid, ref_pos, ins_off, code, rev, haplotype, is_called, is_mod);
                } else {
                    mmc_view_rec_t v; memset(&v, 0, sizeof(v));
                    v.read = ri; v.ref_pos = ref_pos; v.read_pos = fastq_read_pos; v.ins_offset = (uint32_t)ins_off;
                    v.code = (uint8_t)code; v.mod_prob = mod_prob; v.strand = (uint8_t)rev; v.hp = b->hp[ri];
                    view_add(vl, ref_pos, ins_off, code, &v);
                }
            }
        }
        if (skip_counts_len > 0) ml_start_idx = ml_idx + 1;                         /* src/mod.c:1200 */

        if (status_flag == '.') {                                                  /* skipped bases, src/mod.c:1203-1367 */
            /* the two loops of the reference (between calls, then after the last call) visit the ranks
             * prev+1 .. rank-1 for every call and last+1 .. bases_pos_lens[idx]-1 at the end */
            long skip_base_rank = -1, prev = -1;
            for (int pass = 0; pass <= skip_counts_len; pass++) {
                long from, to;
                if (pass < skip_counts_len) { skip_base_rank += (long)skip_counts[pass] + 1; from = prev + 1; to = skip_base_rank; prev = skip_base_rank; }
                else { from = prev + 1; to = bases_pos_lens[idx]; }
                for (long s = from; s < to; s++) {
                    long skip_read_pos;
                    if (modbase == 'N') skip_read_pos = rev ? (long)seq_len - s - 1 : s;
                    else {
                        if (s >= bases_pos_lens[idx]) { rc = fail(c, "read %u: Read pos cannot exceed seq len (skipped rank)", ri); goto done; }
                        skip_read_pos = rev ? bases_pos[idx][bases_pos_lens[idx] - s - 1] : bases_pos[idx][s];
                    }
                    if (!(skip_read_pos >= 0 && skip_read_pos < (long)seq_len)) { rc = fail(c, "read %u: Read pos cannot exceed seq len", ri); goto done; }
                    char skip_read_base = NT16[seqi(seq, (uint32_t)skip_read_pos)];
                    int skip_fastq_read_pos = rev ? (int)(seq_len - skip_read_pos - 1) : (int)skip_read_pos;
                    int skip_ref_pos = aln[skip_fastq_read_pos];
                    if (c->insertions) skip_ref_pos = skip_ref_pos == -1 ? ins[skip_read_pos] : skip_ref_pos;   /* sic: BAM-orientation index, src/mod.c:1234,1314 */
                    if (skip_ref_pos == -1) continue;
                    for (int m = 0; m < mod_codes_len; m++) {
                        const char *mod_code = has_nums ? mod_codes : &mod_codes[m];
                        int req = c->wildcard;
                        if (req < 0) {
                            for (int q = 0; q < c->n_mods; q++) if (!strcmp(c->mods[q].code, mod_code)) { req = q; break; }
                            if (req < 0) continue;
                        }
                        int req_all_contexts = strcmp(c->mods[req].context, "*") == 0;
                        int in_ctx = (rev && ref->is_context_rev[req][skip_ref_pos]) || (!rev && ref->is_context[req][skip_ref_pos]);
                        int matches = req_all_contexts || mb == 'N' || ref->forward[skip_ref_pos] == skip_read_base;
                        if (c->insertions) { }
                        else if (in_ctx && matches) { }
                        else continue;
                        int ins_off = c->insertions ? ins_offset[skip_fastq_read_pos] : 0;
                        int code = code_index(c, mod_code);
                        if (code < 0) { rc = fail(c, "read %u: too many distinct modification codes", ri); goto done; }
                        if (c->subtool == MMC_FREQ) log_update(c, tid, skip_ref_pos, ins_off, code, rev, haplotype, 1, 0);   /* src/mod.c:1279,1359 */
                        else {
                            mmc_view_rec_t v; memset(&v, 0, sizeof(v));
                            v.read = ri; v.ref_pos = skip_ref_pos; v.read_pos = skip_fastq_read_pos; v.ins_offset = (uint32_t)ins_off;
                            v.code = (uint8_t)code; v.mod_prob = 0; v.strand = (uint8_t)rev; v.hp = b->hp[ri];
                            view_add(vl, skip_ref_pos, ins_off, code, &v);
                        }
                    }
                }
            }
        }
        blk_ord++;
    }
    (void)blk_ord;

done:
    free(aln); free(ins); free(ins_offset); free(skip_counts);
    for (int j = 0; j < N_BASES; j++) free(bases_pos[j]);
    return rc;
}

/* process_db() + merge_db() (src/minimod.c:344-386) / output_db()'s collect (src/mod.c:569-593) for one batch */
static int oracle_process_batch4(oracle_ctx *c, const mmc_batch_t *b);

/* The batch's SEQ in transport form (include/minimod_cuda.h, seq_packing == 2) back to BAM's 4-bit bytes, nibble by
 * nibble: code c -> nt16 1 << c, then the exception entries verbatim.  Only the oracle does this on the CPU. */
static uint8_t *seq4_from_transport(const mmc_batch_t *b) {
    uint8_t *s4 = (uint8_t *)calloc(b->seq_used + 64, 1);
    for (uint64_t i = 0; i < 2 * b->seq_used; i++) {                               /* nibble i of the pool */
        const uint32_t code = (b->seq2[i >> 2] >> (6 - 2 * (i & 3))) & 3u;
        s4[i >> 1] |= (uint8_t)((1u << code) << ((~i & 1) << 2));
    }
    for (uint64_t k = 0; k < b->seq_exc_used; k++) {
        const uint64_t i = b->seq_exc[k] >> 4;
        const uint32_t sh = (uint32_t)((~i & 1) << 2);
        s4[i >> 1] = (uint8_t)((s4[i >> 1] & ~(0xfu << sh)) | ((uint32_t)(b->seq_exc[k] & 15u) << sh));
    }
    return s4;
}

/* The batch's CIGARs in the byte form (include/minimod_cuda.h, cigar_packing == 8) back to BAM's 32-bit words, one op at
 * a time with two running cursors into the escape lists.  Only the oracle does this on the CPU. */
static uint32_t *cigar_from_transport(const mmc_batch_t *b) {
    uint32_t *w = (uint32_t *)calloc(b->cigar_used + 64, 4);
    for (uint32_t r = 0; r < b->n_reads; r++) {
        const uint8_t *blob = b->cig8 + b->cig8_off[r];
        const uint32_t n = b->n_cigar[r];
        uint32_t n1; memcpy(&n1, blob, 4);
        const uint8_t *ops = blob + 4, *l1 = ops + ((n + 3u) & ~3u), *l2 = l1 + ((n1 + 3u) & ~3u);
        for (uint32_t i = 0; i < n; i++) {
            uint32_t len = ops[i] >> 4;
            if (len == 15u) {
                const uint32_t v = *l1++;
                if (v < 255u) len = 15u + v; else { memcpy(&len, l2, 4); l2 += 4; }
            }
            w[b->cigar_off[r] + i] = (len << 4) | (ops[i] & 15u);
        }
    }
    return w;
}

int oracle_process_batch(oracle_ctx *c, const mmc_batch_t *b_in) {
    mmc_batch_t tmp = *b_in;
    uint8_t *s4 = NULL;
    uint32_t *cw = NULL;
    if (b_in->seq_packing == 2) { s4 = seq4_from_transport(b_in); tmp.seq4 = s4; }
    if (b_in->cigar_packing == 8) { cw = cigar_from_transport(b_in); tmp.cigar = cw; }
    int rc_all = oracle_process_batch4(c, &tmp);
    free(s4); free(cw);
    return rc_all;
}

static int oracle_process_batch4(oracle_ctx *c, const mmc_batch_t *b) {
    c->n_view = 0;
    for (uint32_t ri = 0; ri < b->n_reads; ri++) {
        vlist_t vl; memset(&vl, 0, sizeof(vl));
        int rc = process_read(c, b, ri, &vl);
        if (rc == 0 && c->subtool == MMC_VIEW && vl.n) {
            qsort(vl.rows, vl.n, sizeof(vrow_t), cmp_vrow);                          /* rows of a read ordered by (pos, code, ins) */
            if (c->n_view + vl.n > c->cap_view) {
                c->cap_view = (c->n_view + vl.n) * 2;
                c->view = (mmc_view_rec_t *)realloc(c->view, c->cap_view * sizeof(mmc_view_rec_t));
            }
            for (size_t i = 0; i < vl.n; i++) c->view[c->n_view++] = vl.rows[i].rec;
        }
        free(vl.rows);
        if (rc) return rc;
    }
    return 0;
}

static int cmp_update(const void *a, const void *b) {
    const update_t *x = (const update_t *)a, *y = (const update_t *)b;
    if (x->tid != y->tid) return x->tid < y->tid ? -1 : 1;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    if (x->strand != y->strand) return x->strand < y->strand ? -1 : 1;
    if (x->code != y->code) return x->code < y->code ? -1 : 1;
    if (x->ins16 != y->ins16) return x->ins16 < y->ins16 ? -1 : 1;
    if (x->hap != y->hap) return x->hap < y->hap ? -1 : 1;
    return 0;
}

/* merge_freq_maps() + the collect/sort of print_freq_output() (src/mod.c:743-774,644-664) */
int oracle_freq_records(oracle_ctx *c, const mmc_freq_rec_t **recs, uint64_t *n) {
    qsort(c->upd, c->n_upd, sizeof(update_t), cmp_update);
    free(c->freq);
    c->freq = (mmc_freq_rec_t *)malloc(sizeof(mmc_freq_rec_t) * (c->n_upd ? c->n_upd : 1));
    c->n_freq = 0;
    for (size_t i = 0; i < c->n_upd;) {
        size_t j = i;
        uint32_t called = 0, mod = 0;
        while (j < c->n_upd && cmp_update(&c->upd[i], &c->upd[j]) == 0) { called += c->upd[j].called; mod += c->upd[j].mod; j++; }
        mmc_freq_rec_t *r = &c->freq[c->n_freq++];
        r->tid = c->upd[i].tid; r->pos = c->upd[i].pos; r->n_called = called; r->n_mod = mod;
        r->ins_offset = c->upd[i].ins16; r->hap = c->upd[i].hap; r->strand = c->upd[i].strand; r->code = c->upd[i].code; r->reserved = 0;
        i = j;
    }
    *recs = c->freq; *n = c->n_freq;
    return 0;
}

int oracle_view_records(oracle_ctx *c, const mmc_view_rec_t **recs, uint64_t *n) {
    *recs = c->view; *n = c->n_view;
    return 0;
}
