/* TEST INFRASTRUCTURE ONLY -- see shim/htslib/sam.h.
 *
 * A BAM reader just big enough for the reference's call sites.  BGZF is a
 * concatenation of gzip members, which zlib's gzread() decodes transparently,
 * so no BGZF framing logic is needed here.  Record layout per SAM spec 4.2:
 *   int32 block_size; then 32 fixed bytes (refID, pos, l_read_name, mapq, bin,
 *   n_cigar_op, flag, l_seq, next_refID, next_pos, tlen); then
 *   qname, cigar, seq (4-bit), qual, aux.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include "htslib/sam.h"

const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";

struct htsFile {
    gzFile gz;
};

static int read_exact(gzFile gz, void *buf, size_t n) {
    uint8_t *p = (uint8_t *)buf;
    while (n > 0) {
        unsigned chunk = n > (1u << 30) ? (1u << 30) : (unsigned)n;
        int got = gzread(gz, p, chunk);
        if (got <= 0) return -1;
        p += got;
        n -= (size_t)got;
    }
    return 0;
}

static uint32_t le32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

htsFile *sam_open(const char *fn, const char *mode) {
    (void)mode;
    gzFile gz = gzopen(fn, "rb");
    if (!gz) return NULL;
    gzbuffer(gz, 1 << 20);
    htsFile *fp = (htsFile *)calloc(1, sizeof(*fp));
    fp->gz = gz;
    return fp;
}

int sam_close(htsFile *fp) {
    if (!fp) return 0;
    gzclose(fp->gz);
    free(fp);
    return 0;
}

int hts_set_threads(htsFile *fp, int n) { (void)fp; (void)n; return 0; }

bam_hdr_t *sam_hdr_read(htsFile *fp) {
    uint8_t b4[4];
    if (read_exact(fp->gz, b4, 4) || memcmp(b4, "BAM\1", 4) != 0) return NULL;
    bam_hdr_t *h = (bam_hdr_t *)calloc(1, sizeof(*h));
    if (read_exact(fp->gz, b4, 4)) { free(h); return NULL; }
    h->l_text = le32(b4);
    h->text = (char *)malloc((size_t)h->l_text + 1);
    if (h->l_text && read_exact(fp->gz, h->text, h->l_text)) return NULL;
    h->text[h->l_text] = 0;
    if (read_exact(fp->gz, b4, 4)) return NULL;
    h->n_targets = (int32_t)le32(b4);
    h->target_name = (char **)calloc((size_t)h->n_targets + 1, sizeof(char *));
    h->target_len = (uint32_t *)calloc((size_t)h->n_targets + 1, sizeof(uint32_t));
    for (int32_t i = 0; i < h->n_targets; i++) {
        if (read_exact(fp->gz, b4, 4)) return NULL;
        uint32_t l_name = le32(b4);
        h->target_name[i] = (char *)malloc((size_t)l_name + 1);
        if (read_exact(fp->gz, h->target_name[i], l_name)) return NULL;
        h->target_name[i][l_name] = 0;
        if (read_exact(fp->gz, b4, 4)) return NULL;
        h->target_len[i] = le32(b4);
    }
    return h;
}

void bam_hdr_destroy(bam_hdr_t *h) {
    if (!h) return;
    for (int32_t i = 0; i < h->n_targets; i++) free(h->target_name[i]);
    free(h->target_name);
    free(h->target_len);
    free(h->text);
    free(h);
}

bam1_t *bam_init1(void) { return (bam1_t *)calloc(1, sizeof(bam1_t)); }

void bam_destroy1(bam1_t *b) {
    if (!b) return;
    free(b->data);
    free(b);
}

/* htslib >= 1.7 (bam_tag2cigar, called by bam_read1): a CIGAR of more than 65535 ops travels as <l_seq>S<ref_len>N plus a
 * CG:B,I tag; the real ops are put back and the tag is removed. */
static void shim_tag2cigar(bam1_t *b) {
    bam1_core_t *c = &b->core;
    if (c->n_cigar == 0 || c->tid < 0 || c->pos < 0) return;
    uint32_t *cigar0 = bam_get_cigar(b);
    if (bam_cigar_op(cigar0[0]) != BAM_CSOFT_CLIP || (int32_t)bam_cigar_oplen(cigar0[0]) != c->l_qseq) return;
    uint8_t *CG = bam_aux_get(b, "CG");
    if (!CG || CG[0] != 'B' || CG[1] != 'I') return;
    uint32_t n = le32(CG + 2);
    if (n < c->n_cigar || n >= (1u << 29)) return;
    size_t fake = 4 * (size_t)c->n_cigar, real = 4 * (size_t)n;
    uint8_t *tag0 = CG - 2, *tag1 = CG + 6 + real, *end = b->data + b->l_data;
    if (tag1 > end) return;
    size_t new_len = (size_t)b->l_data - fake + real - (size_t)(tag1 - tag0) + real * 0;
    uint8_t *d = (uint8_t *)malloc(new_len + real + 16), *p = d;
    size_t qn = (size_t)c->l_qname;
    memcpy(p, b->data, qn); p += qn;
    memcpy(p, CG + 6, real); p += real;
    size_t mid = (size_t)(tag0 - (b->data + qn + fake));
    memcpy(p, b->data + qn + fake, mid); p += mid;
    memcpy(p, tag1, (size_t)(end - tag1)); p += end - tag1;
    memset(p, 0, 8);
    free(b->data);
    b->data = d; b->l_data = (int)(p - d); b->m_data = (uint32_t)(new_len + real + 16);
    c->n_cigar = n;
}

int sam_read1(htsFile *fp, bam_hdr_t *h, bam1_t *b) {
    (void)h;
    uint8_t b4[4], fixed[32];
    int got = gzread(fp->gz, b4, 4);
    if (got == 0) return -1;              /* clean EOF */
    if (got != 4) return -2;
    uint32_t block_size = le32(b4);
    if (block_size < 32) return -3;
    if (read_exact(fp->gz, fixed, 32)) return -4;
    b->core.tid = (int32_t)le32(fixed + 0);
    b->core.pos = (int32_t)le32(fixed + 4);
    b->core.l_qname = fixed[8];
    b->core.qual = fixed[9];
    b->core.bin = le16(fixed + 10);
    b->core.n_cigar = le16(fixed + 12);
    b->core.flag = le16(fixed + 14);
    b->core.l_qseq = (int32_t)le32(fixed + 16);
    b->core.mtid = (int32_t)le32(fixed + 20);
    b->core.mpos = (int32_t)le32(fixed + 24);
    b->core.isize = (int32_t)le32(fixed + 28);
    b->l_data = (int)(block_size - 32);
    if ((uint32_t)b->l_data + 8 > b->m_data) {
        b->m_data = (uint32_t)b->l_data + 8;
        b->m_data += b->m_data >> 1;
        b->data = (uint8_t *)realloc(b->data, b->m_data);
    }
    if (b->l_data && read_exact(fp->gz, b->data, (size_t)b->l_data)) return -5;
    memset(b->data + b->l_data, 0, 8);
    shim_tag2cigar(b);
    return (int)block_size;
}

/* size in bytes of one aux value starting at its type byte; 0 on error */
static size_t aux_size(const uint8_t *s, const uint8_t *end) {
    if (s >= end) return 0;
    switch (*s) {
    case 'A': case 'c': case 'C': return 2;
    case 's': case 'S': return 3;
    case 'i': case 'I': case 'f': return 5;
    case 'd': return 9;
    case 'Z': case 'H': {
        const uint8_t *p = s + 1;
        while (p < end && *p) p++;
        return (size_t)(p - s) + 1;
    }
    case 'B': {
        if (s + 6 > end) return 0;
        size_t esz;
        switch (s[1]) {
        case 'c': case 'C': esz = 1; break;
        case 's': case 'S': esz = 2; break;
        case 'i': case 'I': case 'f': esz = 4; break;
        default: return 0;
        }
        return 6 + esz * (size_t)le32(s + 2);
    }
    default: return 0;
    }
}

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]) {
    const uint8_t *s = bam_get_aux(b);
    const uint8_t *end = b->data + b->l_data;
    while (s + 3 <= end) {
        size_t sz = aux_size(s + 2, end);
        if (sz == 0) return NULL;
        if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return (uint8_t *)(s + 2);
        s += 2 + sz;
    }
    return NULL;
}

char *bam_aux2Z(const uint8_t *s) {
    if (*s == 'Z' || *s == 'H') return (char *)(s + 1);
    return NULL;
}

int64_t bam_aux2i(const uint8_t *s) {
    switch (*s) {
    case 'c': return (int8_t)s[1];
    case 'C': return s[1];
    case 's': return (int16_t)le16(s + 1);
    case 'S': return le16(s + 1);
    case 'i': return (int32_t)le32(s + 1);
    case 'I': return le32(s + 1);
    default: return 0;
    }
}

uint32_t bam_auxB_len(const uint8_t *s) {
    if (s[0] != 'B') return 0;
    return le32(s + 2);
}

int64_t bam_auxB2i(const uint8_t *s, uint32_t idx) {
    const uint8_t *p = s + 6;
    switch (s[1]) {
    case 'c': return (int8_t)p[idx];
    case 'C': return p[idx];
    case 's': return (int16_t)le16(p + 2 * (size_t)idx);
    case 'S': return le16(p + 2 * (size_t)idx);
    case 'i': return (int32_t)le32(p + 4 * (size_t)idx);
    case 'I': return le32(p + 4 * (size_t)idx);
    default: return 0;
    }
}

/* htslib: pos + reference length of the CIGAR (M,D,N,=,X); pos+1 when that is 0 or the read is unmapped */
int32_t bam_endpos(const bam1_t *b) {
    if ((b->core.flag & BAM_FUNMAP) || b->core.n_cigar == 0) return b->core.pos + 1;
    const uint32_t *cigar = bam_get_cigar(b);
    int32_t rlen = 0;
    for (uint32_t k = 0; k < b->core.n_cigar; k++) {
        int op = bam_cigar_op(cigar[k]);
        if (op == BAM_CMATCH || op == BAM_CDEL || op == BAM_CREF_SKIP || op == BAM_CEQUAL || op == BAM_CDIFF)
            rlen += (int32_t)bam_cigar_oplen(cigar[k]);
    }
    return b->core.pos + (rlen ? rlen : 1);
}
