/* TEST INFRASTRUCTURE ONLY -- minimal stand-in for <htslib/sam.h>.
 *
 * htslib 1.9 is not vendored by the reference (scripts/install-hts.sh:10
 * downloads it) and is absent from this image.  The reference uses htslib
 * purely as a BAM container reader; this shim implements the symbols it
 * touches (src/minimod.c:73-89,250; src/mod.c:127-199,780-790,956,978,1118)
 * on top of zlib's gzread, which transparently reads BGZF (a series of gzip
 * members).  It exists so the UNMODIFIED reference sources can be compiled
 * into oracle/_ref/minimod_ref.  Nothing in the product links against it.
 */
#ifndef ORACLE_SHIM_SAM_H
#define ORACLE_SHIM_SAM_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct htsFile htsFile;
typedef struct hts_idx_t hts_idx_t;
typedef struct hts_itr_t hts_itr_t;

typedef struct {
    int32_t n_targets;
    uint32_t *target_len;
    char **target_name;
    char *text;
    uint32_t l_text;
} bam_hdr_t;

typedef struct {
    int32_t tid;
    int32_t pos;
    uint16_t bin;
    uint8_t qual;
    uint8_t l_qname;
    uint16_t flag;
    uint32_t n_cigar;
    int32_t l_qseq;
    int32_t mtid;
    int32_t mpos;
    int32_t isize;
} bam1_core_t;

typedef struct {
    bam1_core_t core;
    int l_data;
    uint32_t m_data;
    uint8_t *data;
} bam1_t;

#define BAM_CMATCH      0
#define BAM_CINS        1
#define BAM_CDEL        2
#define BAM_CREF_SKIP   3
#define BAM_CSOFT_CLIP  4
#define BAM_CHARD_CLIP  5
#define BAM_CPAD        6
#define BAM_CEQUAL      7
#define BAM_CDIFF       8
#define BAM_CBACK       9

#define BAM_FPAIRED        1
#define BAM_FPROPER_PAIR   2
#define BAM_FUNMAP         4
#define BAM_FMUNMAP        8
#define BAM_FREVERSE      16
#define BAM_FMREVERSE     32
#define BAM_FREAD1        64
#define BAM_FREAD2       128
#define BAM_FSECONDARY   256
#define BAM_FQCFAIL      512
#define BAM_FDUP        1024
#define BAM_FSUPPLEMENTARY 2048

#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK  0xf
#define bam_cigar_op(c)    ((c) & BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)

#define bam_is_rev(b)   (((b)->core.flag & BAM_FREVERSE) != 0)
#define bam_get_qname(b) ((char *)(b)->data)
#define bam_get_cigar(b) ((uint32_t *)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b)   ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b)  ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b)   ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_seqi(s, i)   ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)

extern const char seq_nt16_str[];

htsFile *sam_open(const char *fn, const char *mode);
int sam_close(htsFile *fp);
int hts_set_threads(htsFile *fp, int n);
bam_hdr_t *sam_hdr_read(htsFile *fp);
void bam_hdr_destroy(bam_hdr_t *h);
int sam_read1(htsFile *fp, bam_hdr_t *h, bam1_t *b);
bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]);
char *bam_aux2Z(const uint8_t *s);
int64_t bam_aux2i(const uint8_t *s);
uint32_t bam_auxB_len(const uint8_t *s);
int64_t bam_auxB2i(const uint8_t *s, uint32_t idx);
int32_t bam_endpos(const bam1_t *b);

#ifdef __cplusplus
}
#endif
#endif
