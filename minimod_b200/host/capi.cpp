// capi.cpp -- C entry points of libminimod_host.so: the host-side pieces (BAM reading,
// load_db filtering/packing, FASTA, -c/-m parsing, text formatting) exposed so that the Python
// mirror of the reference interface (minimod_b200/api.py) and the tests drive exactly the code
// the `minimod` binary runs.  No compute happens here.
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "bam.h"
#include "fasta.h"
#include "format.h"
#include "fastfmt.h"
#include "modopts.h"
#include "pack.h"

using namespace mmh;

namespace {
void set_err(char *err, int errlen, const std::string &msg) {
    if (err && errlen > 0) { snprintf(err, (size_t)errlen, "%s", msg.c_str()); }
}
struct Loader { BatchLoader *bl; BatchMeta meta; };
struct Fasta { std::vector<FastaRecord> recs; };
}

extern "C" {

typedef struct {
    int32_t total_reads; int32_t n_recs;
    int64_t total_bytes; int64_t processed_bytes; int64_t ml_entries; int64_t bases;
} mmh_stats_t;

void *mmh_bam_open(const char *path, char *err, int errlen) {
    BamFile *b = new BamFile();
    std::string e;
    if (!b->open(path, &e)) { set_err(err, errlen, e); delete b; return nullptr; }
    return b;
}
// the same with an explicit number of BGZF inflate threads (hts_set_threads equivalent)
void *mmh_bam_open_t(const char *path, int threads, char *err, int errlen) {
    BamFile *b = new BamFile();
    std::string e;
    if (!b->open(path, &e, threads)) { if (err && errlen > 0) snprintf(err, errlen, "%s", e.c_str()); delete b; return nullptr; }
    return b;
}
// read every record (ingest throughput measurements): returns the number of records, -1 on a corrupt file
long long mmh_bam_scan(void *h, unsigned long long *bytes) {
    BamFile *b = (BamFile *)h;
    BamRecord r;
    long long n = 0;
    unsigned long long tot = 0;
    for (;;) { int rc = b->next(&r); if (rc == 0) break; if (rc < 0) return -1; ++n; tot += (unsigned long long)r.l_data + 36; }
    if (bytes) *bytes = tot;
    return n;
}
void mmh_bam_close(void *h) { delete (BamFile *)h; }
int mmh_bam_n_targets(void *h) { return (int)((BamFile *)h)->names.size(); }
const char *mmh_bam_target_name(void *h, int i) { return ((BamFile *)h)->names[i].c_str(); }
uint32_t mmh_bam_target_len(void *h, int i) { return ((BamFile *)h)->lens[i]; }

void *mmh_loader_new(void *bam, int32_t batch_size, int64_t batch_bytes, int allow_secondary, int skip_supplementary, int keep_qnames) {
    LoadOpts lo;
    lo.batch_size = batch_size; lo.batch_size_bases = batch_bytes;
    lo.allow_secondary = allow_secondary; lo.skip_supplementary = skip_supplementary; lo.keep_qnames = keep_qnames;
    Loader *l = new Loader();
    l->bl = new BatchLoader((BamFile *)bam, lo);
    return l;
}
void mmh_loader_free(void *h) { Loader *l = (Loader *)h; delete l->bl; delete l; }
int mmh_loader_fill(void *h, mmc_batch_t *b, mmh_stats_t *st, char *err, int errlen) {
    Loader *l = (Loader *)h;
    std::string e;
    int rc = l->bl->fill(b, &l->meta, &e);
    if (rc < 0) set_err(err, errlen, e);
    if (st) {
        st->total_reads = l->meta.stats.total_reads; st->n_recs = l->meta.stats.n_recs;
        st->total_bytes = l->meta.stats.total_bytes; st->processed_bytes = l->meta.stats.processed_bytes;
        st->ml_entries = l->meta.stats.ml_entries; st->bases = l->meta.stats.bases;
    }
    return rc;
}
const char *mmh_loader_qname(void *h, uint32_t i) {
    Loader *l = (Loader *)h;
    return i < l->meta.qname_off.size() ? l->meta.qname(i) : "";
}

void *mmh_fasta_load(const char *path, char *err, int errlen) {
    Fasta *f = new Fasta();
    std::string e;
    if (!read_fasta(path, &f->recs, &e, 8)) { set_err(err, errlen, e); delete f; return nullptr; }
    return f;
}
void mmh_fasta_free(void *h) { delete (Fasta *)h; }
int mmh_fasta_n(void *h) { return (int)((Fasta *)h)->recs.size(); }
const char *mmh_fasta_name(void *h, int i) { return ((Fasta *)h)->recs[i].name.c_str(); }
const char *mmh_fasta_seq(void *h, int i) { return ((Fasta *)h)->recs[i].seq.data(); }
uint64_t mmh_fasta_len(void *h, int i) { return ((Fasta *)h)->recs[i].seq.size(); }

// "-c" (+ "-m" for freq; NULL/"" = default 0.8 each) -> mmc_mod_t table.  Returns n_mods or -1.
int mmh_parse_mods(const char *codes, const char *threshes, int subtool, mmc_mod_t *out, int max_out, char *err, int errlen) {
    std::vector<ModSpec> mods;
    std::string e;
    if (!codes || !*codes) codes = "m";
    if (!parse_mod_codes(codes, &mods, &e)) { set_err(err, errlen, e); return -1; }
    if (subtool == MMC_FREQ) {
        std::string t = threshes ? threshes : "";
        if (t.empty()) for (size_t i = 0; i < mods.size(); ++i) t += i ? ",0.8" : "0.8";
        if (!parse_mod_threshes(t, &mods, &e)) { set_err(err, errlen, e); return -1; }
    }
    std::vector<mmc_mod_t> mm;
    if (!to_mmc_mods(mods, &mm, &e)) { set_err(err, errlen, e); return -1; }
    if ((int)mm.size() > max_out) { set_err(err, errlen, "too many modification codes"); return -1; }
    memcpy(out, mm.data(), sizeof(mmc_mod_t) * mm.size());
    return (int)mm.size();
}

// Format freq rows exactly as print_freq_header()+print_freq_output() would; append=0 truncates.
// fastfmt.h against the C library: returns the number of mismatches of fmt_f6 vs snprintf("%f") over
// (a) every n_mod/n_called and n_mod*100/n_called with 1 <= n_called <= max_den, 0 <= n_mod <= n_called,
// (b) n_random pseudo-random pairs of 32-bit counts (n_mod <= n_called) and raw doubles in [0, 2^40).
long long mmh_fastfmt_selftest(unsigned max_den, unsigned long long n_random, char *first_bad, int first_bad_len) {
    long long bad = 0;
    char a[64], b[64];
    auto check = [&](double x) {
        char *e = fmt_f6(a, x); *e = 0;
        snprintf(b, sizeof b, "%f", x);
        if (strcmp(a, b) != 0) { if (!bad && first_bad) snprintf(first_bad, first_bad_len, "%.17g: fmt_f6=%s printf=%s", x, a, b); ++bad; }
    };
    for (unsigned d = 1; d <= max_den; ++d)
        for (unsigned n = 0; n <= d; ++n) { check((double)n / d); check((double)n * 100 / d); }
    uint64_t s = 0x9e3779b97f4a7c15ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (unsigned long long i = 0; i < n_random; ++i) {
        uint32_t d = (uint32_t)(rnd() >> (rnd() % 32 + 32)) | 1u, n = (uint32_t)(rnd() % ((uint64_t)d + 1));
        check((double)n / d); check((double)n * 100 / d);
        double x = (double)(rnd() >> 24) / (double)(1ull << (rnd() % 40));          // dyadic values: exact ties happen here
        check(x);
        check((double)((rnd() % 2000000) * 2 + 1) / 2097152.0);                      // k/2^21: many exact .5 ties at 6 decimals
    }
    return bad;
}

int mmh_write_freq(const char *path, int bedmethyl, int insertions, int haplotypes, int n_contigs, const char *const *contig_names,
                   const mmc_freq_rec_t *recs, uint64_t n, int n_codes, const char *const *code_names) {
    FILE *fp = fopen(path, "w");
    if (!fp) return -1;
    OutOpts o; o.bedmethyl = bedmethyl; o.insertions = insertions; o.haplotypes = haplotypes;
    std::vector<std::string> names(contig_names, contig_names + n_contigs), codes(code_names, code_names + n_codes);
    print_freq_header(fp, o);
    print_freq_records(fp, o, names, recs, n, codes);
    fclose(fp);
    return 0;
}

// Format the view rows of one batch as print_view_output() would; append=0 writes the header first.
int mmh_write_view(const char *path, int append, int insertions, int haplotypes, int n_contigs, const char *const *contig_names,
                   const mmc_batch_t *batch, void *loader, const mmc_view_rec_t *recs, uint64_t n, int n_codes,
                   const char *const *code_names) {
    FILE *fp = fopen(path, append ? "a" : "w");
    if (!fp) return -1;
    OutOpts o; o.insertions = insertions; o.haplotypes = haplotypes;
    std::vector<std::string> names(contig_names, contig_names + n_contigs), codes(code_names, code_names + n_codes);
    if (!append) print_view_header(fp, o);
    print_view_records(fp, o, names, batch, ((Loader *)loader)->meta, recs, n, codes);
    fclose(fp);
    return 0;
}

}  // extern "C"
