#include "format.h"
#include <algorithm>
#include <string.h>

namespace mmh {

void print_freq_header(FILE *fp, const OutOpts &o) {              // src/mod.c:628-642
    if (o.bedmethyl) return;
    fprintf(fp, "contig\tstart\tend\tstrand\tn_called\tn_mod\tfreq\tmod_code%s%s\n",
            o.insertions ? "\tins_offset" : "", o.haplotypes ? "\thaplotype" : "");
}

void print_view_header(FILE *fp, const OutOpts &o) {              // src/mod.c:546-558
    fprintf(fp, "ref_contig\tref_pos\tstrand\tread_id\tread_pos\tmod_code\tmod_prob%s%s\n",
            o.insertions ? "\tins_offset" : "", o.haplotypes ? "\thaplotype" : "");
}

void print_freq_records(FILE *fp, const OutOpts &o, const std::vector<std::string> &names,
                        const mmc_freq_rec_t *recs, uint64_t n, const std::vector<std::string> &codes) {
    // group boundaries per tid (records are contiguous per tid), then order the groups by contig name
    struct Group { int32_t tid; uint64_t b, e; };
    std::vector<Group> groups;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i;
        while (j < n && recs[j].tid == recs[i].tid) ++j;
        groups.push_back({recs[i].tid, i, j});
        i = j;
    }
    std::stable_sort(groups.begin(), groups.end(), [&](const Group &x, const Group &y) {
        return strcmp(names[x.tid].c_str(), names[y.tid].c_str()) < 0;
    });
    for (const Group &g : groups) {
        const char *contig = names[g.tid].c_str();
        for (uint64_t i = g.b; i < g.e; ++i) {
            const mmc_freq_rec_t &r = recs[i];
            const char *code = codes[r.code].c_str();
            const char strand = r.strand ? '-' : '+';
            if (o.bedmethyl) {                                       // src/mod.c:672-688
                double f = (double)r.n_mod * 100 / r.n_called;
                int end = r.pos + 1;
                fprintf(fp, "%s\t%d\t%d\t%s\t%d\t%c\t%d\t%d\t255,0,0\t%d\t%f\n", contig, r.pos, end, code, (int)r.n_called,
                        strand, r.pos, end, (int)r.n_called, f);
            } else {                                                 // src/mod.c:691-718
                double f = (double)r.n_mod / r.n_called;
                fprintf(fp, "%s\t%d\t%d\t%c\t%d\t%d\t%f\t%s", contig, r.pos, r.pos, strand, (int)r.n_called, (int)r.n_mod, f, code);
                if (o.insertions) fprintf(fp, "\t%d", (int)r.ins_offset);
                if (o.haplotypes) { if (r.hap == -1) fputs("\t*", fp); else fprintf(fp, "\t%d", (int)r.hap); }
                fputc('\n', fp);
            }
        }
    }
}

void print_view_records(FILE *fp, const OutOpts &o, const std::vector<std::string> &names, const mmc_batch_t *batch,
                        const BatchMeta &meta, const mmc_view_rec_t *recs, uint64_t n, const std::vector<std::string> &codes) {
    for (uint64_t i = 0; i < n; ++i) {                               // src/mod.c:595-616
        const mmc_view_rec_t &v = recs[i];
        int32_t tid = batch->tid[v.read];
        const char *tname = tid >= 0 && (size_t)tid < names.size() ? names[tid].c_str() : "*";
        double p = (double)((v.mod_prob + 0.5) / 256.0);
        fprintf(fp, "%s\t%d\t%c\t%s\t%d\t%s\t%f", tname, v.ref_pos, v.strand ? '-' : '+', meta.qname(v.read), v.read_pos,
                codes[v.code].c_str(), p);
        if (o.insertions) fprintf(fp, "\t%d", (int)v.ins_offset);
        if (o.haplotypes) fprintf(fp, "\t%d", (int)v.hp);
        fputc('\n', fp);
    }
}

}  // namespace mmh
