#include "format.h"
#include "fastfmt.h"
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

// rows [b,e) of one contig -> text, appended to *out (integer arithmetic only: fastfmt.h, fmt_f6 == printf("%f") exactly)
void format_freq_rows(std::string *out, const OutOpts &o, const std::string &contig_name, const mmc_freq_rec_t *recs, uint64_t b, uint64_t e,
                      const std::vector<std::string> &codes) {
    const char *contig = contig_name.c_str();
    const size_t contig_len = contig_name.size();
    char buf[1u << 16];
    char *p = buf, *const flush_at = buf + sizeof(buf) - 512;
    for (uint64_t i = b; i < e; ++i) {
        const mmc_freq_rec_t &r = recs[i];
        const std::string &code = codes[r.code];
        const char strand = r.strand ? '-' : '+';
        if (contig_len > 200 || code.size() > 64) {                  // pathological strings: stdio, through a sized buffer
            if (p != buf) { out->append(buf, (size_t)(p - buf)); p = buf; }
            std::string line(contig_len + code.size() + 256, '\0');
            int w;
            if (o.bedmethyl) {
                double f = (double)r.n_mod * 100 / r.n_called;
                w = snprintf(&line[0], line.size(), "%s\t%d\t%d\t%s\t%d\t%c\t%d\t%d\t255,0,0\t%d\t%f\n", contig, r.pos, r.pos + 1, code.c_str(), (int)r.n_called,
                             strand, r.pos, r.pos + 1, (int)r.n_called, f);
            } else {
                double f = (double)r.n_mod / r.n_called;
                w = snprintf(&line[0], line.size(), "%s\t%d\t%d\t%c\t%d\t%d\t%f\t%s", contig, r.pos, r.pos, strand, (int)r.n_called, (int)r.n_mod, f, code.c_str());
                if (o.insertions) w += snprintf(&line[w], line.size() - w, "\t%d", (int)r.ins_offset);
                if (o.haplotypes) { if (r.hap == -1) w += snprintf(&line[w], line.size() - w, "\t*"); else w += snprintf(&line[w], line.size() - w, "\t%d", (int)r.hap); }
                w += snprintf(&line[w], line.size() - w, "\n");
            }
            out->append(line.data(), (size_t)w);
            continue;
        }
        memcpy(p, contig, contig_len); p += contig_len;
        if (o.bedmethyl) {                                           // src/mod.c:672-688
            const double f = (double)r.n_mod * 100 / r.n_called;
            const int32_t end = r.pos + 1;
            *p++ = '\t'; p = fmt_i32(p, r.pos); *p++ = '\t'; p = fmt_i32(p, end); *p++ = '\t';
            memcpy(p, code.data(), code.size()); p += code.size();
            *p++ = '\t'; p = fmt_i32(p, (int32_t)r.n_called); *p++ = '\t'; *p++ = strand;
            *p++ = '\t'; p = fmt_i32(p, r.pos); *p++ = '\t'; p = fmt_i32(p, end);
            p = fmt_str(p, "\t255,0,0\t"); p = fmt_i32(p, (int32_t)r.n_called); *p++ = '\t'; p = fmt_f6(p, f);
        } else {                                                     // src/mod.c:691-718
            const double f = (double)r.n_mod / r.n_called;
            *p++ = '\t'; p = fmt_i32(p, r.pos); *p++ = '\t'; p = fmt_i32(p, r.pos); *p++ = '\t'; *p++ = strand;
            *p++ = '\t'; p = fmt_i32(p, (int32_t)r.n_called); *p++ = '\t'; p = fmt_i32(p, (int32_t)r.n_mod);
            *p++ = '\t'; p = fmt_f6(p, f); *p++ = '\t';
            memcpy(p, code.data(), code.size()); p += code.size();
            if (o.insertions) { *p++ = '\t'; p = fmt_i32(p, (int32_t)r.ins_offset); }
            if (o.haplotypes) { *p++ = '\t'; if (r.hap == -1) *p++ = '*'; else p = fmt_i32(p, (int32_t)r.hap); }
        }
        *p++ = '\n';
        if (p >= flush_at) { out->append(buf, (size_t)(p - buf)); p = buf; }
    }
    if (p != buf) out->append(buf, (size_t)(p - buf));
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
    std::string text;
    const uint64_t step = 1u << 16;                                  // rows per write
    for (const Group &g : groups)
        for (uint64_t b = g.b; b < g.e; b += step) {
            text.clear();
            format_freq_rows(&text, o, names[g.tid], recs, b, std::min(g.e, b + step), codes);
            fwrite(text.data(), 1, text.size(), fp);
        }
}

void print_view_records(FILE *fp, const OutOpts &o, const std::vector<std::string> &names, const mmc_batch_t *batch,
                        const BatchMeta &meta, const mmc_view_rec_t *recs, uint64_t n, const std::vector<std::string> &codes) {
    char prob_txt[256][16];                                          // (ml + 0.5) / 256 has 256 values: format them once
    for (int b = 0; b < 256; ++b) { char *e = fmt_f6(prob_txt[b], (double)((b + 0.5) / 256.0)); *e = 0; }
    for (uint64_t i = 0; i < n; ++i) {                               // src/mod.c:595-616
        const mmc_view_rec_t &v = recs[i];
        int32_t tid = batch->tid[v.read];
        const char *tname = tid >= 0 && (size_t)tid < names.size() ? names[tid].c_str() : "*";
        fprintf(fp, "%s\t%d\t%c\t%s\t%d\t%s\t%s", tname, v.ref_pos, v.strand ? '-' : '+', meta.qname(v.read), v.read_pos,
                codes[v.code].c_str(), prob_txt[v.mod_prob]);
        if (o.insertions) fprintf(fp, "\t%d", (int)v.ins_offset);
        if (o.haplotypes) fprintf(fp, "\t%d", (int)v.hp);
        fputc('\n', fp);
    }
}

}  // namespace mmh
