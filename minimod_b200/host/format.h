// format.h -- text output identical to print_freq_output()/print_view_output()
// (src/mod.c:560-728).  Contigs are ordered by strcmp of their names, as cmp_key_fast()
// orders them (src/mod.c:59-87), not by header order.
#ifndef MMH_FORMAT_H
#define MMH_FORMAT_H
#include <stdio.h>
#include <string>
#include <vector>
#include "minimod_cuda.h"
#include "pack.h"
namespace mmh {
struct OutOpts { int bedmethyl = 0, insertions = 0, haplotypes = 0; };
void print_freq_header(FILE *fp, const OutOpts &o);
void print_view_header(FILE *fp, const OutOpts &o);
// recs sorted by (tid,pos,...) as mmc_freq_finalize() returns them; code_names[code] gives the string
void print_freq_records(FILE *fp, const OutOpts &o, const std::vector<std::string> &contig_names,
                        const mmc_freq_rec_t *recs, uint64_t n, const std::vector<std::string> &code_names);
// rows [b,e) of one contig (all of tid == that contig) appended to *out as text: what print_freq_records() writes for them
void format_freq_rows(std::string *out, const OutOpts &o, const std::string &contig_name, const mmc_freq_rec_t *recs, uint64_t b, uint64_t e,
                      const std::vector<std::string> &code_names);
void print_view_records(FILE *fp, const OutOpts &o, const std::vector<std::string> &contig_names,
                        const mmc_batch_t *batch, const BatchMeta &meta,
                        const mmc_view_rec_t *recs, uint64_t n, const std::vector<std::string> &code_names);
}
#endif
