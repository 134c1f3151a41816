// fasta.h -- FASTA(.gz) reader with the record semantics of the reference's kseq usage
// (load_ref, src/ref.c:46-89): name = header up to the first whitespace, sequence = all lines
// joined (CR stripped); a later record with the same name replaces an earlier one.
#ifndef MMH_FASTA_H
#define MMH_FASTA_H
#include <string>
#include <vector>
namespace mmh {
struct FastaRecord { std::string name, seq; };
// threads > 1: plain FASTA files are mapped and parsed record-parallel; gzip / FASTQ input falls back to the serial reader
bool read_fasta(const std::string &path, std::vector<FastaRecord> *out, std::string *err, int threads = 1);
}
#endif
