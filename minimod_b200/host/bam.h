// bam.h -- minimal BAM container reader for the host side (htslib is not in this image).
//
// Stands in for the htslib calls the reference's loader makes (sam_open, sam_hdr_read,
// sam_read1, bam_aux_get ...; src/minimod.c:73-89,250, src/mod.c:123-202).  BGZF is a series
// of gzip members, which zlib's gzread() decodes transparently.
#ifndef MMH_BAM_H
#define MMH_BAM_H

#include <stdint.h>
#include <string>
#include <vector>
#include <zlib.h>

namespace mmh {

struct BamRecord {
    int32_t tid = -1, pos = -1;
    uint16_t flag = 0;
    uint32_t n_cigar = 0;
    int32_t l_qseq = 0;
    uint8_t l_qname = 0;
    int32_t l_data = 0;                 // block_size - 32, what load_db() budgets with (-B)
    std::vector<uint8_t> data;          // qname, cigar, seq, qual, aux

    const char *qname() const { return (const char *)data.data(); }
    const uint8_t *cigar() const { return data.data() + l_qname; }                    // unaligned u32 LE
    const uint8_t *seq() const { return cigar() + 4 * (size_t)n_cigar; }
    const uint8_t *aux() const { return seq() + ((size_t)l_qseq + 1) / 2 + (size_t)l_qseq; }
    const uint8_t *end() const { return data.data() + l_data; }
    // first aux field with this tag: pointer to its type byte, or nullptr (bam_aux_get)
    const uint8_t *aux_get(const char tag[2]) const;
};

class BamFile {
public:
    ~BamFile();
    bool open(const std::string &path, std::string *err);
    // >0 record read, 0 clean EOF, <0 truncated/corrupt
    int next(BamRecord *rec);
    std::vector<std::string> names;
    std::vector<uint32_t> lens;
private:
    bool read_exact(void *buf, size_t n);
    gzFile gz_ = nullptr;
};

}  // namespace mmh
#endif
