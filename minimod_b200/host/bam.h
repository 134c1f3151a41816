// bam.h -- minimal BAM container reader for the host side (htslib is not in this image).
//
// Stands in for the htslib calls the reference's loader makes (sam_open, sam_hdr_read,
// sam_read1, bam_aux_get ...; src/minimod.c:73-89,250, src/mod.c:123-202).
// BGZF files (gzip members that announce their size in a 'BC' extra field) are inflated block-parallel
// by a small thread pool -- the equivalent of hts_set_threads(), src/minimod.c:76-78 (the
// inflate is the end-to-end limiter of the tool).  Anything else that gzip can read (plain .gz members, as the
// synthetic writer produces) goes through zlib's gzread().
#ifndef MMH_BAM_H
#define MMH_BAM_H

#include <stdint.h>
#include <stdio.h>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
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
    bool no_qual = false;               // QUAL is not in `data` (BamFile::next() skips it: two thirds of a record nobody here reads)
    std::vector<uint8_t> data;          // qname, cigar, seq, [qual,] aux

    const char *qname() const { return (const char *)data.data(); }
    const uint8_t *cigar() const { return data.data() + l_qname; }                    // unaligned u32 LE
    const uint8_t *seq() const { return cigar() + 4 * (size_t)n_cigar; }
    const uint8_t *aux() const { return seq() + ((size_t)l_qseq + 1) / 2 + (no_qual ? 0 : (size_t)l_qseq); }
    const uint8_t *end() const { return data.data() + l_data - (no_qual ? l_qseq : 0); }
    // first aux field with this tag: pointer to its type byte, or nullptr (bam_aux_get)
    const uint8_t *aux_get(const char tag[2]) const;
};

// Ordered, block-parallel BGZF inflate: one reader thread splits the file into blocks, `threads` workers
// inflate them, read() hands the bytes out in file order.
class BgzfReader {
public:
    ~BgzfReader();
    static bool is_bgzf(const std::string &path);
    bool open(const std::string &path, int threads);
    long read(void *buf, size_t n);              // bytes delivered (< n only at EOF), -1 on a corrupt block; buf == nullptr: skipped, not copied
private:
    enum State { EMPTY, FILLED, BUSY, DONE };
    struct Slot { std::vector<uint8_t> in, out; size_t out_len = 0; State st = EMPTY; bool bad = false; };
    void producer();
    void worker();
    FILE *fp_ = nullptr;
    std::vector<Slot> ring_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_;
    uint64_t produced_ = 0, claimed_ = 0, consumed_ = 0;   // block sequence numbers
    bool eof_ = false, stop_ = false, io_error_ = false;
    size_t cur_off_ = 0;                                   // bytes of block `consumed_` already handed out
};

class BamFile {
public:
    ~BamFile();
    bool open(const std::string &path, std::string *err, int threads = 4);
    // >0 record read, 0 clean EOF, <0 truncated/corrupt
    int next(BamRecord *rec);
    std::vector<std::string> names;
    std::vector<uint32_t> lens;
private:
    bool read_exact(void *buf, size_t n);
    long read_some(void *buf, size_t n);
    gzFile gz_ = nullptr;
    BgzfReader *bgzf_ = nullptr;
};

}  // namespace mmh
#endif
