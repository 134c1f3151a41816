// pack.h -- the host packer: load_db()'s read filters and -K/-B batching (src/minimod.c:235-333)
// writing straight into the library's pinned structure-of-arrays batch (mmc_batch_t).
#ifndef MMH_PACK_H
#define MMH_PACK_H
#include <stdint.h>
#include <functional>
#include <string>
#include <vector>
#include "bam.h"
#include "minimod_cuda.h"

namespace mmh {

struct LoadOpts {
    int32_t batch_size = 512;                      // -K
    int64_t batch_size_bases = 20 * 1000 * 1000;   // -B (bytes of l_data, despite the name)
    int allow_secondary = 0;
    int skip_supplementary = 0;
    int keep_qnames = 0;                           // view needs read names on the host
};

struct BatchStats {                                // db_t counters (src/minimod.h:152-155)
    int32_t total_reads = 0;
    int64_t total_bytes = 0;
    int32_t n_recs = 0;
    int64_t processed_bytes = 0;
    int64_t ml_entries = 0;                        // sum of ml_len over packed reads ("calls")
    int64_t bases = 0;
};

struct BatchMeta {                                 // host-only side data of one batch
    BatchStats stats;
    std::vector<uint32_t> qname_off;
    std::string qnames;
    const char *qname(uint32_t i) const { return qnames.c_str() + qname_off[i]; }
};

enum PackResult { kPacked, kSkipped, kNoSpace };

// Apply load_db()'s filters to `rec` and append it to `b` if it passes.
PackResult pack_record(const BamRecord &rec, mmc_batch_t *b, const LoadOpts &opt, BatchMeta *meta);

class BatchLoader {
public:
    BatchLoader(BamFile *bam, const LoadOpts &opt) : bam_(bam), opt_(opt) {}
    // Fill `b` (reset first).  Returns 1 if more records may follow, 0 at end of file, <0 on error.
    int fill(mmc_batch_t *b, BatchMeta *meta, std::string *err);
    // Multi-device runs: owner(record) names the device context a record belongs to (< 0: any).  A batch then ends where the
    // owner changes (a coordinate-sorted BAM changes owner a few hundred times at most), and next_owner() tells the caller
    // which context to acquire the next batch from (-1: none known yet / end of file).
    void set_owner(std::function<int(const BamRecord &)> fn) { owner_ = std::move(fn); }
    int next_owner(std::string *err);
private:
    std::function<int(const BamRecord &)> owner_;
    BamFile *bam_;
    LoadOpts opt_;
    BamRecord pending_;
    bool has_pending_ = false;
    bool eof_ = false;
};

}  // namespace mmh
#endif
