#include "pack.h"
#include <string.h>

namespace mmh {

static inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static inline uint64_t up16(uint64_t x) { return (x + (MMC_ALIGN - 1)) & ~(uint64_t)(MMC_ALIGN - 1); }

// (uint8_t)bam_aux2i() of the HP tag, 0 when absent (get_hp_tag, src/mod.c:188-202)
static uint8_t hp_of(const BamRecord &r) {
    const uint8_t *s = r.aux_get("HP");
    if (!s) return 0;
    int64_t v = 0;
    switch (*s) {
    case 'c': v = (int8_t)s[1]; break;
    case 'C': v = s[1]; break;
    case 's': v = (int16_t)(s[1] | s[2] << 8); break;
    case 'S': v = (uint16_t)(s[1] | s[2] << 8); break;
    case 'i': v = (int32_t)le32(s + 1); break;
    case 'I': v = le32(s + 1); break;
    default: v = 0;
    }
    return (uint8_t)v;
}

// ---- SEQ in its transport form (mmc_batch_t.seq2 / seq_exc, include/minimod_cuda.h): 2 bits per base, first base of
// a byte in bits 7:6, A 0 C 1 G 2 T 3; every other nt16 code travels as 0 plus an exception entry, and so does the
// pad nibble of an odd-length read, so that the device rebuilds BAM's 4-bit bytes exactly.
namespace {
inline bool nib_acgt(uint32_t n) { return n == 1 || n == 2 || n == 4 || n == 8; }
struct Seq2Tables {
    uint8_t code[256], bad[256];                  // per BAM byte (two bases): 4 bits of codes; any base without a code
    Seq2Tables() {
        auto c = [](uint32_t n) { return n == 2 ? 1u : n == 4 ? 2u : n == 8 ? 3u : 0u; };
        for (uint32_t x = 0; x < 256; ++x) {
            code[x] = (uint8_t)(c(x >> 4) << 2 | c(x & 15));
            bad[x] = (uint8_t)(!nib_acgt(x >> 4) || !nib_acgt(x & 15));
        }
    }
};
const Seq2Tables kSeq2;
}  // namespace

// returns the number of exception entries written at seq_exc[seq_exc_used ..) (not yet committed), -1 if they do not fit
static int64_t pack_seq2(mmc_batch_t *b, uint64_t s0, const uint8_t *seq, uint32_t L) {
    uint8_t *dst = b->seq2 + s0 / 2;
    const uint32_t nb = (L + 1) / 2, full = L / 2;            // BAM bytes; bytes that hold two bases
    uint32_t bad = 0, j = 0;
    for (; j + 1 < full; j += 2) {
        dst[j >> 1] = (uint8_t)(kSeq2.code[seq[j]] << 4 | kSeq2.code[seq[j + 1]]);
        bad |= kSeq2.bad[seq[j]] | kSeq2.bad[seq[j + 1]];
    }
    for (; j < nb; j += 2) {                                   // the last one or two bytes (one may hold the pad nibble)
        const uint32_t x0 = seq[j], x1 = j + 1 < nb ? seq[j + 1] : 0x11u;
        dst[j >> 1] = (uint8_t)(kSeq2.code[x0] << 4 | kSeq2.code[x1]);
        if (j < full) bad |= kSeq2.bad[x0];
        if (j + 1 < full) bad |= kSeq2.bad[x1];
    }
    uint64_t *e0 = b->seq_exc + b->seq_exc_used, *e = e0;
    const uint64_t room = b->seq_exc_cap - b->seq_exc_used, base = s0 * 2;
    if (bad || (L & 1u)) {
        const uint32_t first = bad ? 0u : L - 1u;              // only the last byte can matter when every full byte was clean
        for (uint32_t i = first; i < L; ++i) {
            const uint32_t nib = (seq[i >> 1] >> ((~i & 1u) << 2)) & 15u;
            if (nib_acgt(nib)) continue;
            if ((uint64_t)(e - e0) >= room) return -1;
            *e++ = ((base + i) << 4) | nib;
        }
        if (L & 1u) {
            if ((uint64_t)(e - e0) >= room) return -1;
            *e++ = (base + L) << 4;                            // BAM pads with 0
        }
    }
    return (int64_t)(e - e0);
}

// CIGAR in the byte form of include/minimod_cuda.h (cigar_packing == 8): blob at cig8[at], returns its size, or -1 if it
// does not fit the pool
static int64_t pack_cig8(mmc_batch_t *b, uint64_t at, const uint8_t *cig_bytes, uint32_t n) {
    auto word = [&](uint32_t i) { uint32_t w; memcpy(&w, cig_bytes + 4 * (size_t)i, 4); return w; };   // (BAM records are not aligned)
    uint32_t n1 = 0, n2 = 0;
    for (uint32_t i = 0; i < n; ++i) { const uint32_t len = word(i) >> 4; n1 += len >= 15u; n2 += len >= 15u + 255u; }
    const uint64_t size = 4 + (((uint64_t)n + 3) & ~3ull) + (((uint64_t)n1 + 3) & ~3ull) + 4ull * n2;
    if (at + size > b->cig8_cap) return -1;
    uint8_t *p = b->cig8 + at;
    memcpy(p, &n1, 4);
    uint8_t *ops = p + 4, *l1 = ops + ((n + 3u) & ~3u), *l2 = l1 + ((n1 + 3u) & ~3u);
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t w = word(i), len = w >> 4, op = w & 15u;
        if (len < 15u) { ops[i] = (uint8_t)(op | (len << 4)); continue; }
        ops[i] = (uint8_t)(op | 0xf0u);
        if (len < 15u + 255u) { *l1++ = (uint8_t)(len - 15u); continue; }
        *l1++ = 255u;
        memcpy(l2, &len, 4); l2 += 4;
    }
    return (int64_t)size;
}

PackResult pack_record(const BamRecord &rec, mmc_batch_t *b, const LoadOpts &opt, BatchMeta *meta) {
    // ---- filters, in the reference's order (src/minimod.c:260-284)
    if (rec.flag & 4) return kSkipped;                                   // BAM_FUNMAP
    if (!opt.allow_secondary && (rec.flag & 256)) return kSkipped;       // BAM_FSECONDARY
    if (opt.skip_supplementary && (rec.flag & 2048)) return kSkipped;    // BAM_FSUPPLEMENTARY
    if (rec.l_qseq == 0) return kSkipped;
    const uint8_t *mm_aux = rec.aux_get("MM");
    if (!mm_aux || (*mm_aux != 'Z' && *mm_aux != 'H')) return kSkipped;  // get_mm_tag_ptr(), src/mod.c:123-140
    const char *mm = (const char *)(mm_aux + 1);
    const size_t mm_len = strlen(mm);
    // ML must be B,C with len>0 (get_ml_tag, src/mod.c:142-185); otherwise the read is kept with no ML
    const uint8_t *ml = nullptr;
    uint32_t ml_len = 0;
    const uint8_t *ml_aux = rec.aux_get("ML");
    if (ml_aux && ml_aux[0] == 'B' && ml_aux[1] == 'C' && le32(ml_aux + 2) > 0) { ml_len = le32(ml_aux + 2); ml = ml_aux + 6; }

    // ---- capacity
    const uint64_t seq_bytes = ((uint64_t)rec.l_qseq + 1) / 2;
    const uint64_t c0 = up16(b->cigar_used * 4) / 4, s0 = up16(b->seq_used), m0 = up16(b->mm_used), l0 = up16(b->ml_used);
    if (b->n_reads >= b->max_reads || c0 + rec.n_cigar > b->cigar_cap || s0 + seq_bytes > b->seq_cap ||
        m0 + mm_len > b->mm_cap || l0 + ml_len > b->ml_cap)
        return kNoSpace;
    const bool two_bit = b->seq_packing == 2, cig_bytes = b->cigar_packing == 8;
    int64_t n_exc = 0, cig8_size = 0;
    const uint64_t g0 = up16(b->cig8_used);
    if (cig_bytes) {                                // written past cig8_used; committed with the other counters below
        cig8_size = pack_cig8(b, g0, (const uint8_t *)rec.cigar(), rec.n_cigar);
        if (cig8_size < 0) return kNoSpace;
    }
    if (two_bit) {                                  // written past seq_used; committed with the other counters below
        n_exc = pack_seq2(b, s0, rec.seq(), (uint32_t)rec.l_qseq);
        if (n_exc < 0) return kNoSpace;
    }

    const uint32_t i = b->n_reads;
    b->tid[i] = rec.tid; b->pos[i] = rec.pos; b->flag[i] = rec.flag;
    b->l_seq[i] = (uint32_t)rec.l_qseq; b->n_cigar[i] = rec.n_cigar;
    b->mm_len[i] = (uint32_t)mm_len; b->ml_len[i] = ml_len;
    b->hp[i] = hp_of(rec);
    b->cigar_off[i] = c0; b->seq_off[i] = s0; b->mm_off[i] = m0; b->ml_off[i] = l0;
    if (!cig_bytes) memcpy(b->cigar + c0, rec.cigar(), 4 * (size_t)rec.n_cigar);
    else { b->cig8_off[i] = g0; b->cig8_used = g0 + (uint64_t)cig8_size; }
    if (!two_bit) memcpy(b->seq4 + s0, rec.seq(), seq_bytes);
    else b->seq_exc_used += (uint64_t)n_exc;
    memcpy(b->mm + m0, mm, mm_len);
    if (ml_len) memcpy(b->ml + l0, ml, ml_len);
    b->cigar_used = c0 + rec.n_cigar; b->seq_used = s0 + seq_bytes; b->mm_used = m0 + mm_len; b->ml_used = l0 + ml_len;
    b->n_reads = i + 1;
    if (meta) {
        meta->stats.ml_entries += ml_len;
        meta->stats.bases += rec.l_qseq;
        if (opt.keep_qnames) {
            meta->qname_off.push_back((uint32_t)meta->qnames.size());
            meta->qnames.append(rec.qname());
            meta->qnames.push_back('\0');
        }
    }
    return kPacked;
}

int BatchLoader::next_owner(std::string *err) {
    for (;;) {
        if (!has_pending_) {
            if (eof_) return -1;
            int rc = bam_->next(&pending_);
            if (rc == 0) { eof_ = true; return -1; }
            if (rc < 0) { if (err) *err = "truncated or corrupt BAM record"; return -2; }
            has_pending_ = true;
        }
        const int o = owner_ ? owner_(pending_) : 0;
        if (o >= 0 || !owner_) return o;
        return -1;                                      // a record nobody owns (unmapped): it rides along with whatever batch comes next
    }
}

int BatchLoader::fill(mmc_batch_t *b, BatchMeta *meta, std::string *err) {
    b->n_reads = 0; b->cigar_used = b->seq_used = b->mm_used = b->ml_used = 0; b->seq_exc_used = 0; b->cig8_used = 0;
    meta->stats = BatchStats(); meta->qname_off.clear(); meta->qnames.clear();
    BatchStats &st = meta->stats;
    int batch_owner = -1;
    // while (n_bam_recs < cap && processed_bytes < batch_size_bases), src/minimod.c:249
    while (st.n_recs < opt_.batch_size && st.processed_bytes < opt_.batch_size_bases) {
        if (!has_pending_) {
            if (eof_) return 0;
            int rc = bam_->next(&pending_);
            if (rc == 0) { eof_ = true; return 0; }
            if (rc < 0) { if (err) *err = "truncated or corrupt BAM record"; return -1; }
            has_pending_ = true;
        }
        if (owner_) {                                   // the batch ends where the owning device changes
            const int o = owner_(pending_);
            if (o >= 0) {
                if (batch_owner < 0) batch_owner = o;
                else if (o != batch_owner) return 1;
            }
        }
        PackResult pr = pack_record(pending_, b, opt_, meta);
        if (pr == kNoSpace) {
            if (b->n_reads == 0) { if (err) *err = "a single read does not fit the batch buffers; raise -B"; return -1; }
            return 1;                                   // pinned pools full: end the batch early, keep the record
        }
        has_pending_ = false;
        st.total_reads++; st.total_bytes += pending_.l_data;
        if (pr == kPacked) { st.n_recs++; st.processed_bytes += pending_.l_data; }
    }
    return 1;
}

}  // namespace mmh
