#include "bam.h"
#include <string.h>
#include <algorithm>

namespace mmh {

static inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static inline uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }

// ---------------------------------------------------------------------------------------------------------
// BGZF: gzip member = 12-byte header (XLEN at 10) + extra field with subfield 'B','C',2,BSIZE + raw deflate
// + CRC32 + ISIZE; BSIZE + 1 is the size of the whole member (SAM spec 4.1).
// ---------------------------------------------------------------------------------------------------------
static const size_t kRing = 4096;   // blocks in flight: ~256 MB of inflated data can run ahead of the parser (the device start-up and every pause of the consumer are used)

bool BgzfReader::is_bgzf(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    uint8_t h[18];
    const bool ok = fread(h, 1, 18, f) == 18 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && le16(h + 10) >= 6 &&
                    h[12] == 'B' && h[13] == 'C' && le16(h + 14) == 2;
    fclose(f);
    return ok;
}

bool BgzfReader::open(const std::string &path, int threads) {
    fp_ = fopen(path.c_str(), "rb");
    if (!fp_) return false;
    ring_.resize(kRing);
    if (threads < 1) threads = 1;
    threads_.emplace_back(&BgzfReader::producer, this);
    for (int i = 0; i < threads; ++i) threads_.emplace_back(&BgzfReader::worker, this);
    return true;
}

BgzfReader::~BgzfReader() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto &t : threads_) t.join();
    if (fp_) fclose(fp_);
}

void BgzfReader::producer() {
    std::vector<uint8_t> hdr(18);
    for (;;) {
        Slot *s;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || ring_[produced_ % kRing].st == EMPTY; });
            if (stop_) return;
            s = &ring_[produced_ % kRing];
        }
        bool bad = false, end = false;
        size_t got = fread(hdr.data(), 1, 18, fp_);
        if (got == 0) end = true;
        else if (got != 18 || hdr[0] != 0x1f || hdr[1] != 0x8b || hdr[2] != 8 || !(hdr[3] & 4)) bad = true;
        else {
            // find the BC subfield (it is the first one in every BGZF writer, but the spec allows others)
            const size_t xlen = le16(hdr.data() + 10);
            std::vector<uint8_t> extra(xlen);
            memcpy(extra.data(), hdr.data() + 12, 6);
            if (xlen < 6 || (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, fp_) != xlen - 6)) bad = true;
            size_t bsize = 0;
            for (size_t o = 0; !bad && o + 4 <= xlen;) {
                const size_t slen = le16(extra.data() + o + 2);
                if (extra[o] == 'B' && extra[o + 1] == 'C' && slen == 2 && o + 6 <= xlen) { bsize = (size_t)le16(extra.data() + o + 4) + 1; break; }
                o += 4 + slen;
            }
            if (!bad && (bsize < 12 + xlen + 8)) bad = true;
            if (!bad) {
                const size_t rest = bsize - 12 - xlen;                              // deflate data + CRC32 + ISIZE
                s->in.resize(rest);
                if (fread(s->in.data(), 1, rest, fp_) != rest) bad = true;
            }
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (end) { eof_ = true; cv_.notify_all(); return; }
            s->bad = bad; s->st = FILLED;
            ++produced_;
            if (bad) { eof_ = true; cv_.notify_all(); return; }
        }
        cv_.notify_all();
    }
}

void BgzfReader::worker() {
    z_stream zs;
    for (;;) {
        Slot *s;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || (claimed_ < produced_ && ring_[claimed_ % kRing].st == FILLED) || (eof_ && claimed_ == produced_); });
            if (stop_ || (eof_ && claimed_ == produced_)) return;
            s = &ring_[claimed_ % kRing];
            s->st = BUSY;
            ++claimed_;
        }
        bool bad = s->bad;
        size_t out_len = 0;
        if (!bad) {
            const size_t n = s->in.size();
            const uint32_t isize = le32(s->in.data() + n - 4), crc = le32(s->in.data() + n - 8);
            s->out.resize(isize ? isize : 1);
            memset(&zs, 0, sizeof zs);
            if (isize > (1u << 16) || inflateInit2(&zs, -15) != Z_OK) bad = true;
            else {
                zs.next_in = s->in.data(); zs.avail_in = (uInt)(n - 8);
                zs.next_out = s->out.data(); zs.avail_out = (uInt)isize;
                const int rc = isize ? inflate(&zs, Z_FINISH) : Z_STREAM_END;
                if (rc != Z_STREAM_END || zs.total_out != isize) bad = true;
                inflateEnd(&zs);
                if (!bad && isize && (uint32_t)crc32(crc32(0L, Z_NULL, 0), s->out.data(), (uInt)isize) != crc) bad = true;
                out_len = isize;
            }
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            s->bad = bad; s->out_len = out_len; s->st = DONE;
        }
        cv_.notify_all();
    }
}

long BgzfReader::read(void *buf, size_t n) {
    uint8_t *p = (uint8_t *)buf;
    size_t done = 0;
    while (done < n) {
        Slot *s;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return ring_[consumed_ % kRing].st == DONE || (eof_ && consumed_ == produced_); });
            if (ring_[consumed_ % kRing].st != DONE) break;                         // end of file
            s = &ring_[consumed_ % kRing];
        }
        if (s->bad) return -1;
        const size_t take = std::min(n - done, s->out_len - cur_off_);
        if (p) memcpy(p + done, s->out.data() + cur_off_, take);
        done += take; cur_off_ += take;
        if (cur_off_ == s->out_len) {
            { std::lock_guard<std::mutex> lk(mu_); s->st = EMPTY; ++consumed_; }
            cur_off_ = 0;
            cv_.notify_all();
        }
    }
    return (long)done;
}

BamFile::~BamFile() { if (gz_) gzclose(gz_); delete bgzf_; }

long BamFile::read_some(void *buf, size_t n) {
    if (bgzf_) return bgzf_->read(buf, n);
    if (!buf) {                                                   // skip (plain gzip stream): through a scratch buffer
        std::vector<uint8_t> scratch(std::min<size_t>(n, 1u << 16));
        size_t left = n;
        while (left) {
            const long got = read_some(scratch.data(), std::min(left, scratch.size()));
            if (got < 0) return -1;
            if (got == 0) break;
            left -= (size_t)got;
        }
        return (long)(n - left);
    }
    size_t done = 0;
    uint8_t *p = (uint8_t *)buf;
    while (done < n) {
        unsigned chunk = n - done > (1u << 30) ? (1u << 30) : (unsigned)(n - done);
        int got = gzread(gz_, p + done, chunk);
        if (got < 0) return -1;
        if (got == 0) break;
        done += (size_t)got;
    }
    return (long)done;
}

bool BamFile::read_exact(void *buf, size_t n) { return read_some(buf, n) == (long)n; }

bool BamFile::open(const std::string &path, std::string *err, int threads) {
    if (BgzfReader::is_bgzf(path)) {
        bgzf_ = new BgzfReader();
        if (!bgzf_->open(path, threads)) { if (err) *err = "cannot open " + path; return false; }
    } else {
        gz_ = gzopen(path.c_str(), "rb");
        if (!gz_) { if (err) *err = "cannot open " + path; return false; }
        gzbuffer(gz_, 4u << 20);
    }
    uint8_t b4[4];
    if (!read_exact(b4, 4) || memcmp(b4, "BAM\1", 4) != 0) { if (err) *err = path + " is not a BAM file"; return false; }
    if (!read_exact(b4, 4)) { if (err) *err = "truncated BAM header"; return false; }
    std::vector<char> text(le32(b4));
    if (!text.empty() && !read_exact(text.data(), text.size())) { if (err) *err = "truncated BAM header"; return false; }
    if (!read_exact(b4, 4)) { if (err) *err = "truncated BAM header"; return false; }
    uint32_t n_ref = le32(b4);
    names.resize(n_ref); lens.resize(n_ref);
    for (uint32_t i = 0; i < n_ref; ++i) {
        if (!read_exact(b4, 4)) { if (err) *err = "truncated BAM header"; return false; }
        uint32_t l_name = le32(b4);
        std::vector<char> nm(l_name);
        if (l_name && !read_exact(nm.data(), l_name)) { if (err) *err = "truncated BAM header"; return false; }
        names[i] = std::string(nm.data(), l_name ? strnlen(nm.data(), l_name) : 0);
        if (!read_exact(b4, 4)) { if (err) *err = "truncated BAM header"; return false; }
        lens[i] = le32(b4);
    }
    return true;
}

static size_t aux_size(const uint8_t *s, const uint8_t *end);

// BAM's n_cigar field has 16 bits; a CIGAR with more ops (ultra-long ONT reads) is stored as the placeholder
// <l_seq>S<ref_len>N with the real ops in a CG:B,I tag (SAM spec 4.2.2).  htslib puts them back when it reads the
// record (bam_tag2cigar() in sam_read1()), so the reference sees the real CIGAR: do the same, and drop the tag.
static void restore_long_cigar(BamRecord *r) {
    if (r->n_cigar == 0 || r->tid < 0 || r->pos < 0) return;
    uint32_t c0; memcpy(&c0, r->cigar(), 4);
    if ((c0 & 15u) != 4u || (c0 >> 4) != (uint32_t)r->l_qseq) return;
    const uint8_t *cg = r->aux_get("CG");
    if (!cg || cg[0] != 'B' || cg[1] != 'I') return;
    uint32_t n; memcpy(&n, cg + 2, 4);
    if (n < r->n_cigar || n >= (1u << 29)) return;
    const uint8_t *tag0 = cg - 2, *tag1 = cg + 6 + 4 * (size_t)n, *end = r->end();
    if (tag1 > end) return;
    std::vector<uint8_t> d;
    d.reserve((size_t)r->l_data + 4 * (size_t)n + 8);
    const uint8_t *base = r->data.data();
    d.insert(d.end(), base, r->cigar());                                  // qname
    d.insert(d.end(), cg + 6, cg + 6 + 4 * (size_t)n);                    // the real ops
    d.insert(d.end(), r->seq(), tag0);                                    // seq, qual, aux before CG
    d.insert(d.end(), tag1, end);                                         // aux after CG
    r->n_cigar = n;
    r->l_data = (int32_t)(d.size() + (r->no_qual ? (size_t)r->l_qseq : 0));   // as if QUAL were there: the length htslib's record has
    d.resize(d.size() + 8, 0);
    r->data.swap(d);
}

int BamFile::next(BamRecord *r) {
    uint8_t b4[4], fx[32];
    long got = read_some(b4, 4);
    if (got == 0) return 0;
    if (got != 4) return -1;
    uint32_t block = le32(b4);
    if (block < 32 || !read_exact(fx, 32)) return -1;
    r->tid = (int32_t)le32(fx); r->pos = (int32_t)le32(fx + 4);
    r->l_qname = fx[8];
    r->n_cigar = le16(fx + 12); r->flag = le16(fx + 14);
    r->l_qseq = (int32_t)le32(fx + 16);
    r->l_data = (int32_t)(block - 32);
    const size_t lq = (size_t)(r->l_qseq < 0 ? 0 : r->l_qseq);
    const size_t head = (size_t)r->l_qname + 4 * (size_t)r->n_cigar + (lq + 1) / 2, fixed = head + lq;
    if (r->l_qseq < 0 || fixed > (size_t)r->l_data) return -1;
    // qname | cigar | seq, then QUAL is skipped (never copied out of the inflated block), then the aux fields
    r->no_qual = true;
    const size_t kept = (size_t)r->l_data - lq;
    r->data.resize(kept + 8);
    if (head && !read_exact(r->data.data(), head)) return -1;
    if (lq && read_some(nullptr, lq) != (long)lq) return -1;
    if (kept > head && !read_exact(r->data.data() + head, kept - head)) return -1;
    memset(r->data.data() + kept, 0, 8);
    restore_long_cigar(r);
    return 1;
}

static size_t aux_size(const uint8_t *s, const uint8_t *end) {       // size from the type byte on; 0 = malformed
    if (s >= end) return 0;
    switch (*s) {
    case 'A': case 'c': case 'C': return 2;
    case 's': case 'S': return 3;
    case 'i': case 'I': case 'f': return 5;
    case 'd': return 9;
    case 'Z': case 'H': { const uint8_t *p = s + 1; while (p < end && *p) ++p; return (size_t)(p - s) + 1; }
    case 'B': {
        if (s + 6 > end) return 0;
        size_t e;
        switch (s[1]) { case 'c': case 'C': e = 1; break; case 's': case 'S': e = 2; break; case 'i': case 'I': case 'f': e = 4; break; default: return 0; }
        return 6 + e * (size_t)le32(s + 2);
    }
    default: return 0;
    }
}

const uint8_t *BamRecord::aux_get(const char tag[2]) const {
    const uint8_t *s = aux(), *e = end();
    while (s + 3 <= e) {
        size_t sz = aux_size(s + 2, e);
        if (!sz || s + 2 + sz > e) return nullptr;
        if (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]) return s + 2;
        s += 2 + sz;
    }
    return nullptr;
}

}  // namespace mmh
