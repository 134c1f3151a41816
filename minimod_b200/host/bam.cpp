#include "bam.h"
#include <string.h>

namespace mmh {

static inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static inline uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }

BamFile::~BamFile() { if (gz_) gzclose(gz_); }

bool BamFile::read_exact(void *buf, size_t n) {
    uint8_t *p = (uint8_t *)buf;
    while (n) {
        unsigned chunk = n > (1u << 30) ? (1u << 30) : (unsigned)n;
        int got = gzread(gz_, p, chunk);
        if (got <= 0) return false;
        p += got; n -= (size_t)got;
    }
    return true;
}

bool BamFile::open(const std::string &path, std::string *err) {
    gz_ = gzopen(path.c_str(), "rb");
    if (!gz_) { if (err) *err = "cannot open " + path; return false; }
    gzbuffer(gz_, 4u << 20);
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

int BamFile::next(BamRecord *r) {
    uint8_t b4[4], fx[32];
    int got = gzread(gz_, b4, 4);
    if (got == 0) return 0;
    if (got != 4) return -1;
    uint32_t block = le32(b4);
    if (block < 32 || !read_exact(fx, 32)) return -1;
    r->tid = (int32_t)le32(fx); r->pos = (int32_t)le32(fx + 4);
    r->l_qname = fx[8];
    r->n_cigar = le16(fx + 12); r->flag = le16(fx + 14);
    r->l_qseq = (int32_t)le32(fx + 16);
    r->l_data = (int32_t)(block - 32);
    r->data.resize((size_t)r->l_data + 8);
    if (r->l_data && !read_exact(r->data.data(), (size_t)r->l_data)) return -1;
    memset(r->data.data() + r->l_data, 0, 8);
    size_t fixed = (size_t)r->l_qname + 4 * (size_t)r->n_cigar + ((size_t)(r->l_qseq < 0 ? 0 : r->l_qseq) + 1) / 2 + (size_t)(r->l_qseq < 0 ? 0 : r->l_qseq);
    if (r->l_qseq < 0 || fixed > (size_t)r->l_data) return -1;
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
