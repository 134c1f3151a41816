// synth.cpp -- deterministic synthetic workloads of the shapes BASELINE.json names
// (SURVEY.md section 8(d)): a reference genome and coordinate-sorted reads with CIGAR, MM:Z and
// ML:B:C tags, produced as in-memory BAM records.  The same records are either packed straight
// into a pinned mmc_batch_t (GPU arm) or written as a BAM + FASTA for the unmodified reference
// binary (CPU arm), so both arms always see identical input.  All integer, seeded.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "bam.h"
#include "pack.h"

namespace mmh {

struct Rng {                                      // splitmix64 seeding + xorshift64*
    uint64_t s;
    explicit Rng(uint64_t seed) { s = seed + 0x9e3779b97f4a7c15ull; s = mix(s); if (!s) s = 1; }
    static uint64_t mix(uint64_t z) { z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
    uint64_t next() { s ^= s >> 12; s ^= s << 25; s ^= s >> 27; return s * 0x2545f4914f6cdd1dull; }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    bool chance(uint32_t per_million) { return below(1000000u) < per_million; }
    uint32_t geometric(double mean) {             // >= 1
        double p = 1.0 / mean, u = uniform();
        uint32_t k = 1 + (uint32_t)floor(log(1.0 - u) / log(1.0 - p));
        return k < 1 ? 1 : k;
    }
    double normal() { double u1 = uniform(), u2 = uniform(); if (u1 < 1e-300) u1 = 1e-300; return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2); }
};

struct SynthConfig {
    int id;
    // read length model: 0 normal(mean,sd) clipped [lo,hi]; 1 log-normal(median, mean) ; 2 fixed
    int len_model; double len_a, len_b; uint32_t len_lo, len_hi;
    uint32_t ins_ppm, del_ppm, mm_ppm;            // per-base rates (parts per million)
    uint32_t softclip_ppm;                        // per read end
    int tag_style;                                // 0 C+m? CpG; 1 C+h?;C+m? CpG; 2 C+m. all C + A+a. all A + HP; 3 as 1 but '.' status
    double coverage;
};

static SynthConfig config_of(int id) {
    switch (id) {
    case 2: return {2, 0, 15000, 2000, 5000, 25000, 500, 900, 1000, 0, 0, 30.0};
    case 3: return {3, 1, 8000, 10000, 200, 200000, 13000, 27000, 10000, 50000, 1, 30.0};
    case 4: return {4, 2, 50000, 0, 50000, 50000, 13000, 27000, 10000, 50000, 2, 30.0};
    case 5: return {5, 1, 12000, 15000, 200, 300000, 13000, 27000, 10000, 50000, 0, 30.0};
    case 6: return {6, 1, 8000, 10000, 200, 200000, 13000, 27000, 10000, 50000, 3, 30.0};   // C3 variant with '.' status (Q9)
    case 7: return {7, 2, 1300000, 0, 1300000, 1300000, 13000, 27000, 10000, 50000, 0, 30.0};   // ultra-long ONT: > 65535 CIGAR ops (CG:B,I tag)
    default: return {id, 0, 15000, 2000, 5000, 25000, 500, 900, 1000, 0, 0, 30.0};
    }
}

struct Contig { std::string name; std::string seq; std::vector<std::pair<uint32_t, uint32_t>> mappable; uint64_t mappable_len = 0;
                uint32_t gid = 0; uint64_t read_base = 0, n_reads = 0; };   // gid: job-wide contig id (seeds); reads [read_base, read_base + n_reads)

class Synth {
public:
    SynthConfig cfg;
    uint64_t seed;
    std::vector<Contig> contigs;
    uint64_t total_mappable = 0;
    uint64_t n_reads = 0;

    // Every contig is generated from (seed, gid) alone -- its sequence and its reads -- so a job is the same set of
    // reads however its contigs are dealt to instances (contig sharding across GPUs).
    void build_reference(const std::vector<std::pair<std::string, uint32_t>> &table, const uint32_t *gids) {
        contigs.resize(table.size());
        for (size_t ci = 0; ci < table.size(); ++ci) contigs[ci].gid = gids ? gids[ci] : (uint32_t)ci;
        std::vector<std::thread> th;
        const size_t nt = std::max<size_t>(1, std::min<size_t>(table.size(), std::thread::hardware_concurrency() ? std::thread::hardware_concurrency() : 8));
        for (size_t t = 0; t < nt; ++t)
            th.emplace_back([this, t, nt, &table]() { for (size_t ci = t; ci < table.size(); ci += nt) gen_contig(ci, table[ci].first, table[ci].second); });
        for (auto &t : th) t.join();
        total_mappable = 0;
        const double mean_len = cfg.len_model == 1 ? cfg.len_b : cfg.len_a;
        n_reads = 0;
        for (auto &c : contigs) {
            total_mappable += c.mappable_len;
            c.read_base = n_reads;
            c.n_reads = (uint64_t)(cfg.coverage * (double)c.mappable_len / mean_len);
            n_reads += c.n_reads;
        }
    }

    void gen_contig(size_t ci, const std::string &name, uint32_t len) {
        Contig &c = contigs[ci];
        c.name = name;
        c.seq.resize(len);
        Rng r(seed * 1000003ull + (uint64_t)c.gid * 7919ull + 17);
        char *s = &c.seq[0];
        for (uint32_t i = 0; i < len; ++i) {
            uint32_t u = r.below(100);
            s[i] = u < 29 ? 'A' : u < 50 ? 'C' : u < 71 ? 'G' : 'T';
        }
        // CpG depletion to about one CpG per 55 bp: deaminate 59 % of the CpG cytosines (C->T)
        for (uint32_t i = 0; i + 1 < len; ++i)
            if (s[i] == 'C' && s[i + 1] == 'G' && r.below(100) < 59) s[i] = 'T';
        // N runs: 10 kb at the contig start, one 1 Mb run in the middle (scaled down for short contigs)
        uint32_t n0 = std::min<uint32_t>(10000, len / 10), n1 = std::min<uint32_t>(1000000, len / 20);
        uint32_t mid = len / 2;
        memset(s, 'N', n0);
        memset(s + mid, 'N', n1);
        c.mappable.clear();
        if (mid > n0) c.mappable.push_back({n0, mid});
        if (len > mid + n1) c.mappable.push_back({mid + n1, len});
        c.mappable_len = 0;
        for (auto &iv : c.mappable) c.mappable_len += iv.second - iv.first;
    }

    // read `idx` (0..n_reads-1), coordinate-sorted by construction (stratified uniform starts)
    void make_read(uint64_t idx, BamRecord *rec) const {
        size_t ci = 0, iv = 0;
        {
            size_t lo = 0, hi = contigs.size();                 // last contig with read_base <= idx (coordinate order == index order)
            while (hi - lo > 1) { size_t mid = (lo + hi) / 2; if (contigs[mid].read_base <= idx) lo = mid; else hi = mid; }
            ci = lo;
        }
        const Contig &c = contigs[ci];
        const uint64_t j = idx - c.read_base;                   // the contig's j-th read
        Rng r(seed ^ Rng::mix(((uint64_t)c.gid + 1) * 0x9e3779b97f4a7c15ull + j * 2 + 1));
        // ---- length
        uint32_t L;
        if (cfg.len_model == 0) { double v = cfg.len_a + cfg.len_b * r.normal(); L = (uint32_t)std::max<double>(cfg.len_lo, std::min<double>(cfg.len_hi, v)); }
        else if (cfg.len_model == 1) {
            double mu = log(cfg.len_a), sigma = sqrt(2.0 * log(cfg.len_b / cfg.len_a));
            double v = exp(mu + sigma * r.normal());
            L = (uint32_t)std::max<double>(cfg.len_lo, std::min<double>(cfg.len_hi, v));
        } else L = (uint32_t)cfg.len_a;
        // ---- start: stratum j of the contig's mappable space
        double frac = ((double)j + r.uniform()) / (double)std::max<uint64_t>(1, c.n_reads);
        uint64_t off = (uint64_t)(frac * (double)c.mappable_len);
        if (off >= c.mappable_len) off = c.mappable_len ? c.mappable_len - 1 : 0;
        for (iv = 0; iv < c.mappable.size(); ++iv) {
            uint32_t l = c.mappable[iv].second - c.mappable[iv].first;
            if (off < l) break;
            off -= l;
        }
        if (iv == c.mappable.size()) { iv = c.mappable.size() - 1; off = 0; }
        uint32_t iv_b = c.mappable[iv].first, iv_e = c.mappable[iv].second;
        uint32_t start = iv_b + (uint32_t)off;
        if (iv_e - iv_b > 64 && start > iv_e - 64) start = iv_e - 64;

        const bool rev = r.below(2) == 1;
        const bool supp = r.below(100) < 2;
        uint32_t clip5 = 0, clip3 = 0;
        if (supp) { clip5 = 20 + r.below(L / 4 + 1); clip3 = 20 + r.below(L / 4 + 1); }
        else { if (r.chance(cfg.softclip_ppm)) clip5 = 1 + r.below(100); if (r.chance(cfg.softclip_ppm)) clip3 = 1 + r.below(100); }
        if (clip5 + clip3 + 10 > L) { clip5 = 0; clip3 = 0; }

        // ---- alignment walk in reference (== BAM SEQ) orientation
        std::string seq;
        seq.reserve(L + 64);
        std::vector<uint32_t> cigar;
        auto push = [&](uint32_t op, uint32_t len) {
            if (!len) return;
            if (!cigar.empty() && (cigar.back() & 15u) == op) cigar.back() += len << 4; else cigar.push_back(len << 4 | op);
        };
        static const char B[4] = {'A', 'C', 'G', 'T'};
        for (uint32_t i = 0; i < clip5; ++i) seq.push_back(B[r.below(4)]);
        push(4, clip5);
        uint32_t rp = start;
        const uint32_t body = L - clip5 - clip3;
        const char *ref = c.seq.data();
        bool first = true;
        while (seq.size() < clip5 + body && rp < iv_e) {
            uint32_t u = r.below(1000000u);
            if (!first && u < cfg.ins_ppm) {
                uint32_t n = std::min<uint32_t>(r.geometric(1.6), clip5 + body - (uint32_t)seq.size());
                for (uint32_t k = 0; k < n; ++k) seq.push_back(B[r.below(4)]);
                push(1, n);
            } else if (!first && u < cfg.ins_ppm + cfg.del_ppm) {
                uint32_t n = std::min<uint32_t>(r.geometric(1.6), iv_e - rp - 1);
                push(2, n); rp += n;
            } else {
                char b = ref[rp];
                if (r.chance(cfg.mm_ppm)) { char nb; do { nb = B[r.below(4)]; } while (nb == b); b = nb; }
                seq.push_back(b); push(0, 1); ++rp;
            }
            first = false;
        }
        // the last aligned op must be a match so that the record is well formed
        while (!cigar.empty() && ((cigar.back() & 15u) == 2u)) { rp -= cigar.back() >> 4; cigar.pop_back(); }
        for (uint32_t i = 0; i < clip3; ++i) seq.push_back(B[r.below(4)]);
        push(4, clip3);
        L = (uint32_t)seq.size();

        // ---- MM / ML in original-read orientation
        std::string mm;
        std::vector<uint8_t> ml;
        auto prob = [&]() -> uint8_t {
            uint32_t u = r.below(100);
            return (uint8_t)(u < 55 ? 205 + r.below(51) : u < 90 ? r.below(51) : 51 + r.below(154));
        };
        // canonical base `cb` of the original read sits in SEQ as fwd: cb, rev: comp(cb), scanned from the end
        auto emit_block = [&](char cb, const char *hdr, bool cpg_only, std::vector<uint8_t> *mlv) {
            mm += hdr;
            const char want = rev ? (cb == 'C' ? 'G' : cb == 'A' ? 'T' : cb == 'G' ? 'C' : 'A') : cb;
            uint32_t skip = 0;
            char buf[16];
            for (uint32_t k = 0; k < L; ++k) {
                uint32_t q = rev ? L - 1 - k : k;
                if (seq[q] != want) continue;
                bool call = true;
                if (cpg_only) call = rev ? (q > 0 && seq[q - 1] == 'C') : (q + 1 < L && seq[q + 1] == 'G');
                if (call) { int n = snprintf(buf, sizeof(buf), ",%u", skip); mm.append(buf, (size_t)n); mlv->push_back(prob()); skip = 0; }
                else ++skip;
            }
            mm.push_back(';');
        };
        uint8_t hp = 0; bool has_hp = false;
        if (cfg.tag_style == 0) emit_block('C', "C+m?", true, &ml);
        else if (cfg.tag_style == 1 || cfg.tag_style == 3) {
            const bool dot = cfg.tag_style == 3;
            emit_block('C', dot ? "C+h." : "C+h?", true, &ml);
            emit_block('C', dot ? "C+m." : "C+m?", true, &ml);
        } else {
            emit_block('C', "C+m.", false, &ml);
            emit_block('A', "A+a.", false, &ml);
            uint32_t u = r.below(100);
            if (u < 40) { hp = 1; has_hp = true; } else if (u < 80) { hp = 2; has_hp = true; }
        }

        // ---- BAM record
        char qname[40];
        int qn = snprintf(qname, sizeof(qname), "synth%d_%010llu", cfg.id, (unsigned long long)idx) + 1;
        rec->tid = (int32_t)ci; rec->pos = (int32_t)start;
        rec->flag = (uint16_t)((rev ? 16 : 0) | (supp ? 2048 : 0));
        rec->n_cigar = (uint32_t)cigar.size(); rec->l_qseq = (int32_t)L; rec->l_qname = (uint8_t)qn;
        size_t sz = (size_t)qn + 4 * cigar.size() + (L + 1) / 2 + L + 3 + mm.size() + 1 + 8 + ml.size() + (has_hp ? 4 : 0);
        rec->data.assign(sz + 8, 0);
        uint8_t *p = rec->data.data();
        memcpy(p, qname, (size_t)qn); p += qn;
        memcpy(p, cigar.data(), 4 * cigar.size()); p += 4 * cigar.size();
        for (uint32_t i = 0; i < L; ++i) {
            uint8_t code = seq[i] == 'A' ? 1 : seq[i] == 'C' ? 2 : seq[i] == 'G' ? 4 : seq[i] == 'T' ? 8 : 15;
            p[i >> 1] |= (i & 1) ? code : (uint8_t)(code << 4);
        }
        p += (L + 1) / 2;
        memset(p, 0xff, L); p += L;
        p[0] = 'M'; p[1] = 'M'; p[2] = 'Z'; memcpy(p + 3, mm.c_str(), mm.size() + 1); p += 3 + mm.size() + 1;
        p[0] = 'M'; p[1] = 'L'; p[2] = 'B'; p[3] = 'C';
        uint32_t mln = (uint32_t)ml.size();
        memcpy(p + 4, &mln, 4); memcpy(p + 8, ml.data(), ml.size()); p += 8 + ml.size();
        if (has_hp) { p[0] = 'H'; p[1] = 'P'; p[2] = 'C'; p[3] = hp; p += 4; }
        rec->l_data = (int32_t)(p - rec->data.data());
    }
};

}  // namespace mmh

using namespace mmh;

extern "C" {

typedef struct {
    uint64_t n_reads, bases, ml_entries, cigar_ops, mm_bytes, ref_span, seq_bytes;
} mmh_synth_stats_t;

// contig_names/lens: the header table; the reference sequence is generated for every contig.
// gids (optional): job-wide ids of the contigs; a contig's sequence and reads depend on (seed, gid) only.
void *mmh_synth_new2(int config_id, uint64_t seed, int n_contigs, const char *const *names, const uint32_t *lens, const uint32_t *gids, double coverage) {
    Synth *s = new Synth();
    s->cfg = config_of(config_id);
    if (coverage > 0) s->cfg.coverage = coverage;
    s->seed = seed;
    std::vector<std::pair<std::string, uint32_t>> table;
    for (int i = 0; i < n_contigs; ++i) table.push_back({names[i], lens[i]});
    s->build_reference(table, gids);
    return s;
}
void *mmh_synth_new(int config_id, uint64_t seed, int n_contigs, const char *const *names, const uint32_t *lens, double coverage) {
    return mmh_synth_new2(config_id, seed, n_contigs, names, lens, nullptr, coverage);
}
void mmh_synth_free(void *h) { delete (Synth *)h; }
uint64_t mmh_synth_n_reads(void *h) { return ((Synth *)h)->n_reads; }
const char *mmh_synth_ref(void *h, int tid, uint64_t *len) { Synth *s = (Synth *)h; *len = s->contigs[tid].seq.size(); return s->contigs[tid].seq.data(); }

int mmh_synth_write_fasta(void *h, const char *path) {
    Synth *s = (Synth *)h;
    FILE *fp = fopen(path, "w");
    if (!fp) return -1;
    for (auto &c : s->contigs) {
        fprintf(fp, ">%s\n", c.name.c_str());
        fwrite(c.seq.data(), 1, c.seq.size(), fp);
        fputc('\n', fp);
    }
    fclose(fp);
    return 0;
}

static void add_stats(const BamRecord &r, mmh_synth_stats_t *st) {
    if (!st) return;
    st->n_reads++; st->bases += (uint64_t)r.l_qseq; st->cigar_ops += r.n_cigar; st->seq_bytes += ((uint64_t)r.l_qseq + 1) / 2;
    const uint8_t *mm = r.aux_get("MM"), *ml = r.aux_get("ML");
    if (mm) st->mm_bytes += strlen((const char *)mm + 1);
    if (ml) { uint32_t n; memcpy(&n, ml + 2, 4); st->ml_entries += n; }
    const uint8_t *cg = r.cigar();
    for (uint32_t i = 0; i < r.n_cigar; ++i) {
        uint32_t w; memcpy(&w, cg + 4 * i, 4);
        uint32_t op = w & 15u;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) st->ref_span += w >> 4;
    }
}

// Pack reads [first, first+count) into the batch (appending); returns the number packed (stops when full).
int64_t mmh_synth_fill(void *h, mmc_batch_t *b, uint64_t first, uint64_t count, int n_threads, mmh_synth_stats_t *st) {
    Synth *s = (Synth *)h;
    if (first + count > s->n_reads) count = s->n_reads > first ? s->n_reads - first : 0;
    LoadOpts lo;
    const uint64_t chunk = 256;
    if (n_threads < 1) n_threads = 1;
    uint64_t done = 0;
    std::vector<BamRecord> recs(chunk * (uint64_t)n_threads);
    while (done < count) {
        uint64_t n = std::min<uint64_t>(recs.size(), count - done);
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t)
            th.emplace_back([&, t]() { for (uint64_t i = (uint64_t)t; i < n; i += (uint64_t)n_threads) s->make_read(first + done + i, &recs[i]); });
        for (auto &x : th) x.join();
        for (uint64_t i = 0; i < n; ++i) {
            PackResult pr = pack_record(recs[i], b, lo, nullptr);
            if (pr == kNoSpace) return (int64_t)(done + i);
            add_stats(recs[i], st);
        }
        done += n;
    }
    return (int64_t)done;
}

namespace {
// Minimal BGZF writer (SAM spec 4.1): 0xff00-byte blocks, raw deflate level 1, 'BC' extra field, EOF marker.
struct BgzfWriter {
    FILE *fp = nullptr;
    std::vector<uint8_t> buf, out;
    bool open(const char *path) { fp = fopen(path, "wb"); buf.reserve(0xff00); out.resize(0x10000 + 1024); return fp != nullptr; }
    void block(const uint8_t *p, size_t n) {
        z_stream zs; memset(&zs, 0, sizeof zs);
        deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        zs.next_in = const_cast<uint8_t *>(p); zs.avail_in = (uInt)n;
        zs.next_out = out.data() + 18; zs.avail_out = (uInt)(out.size() - 18 - 8);
        deflate(&zs, Z_FINISH);
        const size_t clen = zs.total_out;
        deflateEnd(&zs);
        const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
        memcpy(out.data(), hdr, 16);
        const uint16_t bsize = (uint16_t)(18 + clen + 8 - 1);
        memcpy(out.data() + 16, &bsize, 2);
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), p, (uInt)n), isize = (uint32_t)n;
        memcpy(out.data() + 18 + clen, &crc, 4); memcpy(out.data() + 18 + clen + 4, &isize, 4);
        fwrite(out.data(), 1, 18 + clen + 8, fp);
    }
    void write(const void *p, size_t n) {
        const uint8_t *q = (const uint8_t *)p;
        while (n) {
            const size_t take = std::min(n, (size_t)0xff00 - buf.size());
            buf.insert(buf.end(), q, q + take); q += take; n -= take;
            if (buf.size() == 0xff00) { block(buf.data(), buf.size()); buf.clear(); }
        }
    }
    void close() {
        if (!buf.empty()) { block(buf.data(), buf.size()); buf.clear(); }
        block(nullptr, 0);                                                          // EOF marker: an empty block
        fclose(fp); fp = nullptr;
    }
};
}  // namespace

// Write reads [first, first+count) as a BAM in BGZF blocks (deflate level 1 keeps it fast).
int mmh_synth_write_bam(void *h, const char *path, uint64_t first, uint64_t count, int n_threads, mmh_synth_stats_t *st) {
    Synth *s = (Synth *)h;
    if (first + count > s->n_reads) count = s->n_reads > first ? s->n_reads - first : 0;
    BgzfWriter gz;
    if (!gz.open(path)) return -1;
    auto w32 = [&](uint32_t v) { gz.write(&v, 4); };
    gz.write("BAM\1", 4);
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (auto &c : s->contigs) text += "@SQ\tSN:" + c.name + "\tLN:" + std::to_string(c.seq.size()) + "\n";
    w32((uint32_t)text.size()); gz.write(text.data(), text.size());
    w32((uint32_t)s->contigs.size());
    for (auto &c : s->contigs) { w32((uint32_t)c.name.size() + 1); gz.write(c.name.c_str(), c.name.size() + 1); w32((uint32_t)c.seq.size()); }
    if (n_threads < 1) n_threads = 1;
    const uint64_t chunk = 256;
    std::vector<BamRecord> recs(chunk * (uint64_t)n_threads);
    uint64_t done = 0;
    while (done < count) {
        uint64_t n = std::min<uint64_t>(recs.size(), count - done);
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t)
            th.emplace_back([&, t]() { for (uint64_t i = (uint64_t)t; i < n; i += (uint64_t)n_threads) s->make_read(first + done + i, &recs[i]); });
        for (auto &x : th) x.join();
        for (uint64_t i = 0; i < n; ++i) {
            if (recs[i].n_cigar > 65535u) {                                    // SAM spec 4.2.2: placeholder CIGAR + CG:B,I tag
                BamRecord &w = recs[i];
                uint32_t ref_len = 0;
                for (uint32_t k = 0; k < w.n_cigar; ++k) { uint32_t c; memcpy(&c, w.cigar() + 4 * k, 4); uint32_t op = c & 15u; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += c >> 4; }
                std::vector<uint8_t> d(w.data.begin(), w.data.begin() + w.l_qname);
                const uint32_t fake[2] = {((uint32_t)w.l_qseq << 4) | 4u, (ref_len << 4) | 3u};
                d.insert(d.end(), (const uint8_t *)fake, (const uint8_t *)fake + 8);
                d.insert(d.end(), w.seq(), w.end());
                const uint8_t hdr[4] = {'C', 'G', 'B', 'I'};
                d.insert(d.end(), hdr, hdr + 4);
                const uint32_t n = w.n_cigar;
                d.insert(d.end(), (const uint8_t *)&n, (const uint8_t *)&n + 4);
                d.insert(d.end(), w.cigar(), w.cigar() + 4 * (size_t)n);
                w.l_data = (int32_t)d.size(); d.resize(d.size() + 8, 0);
                w.data.swap(d); w.n_cigar = 2;
            }
            const BamRecord &r = recs[i];
            uint8_t fx[36];
            uint32_t block = 32 + (uint32_t)r.l_data;
            memcpy(fx, &block, 4);
            memcpy(fx + 4, &r.tid, 4); memcpy(fx + 8, &r.pos, 4);
            fx[12] = r.l_qname; fx[13] = 60; uint16_t bin = 4680; memcpy(fx + 14, &bin, 2);
            uint16_t nc = (uint16_t)r.n_cigar; memcpy(fx + 16, &nc, 2); memcpy(fx + 18, &r.flag, 2);
            memcpy(fx + 20, &r.l_qseq, 4);
            int32_t m1 = -1, z = 0; memcpy(fx + 24, &m1, 4); memcpy(fx + 28, &m1, 4); memcpy(fx + 32, &z, 4);
            gz.write(fx, 36);
            gz.write(r.data.data(), (size_t)r.l_data);
            add_stats(r, st);
        }
        done += n;
    }
    gz.close();
    return 0;
}

}  // extern "C"
