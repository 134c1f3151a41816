#include "fasta.h"
#include <ctype.h>
#include <string.h>
#include <algorithm>
#include <unordered_map>
#include <zlib.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <thread>

namespace mmh {

// Plain (uncompressed) FASTA: the file is mapped, record starts are found with one memchr sweep and the records are
// line-joined by a pool of threads (load_ref() is single-threaded kseq, src/ref.c:46-89; a 3.1 Gbp genome spends seconds
// there).  Returns false when the file is not plain FASTA (gzip, FASTQ, unreadable): the caller then reads it serially.
static bool read_fasta_mapped(const std::string &path, std::vector<FastaRecord> *out, int threads) {
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 2 || !S_ISREG(st.st_mode)) { close(fd); return false; }
    const size_t n = (size_t)st.st_size;
    const char *d = (const char *)mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (d == MAP_FAILED) return false;
    if (d[0] != '>') { munmap((void *)d, n); return false; }                     // gzip magic, FASTQ '@', anything else
    std::vector<size_t> starts;
    for (const char *p = d; p;) {
        starts.push_back((size_t)(p - d));
        const char *q = p + 1;
        for (;;) {
            q = (const char *)memchr(q, '>', n - (size_t)(q - d));
            if (!q || q[-1] == '\n') break;
            ++q;
        }
        p = q;
    }
    const size_t nrec = starts.size();
    starts.push_back(n);
    std::vector<FastaRecord> recs(nrec);
    std::atomic<size_t> next(0);
    std::atomic<bool> bad(false);
    auto work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= nrec) break;
            const char *p = d + starts[i] + 1, *e = d + starts[i + 1];
            const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
            const char *hs = nl ? nl : e, *q = p;
            while (q < hs && !isspace((unsigned char)*q)) ++q;
            recs[i].name.assign(p, (size_t)(q - p));
            std::string &seq = recs[i].seq;
            seq.reserve((size_t)(e - hs));
            for (p = nl ? nl + 1 : e; p < e;) {
                if (*p == '+' || *p == '@') { bad = true; return; }               // FASTQ-like content: not handled here
                const char *l = (const char *)memchr(p, '\n', (size_t)(e - p));
                const char *stop = l ? l : e;
                size_t len = (size_t)(stop - p);
                if (len && stop[-1] == '\r') --len;
                seq.append(p, len);
                p = l ? l + 1 : e;
            }
        }
    };
    std::vector<std::thread> th;
    const int nt = std::max(1, std::min<int>(threads, (int)nrec));
    for (int t = 0; t < nt; ++t) th.emplace_back(work);
    for (auto &t : th) t.join();
    munmap((void *)d, n);
    if (bad) return false;
    std::unordered_map<std::string, size_t> index;                               // a later record replaces an earlier one of the same name
    for (size_t i = 0; i < nrec; ++i) {
        auto it = index.find(recs[i].name);
        if (it == index.end()) { index[recs[i].name] = out->size(); out->push_back(std::move(recs[i])); }
        else (*out)[it->second].seq = std::move(recs[i].seq);
    }
    return true;
}

bool read_fasta(const std::string &path, std::vector<FastaRecord> *out, std::string *err, int threads) {
    if (threads > 1 && read_fasta_mapped(path, out, threads)) return true;
    out->clear();
    gzFile fp = gzopen(path.c_str(), "r");
    if (!fp) { if (err) *err = "cannot open " + path; return false; }
    gzbuffer(fp, 4u << 20);
    std::vector<char> buf(8u << 20);
    std::unordered_map<std::string, size_t> index;
    size_t cur = (size_t)-1;
    bool in_header = false, line_start = true, name_done = false, fastq_skip = false;
    std::string header;
    for (;;) {
        int n = gzread(fp, buf.data(), (unsigned)buf.size());
        if (n < 0) { if (err) *err = "read error in " + path; gzclose(fp); return false; }
        if (n == 0) break;
        const char *p = buf.data(), *e = p + n;
        while (p < e) {
            if (in_header) {
                const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
                const char *stop = nl ? nl : e;
                if (!name_done) {
                    const char *q = p;
                    while (q < stop && !isspace((unsigned char)*q)) ++q;
                    header.append(p, (size_t)(q - p));
                    if (q < stop) name_done = true;
                }
                if (nl) {
                    in_header = false; line_start = true;
                    auto it = index.find(header);
                    if (it == index.end()) { index[header] = out->size(); out->push_back(FastaRecord()); cur = out->size() - 1; (*out)[cur].name = header; }
                    else { cur = it->second; (*out)[cur].seq.clear(); }
                    p = nl + 1;
                } else p = e;
                continue;
            }
            if (line_start && (*p == '>' || *p == '@')) { in_header = true; name_done = false; header.clear(); fastq_skip = false; ++p; line_start = false; continue; }
            if (line_start && *p == '+') fastq_skip = true;     // FASTQ quality section: ignore until next header
            const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
            const char *stop = nl ? nl : e;
            if (cur != (size_t)-1 && !fastq_skip && stop > p) (*out)[cur].seq.append(p, (size_t)(stop - p));
            if (nl) {
                if (cur != (size_t)-1 && !fastq_skip && !(*out)[cur].seq.empty() && (*out)[cur].seq.back() == '\r') (*out)[cur].seq.pop_back();
                line_start = true; p = nl + 1;
            } else { line_start = false; p = e; }
        }
    }
    gzclose(fp);
    return true;
}

}  // namespace mmh
