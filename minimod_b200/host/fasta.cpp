#include "fasta.h"
#include <ctype.h>
#include <string.h>
#include <unordered_map>
#include <zlib.h>

namespace mmh {

bool read_fasta(const std::string &path, std::vector<FastaRecord> *out, std::string *err) {
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
