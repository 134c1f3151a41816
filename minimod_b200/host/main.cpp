// main.cpp -- the `minimod` command line of the B200 build: same sub-commands, options,
// defaults, stderr statistics and stdout text as the reference drivers
// (src/main.c:62-98, src/freq_main.c:166-519, src/view_main.c:164-491), with the
// pthread_processor / pthread_post_processor pair replaced by libminimod_cuda's
// stream-pipelined batch slots.  This binary links libminimod_cuda.so directly: without it
// (or without a CUDA device) it cannot run -- there is no CPU path here.
#include <getopt.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>
#include <sys/time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bam.h"
#include "fasta.h"
#include "format.h"
#include "minimod_cuda.h"
#include "modopts.h"
#include "pack.h"

#ifndef MINIMOD_VERSION
#define MINIMOD_VERSION "v0.5.0-b200"
#endif

using namespace mmh;

static int g_log_level = 4;   // LOG_VERB, src/error.c:36

static double realtime() { struct timeval tp; gettimeofday(&tp, NULL); return tp.tv_sec + tp.tv_usec * 1e-6; }
static double cputime() {
    struct rusage r; getrusage(RUSAGE_SELF, &r);
    return r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec);
}
static long peakrss() { struct rusage r; getrusage(RUSAGE_SELF, &r); return r.ru_maxrss * 1024; }

static void log_msg(int level, const char *func, const char *kind, const char *colour, const char *fmt, ...) {
    if (g_log_level < level) return;
    fprintf(stderr, "[%s::%s]%s ", func, kind, colour);
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
    fprintf(stderr, "\033[0m\n");
}
#define INFO(...)    log_msg(3, __func__, "INFO", "\033[1;34m", __VA_ARGS__)
#define WARNING(...) log_msg(2, __func__, "WARNING", "\033[1;33m", __VA_ARGS__)
#define ERROR(...)   log_msg(1, __func__, "ERROR", "\033[1;31m", __VA_ARGS__)

static int64_t mm_parse_num(const char *str) {           // src/misc.c:71-83
    char *p;
    double x = strtod(str, &p);
    if (*p == 'G' || *p == 'g') x *= 1e9;
    else if (*p == 'M' || *p == 'm') x *= 1e6;
    else if (*p == 'K' || *p == 'k') x *= 1e3;
    return (int64_t)(x + .499);
}

struct Opt {
    int subtool = MMC_FREQ;
    int32_t batch_size = 512;
    int64_t batch_size_bases = 20 * 1000 * 1000;
    int32_t num_thread = 8;
    int32_t debug_break = -1;
    int bedmethyl = 0, insertions = 0, haplotypes = 0, allow_secondary = 0, alt_alleles = 0, skip_supplementary = 0;
    int progress_interval = 0;
    int device = 0;
    std::vector<int> devices;                     // --devices: one context per entry, batches dealt by contig owner (or region)
    int shard_regions = 0;                        // --shard-regions: split the longest contig by read start, halo counts reduced at the end
    int64_t sparse_cap = 0;                       // --sparse-cap
    const char *mod_codes = nullptr;
    std::string mod_threshes;
    const char *output_file = nullptr;
    FILE *out = stdout;
};

static void print_help(FILE *fp, const Opt &o, const char *tool) {
    fprintf(fp, "Usage: minimod %s ref.fa reads.bam\n", tool);
    fprintf(fp, "\nbasic options:\n");
    if (o.subtool == MMC_FREQ) fprintf(fp, "   -b                         output in bedMethyl format [%s]\n", o.bedmethyl ? "yes" : "not set");
    fprintf(fp, "   -c STR                     modification code(s) (eg. m, h or mh or as ChEBI) [%s]\n", o.mod_codes ? o.mod_codes : "m");
    if (o.subtool == MMC_FREQ) fprintf(fp, "   -m FLOAT                   min modification threshold(s). Comma separated values for each modification code given in -c [%s]\n", o.mod_threshes.c_str());
    fprintf(fp, "   -t INT                     number of processing threads [%d] (host side only; the per-read work runs on the GPU)\n", o.num_thread);
    fprintf(fp, "   -K INT                     batch size (max number of reads loaded at once) [%d]\n", o.batch_size);
    fprintf(fp, "   -B FLOAT[K/M/G]            max number of bases loaded at once [%.1fM]\n", o.batch_size_bases / (float)(1000 * 1000));
    fprintf(fp, "   -h                         help\n");
    fprintf(fp, "   -p INT                     print progress every INT seconds (0: per batch) [%d]\n", o.progress_interval);
    fprintf(fp, "   -o FILE                    output file [%s]\n", o.output_file == NULL ? "stdout" : o.output_file);
    fprintf(fp, "   --insertions               output modifications in insertions [%s]\n", o.insertions ? "yes" : "no");
    fprintf(fp, "   --haplotypes               output haplotypes [%s]\n", o.haplotypes ? "yes" : "no");
    fprintf(fp, "   --verbose INT              verbosity level [%d]\n", g_log_level);
    fprintf(fp, "   --version                  print version\n");
    fprintf(fp, "   --allow-secondary          allow secondary alignments [%s]\n", o.allow_secondary ? "yes" : "no");
    fprintf(fp, "   --skip-supplementary       skip supplementary alignments [%s]\n", o.skip_supplementary ? "yes" : "no");
    fprintf(fp, "\nadvanced options:\n");
    fprintf(fp, "   --debug-break INT          break after processing the specified no. of batches\n");
    fprintf(fp, "   --device INT               CUDA device ordinal [%d]\n", o.device);
    fprintf(fp, "   --devices LIST             comma separated CUDA devices (e.g. 0,1,2,3): contigs are dealt to them by length (LPT)\n");
    fprintf(fp, "   --shard-regions            with --devices: split the longest contig by read start; boundary counts are summed with NCCL\n");
    fprintf(fp, "   --sparse-cap INT[K/M/G]    initial records of the side buffer for counts without a dense cell (it grows) [auto]\n");
}

static int run_tool(int subtool, int argc, char *argv[]) {
    const double realtime0 = realtime();
    const bool trace = getenv("MINIMOD_TRACE") != nullptr;       // wall-clock stamps of the tool's phases (stderr)
    auto stamp = [&](const char *what) { if (trace) fprintf(stderr, "[trace] %8.3f s  %s\n", realtime() - realtime0, what); };
    const char *tool = subtool == MMC_FREQ ? "freq" : "view";
    const char *func = subtool == MMC_FREQ ? "freq_main" : "view_main";
    static struct option long_options[] = {
        {"bedmethyl", no_argument, 0, 'b'},        {"mod_codes", required_argument, 0, 'c'},
        {"mod_thresh", required_argument, 0, 'm'}, {"threads", required_argument, 0, 't'},
        {"batchsize", required_argument, 0, 'K'},  {"max-bytes", required_argument, 0, 'B'},
        {"verbose", required_argument, 0, 'v'},    {"help", no_argument, 0, 'h'},
        {"version", no_argument, 0, 'V'},          {"prog-interval", required_argument, 0, 'p'},
        {"debug-break", required_argument, 0, 1000}, {"output", required_argument, 0, 'o'},
        {"insertions", no_argument, 0, 1001},      {"haplotypes", no_argument, 0, 1002},
        {"allow-secondary", no_argument, 0, 1003}, {"include-non-ref", no_argument, 0, 1004},
        {"skip-supplementary", no_argument, 0, 1005}, {"device", required_argument, 0, 1006},
        {"devices", required_argument, 0, 1007},   {"shard-regions", no_argument, 0, 1008},
        {"sparse-cap", required_argument, 0, 1009},
        {0, 0, 0, 0}};
    const char *optstring = subtool == MMC_FREQ ? "m:c:t:B:K:v:p:o:hVb" : "c:t:B:K:v:p:o:hV";

    Opt opt;
    opt.subtool = subtool;
    FILE *fp_help = stderr;
    int longindex = 0, c;
    while ((c = getopt_long(argc, argv, optstring, long_options, &longindex)) >= 0) {
        if (c == 'B') {
            opt.batch_size_bases = mm_parse_num(optarg);
            if (opt.batch_size_bases <= 0) { ERROR("%s", "Maximum number of bases should be larger than 0."); exit(EXIT_FAILURE); }
        } else if (c == 'K') {
            opt.batch_size = atoi(optarg);
            if (opt.batch_size < 1) { ERROR("Batch size should larger than 0. You entered %d", opt.batch_size); exit(EXIT_FAILURE); }
        } else if (c == 't') {
            opt.num_thread = atoi(optarg);
            if (opt.num_thread < 1) { ERROR("Number of threads should larger than 0. You entered %d", opt.num_thread); exit(EXIT_FAILURE); }
        } else if (c == 'v') g_log_level = atoi(optarg);
        else if (c == 'p') {
            if (atoi(optarg) < 0) { ERROR("Progress interval should be 0 or positive. You entered %d", atoi(optarg)); exit(EXIT_FAILURE); }
            opt.progress_interval = atoi(optarg);
        } else if (c == 'o') {
            FILE *fp = fopen(optarg, "w");
            if (!fp) { ERROR("Cannot open file %s for writing", optarg); exit(EXIT_FAILURE); }
            opt.output_file = optarg; opt.out = fp;
        } else if (c == 'V') { fprintf(stdout, "minimod %s\n", MINIMOD_VERSION); exit(EXIT_SUCCESS); }
        else if (c == 'h') fp_help = stdout;
        else if (c == 'm' && subtool == MMC_FREQ) opt.mod_threshes = optarg;
        else if (c == 'c') opt.mod_codes = optarg;
        else if (c == 'b' && subtool == MMC_FREQ) opt.bedmethyl = 1;
        else if (c == 1000) opt.debug_break = atoi(optarg);
        else if (c == 1001) opt.insertions = 1;
        else if (c == 1002) opt.haplotypes = 1;
        else if (c == 1003) opt.allow_secondary = 1;
        else if (c == 1004) opt.alt_alleles = 1;
        else if (c == 1005) opt.skip_supplementary = 1;
        else if (c == 1006) opt.device = atoi(optarg);
        else if (c == 1007) {
            for (const char *q = optarg; *q;) { opt.devices.push_back(atoi(q)); while (*q && *q != ',') ++q; if (*q == ',') ++q; }
        } else if (c == 1008) opt.shard_regions = 1;
        else if (c == 1009) opt.sparse_cap = mm_parse_num(optarg);
        else { print_help(fp_help, opt, tool); exit(fp_help == stdout ? EXIT_SUCCESS : EXIT_FAILURE); }
    }

    if (!opt.mod_codes || !*opt.mod_codes) {
        INFO("%s", "Modification codes not provided. Using default modification code m");
        opt.mod_codes = "m";
    }
    std::vector<ModSpec> mods;
    std::string err;
    if (!parse_mod_codes(opt.mod_codes, &mods, &err)) { ERROR("%s", err.c_str()); exit(EXIT_FAILURE); }
    for (const ModSpec &m : mods) {
        if (!m.context_given) INFO("Context not provided for modification code %s in -c argument. Using %s", m.code.c_str(), m.context.c_str());
        if (!is_tested_case(m.code, m.context)) WARNING("Modification code with context %s[%s] has not been tested.", m.code.c_str(), m.context.c_str());
    }
    if (subtool == MMC_FREQ) {
        if (opt.mod_threshes.empty()) {
            INFO("%s", "Modification threshold not provided. Using default threshold 0.8");
            for (size_t i = 0; i < mods.size(); ++i) opt.mod_threshes += i ? ",0.8" : "0.8";
        }
        if (!parse_mod_threshes(opt.mod_threshes, &mods, &err)) { ERROR("%s", err.c_str()); exit(EXIT_FAILURE); }
        for (const ModSpec &m : mods) INFO("Modification code: %s, Context: %s, Threshold: %f", m.code.c_str(), m.context.c_str(), m.thresh);
    } else {
        for (const ModSpec &m : mods) INFO("Modification code: %s, Context: %s", m.code.c_str(), m.context.c_str());
    }

    if (argc - optind != 2 || fp_help == stdout) {
        WARNING("%s", "Missing arguments");
        print_help(fp_help, opt, tool);
        exit(fp_help == stdout ? EXIT_SUCCESS : EXIT_FAILURE);
    }
    const char *ref_file = argv[optind], *bam_file = argv[optind + 1];
    if (access(bam_file, F_OK) == -1) { ERROR("BAM file %s does not exist", bam_file); exit(EXIT_FAILURE); }

    std::vector<mmc_mod_t> mmods;
    if (!to_mmc_mods(mods, &mmods, &err)) { ERROR("%s", err.c_str()); exit(EXIT_FAILURE); }

    // ---- the FASTA is parsed in the background while the BAM header is read and the CUDA contexts come up
    std::vector<FastaRecord> fa;
    std::string fa_err;
    bool fa_ok = false;
    double fa_secs = 0;
    std::thread fa_thread([&]() { const double t = realtime(); fa_ok = read_fasta(ref_file, &fa, &fa_err, opt.num_thread); fa_secs = realtime() - t; });

    // ---- BAM header first: the device context is sized from the contig table
    BamFile bam;
    if (!bam.open(bam_file, &err, opt.num_thread)) { ERROR("%s", err.c_str()); exit(EXIT_FAILURE); }
    std::vector<const char *> names;
    for (const std::string &n : bam.names) names.push_back(n.c_str());

    if (opt.devices.empty()) opt.devices.push_back(opt.device);
    const int ndev = (int)opt.devices.size();
    // ---- who owns what: contigs are dealt to the devices by longest-processing-time bin packing; with
    // --shard-regions the longest contig is instead cut into ndev slices by read start and every device keeps a full copy of it
    const int n_contigs = (int)bam.names.size();
    std::vector<int> owner(n_contigs, 0);
    int big_tid = -1;
    if (ndev > 1) {
        std::vector<int> order(n_contigs);
        for (int i = 0; i < n_contigs; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int a, int b) { return bam.lens[a] != bam.lens[b] ? bam.lens[a] > bam.lens[b] : a < b; });
        if (opt.shard_regions && n_contigs > 0) big_tid = order[0];
        std::vector<uint64_t> load(ndev, 0);
        for (int tid : order) {
            if (tid == big_tid) continue;
            int d = 0;
            for (int k = 1; k < ndev; ++k) if (load[k] < load[d]) d = k;
            owner[tid] = d; load[d] += bam.lens[tid];
        }
    }
    auto owner_of = [&](int32_t tid, int32_t pos) -> int {
        if (tid < 0 || tid >= n_contigs) return -1;
        if (tid == big_tid) { int64_t d = (int64_t)(pos < 0 ? 0 : pos) * ndev / std::max<uint32_t>(1u, bam.lens[tid]); return (int)std::min<int64_t>(d, ndev - 1); }
        return owner[tid];
    };

    std::vector<mmc_ctx *> ctxs(ndev, nullptr);
    for (int d = 0; d < ndev; ++d) {
        mmc_opts_t mo;
        memset(&mo, 0, sizeof(mo));
        mo.struct_size = sizeof(mo);
        mo.subtool = subtool; mo.n_mods = (int32_t)mmods.size(); mo.mods = mmods.data();
        mo.insertions = opt.insertions; mo.haplotypes = opt.haplotypes; mo.device = opt.devices[d];
        mo.n_slots = 3; mo.max_reads = (uint64_t)opt.batch_size; mo.max_bytes = (uint64_t)opt.batch_size_bases;
        mo.sparse_capacity = (uint64_t)opt.sparse_cap;
        mo.seq_packing = 2;                          // SEQ crosses PCIe at 2 bits per base + exceptions
        mo.cigar_packing = 8;                        // CIGARs at a byte per op + escapes
        if (mmc_create(&ctxs[d], &mo, (int32_t)names.size(), names.data(), bam.lens.data()) != MMC_OK) {
            ERROR("%s", mmc_strerror(nullptr)); exit(EXIT_FAILURE);
        }
    }
    mmc_ctx *ctx = ctxs[0];
    stamp("BAM header read, device contexts created");

    // ---- reference: FASTA parsing on the host, packing + context evaluation on the device that owns the contig
    double realtime1 = realtime();
    fprintf(stderr, "[%s] Loading reference genome %s\n", func, ref_file);
    {
        fa_thread.join();
        if (!fa_ok) { ERROR("Could not to open file %s: %s", ref_file, fa_err.c_str()); exit(EXIT_FAILURE); }
        fprintf(stderr, "[%s] Reference genome loaded in %.3f sec (%.3f sec of parsing, overlapped with device start-up)\n", func, realtime() - realtime1, fa_secs);
        double realtime2 = realtime();
        fprintf(stderr, "[%s] Loading contexts in reference\n", func);
        for (const FastaRecord &r : fa) {
            int32_t tid = -1;
            for (size_t i = 0; i < bam.names.size(); ++i) if (bam.names[i] == r.name) { tid = (int32_t)i; break; }
            if (tid < 0) continue;                       // contigs the BAM header does not know can never be hit
            for (int d = 0; d < ndev; ++d) {
                if (tid != big_tid && owner[tid] != d) continue;
                const int rc = mmc_ref_add(ctxs[d], tid, r.seq.data(), (uint32_t)r.seq.size());
                if (rc == MMC_EINVAL && r.seq.size() != bam.lens[tid]) {
                    // a length mismatch is only fatal in the reference when a read maps there (src/mod.c:861)
                    WARNING("%s", mmc_strerror(ctxs[d]));
                } else if (rc != MMC_OK) {               // out of device memory, CUDA failure: nothing sensible can follow
                    ERROR("%s", mmc_strerror(ctxs[d])); exit(EXIT_FAILURE);
                }
            }
        }
        for (int d = 0; d < ndev; ++d)
            if (mmc_ref_commit(ctxs[d]) != MMC_OK) { ERROR("%s", mmc_strerror(ctxs[d])); exit(EXIT_FAILURE); }
        fprintf(stderr, "[%s] Reference contexts loaded in %.3f sec\n", func, realtime() - realtime2);
    }

    stamp("reference packed on the device");
    OutOpts oo; oo.bedmethyl = opt.bedmethyl; oo.insertions = opt.insertions; oo.haplotypes = opt.haplotypes;
    if (subtool == MMC_FREQ) print_freq_header(opt.out, oo); else print_view_header(opt.out, oo);

    LoadOpts lo;
    lo.batch_size = opt.batch_size; lo.batch_size_bases = opt.batch_size_bases;
    lo.allow_secondary = opt.allow_secondary; lo.skip_supplementary = opt.skip_supplementary;
    lo.keep_qnames = 1;                               // read names: view rows, and the reference's per-read fatal messages
    BatchLoader loader(&bam, lo);
    if (ndev > 1) loader.set_owner([&](const BamRecord &r) { return (r.flag & 4) ? -1 : owner_of(r.tid, r.pos); });

    uint64_t total_reads = 0, total_bytes = 0, processed_reads = 0, processed_bytes = 0;
    double load_time = 0, output_time = 0;
    int32_t counter = 0;
    const int n_slots = 3;
    struct Ring { std::vector<mmc_batch_t *> b; std::vector<BatchMeta> meta; int next = 0; };
    std::vector<Ring> rings(ndev);
    for (Ring &r : rings) { r.b.assign(n_slots, nullptr); r.meta.resize(n_slots); }
    auto code_names_of = [&](mmc_ctx *c) {
        std::vector<std::string> v(256);
        for (int i = 0; i < 256; ++i) v[i] = mmc_code_name(c, i);
        return v;
    };
    // a per-read fatal condition: the library names the read by its index in the batch ("\x1f<index>" suffix); the
    // reference prints the read's name (src/mod.c:843,1174)
    auto die_read = [&](mmc_ctx *c, const BatchMeta &meta) {
        std::string msg = mmc_strerror(c);
        const size_t sep = msg.find('\x1f');
        if (sep != std::string::npos) {
            const unsigned long idx = strtoul(msg.c_str() + sep + 1, nullptr, 10);
            msg.resize(sep);
            const size_t colon = msg.find(": ");
            if (colon != std::string::npos && idx < meta.qname_off.size()) {
                std::string text = msg.substr(colon + 2);
                const std::string hc = "Hard clipping found and";
                if (text.compare(0, hc.size(), hc) == 0) text = std::string("Hard clipping found in ") + meta.qname((uint32_t)idx) + " and" + text.substr(hc.size());
                else text = std::string("read_id:") + meta.qname((uint32_t)idx) + " " + text;
                msg = text;
            }
        }
        ERROR("%s", msg.c_str());
        exit(EXIT_FAILURE);
    };
    // ---- freq: rows of finished positions leave the device while the BAM is still being read (mmc_freq_drain: the reads of a
    // coordinate-sorted BAM never come back to positions before a batch's first read) and are turned into text by a worker
    // thread, per contig; at the end the contigs are written in strcmp order (cmp_key_fast, src/mod.c:59-87).  Anything else
    // (--shard-regions, a BAM that turns out not to be sorted) is read back and formatted after the last batch as before.
    const bool stream_rows = subtool == MMC_FREQ && big_tid < 0 && !getenv("MINIMOD_NO_DRAIN");
    struct FmtJob { std::vector<mmc_freq_rec_t> rows; std::vector<std::string> codes; };
    std::vector<std::vector<std::string>> contig_text(stream_rows ? n_contigs : 0);   // per contig: pieces of text in row order
    std::deque<FmtJob> fmt_queue;
    std::mutex fmt_mu;
    std::condition_variable fmt_cv;
    bool fmt_done = false;
    uint64_t fmt_pushed = 0;                              // jobs handed to the worker (main thread) / jobs it has finished
    std::atomic<uint64_t> fmt_finished(0);
    uint64_t rows_early = 0, rows_total = 0;
    double fmt_secs = 0;
    std::thread fmt_thread;
    if (stream_rows) fmt_thread = std::thread([&]() {
        for (;;) {
            FmtJob job;
            {
                std::unique_lock<std::mutex> lk(fmt_mu);
                fmt_cv.wait(lk, [&]() { return fmt_done || !fmt_queue.empty(); });
                if (fmt_queue.empty()) return;
                job = std::move(fmt_queue.front());
                fmt_queue.pop_front();
            }
            const double t0 = realtime();
            const uint64_t n = job.rows.size();
            for (uint64_t i = 0; i < n;) {
                uint64_t j = i;
                while (j < n && job.rows[j].tid == job.rows[i].tid) ++j;
                contig_text[job.rows[i].tid].emplace_back();
                format_freq_rows(&contig_text[job.rows[i].tid].back(), oo, bam.names[job.rows[i].tid], job.rows.data(), i, j, job.codes);
                i = j;
            }
            fmt_secs += realtime() - t0;
            fmt_finished.fetch_add(1, std::memory_order_release);
        }
    });
    auto fmt_push = [&](mmc_ctx *c, const mmc_freq_rec_t *r, uint64_t n) {
        if (!n) return;
        FmtJob job;
        job.rows.assign(r, r + n);                        // out of the library's alternating pinned buffers
        job.codes = code_names_of(c);
        { std::lock_guard<std::mutex> lk(fmt_mu); fmt_queue.push_back(std::move(job)); }
        ++fmt_pushed;
        fmt_cv.notify_one();
    };
    auto fmt_idle = [&]() {                               // every job handed over so far is formatted: contig_text is the caller's
        while (fmt_finished.load(std::memory_order_acquire) != fmt_pushed) std::this_thread::yield();
    };

    double t_release = 0, t_acquire = 0, t_submit = 0, t_drain = 0, t_view = 0;   // MINIMOD_TRACE: where the main loop's wall clock goes
    int more = 1;
    while (more) {
        int d = 0;
        if (ndev > 1) {                                   // the device that owns the next record
            d = loader.next_owner(&err);
            if (d == -2) { ERROR("%s", err.c_str()); exit(EXIT_FAILURE); }
            if (d < 0) d = 0;
        }
        mmc_ctx *c = ctxs[d];
        Ring &rg = rings[d];
        const int si = rg.next; rg.next = (rg.next + 1) % n_slots;
        double tr0 = realtime();
        if (rg.b[si]) {                                   // recycle the oldest slot (joins the "previous processor")
            if (mmc_batch_release(c, rg.b[si]) != MMC_OK) die_read(c, rg.meta[si]);
            rg.b[si] = nullptr;
        }
        t_release += realtime() - tr0; tr0 = realtime();
        mmc_batch_t *b = nullptr;
        if (mmc_batch_acquire(c, &b) != MMC_OK) { ERROR("%s", mmc_strerror(c)); exit(EXIT_FAILURE); }
        t_acquire += realtime() - tr0;
        double t0 = realtime();
        more = loader.fill(b, &rg.meta[si], &err);
        load_time += realtime() - t0;
        if (more < 0) { ERROR("%s", err.c_str()); exit(EXIT_FAILURE); }
        const BatchStats &st = rg.meta[si].stats;
        fprintf(stderr, "[%s::%.3f*%.2f] %d Entries (%.1fM bases) loaded\n", func, realtime() - realtime0,
                cputime() / (realtime() - realtime0), st.n_recs, st.processed_bytes / (1000.0 * 1000.0));
        tr0 = realtime();
        if (mmc_batch_submit(c, b) != MMC_OK) { ERROR("%s", mmc_strerror(c)); exit(EXIT_FAILURE); }
        t_submit += realtime() - tr0; tr0 = realtime();
        rg.b[si] = b;
        if (stream_rows && b->n_reads && b->tid[0] >= 0) {
            // the batches submitted before this one hold every read that starts before its first read
            for (int i = 1; i < n_slots; ++i) {
                const int oi = (si + i) % n_slots;
                if (rg.b[oi] && mmc_batch_wait(c, rg.b[oi]) != MMC_OK) die_read(c, rg.meta[oi]);
            }
            const mmc_freq_rec_t *recs = nullptr; uint64_t n = 0;
            if (mmc_freq_drain(c, b->tid[0], (uint32_t)std::max<int32_t>(0, b->pos[0]), &recs, &n) != MMC_OK) { ERROR("%s", mmc_strerror(c)); exit(EXIT_FAILURE); }
            rows_early += n; rows_total += n;
            fmt_push(c, recs, n);
        }
        t_drain += realtime() - tr0;
        if (subtool == MMC_VIEW) {
            const mmc_view_rec_t *recs = nullptr; uint64_t n = 0;
            if (mmc_view_fetch(c, b, &recs, &n) != MMC_OK) die_read(c, rg.meta[si]);
            double o0 = realtime();
            print_view_records(opt.out, oo, bam.names, b, rg.meta[si], recs, n, code_names_of(c));
            output_time += realtime() - o0;
        }
        fprintf(stderr, "[%s::%.3f*%.2f] %d Entries (%.1fM bytes) processed\t%d Entries (%.1fM bytes) skipped\n", func,
                realtime() - realtime0, cputime() / (realtime() - realtime0), st.n_recs, st.total_bytes / (1000.0 * 1000.0),
                st.total_reads - st.n_recs, (st.total_bytes - st.processed_bytes) / (1000.0 * 1000.0));
        total_reads += st.total_reads; total_bytes += st.total_bytes;
        processed_reads += st.n_recs; processed_bytes += st.processed_bytes;
        uint64_t skipped = total_reads - processed_reads;
        if (skipped > 0.9 * total_reads)
            WARNING("%s", "90% of the reads are skipped. Possible causes: unmapped bam, zero sequence lengths, or missing MM, ML tags (not performed base modification aware basecalling). Refer https://github.com/warp9seq/minimod for more information.");
        if (skipped == total_reads)
            ERROR("%s", "All reads are skipped. Quitting. Possible causes: unmapped bam, zero sequence lengths, or missing MM, ML tags (not performed base modification aware basecalling). Refer https://github.com/warp9seq/minimod for more information.");
        if (opt.debug_break == counter) break;
        counter++;
    }
    for (int d = 0; d < ndev; ++d)
        for (int i = 0; i < n_slots; ++i)
            if (rings[d].b[i] && mmc_batch_release(ctxs[d], rings[d].b[i]) != MMC_OK) die_read(ctxs[d], rings[d].meta[i]);

    stamp("last batch submitted and released");
    if (trace) fprintf(stderr, "[trace] main loop: fill %.3f s, slot release (waits for the slot's batch) %.3f s, acquire %.3f s, submit %.3f s, drain %.3f s\n",
                       load_time, t_release, t_acquire, t_submit, t_drain);
    (void)t_view;
    double sort_time = 0, halo_ms = 0;
    uint64_t halo_bytes = 0;
    if (stream_rows) {
        double s0 = realtime();
        for (int d = 0; d < ndev; ++d) {
            const mmc_freq_rec_t *recs = nullptr; uint64_t n = 0;
            int rc = mmc_freq_finalize(ctxs[d], &recs, &n);
            if (rc == MMC_EORDER) {
                // the BAM was not coordinate-sorted after all: nothing is lost (drains never clear counts), the rows this
                // device handed out early are dropped and its whole table is read back
                WARNING("%s", "reads are not in coordinate order: the rows written out early are discarded and the table is rebuilt at the end");
                fmt_idle();
                { std::lock_guard<std::mutex> lk(fmt_mu); }
                for (int t = 0; t < n_contigs; ++t) if (owner[t] == d) std::vector<std::string>().swap(contig_text[t]);
                rows_total = 0; rows_early = 0;           // (reported figures only)
                if (mmc_freq_undrain(ctxs[d]) != MMC_OK) { ERROR("%s", mmc_strerror(ctxs[d])); exit(EXIT_FAILURE); }
                rc = mmc_freq_finalize(ctxs[d], &recs, &n);
            }
            if (rc != MMC_OK) { ERROR("%s", mmc_strerror(ctxs[d])); exit(EXIT_FAILURE); }
            rows_total += n;
            sort_time += realtime() - s0;
            // what is left (everything, when nothing could leave early: --insertions, unsorted input) is formatted by -t threads:
            // chunks of rows that stay inside a contig, claimed by an atomic counter, their text kept in row order
            const double o1 = realtime();
            fmt_idle();
            { std::lock_guard<std::mutex> lk(fmt_mu); }                     // (the worker is between jobs: contig_text is ours)
            struct Chunk { int32_t tid; uint64_t b, e; std::string text; };
            std::vector<Chunk> chunks;
            const uint64_t per = 1u << 18;
            for (uint64_t i = 0; i < n;) {
                uint64_t j = i;
                while (j < n && j - i < per && recs[j].tid == recs[i].tid) ++j;
                chunks.push_back({recs[i].tid, i, j, std::string()});
                i = j;
            }
            const std::vector<std::string> cn = code_names_of(ctxs[d]);
            std::atomic<size_t> next_chunk(0);
            auto work = [&]() {
                for (size_t k; (k = next_chunk.fetch_add(1)) < chunks.size();)
                    format_freq_rows(&chunks[k].text, oo, bam.names[chunks[k].tid], recs, chunks[k].b, chunks[k].e, cn);
            };
            std::vector<std::thread> pool;
            const int nt = (int)std::min<size_t>((size_t)std::max(1, opt.num_thread), chunks.size());
            for (int t = 1; t < nt; ++t) pool.emplace_back(work);
            work();
            for (std::thread &t : pool) t.join();
            for (Chunk &ck : chunks) contig_text[ck.tid].push_back(std::move(ck.text));
            output_time += realtime() - o1;
            s0 = realtime();
        }
        double o0 = realtime();
        { std::lock_guard<std::mutex> lk(fmt_mu); fmt_done = true; }
        fmt_cv.notify_all();
        fmt_thread.join();
        std::vector<int> order(n_contigs);
        for (int i = 0; i < n_contigs; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return strcmp(bam.names[a].c_str(), bam.names[b].c_str()) < 0; });
        for (int t : order) for (const std::string &piece : contig_text[t]) fwrite(piece.data(), 1, piece.size(), opt.out);
        output_time += realtime() - o0;
    } else if (subtool == MMC_FREQ) {
        double s0 = realtime();
        if (big_tid >= 0 && mmc_region_reduce(ctxs.data(), ndev, big_tid, &halo_ms, &halo_bytes) != MMC_OK) {
            ERROR("%s", mmc_strerror(ctxs[0])); exit(EXIT_FAILURE);
        }
        // every device's table; code ids (dictionary order with a '*' code) are translated to one list
        std::vector<mmc_freq_rec_t> all;
        std::vector<std::string> codes;
        const mmc_freq_rec_t *recs = nullptr; uint64_t n = 0;
        for (int d = 0; d < ndev; ++d) {
            if (mmc_freq_finalize(ctxs[d], &recs, &n) != MMC_OK) { ERROR("%s", mmc_strerror(ctxs[d])); exit(EXIT_FAILURE); }
            if (ndev == 1) break;
            const std::vector<std::string> cn = code_names_of(ctxs[d]);
            std::vector<int> remap(256, -1);
            const size_t at = all.size();
            all.insert(all.end(), recs, recs + n);
            for (size_t i = at; i < all.size(); ++i) {
                int &m = remap[all[i].code];
                if (m < 0) {
                    size_t k = 0;
                    while (k < codes.size() && codes[k] != cn[all[i].code]) ++k;
                    if (k == codes.size()) codes.push_back(cn[all[i].code]);
                    if (k > 255) { ERROR("%s", "more than 256 distinct modification codes across devices"); exit(EXIT_FAILURE); }
                    m = (int)k;
                }
                all[i].code = (uint8_t)m;
            }
        }
        if (ndev > 1) {
            auto key_less = [](const mmc_freq_rec_t &x, const mmc_freq_rec_t &y) {
                if (x.tid != y.tid) return x.tid < y.tid;
                if (x.pos != y.pos) return x.pos < y.pos;
                if (x.strand != y.strand) return x.strand < y.strand;
                if (x.code != y.code) return x.code < y.code;
                if (x.ins_offset != y.ins_offset) return x.ins_offset < y.ins_offset;
                return x.hap < y.hap;
            };
            if (big_tid >= 0) {
                // region sharding: rows of the cut contig come from several devices and sparse rows (insertions, exotic
                // haplotypes) of a boundary may exist on both sides: order everything and add rows with equal keys
                std::stable_sort(all.begin(), all.end(), key_less);
                size_t w = 0;
                for (size_t i = 0; i < all.size(); ++i) {
                    if (w && !key_less(all[w - 1], all[i]) && !key_less(all[i], all[w - 1])) { all[w - 1].n_called += all[i].n_called; all[w - 1].n_mod += all[i].n_mod; }
                    else all[w++] = all[i];
                }
                all.resize(w);
            }
            codes.resize(256);
            recs = all.data(); n = all.size();
        }
        sort_time = realtime() - s0;
        double o0 = realtime();
        print_freq_records(opt.out, oo, bam.names, recs, n, ndev > 1 ? codes : code_names_of(ctx));
        output_time += realtime() - o0;
    }
    if (opt.out != stdout) fclose(opt.out); else fflush(stdout);
    stamp("table finalized, formatted and written");

    mmc_timers_t tm;
    memset(&tm, 0, sizeof(tm));
    for (int d = 0; d < ndev; ++d) {
        mmc_timers_t t1;
        mmc_get_timers(ctxs[d], &t1);
        tm.decode_ms = std::max(tm.decode_ms, t1.decode_ms); tm.finalize_ms = std::max(tm.finalize_ms, t1.finalize_ms);
        tm.h2d_ms = std::max(tm.h2d_ms, t1.h2d_ms); tm.h2d_bytes += t1.h2d_bytes; tm.kernel_launches += t1.kernel_launches;
    }
    fprintf(stderr, "[%s] total entries: %ld", func, (long)total_reads);
    fprintf(stderr, "\n[%s] total bytes: %.1f M", func, total_bytes / (float)(1000 * 1000));
    fprintf(stderr, "\n[%s] total skipped entries: %ld", func, (long)(total_reads - processed_reads));
    fprintf(stderr, "\n[%s] total skipped bytes: %.1f M", func, (total_bytes - processed_bytes) / (float)(1000 * 1000));
    fprintf(stderr, "\n[%s] total processed entries: %ld", func, (long)processed_reads);
    fprintf(stderr, "\n[%s] total processed bytes: %.1f M", func, processed_bytes / (float)(1000 * 1000));
    fprintf(stderr, "\n[%s] Data loading time: %.3f sec", func, load_time);
    fprintf(stderr, "\n[%s] Data processing time: %.3f sec", func, tm.decode_ms / 1e3);
    if (subtool == MMC_FREQ) {
        fprintf(stderr, "\n[%s] Data merging time: %.3f sec", func, tm.finalize_ms / 1e3);
        fprintf(stderr, "\n[%s] Data sorting time: %.3f sec", func, sort_time);
    }
    fprintf(stderr, "\n[%s] Data output time: %.3f sec", func, output_time);
    fprintf(stderr, "\n[%s] Device: H2D %.3f sec (%.1f MB), %lu kernel launches", func, tm.h2d_ms / 1e3, tm.h2d_bytes / 1e6,
            (unsigned long)tm.kernel_launches);
    if (stream_rows) fprintf(stderr, "\n[%s] Rows: %lu, of which %lu left the device and were formatted while the BAM was being read (%.3f sec of formatting on a worker thread)",
                             func, (unsigned long)rows_total, (unsigned long)rows_early, fmt_secs);
    if (ndev > 1) fprintf(stderr, "\n[%s] Devices: %d (%s); boundary reduce %.3f ms, %.1f MB", func, ndev,
                          big_tid >= 0 ? "longest contig cut by read start, the others dealt by length" : "contigs dealt by length", halo_ms, halo_bytes / 1e6);
    fprintf(stderr, "\n");
    // The process is about to end: an orderly tear-down (cudaFree / cudaFreeHost of every buffer, each a device-wide
    // synchronisation) costs 0.1-0.5 s that the driver spends again when the process exits.  MINIMOD_CLEAN_EXIT=1 keeps it
    // (leak checkers); MMC_TRACE_CREATE wants the library's closing statistics.
    if (getenv("MINIMOD_CLEAN_EXIT") || getenv("MMC_TRACE_CREATE")) {
        for (mmc_ctx *c : ctxs) mmc_destroy(c);
        stamp("device contexts destroyed");
        return 0;
    }
    return -1000;                                     // main(): print the closing lines, then _exit
}

static int print_usage(FILE *fp) {
    fprintf(fp, "Usage: minimod <command> [options]\n\n");
    fprintf(fp, "command:\n");
    fprintf(fp, "         view       view base modifications\n");
    fprintf(fp, "         freq       output base modification frequencies\n");
    fprintf(fp, "         summary    output summary (not part of the B200 build; use the reference)\n");
    return fp == stdout ? EXIT_SUCCESS : EXIT_FAILURE;
}

int main(int argc, char *argv[]) {
    double realtime0 = realtime();
    int ret = 1;
    if (argc < 2) return print_usage(stderr);
    else if (strcmp(argv[1], "view") == 0) ret = run_tool(MMC_VIEW, argc - 1, argv + 1);
    else if (strcmp(argv[1], "mod-freq") == 0) { WARNING("%s", "mod-freq is deprecated. Use freq instead"); ret = run_tool(MMC_FREQ, argc - 1, argv + 1); }
    else if (strcmp(argv[1], "freq") == 0) ret = run_tool(MMC_FREQ, argc - 1, argv + 1);
    else if (strcmp(argv[1], "summary") == 0) {
        fprintf(stderr, "[minimod] 'summary' only parses tag headers and is not on the accelerated path; it is not part of this build\n");
        return EXIT_FAILURE;
    } else if (strcmp(argv[1], "--version") == 0 || strcmp(argv[1], "-V") == 0) { fprintf(stdout, "minimod %s\n", MINIMOD_VERSION); exit(EXIT_SUCCESS); }
    else if (strcmp(argv[1], "--help") == 0 || strcmp(argv[1], "-h") == 0) return print_usage(stdout);
    else { fprintf(stderr, "[minimod] Unrecognised command %s\n", argv[1]); return print_usage(stderr); }

    fprintf(stderr, "[%s] Version: %s\n", __func__, MINIMOD_VERSION);
    fprintf(stderr, "[%s] CMD:", __func__);
    for (int i = 0; i < argc; ++i) fprintf(stderr, " %s", argv[i]);
    const bool fast_exit = ret == -1000;
    if (fast_exit) ret = 0;
    fprintf(stderr, "\n[%s] Real time: %.3f sec; CPU time: %.3f sec; Peak RAM: %.3f GB\n\n", __func__, realtime() - realtime0, cputime(),
            peakrss() / 1024.0 / 1024.0 / 1024.0);
    if (fast_exit) { fflush(stdout); fflush(stderr); _exit(0); }
    return ret;
}
