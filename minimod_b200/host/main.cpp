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

#include <string>
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
}

static int run_tool(int subtool, int argc, char *argv[]) {
    const double realtime0 = realtime();
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

    // ---- BAM header first: the device context is sized from the contig table
    BamFile bam;
    if (!bam.open(bam_file, &err, opt.num_thread)) { ERROR("%s", err.c_str()); exit(EXIT_FAILURE); }
    std::vector<const char *> names;
    for (const std::string &n : bam.names) names.push_back(n.c_str());

    mmc_opts_t mo;
    memset(&mo, 0, sizeof(mo));
    mo.struct_size = sizeof(mo);
    mo.subtool = subtool; mo.n_mods = (int32_t)mmods.size(); mo.mods = mmods.data();
    mo.insertions = opt.insertions; mo.haplotypes = opt.haplotypes; mo.device = opt.device;
    mo.n_slots = 3; mo.max_reads = (uint64_t)opt.batch_size; mo.max_bytes = (uint64_t)opt.batch_size_bases;
    mo.seq_packing = 2;                          // SEQ crosses PCIe at 2 bits per base + exceptions
    mmc_ctx *ctx = nullptr;
    if (mmc_create(&ctx, &mo, (int32_t)names.size(), names.data(), bam.lens.data()) != MMC_OK) {
        ERROR("%s", mmc_strerror(nullptr)); exit(EXIT_FAILURE);
    }

    // ---- reference: FASTA parsing on the host, packing + context evaluation on the device
    double realtime1 = realtime();
    fprintf(stderr, "[%s] Loading reference genome %s\n", func, ref_file);
    {
        std::vector<FastaRecord> fa;
        if (!read_fasta(ref_file, &fa, &err)) { ERROR("Could not to open file %s: %s", ref_file, err.c_str()); exit(EXIT_FAILURE); }
        fprintf(stderr, "[%s] Reference genome loaded in %.3f sec\n", func, realtime() - realtime1);
        double realtime2 = realtime();
        fprintf(stderr, "[%s] Loading contexts in reference\n", func);
        for (const FastaRecord &r : fa) {
            int32_t tid = -1;
            for (size_t i = 0; i < bam.names.size(); ++i) if (bam.names[i] == r.name) { tid = (int32_t)i; break; }
            if (tid < 0) continue;                       // contigs the BAM header does not know can never be hit
            if (mmc_ref_add(ctx, tid, r.seq.data(), (uint32_t)r.seq.size()) != MMC_OK) {
                // a length mismatch is only fatal in the reference when a read maps there (src/mod.c:861)
                WARNING("%s", mmc_strerror(ctx));
            }
        }
        if (mmc_ref_commit(ctx) != MMC_OK) { ERROR("%s", mmc_strerror(ctx)); exit(EXIT_FAILURE); }
        fprintf(stderr, "[%s] Reference contexts loaded in %.3f sec\n", func, realtime() - realtime2);
    }

    OutOpts oo; oo.bedmethyl = opt.bedmethyl; oo.insertions = opt.insertions; oo.haplotypes = opt.haplotypes;
    if (subtool == MMC_FREQ) print_freq_header(opt.out, oo); else print_view_header(opt.out, oo);

    LoadOpts lo;
    lo.batch_size = opt.batch_size; lo.batch_size_bases = opt.batch_size_bases;
    lo.allow_secondary = opt.allow_secondary; lo.skip_supplementary = opt.skip_supplementary;
    lo.keep_qnames = subtool == MMC_VIEW;
    BatchLoader loader(&bam, lo);

    uint64_t total_reads = 0, total_bytes = 0, processed_reads = 0, processed_bytes = 0;
    double load_time = 0, output_time = 0;
    int32_t counter = 0;
    const int n_slots = 3;
    std::vector<mmc_batch_t *> ring(n_slots, nullptr);
    std::vector<BatchMeta> metas(n_slots);
    auto code_names = [&]() {
        std::vector<std::string> v(256);
        for (int i = 0; i < 256; ++i) v[i] = mmc_code_name(ctx, i);
        return v;
    };
    auto die_read = [&](mmc_batch_t *b, const BatchMeta &meta) {
        (void)b; (void)meta;
        ERROR("%s", mmc_strerror(ctx));
        exit(EXIT_FAILURE);
    };
    int more = 1;
    while (more) {
        const int si = counter % n_slots;
        if (ring[si]) {                                   // recycle the oldest slot (joins the "previous processor")
            if (mmc_batch_release(ctx, ring[si]) != MMC_OK) die_read(ring[si], metas[si]);
            ring[si] = nullptr;
        }
        mmc_batch_t *b = nullptr;
        if (mmc_batch_acquire(ctx, &b) != MMC_OK) { ERROR("%s", mmc_strerror(ctx)); exit(EXIT_FAILURE); }
        double t0 = realtime();
        more = loader.fill(b, &metas[si], &err);
        load_time += realtime() - t0;
        if (more < 0) { ERROR("%s", err.c_str()); exit(EXIT_FAILURE); }
        const BatchStats &st = metas[si].stats;
        fprintf(stderr, "[%s::%.3f*%.2f] %d Entries (%.1fM bases) loaded\n", func, realtime() - realtime0,
                cputime() / (realtime() - realtime0), st.n_recs, st.processed_bytes / (1000.0 * 1000.0));
        if (mmc_batch_submit(ctx, b) != MMC_OK) { ERROR("%s", mmc_strerror(ctx)); exit(EXIT_FAILURE); }
        ring[si] = b;
        if (subtool == MMC_VIEW) {
            const mmc_view_rec_t *recs = nullptr; uint64_t n = 0;
            if (mmc_view_fetch(ctx, b, &recs, &n) != MMC_OK) die_read(b, metas[si]);
            double o0 = realtime();
            print_view_records(opt.out, oo, bam.names, b, metas[si], recs, n, code_names());
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
    for (int i = 0; i < n_slots; ++i)
        if (ring[i] && mmc_batch_release(ctx, ring[i]) != MMC_OK) die_read(ring[i], metas[i]);

    double sort_time = 0;
    if (subtool == MMC_FREQ) {
        const mmc_freq_rec_t *recs = nullptr; uint64_t n = 0;
        double s0 = realtime();
        if (mmc_freq_finalize(ctx, &recs, &n) != MMC_OK) { ERROR("%s", mmc_strerror(ctx)); exit(EXIT_FAILURE); }
        sort_time = realtime() - s0;
        double o0 = realtime();
        print_freq_records(opt.out, oo, bam.names, recs, n, code_names());
        output_time += realtime() - o0;
    }
    if (opt.out != stdout) fclose(opt.out); else fflush(stdout);

    mmc_timers_t tm;
    mmc_get_timers(ctx, &tm);
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
    fprintf(stderr, "\n");
    mmc_destroy(ctx);
    return 0;
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
    fprintf(stderr, "\n[%s] Real time: %.3f sec; CPU time: %.3f sec; Peak RAM: %.3f GB\n\n", __func__, realtime() - realtime0, cputime(),
            peakrss() / 1024.0 / 1024.0 / 1024.0);
    return ret;
}
