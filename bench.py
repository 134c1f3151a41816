#!/usr/bin/env python3
"""bench.py -- minimod freq decode+aggregate throughput on B200 (see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 5] [--impl reference]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, NCCL)

Headline workload (every N): BASELINE.json configs[4], "synthetic whole-human GRCh38-shaped ONT 5mC, contig-sharded
across 1/2/4/8 B200": the 195-contig GRCh38 table (header of the reference's own fixture BAMs), ONT log-normal reads
with C+m? CpG tags, `freq -c m[CG] -m 0.8`, at a depth that fits one GPU and keeps the run within minutes (--coverage,
default 2x; 30x is the same per-read work 15 times over).  The contigs are dealt to the N ranks by LPT bin packing
(minimod_b200.shard.lpt_partition): a FIXED TOTAL JOB split over N ranks -- strong scaling, no data-path collective.
One step = one pass of the hot path (decode stage: MM/ML decode, CIGAR mapping, context check, threshold, dense
aggregation) over the rank's whole shard.
  value    reads/s of the whole job: W warm-up + exactly K timed steps between barrier+synchronize pairs, max over ranks;
           inputs are HBM-resident and far larger than the 126 MB L2.
  value_incl_finalize   the same with the compaction of the dense counts (merge + sort equivalent) added to every step.
  e2e      reads/s through the C ABI with HOST (pinned) buffers: per step reset counts, H2D of every batch chunk +
           kernels on pipelined streams, finalize (compaction) and D2H of the rows.
  roofline / cpu_baseline / clocks / gpu_launches: see the JSON keys.
At N=1 the line also carries `configs`: the same measurement of BASELINE configs 2, 3 and 4 (chr22 30x: HiFi,
ONT multi-mod --insertions, 50 kb all-context --haplotypes), each with its own value / roofline / e2e.
At N>1 it carries `region_shard`: config 2 (one contig) split by read start over the N ranks, whose boundary (halo)
count slices are summed with one NCCL all-reduce per boundary, timed on the device.

Reference arm (--impl reference): the UNMODIFIED reference (oracle/_ref/minimod_ref, built from /root/reference by
oracle/Makefile) on the host cores, `-t nproc -K 4092 -B 100M`, on a bounded sample of the same workload;
decode+aggregate seconds = its own "Data processing time" + "Data merging time".
"""
import argparse
import ctypes as C
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    2: "synthetic chr22 PacBio HiFi 30x (~15 kb reads), C+m? CpG tags, freq -c m[CG] -m 0.8 -b",
    3: "synthetic chr22 ONT 30x (~10 kb reads), C+h?/C+m? CpG tags, freq -c m[CG],h[CG] -m 0.8,0.7 --insertions",
    4: "synthetic chr22 ONT 30x 50 kb reads, C+m./A+a. all-context tags + HP, freq -c m[*],a[A] --haplotypes",
    5: "synthetic whole-human GRCh38-shaped (195 contigs, 3.1 Gbp) ONT 5mC (~15 kb log-normal reads, C+m? CpG tags), "
       "freq -c m[CG] -m 0.8, contig-sharded (LPT) across the GPUs",
}
MEAN_LEN = {2: 15000, 3: 10000, 4: 50000, 5: 15000}
# per-base pool budgets (cigar words, MM bytes, ML bytes) with head-room over the read models of synth.cpp
RATIO = {2: (0.012, 0.08, 0.04), 3: (0.10, 0.14, 0.07), 4: (0.10, 1.3, 0.6), 5: (0.10, 0.08, 0.04)}
CPU_SAMPLE_READS = {2: 24000, 3: 24000, 4: 600, 5: 24000}
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "minimod_ref")
FIXTURE_BAM = os.path.join(ROOT, "tests", "golden", "data", "example-ont.bam")


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def grch38_table(host):
    """The 195-contig hg38 table: header of the reference's fixture BAMs."""
    err = C.create_string_buffer(512)
    h = host.mmh_bam_open(os.fsencode(FIXTURE_BAM), err, 512)
    if not h:
        raise SystemExit("cannot read the contig table from " + FIXTURE_BAM + ": " + err.value.decode())
    tab = [(host.mmh_bam_target_name(h, i).decode(), int(host.mmh_bam_target_len(h, i))) for i in range(host.mmh_bam_n_targets(h))]
    host.mmh_bam_close(h)
    return tab


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, windows):
        """Samples taken inside any (a,b) of `windows` (the timed regions; the GPU idles in between)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if any(a - 0.05 <= t <= b + 0.15 for a, b in windows)] or [r for _, r in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference binary on a bounded sample
# ---------------------------------------------------------------------------------------------------------------
def run_reference_binary(fa, bam, args, threads):
    """Returns (reads processed, process+merge seconds, stderr)."""
    cmd = [REF_BIN, "freq"] + args + ["-t", str(threads), "-K", "4092", "-B", "100M", fa, bam]
    res = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    if res.returncode != 0:
        raise RuntimeError(res.stderr[-2000:])
    proc = float(re.search(r"Data processing time: ([0-9.]+) sec", res.stderr).group(1))
    merge = float(re.search(r"Data merging time: ([0-9.]+) sec", res.stderr).group(1))
    reads = int(re.search(r"total processed entries: (\d+)", res.stderr).group(1))
    return reads, proc + merge, res.stderr


def cpu_sample_synth(config, host):
    """The synthetic input of the CPU arm: config 5 -> two whole chromosomes of the GRCh38 table at the bench depth
    (the reference loads only what the FASTA holds); configs 2-4 -> a slice from the middle of the chr22 job."""
    from minimod_b200.synth import Synth
    if config == 5:
        tab = grch38_table(host)
        pick = [i for i, (n, _) in enumerate(tab) if n in ("chr21", "chr22")]
        s = Synth(5, contigs=[tab[i] for i in pick], gids=pick, coverage=4.0)
        return s, 0, s.n_reads, "chr21 + chr22 of the GRCh38-shaped job at 4x"
    s = Synth(config)
    n = min(CPU_SAMPLE_READS[config], s.n_reads)
    return s, (s.n_reads - n) // 2, n, "a slice from the middle of the chr22 job"


def cpu_sample(config, host, tmpdir, threads):
    from minimod_b200.synth import cli_args
    synth, first, n, what = cpu_sample_synth(config, host)
    fa, bam = os.path.join(tmpdir, "ref.fa"), os.path.join(tmpdir, "sample.bam")
    synth.write_fasta(fa)
    st = synth.write_bam(bam, first, n, threads=threads)
    synth.close()
    return fa, bam, cli_args(config), st, what


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from minimod_b200 import _native as N
    threads = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "freq decode+aggregate throughput", "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8/int32", "data": "synthetic", "config": {"workload": WORKLOADS[args.config]}}
    if not os.path.exists(REF_BIN):
        line["unavailable"] = "oracle/_ref/minimod_ref missing (it is built from /root/reference in the dev container)"
        print(json.dumps(line)); return 0
    host = N.load_host()
    with tempfile.TemporaryDirectory(prefix="mmbench_") as td:
        fa, bam, cargs, st, what = cpu_sample(args.config, host, td, threads)
        secs, reads = [], 0
        for i in range(args.warmup + args.steps):
            reads, s, _ = run_reference_binary(fa, bam, cargs, threads)
            if i >= args.warmup:
                secs.append(s)
    total = sum(secs)
    v = reads * len(secs) / total
    sample = f"{reads} reads ({st['bases'] / 1e6:.0f} Mbase, {st['ml_entries']} ML entries) per step: {what}"
    line.update({"value": v, "ms_per_step": 1e3 * total / len(secs), "calls_per_s": st["ml_entries"] * len(secs) / total,
                 "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "reference", "sample": sample},
                 "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    line["config"].update({"timed": "reference's own Data processing time + Data merging time", "threads": threads,
                           "sample": sample, "cmd": "minimod_ref freq " + " ".join(cargs) + f" -t {threads} -K 4092 -B 100M"})
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
class Env:
    """What every measurement needs: libraries, rank layout, collectives."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from minimod_b200 import _native as N
        self.torch, self.dist, self.N, self.args = torch, dist, N, args
        self.rank, self.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.lib, self.host = N.load_cuda(), N.load_host()
        self.host_threads = max(1, (os.cpu_count() or 8) // max(1, self.world))
        self.windows = []                          # timed regions, for the clock sampler

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def chk(self, ctx, rc):
        if rc != 0:
            raise SystemExit("libminimod_cuda: " + self.lib.mmc_strerror(ctx).decode())


_KEEP = []


def make_ctx(env, config, synth, n_slots, reads_cap, bases_cap):
    from minimod_b200.synth import CONFIG_ARGS
    N, lib, host = env.N, env.lib, env.host
    ca = CONFIG_ARGS[config]
    mods = (N.MmcMod * N.MMC_MAX_MODS)()
    err = C.create_string_buffer(1024)
    n_mods = host.mmh_parse_mods(ca["mod_codes"].encode(), (ca.get("mod_thresh") or "").encode(), N.MMC_FREQ, mods, N.MMC_MAX_MODS, err, 1024)
    assert n_mods > 0, err.value
    ratio = RATIO[config]
    o = N.MmcOpts()
    o.struct_size = C.sizeof(N.MmcOpts)
    o.subtool, o.n_mods, o.mods = N.MMC_FREQ, n_mods, mods
    o.insertions, o.haplotypes = int(bool(ca.get("insertions"))), int(bool(ca.get("haplotypes")))
    o.device, o.n_slots, o.max_reads, o.max_bytes = env.local, n_slots, reads_cap, bases_cap
    o.cap_seq_bytes = int(bases_cap * 0.58) + 16 * reads_cap
    o.cap_cigar_words = int(bases_cap * ratio[0]) + 16 * reads_cap
    o.cap_mm_bytes = int(bases_cap * ratio[1]) + 16 * reads_cap
    o.cap_ml_bytes = int(bases_cap * ratio[2]) + 16 * reads_cap
    o.sparse_capacity = 1 << 26
    o.seq_packing = env.args.seq_packing
    o.cigar_packing = env.args.cigar_packing
    nc = len(synth.names)
    names = (C.c_char_p * max(1, nc))(*synth.names)
    lens = (C.c_uint32 * max(1, nc))(*synth.lens)
    ctx = C.c_void_p()
    if lib.mmc_create(C.byref(ctx), C.byref(o), nc, names, lens) != 0:
        raise SystemExit("mmc_create: " + lib.mmc_strerror(None).decode())
    for tid in range(nc):
        p, n = synth.ref(tid)
        if lib.mmc_ref_add(ctx, tid, C.cast(p, C.c_char_p), n) != 0:
            raise SystemExit("reference: " + lib.mmc_strerror(ctx).decode())
    if lib.mmc_ref_commit(ctx) != 0:
        raise SystemExit("reference: " + lib.mmc_strerror(ctx).decode())
    _KEEP.append((mods, names, lens))              # ctypes arrays must outlive the context
    return ctx


def measure(env, config, synth, steps, warmup, first=0, count=None, shard_note=None, halo=None):
    """value / roofline / e2e of reads [first, first+count) of `synth` on this rank's GPU.  Collective: every rank calls it.
    halo: optional callable(ctx) run (and timed) after the decode passes -- the region-sharding exchange."""
    import numpy as np
    from minimod_b200.synth import CONFIG_ARGS
    N, lib, host, args = env.N, env.lib, env.host, env.args
    ca = CONFIG_ARGS[config]
    n_reads = synth.n_reads - first if count is None else count
    t_gen = time.time()
    job_bases = int(n_reads * MEAN_LEN[config] * 1.10) + (4 << 20)

    # ---- context A: the whole shard as ONE HBM-resident batch (value / roofline)
    ctxA = make_ctx(env, config, synth, 1, n_reads + 16, job_bases)
    bA = C.POINTER(N.MmcBatch)()
    env.chk(ctxA, lib.mmc_batch_acquire(ctxA, C.byref(bA)))
    st = N.MmhSynthStats()
    packed = host.mmh_synth_fill(synth.h, bA, first, n_reads, env.host_threads, C.byref(st))
    assert packed == n_reads, (packed, n_reads)
    gen_s = time.time() - t_gen
    env.chk(ctxA, lib.mmc_batch_upload(ctxA, bA))

    # algorithmic bytes per pass: 32 + 4*n_cigar + ceil(L/2) + |MM| + |ML| + ctx*ceil(span/4) per read, + 8 per emitted update
    ctx_flag = 0 if ca.get("insertions") else int(any(c.split("[")[1] != "*]" for c in ca["mod_codes"].split(",")))
    env.chk(ctxA, lib.mmc_batch_launch(ctxA, bA)); env.chk(ctxA, lib.mmc_sync(ctxA))
    recs, nrec = C.POINTER(N.MmcFreqRec)(), C.c_uint64()
    env.chk(ctxA, lib.mmc_freq_finalize(ctxA, C.byref(recs), C.byref(nrec)))
    n_rows = int(nrec.value)
    emitted, checksum = 0, 0
    if n_rows:
        rows = np.frombuffer((N.MmcFreqRec * n_rows).from_address(C.addressof(recs.contents)), dtype=N.FREQ_DTYPE)
        called = rows["n_called"].astype(np.int64)
        emitted = int(called.sum())
        # order-independent checksum of the table (counts weighted by position): equal for any sharding of the same job
        checksum = int(((rows["pos"].astype(np.int64) + 1) * (called + 3 * rows["n_mod"].astype(np.int64))).sum() % (1 << 61))
    alg_bytes = (32 * st.n_reads + 4 * st.cigar_ops + st.seq_bytes + st.mm_bytes + st.ml_entries + ctx_flag * ((st.ref_span + 3) // 4)
                 + 8 * emitted * (2 if ca.get("haplotypes") else 1))
    env.chk(ctxA, lib.mmc_freq_reset(ctxA))

    # ---- value: W warm-up + exactly K timed launches on HBM-resident inputs
    for _ in range(warmup):
        env.chk(ctxA, lib.mmc_batch_launch(ctxA, bA))
    env.chk(ctxA, lib.mmc_sync(ctxA))
    tm0 = N.MmcTimers(); lib.mmc_get_timers(ctxA, C.byref(tm0))
    env.barrier()
    t0 = time.time()
    kernel_ms = []
    for _ in range(steps):
        env.chk(ctxA, lib.mmc_batch_launch(ctxA, bA))
        ms = C.c_double()
        env.chk(ctxA, lib.mmc_last_decode_ms(ctxA, bA, C.byref(ms)))   # waits for the launch; CUDA events on the launch stream
        kernel_ms.append(ms.value)
    env.barrier()
    t1 = time.time()
    env.windows.append((t0, t1))
    tm1 = N.MmcTimers(); lib.mmc_get_timers(ctxA, C.byref(tm1))
    launches = int(tm1.kernel_launches - tm0.kernel_launches)
    wall = t1 - t0
    dev_s = sum(kernel_ms) / 1e3

    # the table of exactly ONE pass for what follows (the sparse side buffer holds a record per call outside the dense cells, so
    # its finalize would otherwise be timed on warmup + K passes' worth of records)
    env.chk(ctxA, lib.mmc_freq_reset(ctxA))
    env.chk(ctxA, lib.mmc_batch_launch(ctxA, bA)); env.chk(ctxA, lib.mmc_sync(ctxA))
    halo_res = None
    if halo is not None:                                   # region sharding: move the boundary counts to their owners
        halo_res = halo(ctxA)

    # finalize (compaction of the dense arrays = merge + sort equivalent), device-timed, three times after a warm-up
    fin_ms = []
    for i in range(4):
        tfa = N.MmcTimers(); lib.mmc_get_timers(ctxA, C.byref(tfa))
        env.chk(ctxA, lib.mmc_freq_finalize(ctxA, C.byref(recs), C.byref(nrec)))
        tfb = N.MmcTimers(); lib.mmc_get_timers(ctxA, C.byref(tfb))
        if i:
            fin_ms.append(tfb.finalize_ms - tfa.finalize_ms)
    fin = statistics.mean(fin_ms)
    desc = lib.mmc_describe(ctxA).decode()
    lib.mmc_batch_release(ctxA, bA)
    lib.mmc_destroy(ctxA)

    # ---- e2e: the same shard through the public C ABI from pinned HOST buffers, chunked over pipelined slots
    chunks = max(1, args.chunks)
    per = (n_reads + chunks - 1) // chunks
    ctxB = make_ctx(env, config, synth, chunks, per + 16, job_bases // chunks + (8 << 20))
    held = []
    for k in range(chunks):
        b = C.POINTER(N.MmcBatch)()
        env.chk(ctxB, lib.mmc_batch_acquire(ctxB, C.byref(b)))
        f0 = k * per
        cnt = max(0, min(per, n_reads - f0))
        got = host.mmh_synth_fill(synth.h, b, first + f0, cnt, env.host_threads, None)
        assert got == cnt, (got, cnt)
        held.append(b)

    # the reads are coordinate-sorted, so the rows before a batch's first read are final once the batches before it are
    # decoded: mmc_freq_drain() compacts and reads them back while the later batches are still crossing PCIe the other way
    lag = max(1, args.drain_lag)
    marks = [(int(b.contents.tid[0]), int(b.contents.pos[0])) if b.contents.n_reads else None for b in held]
    use_drain = not args.no_drain and not os.environ.get("BENCH_NO_DRAIN") and halo is None

    trace = [] if os.environ.get("BENCH_TRACE") else None

    def e2e_step(collect=None):
        t_a = time.time()
        env.chk(ctxB, lib.mmc_freq_reset(ctxB))
        if trace is not None:
            trace.append(("reset", 0, time.time() - t_a, 0))
        total = 0
        for k, b in enumerate(held):
            t_a = time.time()
            env.chk(ctxB, lib.mmc_batch_submit(ctxB, b))            # async H2D + kernels on the slot's stream
            t_b = time.time()
            if use_drain and k >= lag and marks[k - lag + 1] is not None:
                env.chk(ctxB, lib.mmc_freq_drain(ctxB, marks[k - lag + 1][0], marks[k - lag + 1][1], C.byref(recs), C.byref(nrec)))
                total += int(nrec.value)
                if collect is not None and nrec.value:
                    collect(recs, int(nrec.value))
            if trace is not None:
                trace.append(("submit+drain", k, t_b - t_a, time.time() - t_b))
        t_a = time.time()
        env.chk(ctxB, lib.mmc_freq_finalize(ctxB, C.byref(recs), C.byref(nrec)))   # waits, compacts, D2H of the remaining rows
        if trace is not None:
            trace.append(("finalize", 0, time.time() - t_a, 0))
        if collect is not None and nrec.value:
            collect(recs, int(nrec.value))
        return total + int(nrec.value)

    def row_checksum(acc):
        def add(r, n):
            rows = np.frombuffer((N.MmcFreqRec * n).from_address(C.addressof(r.contents)), dtype=N.FREQ_DTYPE)
            called = rows["n_called"].astype(np.int64)
            acc[0] = (acc[0] + int(((rows["pos"].astype(np.int64) + 1) * (called + 3 * rows["n_mod"].astype(np.int64))).sum())) % (1 << 61)
        return add

    for i in range(max(1, min(warmup, 3))):
        acc = [0]
        rows_e2e = e2e_step(row_checksum(acc) if i == 0 else None)   # (untimed) the drained + remaining rows are the table of `value`
        if i == 0 and halo is None:
            assert rows_e2e == n_rows and acc[0] == checksum, (rows_e2e, n_rows, acc[0], checksum)
    e2e_steps = max(3, min(steps, 10))
    lib.mmc_reset_timers(ctxB)
    env.barrier()
    te0 = time.time()
    for _ in range(e2e_steps):
        rows_e2e = e2e_step()
    env.barrier()
    te1 = time.time()
    env.windows.append((te0, te1))
    tmB = N.MmcTimers(); lib.mmc_get_timers(ctxB, C.byref(tmB))
    if halo is None:
        assert rows_e2e == n_rows, (rows_e2e, n_rows)
    e2e_wall = te1 - te0
    if trace is not None and env.rank == 0:
        for what, k, a, b in trace[-(len(held) + 2):]:
            print(f"[e2e trace] {what} {k}: {1e3 * a:.2f} ms, {1e3 * b:.2f} ms", file=sys.stderr)
    for b in held:
        lib.mmc_batch_release(ctxB, b)
    lib.mmc_destroy(ctxB)

    # ---- reduce over ranks: max time, sum of units
    wall_max, dev_max, e2e_max = env.allmax(wall), env.allmax(dev_s), env.allmax(e2e_wall)
    fin_max = env.allmax(fin)
    reads_all, calls_all = env.allsum(float(n_reads)), env.allsum(float(st.ml_entries))
    rows_all, check_all = env.allsum(float(n_rows)), env.allsum(float(checksum % (1 << 40)))
    h2d_all, d2h_all = env.allsum(float(tmB.h2d_bytes)), env.allsum(float(tmB.d2h_bytes))
    bases_all = env.allsum(float(st.bases))
    km = statistics.mean(kernel_ms)
    peak, peak_src = peak_hbm()
    achieved = alg_bytes / (km / 1e3) / 1e9                       # this rank's kernels on this rank's bytes
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh).get(f"config{config}", {})
            traffic = tj.get("dram_bytes_per_launch") if env.world == 1 and tj.get("reads") in (None, int(n_reads)) else None
    except Exception:
        pass
    res = {
        "value": reads_all * steps / wall_max, "unit": "reads/s", "ms_per_step": 1e3 * wall_max / steps,
        "calls_per_s": calls_all * steps / wall_max,
        "value_incl_finalize": reads_all / (dev_max / steps + fin_max / 1e3),
        "device_timed": {"reads_per_s": reads_all * steps / dev_max, "calls_per_s": calls_all * steps / dev_max,
                         "kernel_ms_mean": km, "kernel_ms_min": min(kernel_ms), "finalize_ms": fin},
        "config": {"workload": WORKLOADS[config], "reads": int(reads_all), "bases": int(bases_all), "ml_entries": int(calls_all),
                   "rows": int(rows_all), "rows_checksum": int(check_all), "reads_this_rank": int(n_reads),
                   "emitted_updates_this_rank": emitted,
                   "l2": "inputs (%.2f GB per pass on this rank) exceed the 126 MB L2" % (alg_bytes / 1e9),
                   "gen_s": gen_s, "e2e_chunks": chunks, "e2e_steps": e2e_steps,
                   "e2e_read_back": (f"mmc_freq_drain after each batch submit (watermark = first read of the batch submitted {lag - 1} before it), "
                                     "remainder by mmc_freq_finalize; row count and checksum equal to the single-finalize table") if use_drain
                                    else "one mmc_freq_finalize after the last batch",
                   "seq_transport": "2 bits per base + exception list in the pinned host buffers, expanded to BAM's 4-bit form on upload "
                                    "(inside e2e; `value` starts from the expanded, HBM-resident batch)" if args.seq_packing == 2 else "BAM 4-bit nibbles",
                   "cigar_transport": "a byte per op + escape lists in the pinned host buffers, expanded to BAM's 32-bit words on upload (inside e2e)"
                                      if args.cigar_packing == 8 else "BAM 32-bit words"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                     "kernel": "decode stage = " + desc + "; CUDA events on the launch stream bracket the whole stage (rank 0's shard)",
                     "reads_deferred_to_fallback_kernels": int(tm1.flat_deferred_reads - tm0.flat_deferred_reads)},
        "e2e": {"value": reads_all * e2e_steps / e2e_max, "unit": "reads/s",
                "h2d_bytes_per_step": int(h2d_all // e2e_steps), "d2h_bytes_per_step": int(d2h_all // e2e_steps),
                "ms_per_step": 1e3 * e2e_max / e2e_steps},
        "gpu_launches": launches,
    }
    if shard_note:
        res["config"]["sharding"] = shard_note
    if halo_res is not None:
        res["halo"] = halo_res
    return res


def region_halo(env, tid=0):
    """Returns the callable measure() runs after the decode passes of a region-sharded contig: one NCCL all-reduce per
    boundary over the dense count cells past it (minimod_b200.shard), timed with CUDA events on torch's stream."""
    from minimod_b200 import shard
    torch, dist, lib = env.torch, env.dist, env.lib

    def run(ctx):
        stats = {}
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        width = shard.exchange_halos(lib, ctx, tid, None, env.rank, dist, cuda=True, stats=stats)   # bounds: from each rank's first read
        ev1.record()
        torch.cuda.synchronize()
        return {"allreduce_ms": ev0.elapsed_time(ev1), "bytes": int(stats.get("bytes", 0)), "max_halo_positions": int(width),
                "collective": "ncclAllReduce(sum, int64 view of the n_called|n_mod cells) per boundary, over NVLink (minimod_b200.shard.exchange_halos)"}
    return run


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=5, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--coverage", type=float, default=0.0, help="depth of the job (default: 2x for config 5, 30x for configs 2-4)")
    ap.add_argument("--chunks", type=int, default=8, help="batches per job on the e2e path")
    ap.add_argument("--seq-packing", type=int, default=2, choices=(2, 4), help="bits per base of SEQ in the host buffers (2: + exception list, expanded on the device)")
    ap.add_argument("--cigar-packing", type=int, default=8, choices=(8, 32), help="CIGARs in the host buffers: a byte per op + escape lists (expanded on the device), or BAM's 32-bit words")
    ap.add_argument("--no-drain", action="store_true", help="e2e: read all rows back after the last batch instead of draining finished positions early")
    ap.add_argument("--drain-lag", type=int, default=1, help="e2e: batches kept in flight behind the drain watermark")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only", action="store_true", help="only the headline workload: no per-config sub-results / region-sharding leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    from minimod_b200 import shard
    from minimod_b200.synth import SEED0, Synth
    env = Env(args)
    rank, world = env.rank, env.world
    sampler = ClockSampler(env.local) if rank == 0 else None

    # ---- headline workload
    if args.config == 5:
        tab = grch38_table(env.host)
        bins, load = shard.lpt_partition([l for _, l in tab], world)
        own = bins[rank]
        cov = args.coverage or 2.0
        synth = Synth(5, contigs=[tab[i] for i in own], gids=own, coverage=cov)
        note = (f"strong scaling: the {len(tab)} contigs dealt to {world} rank(s) by LPT bin packing (rank 0: {len(own)} contigs, "
                f"{load[rank] / 1e6:.0f} Mbp; max/mean load {max(load) * world / sum(load):.3f}); no data-path collective; depth {cov:g}x")
        scaling = "strong"
    else:
        synth = Synth(args.config, coverage=args.coverage, seed=SEED0 + args.config + 1000 * rank)
        note = "one chr22-shaped contig per GPU (weak scaling), no data-path collective"
        scaling = "weak"
    head = measure(env, args.config, synth, args.steps, args.warmup, shard_note=note)
    synth.close()

    line = {"metric": "freq decode+aggregate throughput", "value": head["value"], "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic"}
    for k in ("calls_per_s", "value_incl_finalize", "device_timed", "config", "roofline", "e2e", "gpu_launches"):
        line[k] = head[k]

    # ---- N=1: the other BASELINE configs; N>1: config 2 region-sharded with the NCCL halo reduce
    if not args.only and world == 1:
        line["configs"] = {}
        for c in (2, 3, 4):
            s = Synth(c)
            r = measure(env, c, s, max(5, min(args.steps, 10)), args.warmup, shard_note="single contig, one GPU")
            s.close()
            line["configs"][f"c{c}"] = r
    if not args.only and world > 1:
        s = Synth(2)                                            # the same chr22 job on every rank; rank r takes the r-th slice of the starts
        n = s.n_reads
        f0, f1 = n * rank // world, n * (rank + 1) // world
        r = measure(env, 2, s, max(5, min(args.steps, 10)), args.warmup, first=f0, count=f1 - f0,
                    shard_note=f"region sharding: reads dealt to {world} ranks by start position (equal counts), halo cells summed "
                               "with one NCCL all-reduce per boundary after the decode passes", halo=region_halo(env))
        s.close()
        hm = env.allmax(r["halo"]["allreduce_ms"])
        r["halo"]["allreduce_ms_max_over_ranks"] = hm
        r["value_incl_halo_reduce"] = r["config"]["reads"] / (r["config"]["reads"] / r["device_timed"]["reads_per_s"] + hm / 1e3)
        line["region_shard"] = r

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(REF_BIN):
        threads = os.cpu_count() or 1
        cpu = {}
        for c in ([args.config] if args.only else [args.config, 2, 3, 4]):
            with tempfile.TemporaryDirectory(prefix="mmbench_") as td:
                fa, bam, cargs, sst, what = cpu_sample(c, env.host, td, threads)
                r_reads, r_secs, _ = run_reference_binary(fa, bam, cargs, threads)
            cpu[c] = {"value": r_reads / r_secs, "unit": "reads/s", "cores": threads, "kind": "reference",
                      "calls_per_s": sst["ml_entries"] / r_secs,
                      "sample": f"{r_reads} reads ({sst['bases'] / 1e6:.0f} Mbase): {what}; minimod_ref -t {threads} -K 4092 -B 100M, "
                                f"Data processing + merging time {r_secs:.2f} s"}
    elif rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = {args.config: {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/minimod_ref not present"}}

    if rank == 0:
        line["clocks"] = sampler.stop(env.windows) if sampler else None
        if cpu is not None:
            line["cpu_baseline"] = cpu[args.config]
            for c, v in cpu.items():
                if c != args.config and f"c{c}" in line.get("configs", {}):
                    line["configs"][f"c{c}"]["cpu_baseline"] = v
        print(json.dumps(line))
    if world > 1:
        env.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
