#!/usr/bin/env python3
"""bench.py -- minimod freq decode+aggregate throughput on B200 (see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2] [--impl reference]

Own arm.  Workload = BASELINE.json configs[1]: synthetic chr22 PacBio HiFi 30x (~15 kb reads) with
MM/ML 5mC CG tags, `freq -c m[CG] -m 0.8 -b`.  One step = one pass of the hot path (decode stage:
MM/ML decode, CIGAR mapping, context check, threshold, dense aggregation) over the whole batch.
  value   reads/s with inputs already resident in HBM (W warm-up + exactly K timed steps between
          barrier+synchronize pairs; max over ranks); inputs (~1 GB) exceed the 126 MB L2.  One step launches
          the decode stage: k_flat_setup, k_decode_warp<PRE> and the two fallback kernels (4 launches).
  e2e     reads/s through the C ABI with HOST (pinned) buffers: per step reset counts, H2D of every
          batch chunk + kernels on pipelined streams, finalize (compaction) and D2H of the rows.
  roofline / cpu_baseline / clocks / gpu_launches: see the JSON keys.
With N>1 (torchrun) every rank owns one chr22-shaped contig of its own (contig sharding, no
data-path collective): weak scaling.

Reference arm (--impl reference): the UNMODIFIED reference (oracle/_ref/minimod_ref, built from
/root/reference by oracle/Makefile) on the host cores, `-t nproc -K 4092 -B 100M`, on a bounded sample
of the same workload; decode+aggregate seconds = its own "Data processing time" + "Data merging time".
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
}
CPU_SAMPLE_READS = {2: 24000, 3: 24000, 4: 600}
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "minimod_ref")


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
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

    def stop(self, t0, t1, more=()):
        """Samples taken inside [t0,t1] or any (a,b) of `more` (the timed regions; the GPU idles in between)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        wins = [(t0, t1)] + list(more)
        rows = [r for t, r in self.rows if any(a - 0.05 <= t <= b + 0.15 for a, b in wins)] or [r for _, r in self.rows]
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


def cpu_sample(synth, config, tmpdir, threads):
    from minimod_b200.synth import cli_args
    fa, bam = os.path.join(tmpdir, "ref.fa"), os.path.join(tmpdir, "sample.bam")
    if not os.path.exists(fa):
        synth.write_fasta(fa)
    n = min(CPU_SAMPLE_READS[config], synth.n_reads)
    first = (synth.n_reads - n) // 2                        # a slice from the middle of the contig
    st = synth.write_bam(bam, first, n, threads=threads)
    return fa, bam, cli_args(config), st


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from minimod_b200.synth import Synth
    threads = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "freq decode+aggregate throughput", "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32", "data": "synthetic", "config": {"workload": WORKLOADS[args.config]}}
    if not os.path.exists(REF_BIN):
        line["unavailable"] = "oracle/_ref/minimod_ref missing (it is built from /root/reference in the dev container)"
        print(json.dumps(line)); return 0
    synth = Synth(args.config)
    with tempfile.TemporaryDirectory(prefix="mmbench_") as td:
        fa, bam, cargs, st = cpu_sample(synth, args.config, td, threads)
        secs, reads = [], 0
        for i in range(args.warmup + args.steps):
            reads, s, _ = run_reference_binary(fa, bam, cargs, threads)
            if i >= args.warmup:
                secs.append(s)
    total = sum(secs)
    v = reads * len(secs) / total
    sample = f"{reads} reads ({st['bases'] / 1e6:.0f} Mbase, {st['ml_entries']} ML entries) from the middle of the contig per step"
    line.update({"value": v, "ms_per_step": 1e3 * total / len(secs), "calls_per_s": st["ml_entries"] * len(secs) / total,
                 "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "reference", "sample": sample},
                 "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    line["config"].update({"timed": "reference's own Data processing time + Data merging time", "threads": threads,
                           "cmd": "minimod_ref freq " + " ".join(cargs) + f" -t {threads} -K 4092 -B 100M"})
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--coverage", type=float, default=0.0, help="override the 30x depth (debugging only)")
    ap.add_argument("--chunks", type=int, default=8, help="batches per job on the e2e path")
    ap.add_argument("--seq-packing", type=int, default=2, choices=(2, 4), help="bits per base of SEQ in the host buffers (2: + exception list, expanded on the device)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from minimod_b200 import _native as N
    from minimod_b200.synth import CONFIG_ARGS, Synth

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib, host = N.load_cuda(), N.load_host()
    host_threads = max(1, (os.cpu_count() or 8) // max(1, world))

    # ---- workload: every rank owns one chr22-shaped contig (contig sharding)
    t_gen = time.time()
    from minimod_b200.synth import SEED0
    synth = Synth(args.config, coverage=args.coverage, seed=SEED0 + args.config + 1000 * rank)
    n_reads = synth.n_reads
    ca = CONFIG_ARGS[args.config]
    mods = (N.MmcMod * N.MMC_MAX_MODS)()
    err = C.create_string_buffer(1024)
    n_mods = host.mmh_parse_mods(ca["mod_codes"].encode(), (ca.get("mod_thresh") or "").encode(), N.MMC_FREQ, mods, N.MMC_MAX_MODS, err, 1024)
    assert n_mods > 0, err.value

    # per-base pool budgets (cigar words, MM bytes, ML bytes) with head-room over the read models of synth.cpp
    RATIO = {2: (0.012, 0.08, 0.04), 3: (0.10, 0.14, 0.07), 4: (0.10, 1.3, 0.6)}[args.config]

    def make_ctx(n_slots, reads_cap, bases_cap):
        o = N.MmcOpts()
        o.struct_size = C.sizeof(N.MmcOpts)
        o.subtool, o.n_mods, o.mods = N.MMC_FREQ, n_mods, mods
        o.insertions, o.haplotypes = int(bool(ca.get("insertions"))), int(bool(ca.get("haplotypes")))
        o.device, o.n_slots, o.max_reads, o.max_bytes = local, n_slots, reads_cap, bases_cap
        o.cap_seq_bytes = int(bases_cap * 0.58) + 16 * reads_cap
        o.cap_cigar_words = int(bases_cap * RATIO[0]) + 16 * reads_cap
        o.cap_mm_bytes = int(bases_cap * RATIO[1]) + 16 * reads_cap
        o.cap_ml_bytes = int(bases_cap * RATIO[2]) + 16 * reads_cap
        o.sparse_capacity = 1 << 26
        o.seq_packing = args.seq_packing
        names = (C.c_char_p * 1)(*synth.names)
        lens = (C.c_uint32 * 1)(*synth.lens)
        ctx = C.c_void_p()
        if lib.mmc_create(C.byref(ctx), C.byref(o), 1, names, lens) != 0:
            raise SystemExit("mmc_create: " + lib.mmc_strerror(None).decode())
        p, n = synth.ref(0)
        if lib.mmc_ref_add(ctx, 0, C.cast(p, C.c_char_p), n) != 0 or lib.mmc_ref_commit(ctx) != 0:
            raise SystemExit("reference: " + lib.mmc_strerror(ctx).decode())
        return ctx

    def chk(ctx, rc):
        if rc != 0:
            raise SystemExit("libminimod_cuda: " + lib.mmc_strerror(ctx).decode())

    # bytes of payload for the whole job: seq L/2 + MM + ML + CIGAR; sized generously from the read model
    mean_len = {2: 15000, 3: 10000, 4: 50000}[args.config]
    job_bases = int(n_reads * mean_len * 1.08) + (4 << 20)

    # ---- context A: the whole job as ONE HBM-resident batch (value / roofline)
    ctxA = make_ctx(1, n_reads + 16, job_bases)
    bA = C.POINTER(N.MmcBatch)()
    chk(ctxA, lib.mmc_batch_acquire(ctxA, C.byref(bA)))
    st = N.MmhSynthStats()
    packed = host.mmh_synth_fill(synth.h, bA, 0, n_reads, host_threads, C.byref(st))
    assert packed == n_reads, (packed, n_reads)
    gen_s = time.time() - t_gen
    chk(ctxA, lib.mmc_batch_upload(ctxA, bA))

    # algorithmic bytes per pass: 32 + 4*n_cigar + ceil(L/2) + |MM| + |ML| + ctx*ceil(span/4) per read, + 8 per emitted update
    ctx_flag = 0 if ca.get("insertions") else int(any(c.split("[")[1] != "*]" for c in ca["mod_codes"].split(",")))
    chk(ctxA, lib.mmc_batch_launch(ctxA, bA)); chk(ctxA, lib.mmc_sync(ctxA))
    recs, nrec = C.POINTER(N.MmcFreqRec)(), C.c_uint64()
    chk(ctxA, lib.mmc_freq_finalize(ctxA, C.byref(recs), C.byref(nrec)))
    import numpy as np
    rows = np.frombuffer((N.MmcFreqRec * max(1, nrec.value)).from_address(C.addressof(recs.contents)), dtype=N.FREQ_DTYPE)[:nrec.value]
    emitted = int(rows["n_called"].astype(np.int64).sum())
    n_rows = int(nrec.value)
    alg_bytes = (32 * st.n_reads + 4 * st.cigar_ops + st.seq_bytes + st.mm_bytes + st.ml_entries + ctx_flag * ((st.ref_span + 3) // 4)
                 + 8 * emitted)
    chk(ctxA, lib.mmc_freq_reset(ctxA))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: W warm-up + exactly K timed launches on HBM-resident inputs
    for _ in range(args.warmup):
        chk(ctxA, lib.mmc_batch_launch(ctxA, bA))
    chk(ctxA, lib.mmc_sync(ctxA))
    tm0 = N.MmcTimers(); lib.mmc_get_timers(ctxA, C.byref(tm0))
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t0 = time.time()
    kernel_ms = []
    for _ in range(args.steps):
        chk(ctxA, lib.mmc_batch_launch(ctxA, bA))
        ms = C.c_double()
        chk(ctxA, lib.mmc_last_decode_ms(ctxA, bA, C.byref(ms)))   # waits for the launch; CUDA events on the launch stream
        kernel_ms.append(ms.value)
    barrier()
    t1 = time.time()
    tm1 = N.MmcTimers(); lib.mmc_get_timers(ctxA, C.byref(tm1))
    launches = int(tm1.kernel_launches - tm0.kernel_launches)
    wall = t1 - t0
    dev_s = sum(kernel_ms) / 1e3

    # finalize cost once (compaction of the dense arrays), reported beside the step
    tf0 = time.time()
    chk(ctxA, lib.mmc_freq_finalize(ctxA, C.byref(recs), C.byref(nrec)))
    fin_wall = time.time() - tf0
    tmf = N.MmcTimers(); lib.mmc_get_timers(ctxA, C.byref(tmf))
    lib.mmc_batch_release(ctxA, bA)
    lib.mmc_destroy(ctxA)

    # ---- e2e: the same job through the public C ABI from pinned HOST buffers, chunked over pipelined slots
    chunks = max(1, args.chunks)
    per = (n_reads + chunks - 1) // chunks
    ctxB = make_ctx(chunks, per + 16, job_bases // chunks + (8 << 20))
    held = []
    for k in range(chunks):
        b = C.POINTER(N.MmcBatch)()
        chk(ctxB, lib.mmc_batch_acquire(ctxB, C.byref(b)))
        first = k * per
        cnt = max(0, min(per, n_reads - first))
        got = host.mmh_synth_fill(synth.h, b, first, cnt, host_threads, None)
        assert got == cnt, (got, cnt)
        held.append(b)

    def e2e_step():
        chk(ctxB, lib.mmc_freq_reset(ctxB))
        for b in held:
            chk(ctxB, lib.mmc_batch_submit(ctxB, b))            # async H2D + kernels on the slot's stream
        chk(ctxB, lib.mmc_freq_finalize(ctxB, C.byref(recs), C.byref(nrec)))   # waits, compacts, D2H of the rows
        return int(nrec.value)

    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    e2e_steps = max(3, min(args.steps, 10))
    lib.mmc_reset_timers(ctxB)
    barrier()
    te0 = time.time()
    for _ in range(e2e_steps):
        rows_e2e = e2e_step()
    barrier()
    te1 = time.time()
    clocks = sampler.stop(t0, t1, more=[(te0, te1)]) if sampler else None   # both timed regions (value, e2e)
    tmB = N.MmcTimers(); lib.mmc_get_timers(ctxB, C.byref(tmB))
    assert rows_e2e == n_rows, (rows_e2e, n_rows)
    e2e_wall = te1 - te0
    for b in held:
        lib.mmc_batch_release(ctxB, b)
    lib.mmc_destroy(ctxB)

    # ---- reduce over ranks: max time, sum of units
    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t.item())

    wall_max, dev_max, e2e_max = allmax(wall), allmax(dev_s), allmax(e2e_wall)
    reads_all, calls_all, bytes_all = allsum(float(n_reads)), allsum(float(st.ml_entries)), allsum(float(alg_bytes))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(REF_BIN):
        threads = os.cpu_count() or 1
        with tempfile.TemporaryDirectory(prefix="mmbench_") as td:
            fa, bam, cargs, sst = cpu_sample(synth, args.config, td, threads)
            r_reads, r_secs, _ = run_reference_binary(fa, bam, cargs, threads)
        cpu = {"value": r_reads / r_secs, "unit": "reads/s", "cores": threads, "kind": "reference",
               "calls_per_s": sst["ml_entries"] / r_secs,
               "sample": f"{r_reads} reads ({sst['bases'] / 1e6:.0f} Mbase) from the middle of the contig, minimod_ref -t {threads} -K 4092 -B 100M, "
                         f"Data processing + merging time {r_secs:.2f} s"}
    elif rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/minimod_ref not present"}

    if rank == 0:
        peak, peak_src = peak_hbm()
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get(f"config{args.config}", {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        achieved = alg_bytes / (statistics.mean(kernel_ms) / 1e3) / 1e9     # rank 0's kernel
        line = {
            "metric": "freq decode+aggregate throughput", "value": reads_all * args.steps / wall_max, "unit": "reads/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
            "calls_per_s": calls_all * args.steps / wall_max,
            "device_timed": {"reads_per_s": reads_all * args.steps / dev_max, "calls_per_s": calls_all * args.steps / dev_max,
                             "kernel_ms_mean": statistics.mean(kernel_ms), "kernel_ms_min": min(kernel_ms)},
            "config": {"workload": WORKLOADS[args.config], "reads_per_gpu": int(n_reads), "bases_per_gpu": int(st.bases),
                       "ml_entries_per_gpu": int(st.ml_entries), "rows": n_rows, "emitted_updates": emitted,
                       "l2": "inputs (%.2f GB per pass) exceed the 126 MB L2" % (alg_bytes / 1e9),
                       "sharding": "one chr22-shaped contig per GPU, no data-path collective",
                       "finalize_ms_once": tmf.finalize_ms, "finalize_wall_ms_once": 1e3 * fin_wall, "gen_s": gen_s,
                       "e2e_chunks": chunks, "e2e_steps": e2e_steps,
                       "seq_transport": "2 bits per base + exception list in the pinned host buffers, expanded to BAM's 4-bit form "
                                        "by k_unpack_seq2 on upload (inside e2e; `value` starts from the expanded, HBM-resident batch)"
                                        if args.seq_packing == 2 else "BAM 4-bit nibbles"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                         "kernel": "decode stage = k_flat_setup + k_decode_warp<3,PRE> (dominant, ~80% of the stage) + the two "
                                   "fallback kernels for deferred reads; CUDA events on the launch stream bracket the whole stage",
                         "reads_deferred_to_fallback_kernels": int(tm1.flat_deferred_reads - tm0.flat_deferred_reads)},
            "e2e": {"value": reads_all * e2e_steps / e2e_max, "unit": "reads/s",
                    "h2d_bytes_per_step": int(tmB.h2d_bytes // e2e_steps), "d2h_bytes_per_step": int(tmB.d2h_bytes // e2e_steps),
                    "ms_per_step": 1e3 * e2e_max / e2e_steps},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
