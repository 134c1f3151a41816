"""GPU: synthetic workloads of every BASELINE config shape against the oracle (sizes the oracle finishes in
seconds), plus size-independent properties at a larger size: linearity of the counts in the number of passes,
and freq == reduce(view) (the reference's own self-consistency test, test/test.sh:573-585)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, have_ref_bin, sorted_lines, REF_BIN
from minimod_b200 import _native as N
from minimod_b200.synth import CONFIG_ARGS, Synth, cli_args
from parity import Pair
from synth_suite import run_synth

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "minimod_b200", "bin", "minimod")


@pytest.mark.parametrize("config,cov", [(2, 4.0), (3, 3.0), (6, 2.0), (4, 1.0), (5, 3.0)])
@pytest.mark.parametrize("sub", ["freq", "view"])
def test_synthetic_parity_cuda(cuda_lib, config, cov, sub):
    n, st = run_synth(cuda_lib, config, 1500000, cov, sub)
    assert n > 0


@pytest.mark.parametrize("small_smem", ["0", "1"])
def test_synthetic_parity_cuda_scratch(cuda_lib, monkeypatch, small_smem):
    monkeypatch.setenv("MMC_TEST_SMALL_SMEM", small_smem)
    monkeypatch.setenv("MMC_DECODE_PATH", "general")
    run_synth(cuda_lib, 3, 800000, 2.0, "freq")
    run_synth(cuda_lib, 4, 800000, 0.5, "freq")


@pytest.mark.parametrize("path,arena", [("stream", None), ("split", None), ("warp", None), ("split", "5120"), ("general", None)])
@pytest.mark.parametrize("config,cov", [(2, 3.0), (3, 2.0), (6, 1.5), (4, 0.5)])
def test_synthetic_parity_cuda_paths(cuda_lib, monkeypatch, path, arena, config, cov):
    """Every decode path against the oracle on every config shape (50 kb reads of config 4 are mostly deferred
    by the warp kernels; the small arena forces CIGAR / index sampling and more deferral)."""
    monkeypatch.setenv("MMC_DECODE_PATH", path)
    if arena:
        monkeypatch.setenv("MMC_WARP_ARENA", arena)
    n, st = run_synth(cuda_lib, config, 1000000, cov, "freq")
    assert n > 0


@pytest.mark.parametrize("sub", ["freq", "view"])
@pytest.mark.parametrize("config,cov", [(3, 2.0), (6, 1.5), (4, 0.5), (2, 2.0)])
def test_synthetic_parity_cuda_split_blocks(cuda_lib, monkeypatch, config, cov, sub):
    """k_decode_stream with the even and the odd MM blocks of a read on different warps (first ML index of every block from
    k_flat_setup): two-block reads (configs 3, 4, 6), one-block reads (config 2: the odd work unit is a no-op)."""
    monkeypatch.setenv("MMC_DECODE_PATH", "stream")
    monkeypatch.setenv("MMC_STREAM_SPLIT", "1")
    n, st = run_synth(cuda_lib, config, 1000000, cov, sub)
    assert n > 0


@pytest.mark.parametrize("config,cov", [(3, 3.0), (6, 2.0)])
def test_synthetic_parity_cuda_sparse_on_device(cuda_lib, monkeypatch, config, cov):
    """--insertions (config 3) and haplotype (config 6) rows with the sparse side buffer sorted / reduced / merged
    on the device (mmc_sparse.cuh) whatever its size."""
    monkeypatch.setenv("MMC_SPARSE_DEVICE_MIN", "0")
    n, st = run_synth(cuda_lib, config, 1500000, cov, "freq")
    assert n > 0


@pytest.mark.parametrize("config,cov", [(2, 3.0), (3, 2.0), (6, 1.5), (4, 0.5)])
def test_synthetic_parity_cuda_two_bit_seq(cuda_lib, monkeypatch, config, cov):
    """The synthetic configs with SEQ crossing PCIe at 2 bits per base (the CLI's and bench.py's setting)."""
    monkeypatch.setenv("MMC_SEQ_PACKING", "2")
    n, st = run_synth(cuda_lib, config, 1000000, cov, "freq")
    assert n > 0


def test_counts_are_linear_in_passes(cuda_lib):
    """k passes over the same batch give exactly k times the counts of one pass (aggregation is a pure sum)."""
    s = Synth(2, contigs=(("chrS", 6000000),), coverage=8.0)
    ca = CONFIG_ARGS[2]
    p, n = s.ref(0)
    pair = Pair(cuda_lib, "freq", [("chrS", C.string_at(p, n))], ca["mod_codes"], ca["mod_thresh"], max_reads=s.n_reads + 8,
                max_bytes=int(s.n_reads * 20000 * 2))
    try:
        got, _ = s.fill(pair.batch, 0, s.n_reads, 8)
        assert got == s.n_reads
        lib, ctx, b = pair.lib, pair.ctx, pair.batch
        assert lib.mmc_batch_upload(ctx, b) == 0
        assert lib.mmc_batch_launch(ctx, b) == 0 and lib.mmc_sync(ctx) == 0
        one = pair.device_freq()
        for _ in range(4):
            assert lib.mmc_batch_launch(ctx, b) == 0
        assert lib.mmc_sync(ctx) == 0
        five = pair.device_freq()
        assert len(one) == len(five) > 1000
        assert [r[:6] + (5 * r[6], 5 * r[7]) for r in one] == five
        assert lib.mmc_freq_reset(ctx) == 0
        assert pair.device_freq() == []
        assert lib.mmc_batch_launch(ctx, b) == 0 and lib.mmc_sync(ctx) == 0
        assert pair.device_freq() == one
    finally:
        pair.close(); s.close()


def test_freq_equals_reduced_view(cuda_lib):
    """Re-aggregating `view` rows with the threshold rule must give `freq` (test/freq.sh:36-76)."""
    s = Synth(3, contigs=(("chrS", 2000000),), coverage=4.0)
    p, n = s.ref(0)
    ref = C.string_at(p, n)
    res = {}
    for sub in ("freq", "view"):
        pair = Pair(cuda_lib, sub, [("chrS", ref)], "m[CG],h[CG]", "0.8,0.7" if sub == "freq" else None, max_reads=s.n_reads + 8,
                    max_bytes=int(s.n_reads * 30000 * 2))
        try:
            s.fill(pair.batch, 0, s.n_reads, 8)
            rc, msg = pair.run_device()
            assert rc == 0, msg
            res[sub] = pair.device_freq() if sub == "freq" else pair.device_view()
        finally:
            pair.close()
    s.close()
    agg = {}
    lim = {"m": (205, 50), "h": (179, 76)}          # (p+0.5)/256 >= t  /  <= 1-t for t = 0.8, 0.7
    for read, ref_pos, read_pos, code, ins, prob, strand, hp in res["view"]:
        hi, lo = lim[code]
        if prob >= hi or prob <= lo:
            k = (0, ref_pos, strand, code, 0, -1)
            c = agg.setdefault(k, [0, 0])
            c[0] += 1; c[1] += prob >= hi
    assert sorted(k + (v[0], v[1]) for k, v in agg.items()) == res["freq"]


@pytest.mark.skipif(not have_ref_bin(), reason="oracle/_ref/minimod_ref not present")
@pytest.mark.parametrize("config", [2, 3, 4, 5])
def test_multi_contig_cli_vs_reference_binary(tmp_path, config):
    """The `minimod` binary vs the unmodified reference on a multi-contig synthetic BAM whose contig names do
    not sort in header order (Q14): raw bytes for the tie-free configs, sorted lines otherwise."""
    contigs = (("t2", 400000), ("t10", 300000), ("t1_x", 200000))
    s = Synth(config, contigs=contigs, coverage=2.0 if config != 4 else 0.6)
    fa, bam = str(tmp_path / "ref.fa"), str(tmp_path / "reads.bam")
    s.write_fasta(fa); s.write_bam(bam); s.close()
    args = cli_args(config)
    mine = subprocess.run([CLI, "freq"] + args + ["-K", "97", fa, bam], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    ref = subprocess.run([REF_BIN, "freq"] + args + ["-t", "8", fa, bam], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert mine.returncode == 0, mine.stderr.decode()[-1500:]
    assert ref.returncode == 0
    if config in (2, 5):
        assert mine.stdout == ref.stdout
    else:
        assert sorted_lines(mine.stdout) == sorted_lines(ref.stdout)
    order = [l.split(b"\t")[0] for l in mine.stdout.splitlines() if not l.startswith(b"contig")]
    assert [c for i, c in enumerate(order) if i == 0 or order[i - 1] != c] == [b"t10", b"t1_x", b"t2"]


GRCH38_PRIMARY = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
                  135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
                  46709983, 50818468, 156040895, 57227415, 16569]


@pytest.mark.skipif(not have_ref_bin(), reason="oracle/_ref/minimod_ref not present")
@pytest.mark.parametrize("devices", [None, "all"])
def test_config5_grch38_shaped_vs_reference_binary(tmp_path, devices):
    """BASELINE config 5 at test size: the 25 primary GRCh38 contigs (lengths / 250) + 6 small ones, ONT 5mC reads at 2x,
    `freq -c m[CG] -m 0.8` -- the `minimod` binary (one device, and contig-sharded over every device of the box) must be
    byte-identical to the unmodified reference binary (single-mod context-checked output is tie-free)."""
    names = ["chr%d" % (i + 1) for i in range(22)] + ["chrX", "chrY", "chrM"] + ["chrUn_%d" % i for i in range(6)]
    lens = [max(16569, l // 250) for l in GRCH38_PRIMARY] + [40000 + 3000 * i for i in range(6)]
    s = Synth(5, contigs=tuple(zip(names, lens)), coverage=2.0)
    assert s.n_reads > 1000
    fa, bam = str(tmp_path / "ref.fa"), str(tmp_path / "reads.bam")
    s.write_fasta(fa); s.write_bam(bam); s.close()
    args = cli_args(5)
    extra = []
    if devices:
        import torch
        n = torch.cuda.device_count()
        extra = ["--devices", ",".join(str(i) for i in range(n)) if n > 1 else "0,0"]
    mine = subprocess.run([CLI, "freq"] + args + ["-K", "512"] + extra + [fa, bam], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    ref = subprocess.run([REF_BIN, "freq"] + args + ["-t", "16", "-K", "4092", "-B", "100M", fa, bam], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert mine.returncode == 0, mine.stderr.decode()[-1500:]
    assert ref.returncode == 0
    assert mine.stdout == ref.stdout and len(mine.stdout.splitlines()) > 100000
