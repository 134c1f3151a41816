"""One process, several device contexts (minimod --devices): batches are dealt to the contexts by contig owner (LPT) or,
with --shard-regions, by read start inside the longest contig; the boundary counts are summed by mmc_region_reduce()
(ncclAllReduce over NVLink between devices; a device-local add when the contexts share a device, which is what lets the
CPU-only suite run the same host logic on the SIMT emulator).  The output must be byte-identical to the single-device run.
Reference: one process drives everything, /root/reference/src/freq_main.c:404-474."""
import os
import subprocess

import pytest

from helpers import ROOT
from minimod_b200.synth import Synth, cli_args

EMUL_CLI = os.path.join(ROOT, "tests", "kernel_emul", "_build", "minimod_emul")
CLI = os.path.join(ROOT, "minimod_b200", "bin", "minimod")

CASES = [  # config, contigs, coverage
    (5, (("t2", 300000), ("t10", 200000), ("t1_x", 150000), ("t7", 90000)), 2.0),
    (3, (("big", 400000), ("s1", 100000), ("s2", 60000)), 2.0),          # --insertions: sparse rows on both sides of a boundary
    (6, (("big", 300000), ("s1", 80000)), 1.5),                          # '.' status blocks
    (4, (("big", 260000), ("s1", 120000)), 0.7),                         # --haplotypes, 50 kb reads: halos wider than a slice
]


def run(cli, sub, args, fa, bam, extra):
    r = subprocess.run([cli, sub] + args + ["-K", "37"] + extra + [fa, bam], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-1500:]
    return r.stdout


def check(cli, tmp_path, config, contigs, cov, devices):
    s = Synth(config, contigs=contigs, coverage=cov)
    fa, bam = str(tmp_path / "ref.fa"), str(tmp_path / "reads.bam")
    s.write_fasta(fa); s.write_bam(bam); s.close()
    args = cli_args(config)
    one = run(cli, "freq", args, fa, bam, [])
    assert len(one.splitlines()) > 100
    assert run(cli, "freq", args, fa, bam, ["--devices", devices]) == one
    assert run(cli, "freq", args, fa, bam, ["--devices", devices, "--shard-regions"]) == one
    v1 = run(cli, "view", [a for a in args if a not in ("-b",)][:2] + (["--insertions"] if "--insertions" in args else []), fa, bam, [])
    assert run(cli, "view", [a for a in args if a not in ("-b",)][:2] + (["--insertions"] if "--insertions" in args else []), fa, bam,
               ["--devices", devices]) == v1


@pytest.mark.skipif(not os.path.exists(EMUL_CLI), reason="emulator CLI not built")
@pytest.mark.parametrize("config,contigs,cov", CASES, ids=[f"config{c[0]}" for c in CASES])
def test_multi_context_cli_emulated(emul_lib, tmp_path, config, contigs, cov):
    check(EMUL_CLI, tmp_path, config, contigs, cov, "0,0,0")


@pytest.mark.gpu
@pytest.mark.parametrize("config,contigs,cov", CASES, ids=[f"config{c[0]}" for c in CASES])
def test_multi_device_cli_cuda(tmp_path, config, contigs, cov):
    """On a box with several GPUs: real devices and NCCL; on one GPU: three contexts on device 0."""
    import torch
    n = torch.cuda.device_count()
    devices = ",".join(str(i) for i in range(min(n, 4))) if n > 1 else "0,0,0"
    check(CLI, tmp_path, config, contigs, cov, devices)
