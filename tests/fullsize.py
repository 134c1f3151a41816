"""Full-size jobs of BASELINE.json (configs 2-5) as files + the digest of the freq table, shared by
tools/make_fullsize_checksums.py (runs the UNMODIFIED reference binary here, where /root/reference exists, and commits
tests/golden/fullsize_checksums.json) and tests/test_gpu_fullsize.py (runs the CUDA tool on the GPU box on the same
deterministic synthetic files and compares digests).  The digest (tools/linesum.c) is order-independent: rows sharing
(contig,pos) have no defined order in the reference."""
import ctypes as C
import json
import os
import subprocess

from helpers import DATA, GOLDEN, ROOT
from minimod_b200 import _native as N
from minimod_b200.synth import Synth, cli_args

CHECKSUMS = os.path.join(GOLDEN, "fullsize_checksums.json")
# config -> (coverage; 0 = the config's own 30x)          config 5: the 195-contig GRCh38-shaped table at 1x (3.1 Gbp, ~205 k reads)
JOBS = {2: 0.0, 3: 0.0, 4: 0.0, 5: 1.0}


def grch38_table():
    host = N.load_host()
    err = C.create_string_buffer(512)
    h = host.mmh_bam_open(os.fsencode(os.path.join(DATA, "example-ont.bam")), err, 512)
    assert h, err.value
    tab = [(host.mmh_bam_target_name(h, i).decode(), int(host.mmh_bam_target_len(h, i))) for i in range(host.mmh_bam_n_targets(h))]
    host.mmh_bam_close(h)
    return tab


def write_job(config, outdir, threads=8):
    """-> (fasta, bam, cli args, stats)"""
    cov = JOBS[config]
    s = Synth(config, contigs=grch38_table(), coverage=cov) if config == 5 else Synth(config, coverage=cov)
    fa, bam = os.path.join(outdir, f"c{config}.fa"), os.path.join(outdir, f"c{config}.bam")
    s.write_fasta(fa)
    st = s.write_bam(bam, threads=threads)
    s.close()
    return fa, bam, cli_args(config), st


def linesum_bin():
    exe = os.path.join(ROOT, "oracle", "_build", "linesum")
    src = os.path.join(ROOT, "tools", "linesum.c")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-o", exe, src], check=True)
    return exe


def digest_of_command(cmd, env=None):
    """Runs cmd, pipes its stdout through linesum; -> dict(lines, sum, xor), stderr text."""
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    q = subprocess.Popen([linesum_bin()], stdin=p.stdout, stdout=subprocess.PIPE)
    p.stdout.close()
    out = q.communicate()[0].decode().split()
    err = p.stderr.read().decode()
    assert p.wait() == 0, err[-2000:]
    return dict(lines=int(out[0]), sum=out[1], xor=out[2]), err


def load_checksums():
    with open(CHECKSUMS) as fh:
        return {int(k): v for k, v in json.load(fh).items()}
