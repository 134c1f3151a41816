"""Shared helpers for the test-suite: fixture paths, pseudo reference genomes, golden files,
and runners for the CPU checkers (oracle/_ref/minimod_ref and the C restatement)."""
import gzip
import os
import shlex
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(GOLDEN, "data")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "minimod_ref")
CACHE = os.environ.get("MINIMOD_B200_TEST_CACHE", "/tmp/minimod_b200_test_cache")

CONTIG_LEN = {"chr22": 50818468, "chr1": 248956422}

# (name, subtool, extra args, bam, contig) -- the commands of /root/reference/test/test.sh:66-250
GOLDEN_CASES = [
    ("test1.tsv", "view", "-c m[CG]", "example-hifi.bam", "chr22"),
    ("test2.tsv", "view", "-c m[CG]", "example-ont.bam", "chr22"),
    ("test2a.tsv", "view", "-c m[CG] --insertions", "example-ont.bam", "chr22"),
    ("test2b.tsv", "view", "-c m[*]", "example-ont.bam", "chr22"),
    ("test2c_wild.tsv", "view", "-c *", "example-ont.bam", "chr22"),
    ("test2c.tsv", "view", "-c m[CG] --haplotypes", "hap.bam", "chr1"),
    ("test3.tsv", "freq", "", "example-hifi.bam", "chr22"),
    ("test4.bedmethyl", "freq", "-b -K 1", "example-hifi.bam", "chr22"),
    ("test5.tsv", "freq", "", "example-ont.bam", "chr22"),
    ("test5a.tsv", "freq", "--insertions", "example-ont.bam", "chr22"),
    ("test5b.tsv", "freq", "-c m[*]", "example-ont.bam", "chr22"),
    ("test5c.tsv", "freq", "--haplotypes", "hap.bam", "chr1"),
    ("test6.bedmethyl", "freq", "-b", "example-ont.bam", "chr22"),
    ("test7.tsv", "freq", "-m 0.8", "example-ont.bam", "chr22"),          # BASELINE config 1
    ("test8.tsv", "freq", "-c m,h -m 0.8,0.8", "example-ont.bam", "chr22"),
    ("test9.tsv", "freq", "-c h", "example-ont.bam", "chr22"),
    ("test10.tsv", "view", "", "example-ont.bam", "chr22"),
    ("test11.tsv", "view", "-c m,h", "example-ont.bam", "chr22"),
    ("test12.tsv", "freq", "-c m,h -m 0.8,0.5", "example-ont.bam", "chr22"),
    ("test15.tsv", "view", "-c e,b", "eb.bam", "chr1"),
    ("test16.tsv", "freq", "-c e,b -m 0.5", "eb.bam", "chr1"),
    ("test17a.tsv", "view", "-c 17802[*]", "dRNA.bam", "chr22"),
]
FREQ_CASES = [c for c in GOLDEN_CASES if c[1] == "freq"]
VIEW_CASES = [c for c in GOLDEN_CASES if c[1] == "view"]


def golden_bytes(name):
    with gzip.open(os.path.join(GOLDEN, "expected", name + ".gz"), "rb") as fh:
        return fh.read()


def sorted_lines(blob):
    """LC_ALL=C sort of the lines (the reference's tests sort both sides before diffing,
    /root/reference/test/test.sh:119-121; rows sharing (contig,pos) have no defined order)."""
    return sorted(blob.split(b"\n"))


def pseudo_fasta(contig):
    """Write (once) the pseudo reference for `contig`: all N with the poke lists applied."""
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, f"pseudo_{contig}.fa")
    if os.path.exists(path):
        return path
    pokes = np.load(os.path.join(GOLDEN, "pseudo_ref.npz"))
    seq = np.full(CONTIG_LEN[contig], ord("N"), dtype=np.uint8)
    for base in "ACGT":
        seq[pokes[f"{contig}_{base}"]] = ord(base)
    tmp = path + f".tmp{os.getpid()}"
    with open(tmp, "wb") as fh:
        fh.write(f">{contig}\n".encode())
        seq.tofile(fh)
        fh.write(b"\n")
    os.replace(tmp, path)
    return path


def have_ref_bin():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_ref(subtool, args, fasta, bam, threads=4):
    """Run the unmodified reference (oracle/_ref/minimod_ref); returns stdout bytes."""
    cmd = [REF_BIN, subtool] + shlex.split(args) + ["-t", str(threads), fasta, bam]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if res.returncode != 0:
        raise RuntimeError(f"{cmd} failed ({res.returncode}): {res.stderr.decode()[-2000:]}")
    return res.stdout
