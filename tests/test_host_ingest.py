"""Host ingest: the block-parallel BGZF reader of minimod_b200/host/bam.cpp must deliver the same
records whatever the thread count, agree with the reference binary reading the same file through zlib, and fail
cleanly on a corrupt block."""
import os
import subprocess

import pytest

from helpers import ROOT, REF_BIN, have_ref_bin, sorted_lines
from minimod_b200.synth import Synth, cli_args

EMUL_CLI = os.path.join(ROOT, "tests", "kernel_emul", "_build", "minimod_emul")


@pytest.fixture(scope="module")
def synth_files(tmp_path_factory, host_lib):
    d = tmp_path_factory.mktemp("ingest")
    s = Synth(3, contigs=(("chrS", 200000), ("chrT", 90000)), coverage=2.0)
    fa, bam = str(d / "ref.fa"), str(d / "reads.bam")
    s.write_fasta(fa)
    st = s.write_bam(bam, 0, None, threads=2)
    s.close()
    assert st["n_reads"] > 20
    head = open(bam, "rb").read(16)
    assert head[:4] == b"\x1f\x8b\x08\x04" and head[12:14] == b"BC"          # BGZF
    return fa, bam


def run_cli(args, fa, bam):
    if not os.path.exists(EMUL_CLI):
        subprocess.run(["make", "-C", ROOT, "emul-cli"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return subprocess.run([EMUL_CLI, "freq"] + args + [fa, bam], stdout=subprocess.PIPE, stderr=subprocess.PIPE)


def test_thread_count_does_not_change_the_output(emul_lib, synth_files):
    fa, bam = synth_files
    outs = [run_cli(cli_args(3) + ["-t", t, "-K", "7"], fa, bam) for t in ("1", "3", "16")]
    for r in outs:
        assert r.returncode == 0, r.stderr.decode()[-2000:]
    assert outs[0].stdout == outs[1].stdout == outs[2].stdout and len(outs[0].stdout) > 1000


@pytest.mark.skipif(not have_ref_bin(), reason="oracle/_ref/minimod_ref not present")
def test_same_file_through_the_reference_binary(emul_lib, synth_files):
    fa, bam = synth_files
    mine = run_cli(cli_args(3) + ["-t", "4"], fa, bam)
    assert mine.returncode == 0, mine.stderr.decode()[-2000:]
    ref = subprocess.run([REF_BIN, "freq"] + cli_args(3) + ["-t", "2", fa, bam], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert ref.returncode == 0, ref.stderr.decode()[-2000:]
    assert sorted_lines(mine.stdout) == sorted_lines(ref.stdout)


def test_corrupt_block_is_an_error_not_a_crash(emul_lib, synth_files, tmp_path):
    fa, bam = synth_files
    blob = bytearray(open(bam, "rb").read())
    blob[len(blob) // 2] ^= 0x5a                                              # somewhere inside a deflate stream
    bad = str(tmp_path / "bad.bam")
    open(bad, "wb").write(bytes(blob))
    r = run_cli(cli_args(3) + ["-t", "4"], fa, bad)
    assert r.returncode != 0 and r.returncode > 0                             # exit(EXIT_FAILURE), no signal


def test_fasta_mapped_parser_matches_kseq_semantics(host_lib, tmp_path):
    """Plain FASTA goes through the mapped, record-parallel parser; gzip through the serial one: same records
    (name up to the first blank, lines joined, CR stripped, a later record of the same name replaces the earlier)."""
    import ctypes as C
    import gzip
    text = (b">ctgA some description\r\nACGTacgt\r\nNNNN\r\n\r\n>ctgB\nTTTT\n>empty\n>ctgA again\nGGGGCCCC\nAAAA" )
    plain, gz = tmp_path / "x.fa", tmp_path / "x.fa.gz"
    plain.write_bytes(text)
    with gzip.open(gz, "wb") as fh:
        fh.write(text)
    got = []
    for path in (plain, gz):
        err = C.create_string_buffer(256)
        h = host_lib.mmh_fasta_load(os.fsencode(str(path)), err, 256)
        assert h, err.value
        recs = [(host_lib.mmh_fasta_name(h, i), C.string_at(host_lib.mmh_fasta_seq(h, i), host_lib.mmh_fasta_len(h, i)))
                for i in range(host_lib.mmh_fasta_n(h))]
        host_lib.mmh_fasta_free(h)
        got.append(recs)
    assert got[0] == got[1] == [(b"ctgA", b"GGGGCCCCAAAA"), (b"ctgB", b"TTTT"), (b"empty", b"")]
