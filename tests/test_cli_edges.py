"""Edge cases of the whole tool that no reference golden pins (last bullet), checked against the unmodified
reference binary on the same files:
  Q15  -b together with --haplotypes / --insertions (duplicate-looking bedMethyl rows, /root/reference/src/mod.c:685-703)
  Q10  ins_offset >= 65536: freq truncates it to uint16 in the key (src/mod.c:428), view prints it whole (src/mod.c:608)
  CG   CIGARs of more than 65535 ops travel in a CG:B,I tag and are restored like htslib's bam_tag2cigar()
  per-read fatal messages carry the read name (src/mod.c:843,1174)
  the sparse side buffer grows instead of overflowing at the end of the run
CPU: the CLI linked to the SIMT emulation of the kernels; GPU: the real binary."""
import gzip
import os
import struct
import subprocess

import pytest

from helpers import DATA, REF_BIN, ROOT, have_ref_bin, pseudo_fasta, sorted_lines
from minimod_b200.synth import Synth, cli_args

EMUL_CLI = os.path.join(ROOT, "tests", "kernel_emul", "_build", "minimod_emul")
CUDA_CLI = os.path.join(ROOT, "minimod_b200", "bin", "minimod")
CLIS = [pytest.param(EMUL_CLI, id="emul"), pytest.param(CUDA_CLI, id="cuda", marks=pytest.mark.gpu)]
needs_ref = pytest.mark.skipif(not have_ref_bin(), reason="oracle/_ref/minimod_ref not present")


def run(binary, args, check=True):
    r = subprocess.run([binary] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if check:
        assert r.returncode == 0, r.stderr.decode()[-1500:]
    return r


def write_bam(path, contigs, records):
    """Minimal BAM (one gzip member): records = dicts(tid,pos,flag,qname,cigar[(op,len)],seq,aux bytes)."""
    nt16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    out = bytearray(b"BAM\1")
    text = b"@HD\tVN:1.6\tSO:coordinate\n" + b"".join(b"@SQ\tSN:%s\tLN:%d\n" % (n.encode(), l) for n, l in contigs)
    out += struct.pack("<I", len(text)) + text + struct.pack("<I", len(contigs))
    for n, l in contigs:
        out += struct.pack("<I", len(n) + 1) + n.encode() + b"\0" + struct.pack("<I", l)
    for r in records:
        qn = r["qname"].encode() + b"\0"
        cig = b"".join(struct.pack("<I", (ln << 4) | "MIDNSHP=X".index(op)) for op, ln in r["cigar"])
        seq = r["seq"]
        packed = bytearray((len(seq) + 1) // 2)
        for i, c in enumerate(seq):
            packed[i >> 1] |= nt16[c] << (0 if i & 1 else 4)
        data = qn + cig + bytes(packed) + b"\xff" * len(seq) + r["aux"]
        fixed = struct.pack("<iiBBHHHiiii", r["tid"], r["pos"], len(qn), 60, 4680, len(r["cigar"]), r["flag"], len(seq), -1, -1, 0)
        out += struct.pack("<I", len(fixed) + len(data)) + fixed + data
    with gzip.open(path, "wb", compresslevel=1) as fh:
        fh.write(bytes(out))


def mm_ml(mm, ml):
    return b"MMZ" + mm.encode() + b"\0" + b"MLBC" + struct.pack("<I", len(ml)) + bytes(ml)


@needs_ref
@pytest.mark.parametrize("cli", CLIS)
@pytest.mark.parametrize("args,bam,contig", [("-b --haplotypes", "hap.bam", "chr1"), ("-b --insertions", "example-ont.bam", "chr22"),
                                             ("-b --insertions --haplotypes -c m,h -m 0.8,0.5", "example-ont.bam", "chr22")])
def test_q15_bedmethyl_with_haplotypes_insertions(cli, args, bam, contig):
    fa, path = pseudo_fasta(contig), os.path.join(DATA, bam)
    mine = run(cli, ["freq"] + args.split() + [fa, path]).stdout
    ref = run(REF_BIN, ["freq"] + args.split() + ["-t", "4", fa, path]).stdout
    assert len(mine) > 1000 and sorted_lines(mine) == sorted_lines(ref)


@needs_ref
@pytest.mark.parametrize("cli", CLIS)
def test_q10_ins_offset_beyond_uint16(cli, tmp_path):
    n_ins = 70000
    ref = "ACGT" * 250
    seq = ref[100:110] + "C" * n_ins + ref[110:120]
    # calls on C's: the first C of the insertion is C number (C's in ref[100:110]) -> skip lists below
    c_before = seq[:10].count("C")
    skips = [c_before + 10, 65500, 30, 100, 4000]             # insertion offsets 11, 65512, 65543, 65644, 69645
    ml = [250, 250, 10, 250, 240]
    fa, bam = str(tmp_path / "r.fa"), str(tmp_path / "r.bam")
    with open(fa, "w") as fh:
        fh.write(">ctg\n" + ref + "\n")
    recs = [dict(tid=0, pos=100, flag=0, qname="longins", cigar=[("M", 10), ("I", n_ins), ("M", 10)], seq=seq,
                 aux=mm_ml("C+m?," + ",".join(str(s) for s in skips) + ";", ml)),
            dict(tid=0, pos=300, flag=16, qname="longins_rev", cigar=[("M", 10), ("I", n_ins), ("M", 10)], seq=ref[300:310] + "G" * n_ins + ref[310:320],
                 aux=mm_ml("C+m?,3,65600,10;", [255, 0, 255]))]
    write_bam(bam, [("ctg", len(ref))], recs)
    for sub, extra in (("freq", ["-c", "m[*]", "--insertions"]), ("view", ["-c", "m[*]", "--insertions"]), ("freq", ["-c", "m[*]", "--insertions", "-b"])):
        mine = run(cli, [sub] + extra + [fa, bam]).stdout
        want = run(REF_BIN, [sub] + extra + ["-t", "2", fa, bam]).stdout
        assert sorted_lines(mine) == sorted_lines(want), (sub, extra)
        assert len(mine.splitlines()) >= 5
    view = run(cli, ["view", "-c", "m[*]", "--insertions", fa, bam]).stdout
    assert any(int(l.split(b"\t")[7]) >= 65536 for l in view.splitlines()[1:])     # printed whole in view


@needs_ref
@pytest.mark.parametrize("cli", CLIS)
def test_long_cigar_in_cg_tag(cli, tmp_path):
    """Reads of 1.3 Mb with ONT error rates: ~74 k CIGAR ops each, written as the placeholder CIGAR + CG:B,I."""
    s = Synth(7, contigs=(("chrU", 9000000),), coverage=0.6)
    fa, bam = str(tmp_path / "u.fa"), str(tmp_path / "u.bam")
    s.write_fasta(fa); st = s.write_bam(bam); s.close()
    assert st["cigar_ops"] == 2 * st["n_reads"]                                    # every read went out in the CG form
    mine = run(cli, ["freq"] + cli_args(7) + [fa, bam]).stdout
    want = run(REF_BIN, ["freq"] + cli_args(7) + ["-t", "4", fa, bam]).stdout
    assert mine == want and len(mine.splitlines()) > 10000


@pytest.mark.parametrize("cli", CLIS)
def test_fatal_messages_name_the_read(cli, tmp_path):
    ref = "ACGT" * 250
    fa = str(tmp_path / "r.fa")
    with open(fa, "w") as fh:
        fh.write(">ctg\n" + ref + "\n")
    ok = dict(tid=0, pos=10, flag=0, qname="fine_read", cigar=[("M", 20)], seq=ref[10:30], aux=mm_ml("C+m?,0,1;", [200, 10]))
    clipped = dict(tid=0, pos=40, flag=0, qname="hardclipped_read", cigar=[("H", 5), ("M", 20)], seq=ref[40:60], aux=mm_ml("C+m?,0;", [200]))
    short_ml = dict(tid=0, pos=80, flag=0, qname="short_ml_read", cigar=[("M", 20)], seq=ref[80:100], aux=mm_ml("C+m?,0,0,0;", [200]))
    for bad, needle in ((clipped, b"Hard clipping found in hardclipped_read"), (short_ml, b"read_id:short_ml_read")):
        bam = str(tmp_path / (bad["qname"] + ".bam"))
        write_bam(bam, [("ctg", len(ref))], [ok, bad])
        r = run(cli, ["freq", "-c", "m[*]", fa, bam], check=False)
        assert r.returncode != 0 and needle in r.stderr, r.stderr.decode()[-800:]


@needs_ref
@pytest.mark.parametrize("cli", CLIS)
def test_sparse_side_buffer_grows(cli, tmp_path):
    """--insertions with a side buffer that starts far too small for the run (it used to overflow at the very end)."""
    s = Synth(3, contigs=(("big", 300000), ("s1", 60000)), coverage=2.0)
    fa, bam = str(tmp_path / "ref.fa"), str(tmp_path / "reads.bam")
    s.write_fasta(fa); s.write_bam(bam); s.close()
    mine = run(cli, ["freq"] + cli_args(3) + ["-K", "11", "--sparse-cap", "64", fa, bam]).stdout
    want = run(REF_BIN, ["freq"] + cli_args(3) + ["-t", "4", fa, bam]).stdout
    assert sorted_lines(mine) == sorted_lines(want) and len(mine.splitlines()) > 5000


@needs_ref
@pytest.mark.parametrize("cli", CLIS)
def test_rows_leave_early_and_the_text_is_identical(cli, tmp_path):
    """Coordinate-sorted, several contigs (names out of strcmp order), many small batches: most rows are drained and formatted
    while the BAM is still being read (mmc_freq_drain); the output is byte-identical to the reference's."""
    s = Synth(2, contigs=(("zeta", 200000), ("alpha", 120000), ("chrM", 16569), ("beta", 90000)), coverage=3.0)
    fa, bam = str(tmp_path / "ref.fa"), str(tmp_path / "reads.bam")
    s.write_fasta(fa); s.write_bam(bam); s.close()
    r = run(cli, ["freq"] + cli_args(2) + ["-K", "9", fa, bam])
    want = run(REF_BIN, ["freq"] + cli_args(2) + ["-t", "4", fa, bam]).stdout
    assert r.stdout == want and len(want.splitlines()) > 3000
    line = [l for l in r.stderr.decode().splitlines() if "left the device" in l][0]
    total, early = int(line.split("Rows: ")[1].split(",")[0]), int(line.split("of which ")[1].split()[0])
    assert total == len(want.splitlines()) and early > total // 2, line
    env = dict(os.environ, MINIMOD_NO_DRAIN="1")
    plain = subprocess.run([cli, "freq"] + cli_args(2) + ["-K", "9", fa, bam], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    assert plain.returncode == 0 and plain.stdout == want


@needs_ref
@pytest.mark.parametrize("cli", CLIS)
def test_unsorted_bam_falls_back_to_one_read_back(cli, tmp_path):
    """Reads out of coordinate order (the header even claims SO:coordinate): rows drained early are dropped, the table is
    rebuilt at the end, the output equals the reference's."""
    import random
    rnd = random.Random(7)
    ref = "".join(rnd.choice("ACGT") for _ in range(6000))
    fa, bam = str(tmp_path / "r.fa"), str(tmp_path / "r.bam")
    with open(fa, "w") as fh:
        fh.write(">c1\n" + ref + "\n>c0\n" + ref[::-1] + "\n")
    recs = []
    for i in range(60):
        tid = i % 2
        src = ref if tid == 0 else ref[::-1]
        pos = rnd.randrange(0, 5000)
        seq = src[pos:pos + 400]
        ncs = seq.count("C")
        k = min(ncs, 12)
        recs.append(dict(tid=tid, pos=pos, flag=0, qname=f"r{i}", cigar=[("M", 400)], seq=seq,
                         aux=mm_ml("C+m?," + ",".join("1" for _ in range(k)) + ";", [rnd.choice((5, 250)) for _ in range(k)])))
    write_bam(bam, [("c1", len(ref)), ("c0", len(ref))], recs)
    r = run(cli, ["freq", "-c", "m[*]", "-K", "4", fa, bam])
    want = run(REF_BIN, ["freq", "-c", "m[*]", "-t", "2", fa, bam]).stdout
    assert sorted_lines(r.stdout) == sorted_lines(want) and len(want.splitlines()) > 200
    assert b"not in coordinate order" in r.stderr
