"""Edge-case table shared by the CPU (emulated kernels) and GPU tests."""
import random

from mkbatch import add_read
from parity import Pair

random.seed(7)
REF = "".join(random.choice("ACGT") for _ in range(4000))
# sprinkle CpGs and a run of N / IUPAC letters
REF = REF[:100] + "CGCGACGTTCGA" + REF[112:600] + "NNNNNNNNNN" + REF[610:900] + "RYCGM" + REF[905:]
REF2 = REF[::-1]


def comp(s):
    return s.translate(str.maketrans("ACGTN", "TGCAN"))


def fwd_read(start, n):
    return REF[start:start + n]


def mm_for(seq, base, code, status="?", every=1, rev=False):
    """MM block listing every `every`-th occurrence of `base` in the ORIGINAL read orientation."""
    orig = comp(seq)[::-1] if rev else seq
    skips, skip, k = [], 0, 0
    for ch in orig:
        if ch == base:
            if k % every == 0:
                skips.append(skip); skip = 0
            else:
                skip += 1
            k += 1
    return f"{base}+{code}{status}" + "".join(f",{s}" for s in skips) + ";", len(skips)


def ml_bytes(n, seed=1):
    r = random.Random(seed)
    return bytes(r.choice([0, 10, 50, 51, 76, 77, 127, 128, 178, 179, 204, 205, 229, 230, 255]) for _ in range(n))


def build_basic(bp, rev=False, code="m", status="?", every=1, cigar=None, start=90, n=300, flag_extra=0, hp=0, K=1):
    seq = fwd_read(start, n)
    mm, cnt = mm_for(seq, "C", code, status, every, rev)
    add_read(bp, 0, start, (16 if rev else 0) | flag_extra, seq, cigar or f"{n}M", mm, ml_bytes(cnt * K), hp)


def case(id, build, **kw):
    d = dict(id=id, build=build, subtools=("freq", "view"), codes="m", thresh=None, insertions=False, haplotypes=False, opts={})
    d.update(kw)
    return d


def b_two_blocks(bp):
    seq = fwd_read(90, 400)
    m1, c1 = mm_for(seq, "C", "h", "?")
    m2, c2 = mm_for(seq, "C", "m", "?")
    add_read(bp, 0, 90, 0, seq, "400M", m1 + m2, ml_bytes(c1 + c2))
    rs = fwd_read(500, 350)
    m1, c1 = mm_for(rs, "C", "h", ".", 2, rev=True)
    m2, c2 = mm_for(rs, "C", "m", ".", 3, rev=True)
    add_read(bp, 0, 500, 16, rs, "350M", m1 + m2, ml_bytes(c1 + c2, 3))


def b_combined(bp):
    for rev in (False, True):
        seq = fwd_read(95, 300)
        mm, cnt = mm_for(seq, "C", "mh", "?", 1, rev)
        add_read(bp, 0, 95, 16 if rev else 0, seq, "300M", mm, ml_bytes(cnt * 2, 5))


def b_cigar_mix(bp):
    seq = fwd_read(80, 100) + "ACGCG" + fwd_read(180, 120) + fwd_read(310, 80) + "TTCGT"
    cig = "100M5I120M10D80M5S"
    for rev in (False, True):
        for status in ("?", "."):
            mm, cnt = mm_for(seq, "C", "m", status, 2, rev)
            add_read(bp, 0, 80, 16 if rev else 0, seq, cig, mm, ml_bytes(cnt, 9))


def b_leading_ins(bp):
    seq = "CGCG" + fwd_read(0, 60)
    mm, cnt = mm_for(seq, "C", "m", "?")
    add_read(bp, 0, 0, 0, seq, "4I60M", mm, ml_bytes(cnt))       # Q11: insertion at pos 0 -> ins = -1
    seq2 = "ACG" + fwd_read(50, 60)
    mm, cnt = mm_for(seq2, "C", "m", ".")
    add_read(bp, 0, 50, 0, seq2, "3S60M", mm, ml_bytes(cnt))


def b_n_base(bp):
    seq = fwd_read(590, 60)                                       # covers the N run
    for rev in (False, True):
        for status in ("?", "."):
            add_read(bp, 0, 590, 16 if rev else 0, seq, "60M", f"N+e{status},3,0,10,5;N+b{status},1,1;", ml_bytes(6, 4))


def b_iupac(bp):
    seq = list(fwd_read(880, 60))
    seq[20], seq[21] = "R", "Y"                                  # read letters outside ACGTN count as class A (Q5)
    seq = "".join(seq)
    mm, cnt = mm_for(seq.replace("R", "A").replace("Y", "A"), "A", "a", ".", 2)
    add_read(bp, 0, 880, 0, seq, "60M", mm, ml_bytes(cnt))
    mm, cnt = mm_for(seq, "C", "m", "?")
    add_read(bp, 0, 880, 0, seq, "60M", mm, ml_bytes(cnt))


def b_lower_and_chebi(bp):
    seq = fwd_read(100, 200)
    mm, cnt = mm_for(seq, "C", "m", "?")
    add_read(bp, 0, 100, 0, seq, "200M", "c" + mm[1:], ml_bytes(cnt))
    mm2, cnt2 = mm_for(seq, "C", "21839", ".", 2, rev=True)
    add_read(bp, 0, 100, 16, seq, "200M", mm2.replace("+", "-"), ml_bytes(cnt2))
    mmU, cntU = mm_for(seq, "T", "17802", "?", 3)
    add_read(bp, 0, 100, 0, seq, "200M", "U" + mmU[1:], ml_bytes(cntU))


def b_empty_and_odd(bp):
    seq = fwd_read(120, 101)                                      # odd length
    add_read(bp, 0, 120, 0, seq, "101M", "", b"")                  # empty MM: nothing to do
    add_read(bp, 0, 120, 0, seq, "101M", "C+m?;", b"")             # block with no calls
    add_read(bp, 0, 120, 0, seq, "101M", "C+m.;", b"")             # '.' with no calls: every C implicit
    mm, cnt = mm_for(seq, "C", "m", "?")
    add_read(bp, 0, 120, 0, seq, "101M", mm[:-1], ml_bytes(cnt))   # unterminated last block
    add_read(bp, 0, 120, 0, seq, "101M", "C+m" + mm[4:], ml_bytes(cnt))   # no status flag -> '.'
    add_read(bp, 0, 120, 0, seq, "50M51S", "C+x?,0;" + mm, ml_bytes(cnt + 1))   # unrequired code still consumes ML


def b_haps(bp):
    for hp in (0, 1, 2, 3, 4, 7, 255):
        build_basic(bp, rev=hp % 2 == 1, hp=hp, start=90 + hp, n=200)


def b_many_blocks(bp):
    seq = fwd_read(90, 500)
    for rev in (False, True):
        mm, ml = "", b""
        for k in range(40):                                       # > 32 blocks: exercises the multi-round path
            code = "m" if k % 2 == 0 else "h"
            m, c = mm_for(seq, "C", code, "?" if k % 3 else ".", 5 + k % 4, rev)
            mm += m; ml += ml_bytes(c, k)
        add_read(bp, 0, 90, 16 if rev else 0, seq, "500M", mm, ml)


def b_long_read(bp):
    rng = random.Random(3)
    seq = "".join(rng.choice("ACGT") for _ in range(70000))       # longer than every shared-memory capacity
    ops, left = [], 70000 - 200
    while left > 0:
        k = min(left, rng.randint(1, 30)); ops.append(f"{k}M"); left -= k
        if left > 0:
            ops.append(f"{rng.randint(1, 3)}D")
    cig = "100S" + "".join(ops) + "100S"
    for rev in (False, True):
        mm, cnt = mm_for(seq, "C", "m", ".", 3, rev)
        add_read(bp, 1, 1000, 16 if rev else 0, seq, cig, mm, ml_bytes(cnt, 11))


LONG_REF = "".join(random.Random(5).choice("ACGT") for _ in range(90000))


def b_cigar_len_boundaries(bp):
    """Op lengths on the boundaries of the CIGAR byte form (include/minimod_cuda.h): 14 | 15 (first escape list) ... 269 | 270
    (second list), in every op kind, more than 32 ops per read so that the ranks into the escape lists carry over steps."""
    lens = [14, 15, 16, 1, 269, 270, 271, 2, 15, 255, 256, 14, 270, 3, 269, 15]
    ops, q, r = [], 0, 0
    for rep in range(3):
        for k, n in enumerate(lens):
            ops.append(f"{n}M"); q += n; r += n
            kind = "IDN"[(k + rep) % 3]
            m = lens[(k * 7 + rep) % len(lens)]
            ops.append(f"{m}{kind}")
            if kind == "I":
                q += m
            else:
                r += m
    seq = "".join(random.Random(11).choice("ACGT") for _ in range(q))
    assert r + 2000 < len(LONG_REF)
    for rev in (False, True):
        for status in ("?", "."):
            mm, cnt = mm_for(seq, "C", "m", status, 2, rev)
            add_read(bp, 1, 2000, 16 if rev else 0, seq, "".join(ops), mm, ml_bytes(cnt, 13))

def b_huge_ops(bp):
    """CIGAR lengths >= 2^26: the warp kernels' plain prefix scans leave such reads to k_decode (saturating scans).
    Forward strand only: a CIGAR longer than the read on a reverse read is fatal in the reference (its reversed walk starts
    with the clip) and a documented deviation here."""
    seq = fwd_read(90, 100)
    mm, cnt = mm_for(seq, "C", "m", "?")
    add_read(bp, 0, 90, 0, seq, "100M100000000S", mm, ml_bytes(cnt))      # clip far beyond the read: not checked (src/mod.c:841-861)
    add_read(bp, 0, 90, 0, seq, "100M", mm, ml_bytes(cnt))


CASES = [
    case("fwd_cpg", lambda bp: build_basic(bp)),
    case("rev_cpg", lambda bp: build_basic(bp, rev=True)),
    case("dot_status", lambda bp: [build_basic(bp, status=".", every=2), build_basic(bp, rev=True, status=".", every=3)]),
    case("two_blocks_mh", b_two_blocks, codes="m,h", thresh="0.8,0.5"),
    case("only_h", b_two_blocks, codes="h"),
    case("combined_code_suffix", b_combined, codes="m,h,mh"),
    case("combined_wildcard", b_combined, codes="*"),
    case("all_context", lambda bp: [build_basic(bp), build_basic(bp, rev=True, status=".")], codes="m[*]"),
    case("context_C", lambda bp: [build_basic(bp), build_basic(bp, rev=True)], codes="m[C]"),
    case("context_long", lambda bp: [build_basic(bp), build_basic(bp, rev=True)], codes="m[CGNACGT]"),
    case("context_N", b_n_base, codes="e[N],b[N]", thresh="0.5"),
    case("n_base_T_context", b_n_base, codes="e,b", thresh="0.5"),
    case("n_base_all", b_n_base, codes="e[*],b[*]"),
    case("cigar_mix", b_cigar_mix),
    case("cigar_mix_insertions", b_cigar_mix, insertions=True),
    case("leading_insertion", b_leading_ins, insertions=True),
    case("iupac", b_iupac, codes="a[A],m[CG]"),
    case("iupac_all", b_iupac, codes="*"),
    case("lowercase_chebi_U", b_lower_and_chebi, codes="m[C],21839[*],17802[T]"),
    case("empty_and_odd", b_empty_and_odd),
    case("empty_and_odd_all", b_empty_and_odd, codes="m[*]", insertions=True),
    case("haplotypes_dense_and_sparse", b_haps, haplotypes=True),
    case("haplotypes_insertions", b_cigar_mix, haplotypes=True, insertions=True),
    case("many_blocks", b_many_blocks, codes="m,h"),
    case("long_read_scratch", b_long_read, contigs="long", codes="m[*]"),
    case("long_read_insertions", b_long_read, contigs="long", codes="m[C]", insertions=True),
    case("wild_dense_codes_overflow", b_lower_and_chebi, codes="*", opts=dict(dense_codes=1)),
    case("huge_cigar_ops", b_huge_ops),
    case("cigar_len_boundaries", b_cigar_len_boundaries, contigs="long", codes="m[*]"),
    case("cigar_len_boundaries_insertions", b_cigar_len_boundaries, contigs="long", codes="m[C]", insertions=True),
    case("no_reads", lambda bp: None),
]


def contigs_of(c):
    if c.get("contigs") == "long":
        return [("ctgA", REF), ("ctgL", LONG_REF)]
    return [("ctgA", REF), ("ctgB", REF2)]


def run_case(lib, c):
    for sub in c["subtools"]:
        p = Pair(lib, sub, contigs_of(c), c["codes"], c["thresh"] if sub == "freq" else None, c["insertions"], c["haplotypes"],
                 max_reads=64, max_bytes=1 << 20, **c["opts"])
        try:
            c["build"](p.batch)
            rc, msg = p.run_device()
            assert rc == 0, msg
            orc, omsg = p.run_oracle()
            assert orc == 0, omsg
            if sub == "freq":
                assert p.device_freq() == p.oracle_freq()
            else:
                assert p.device_view() == p.oracle_view()
        finally:
            p.close()


def f_huge_refskip(bp):
    seq = fwd_read(90, 100)
    mm, cnt = mm_for(seq, "C", "m", "?")
    build_basic(bp)
    add_read(bp, 0, 90, 0, seq, "50M100000000N50M", mm, ml_bytes(cnt))                      # lands far past the contig end


def f_huge_leading_clip(bp):
    seq = fwd_read(90, 100)
    mm, cnt = mm_for(seq, "C", "m", "?")
    add_read(bp, 0, 90, 0, seq, "70000000S100M", mm, ml_bytes(cnt))                         # the M op starts beyond the read


def f_hardclip(bp):
    seq = fwd_read(90, 100)
    mm, cnt = mm_for(seq, "C", "m", "?")
    build_basic(bp)
    add_read(bp, 0, 90, 0, seq, "10H100M", mm, ml_bytes(cnt))


def f_bad_op(bp):
    seq = fwd_read(90, 100)
    mm, cnt = mm_for(seq, "C", "m", "?")
    add_read(bp, 0, 90, 0, seq, "50M2P50M", mm, ml_bytes(cnt))


def f_short_ml(bp):
    seq = fwd_read(90, 300)
    mm, cnt = mm_for(seq, "C", "m", "?")
    add_read(bp, 0, 90, 0, seq, "300M", mm, ml_bytes(cnt - 1))


def f_no_ml(bp):
    seq = fwd_read(90, 300)
    mm, cnt = mm_for(seq, "C", "m", "?")
    add_read(bp, 0, 90, 0, seq, "300M", mm, b"")                 # Q13


def f_bad_base(bp):
    add_read(bp, 0, 90, 0, fwd_read(90, 50), "50M", "X+m?,1;", b"\x10")


def f_bad_strand(bp):
    add_read(bp, 0, 90, 0, fwd_read(90, 50), "50M", "C*m?,1;", b"\x10")


def f_empty_block(bp):
    add_read(bp, 0, 90, 0, fwd_read(90, 50), "50M", "C+m?,1;;", b"\x10")


def f_mixed_code(bp):
    add_read(bp, 0, 90, 0, fwd_read(90, 50), "50M", "C+m5?,1;", b"\x10")


def f_rank_overflow(bp):
    add_read(bp, 0, 90, 0, fwd_read(90, 50), "50M", "C+m?,1,500;", b"\x10\x20")


def f_unknown_contig(bp):
    build_basic(bp)
    seq = fwd_read(90, 50)
    mm, cnt = mm_for(seq, "C", "m", "?")
    add_read(bp, 2, 90, 0, seq, "50M", mm, ml_bytes(cnt))        # contig 2 has no reference


def f_past_contig_end(bp):
    seq = fwd_read(3900, 100)
    mm, cnt = mm_for(seq, "C", "m", "?")
    add_read(bp, 0, 3950, 0, seq, "100M", mm, ml_bytes(cnt))


FATAL = [dict(id=f.__name__[2:], build=f) for f in
         (f_hardclip, f_bad_op, f_short_ml, f_no_ml, f_bad_base, f_bad_strand, f_empty_block, f_mixed_code, f_rank_overflow,
          f_unknown_contig, f_past_contig_end, f_huge_refskip, f_huge_leading_clip)]


def run_fatal(lib, c):
    p = Pair(lib, "freq", [("ctgA", REF), ("ctgB", REF2), ("ctgC", None)], "m", None, max_reads=64, max_bytes=1 << 20)
    try:
        c["build"](p.batch)
        rc, msg = p.run_device()
        orc, omsg = p.run_oracle()
        assert orc != 0, "oracle accepted an input the reference treats as fatal"
        assert rc == -4, (rc, msg)          # MMC_EREAD
        assert "read #" in msg
    finally:
        p.close()
