"""Hand-built reads for edge-case tests: write BAM-like fields straight into an mmc_batch_t."""
import ctypes as C
import struct

NT16 = "=ACMGRSVTWYHKDBN"
OPS = "MIDNSHP=XB"


def up16(x):
    return (x + 15) & ~15


def add_read(bp, tid, pos, flag, seq, cigar, mm, ml, hp=0):
    """bp: POINTER(MmcBatch).  cigar: string like '10M2I5M'.  ml: bytes or None."""
    b = bp.contents
    i = b.n_reads
    assert i < b.max_reads
    ops, num = [], ""
    for ch in cigar:
        if ch.isdigit():
            num += ch
        else:
            ops.append((int(num) << 4) | OPS.index(ch)); num = ""
    c0, s0, m0, l0 = up16(b.cigar_used * 4) // 4, up16(b.seq_used), up16(b.mm_used), up16(b.ml_used)
    if b.cigar_packing == 8:        # transport form (include/minimod_cuda.h): a byte per op + two escape lists
        l1 = [min(255, (w >> 4) - 15) for w in ops if (w >> 4) >= 15]
        l2 = [w >> 4 for w in ops if (w >> 4) >= 15 + 255]
        pad = lambda x: x + bytes(-len(x) % 4)
        blob = struct.pack("<I", len(l1)) + pad(bytes((w & 15) | (min(w >> 4, 15) << 4) for w in ops)) + pad(bytes(l1)) + struct.pack(f"<{len(l2)}I", *l2)
        g0 = up16(b.cig8_used)
        assert g0 + len(blob) <= b.cig8_cap
        C.memmove(C.addressof(b.cig8.contents) + g0, blob, len(blob))
        b.cig8_off[i] = g0
        b.cig8_used = g0 + len(blob)
    else:
        for k, w in enumerate(ops):
            b.cigar[c0 + k] = w
    nb = (len(seq) + 1) // 2
    if b.seq_packing == 2:          # transport form (include/minimod_cuda.h): 2 bits per base + exceptions
        code = {1: 0, 2: 1, 4: 2, 8: 3}
        nibs = [NT16.index(ch) for ch in seq] + ([0] if len(seq) & 1 else [])
        for k in range((len(seq) + 3) // 4):
            v = 0
            for j in range(4):
                n = nibs[4 * k + j] if 4 * k + j < len(nibs) else 1
                v |= code.get(n, 0) << (6 - 2 * j)
            b.seq2[s0 // 2 + k] = v
        for k, n in enumerate(nibs):
            if n not in code:
                assert b.seq_exc_used < b.seq_exc_cap
                b.seq_exc[b.seq_exc_used] = ((2 * s0 + k) << 4) | n
                b.seq_exc_used += 1
    else:
        for k in range(nb):
            hi = NT16.index(seq[2 * k])
            lo = NT16.index(seq[2 * k + 1]) if 2 * k + 1 < len(seq) else 0
            b.seq4[s0 + k] = (hi << 4) | lo
    mmb = mm.encode()
    C.memmove(C.addressof(b.mm.contents) + m0, mmb, len(mmb))
    mlb = bytes(ml or b"")
    for k, v in enumerate(mlb):
        b.ml[l0 + k] = v
    b.tid[i], b.pos[i], b.flag[i], b.hp[i] = tid, pos, flag, hp
    b.l_seq[i], b.n_cigar[i], b.mm_len[i], b.ml_len[i] = len(seq), len(ops), len(mmb), len(mlb)
    b.cigar_off[i], b.seq_off[i], b.mm_off[i], b.ml_off[i] = c0, s0, m0, l0
    b.cigar_used, b.seq_used, b.mm_used, b.ml_used = c0 + len(ops), s0 + nb, m0 + len(mmb), l0 + len(mlb)
    b.n_reads = i + 1
