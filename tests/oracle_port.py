"""Test-side driver of the CPU oracle (oracle/modcall_oracle.c, built to oracle/_build/liboracle.so).
Reads BAM/FASTA with the product's host library, packs reads into a plain host-memory batch with the
same layout the CUDA path consumes, runs the oracle, and formats with the same writer -- so a text
diff against the reference isolates the oracle's arithmetic."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

from helpers import ROOT
from minimod_b200 import _native as N

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "restatement"], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(path)
        vp, P = C.c_void_p, C.POINTER
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [C.c_int, C.c_int, P(N.MmcMod), P(C.c_double), C.c_int, C.c_int, C.c_int, P(C.c_uint32)]
        L.oracle_destroy.argtypes = [vp]
        L.oracle_ref_add.argtypes = [vp, C.c_int, C.c_char_p, C.c_uint32]
        L.oracle_process_batch.argtypes = [vp, P(N.MmcBatch)]
        L.oracle_freq_records.argtypes = [vp, P(P(N.MmcFreqRec)), P(C.c_uint64)]
        L.oracle_view_records.argtypes = [vp, P(P(N.MmcViewRec)), P(C.c_uint64)]
        L.oracle_code_name.restype = C.c_char_p
        L.oracle_code_name.argtypes = [vp, C.c_int]
        L.oracle_strerror.restype = C.c_char_p
        L.oracle_strerror.argtypes = [vp]
        _lib = L
    return _lib


class HostBatch:
    """An mmc_batch_t backed by numpy arrays (what mmc_batch_acquire() hands out, minus the pinning)."""

    def __init__(self, max_reads, max_bytes):
        self.arrs = {}
        b = N.MmcBatch()
        b.n_reads, b.max_reads = 0, max_reads

        def mk(name, dtype, n, ctype):
            a = np.zeros(n, dtype=dtype)
            self.arrs[name] = a
            setattr(b, name, a.ctypes.data_as(C.POINTER(ctype)))
        for name, dt, ct in (("tid", np.int32, C.c_int32), ("pos", np.int32, C.c_int32), ("l_seq", np.uint32, C.c_uint32),
                             ("n_cigar", np.uint32, C.c_uint32), ("mm_len", np.uint32, C.c_uint32), ("ml_len", np.uint32, C.c_uint32),
                             ("cigar_off", np.uint64, C.c_uint64), ("seq_off", np.uint64, C.c_uint64), ("mm_off", np.uint64, C.c_uint64),
                             ("ml_off", np.uint64, C.c_uint64), ("flag", np.uint16, C.c_uint16), ("hp", np.uint8, C.c_uint8)):
            mk(name, dt, max_reads, ct)
        mk("cigar", np.uint32, max_bytes // 4 + 64, C.c_uint32); b.cigar_cap = max_bytes // 4
        mk("seq4", np.uint8, max_bytes + 64, C.c_uint8); b.seq_cap = max_bytes
        mk("mm", np.uint8, max_bytes + 64, C.c_char); b.mm_cap = max_bytes
        mk("ml", np.uint8, max_bytes + 64, C.c_uint8); b.ml_cap = max_bytes
        self.b = b

    def ptr(self):
        return C.pointer(self.b)


def parse_thresholds(mod_thresh, n_mods):
    vals = [float(x) for x in mod_thresh.split(",")] if mod_thresh else [0.8] * n_mods
    if len(vals) == 1:
        vals = vals * n_mods
    return vals


def run(subtool, ref_fa, bam, mod_codes="m", mod_thresh=None, bedmethyl=False, insertions=False, haplotypes=False,
        batch_size=512, max_bytes=20 * 1000 * 1000, records=False):
    """Oracle output text (bytes) of `minimod <subtool>`; with records=True the raw record arrays + code names."""
    L, host = lib(), N.load_host()
    st = N.MMC_FREQ if subtool == "freq" else N.MMC_VIEW
    err = C.create_string_buffer(1024)
    mods = (N.MmcMod * N.MMC_MAX_MODS)()
    n_mods = host.mmh_parse_mods((mod_codes or "m").encode(), (mod_thresh or "").encode(), st, mods, N.MMC_MAX_MODS, err, 1024)
    assert n_mods > 0, err.value
    th = (C.c_double * n_mods)(*parse_thresholds(mod_thresh, n_mods))
    bamh = host.mmh_bam_open(os.fsencode(bam), err, 1024)
    assert bamh, err.value
    nt = host.mmh_bam_n_targets(bamh)
    names = [host.mmh_bam_target_name(bamh, i) for i in range(nt)]
    lens = (C.c_uint32 * max(1, nt))(*[host.mmh_bam_target_len(bamh, i) for i in range(nt)])
    ctx = L.oracle_create(st, n_mods, mods, th, int(insertions), int(haplotypes), nt, lens)
    fa = host.mmh_fasta_load(os.fsencode(ref_fa), err, 1024)
    assert fa, err.value
    index = {n: i for i, n in enumerate(names)}
    for i in range(host.mmh_fasta_n(fa)):
        tid = index.get(host.mmh_fasta_name(fa, i))
        if tid is not None and host.mmh_fasta_len(fa, i) == lens[tid]:
            assert L.oracle_ref_add(ctx, tid, C.cast(host.mmh_fasta_seq(fa, i), C.c_char_p), host.mmh_fasta_len(fa, i)) == 0
    host.mmh_fasta_free(fa)
    loader = host.mmh_loader_new(bamh, batch_size, int(max_bytes), 0, 0, int(st == N.MMC_VIEW))
    hb = HostBatch(batch_size, int(max_bytes) + 4096)
    cnames = (C.c_char_p * max(1, nt))(*names)
    out_path = tempfile.NamedTemporaryFile(suffix=".tsv", delete=False).name
    view_chunks, first = [], True
    try:
        more = 1
        while more > 0:
            more = host.mmh_loader_fill(loader, hb.ptr(), None, err, 1024)
            assert more >= 0, err.value
            if L.oracle_process_batch(ctx, hb.ptr()) != 0:
                raise RuntimeError(L.oracle_strerror(ctx).decode())
            if st == N.MMC_VIEW:
                recs, n = C.POINTER(N.MmcViewRec)(), C.c_uint64()
                L.oracle_view_records(ctx, C.byref(recs), C.byref(n))
                codes = (C.c_char_p * 256)(*[L.oracle_code_name(ctx, i) for i in range(256)])
                if records:
                    if n.value:
                        a = np.frombuffer((N.MmcViewRec * n.value).from_address(C.addressof(recs.contents)), dtype=N.VIEW_DTYPE).copy()
                        view_chunks.append(a)
                else:
                    assert host.mmh_write_view(out_path.encode(), int(not first), int(insertions), int(haplotypes), nt, cnames,
                                               hb.ptr(), loader, recs, n.value, 256, codes) == 0
                first = False
        code_names = [L.oracle_code_name(ctx, i) for i in range(256)]
        if st == N.MMC_FREQ:
            recs, n = C.POINTER(N.MmcFreqRec)(), C.c_uint64()
            L.oracle_freq_records(ctx, C.byref(recs), C.byref(n))
            if records:
                a = np.zeros(0, dtype=N.FREQ_DTYPE)
                if n.value:
                    a = np.frombuffer((N.MmcFreqRec * n.value).from_address(C.addressof(recs.contents)), dtype=N.FREQ_DTYPE).copy()
                return a, code_names
            codes = (C.c_char_p * 256)(*code_names)
            assert host.mmh_write_freq(out_path.encode(), int(bedmethyl), int(insertions), int(haplotypes), nt, cnames, recs,
                                       n.value, 256, codes) == 0
        elif records:
            return (np.concatenate(view_chunks) if view_chunks else np.zeros(0, dtype=N.VIEW_DTYPE)), code_names
        with open(out_path, "rb") as fh:
            return fh.read()
    finally:
        os.unlink(out_path)
        L.oracle_destroy(ctx)
        host.mmh_loader_free(loader)
        host.mmh_bam_close(bamh)
