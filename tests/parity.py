"""Record-level comparison helpers: CUDA (or emulated) C ABI vs the CPU oracle on identical batches."""
import ctypes as C

import numpy as np

import oracle_port
from minimod_b200 import _native as N


def make_mods(codes, thresh, subtool):
    host = N.load_host()
    err = C.create_string_buffer(1024)
    mods = (N.MmcMod * N.MMC_MAX_MODS)()
    n = host.mmh_parse_mods(codes.encode(), (thresh or "").encode(), subtool, mods, N.MMC_MAX_MODS, err, 1024)
    assert n > 0, err.value
    return mods, n


class Pair:
    """One device context + one oracle context configured identically, over given contigs (name, seq)."""

    def __init__(self, lib, subtool, contigs, codes="m", thresh=None, insertions=False, haplotypes=False,
                 max_reads=4096, max_bytes=64 << 20, **opts):
        self.lib, self.O = lib, oracle_port.lib()
        self.subtool = N.MMC_FREQ if subtool == "freq" else N.MMC_VIEW
        self.mods, self.n_mods = make_mods(codes, thresh, self.subtool)
        self.contigs = contigs
        # entries: (name, sequence) or (name, None, length) for a header contig whose reference is not loaded here
        contigs = [(c[0], c[1], len(c[1]) if c[1] is not None else (c[2] if len(c) > 2 else 1000)) for c in contigs]
        names = (C.c_char_p * len(contigs))(*[c[0].encode() for c in contigs])
        lens = (C.c_uint32 * len(contigs))(*[c[2] for c in contigs])
        o = N.MmcOpts()
        o.struct_size = C.sizeof(N.MmcOpts)
        o.subtool, o.n_mods, o.mods = self.subtool, self.n_mods, self.mods
        o.insertions, o.haplotypes, o.n_slots = int(insertions), int(haplotypes), 1
        o.max_reads, o.max_bytes = max_reads, max_bytes
        for k, v in opts.items():
            setattr(o, k, v)
        self.ctx = C.c_void_p()
        assert lib.mmc_create(C.byref(self.ctx), C.byref(o), len(contigs), names, lens) == 0, lib.mmc_strerror(None)
        th = (C.c_double * self.n_mods)(*oracle_port.parse_thresholds(thresh, self.n_mods))
        self.octx = self.O.oracle_create(self.subtool, self.n_mods, self.mods, th, int(insertions), int(haplotypes), len(contigs), lens)
        for tid, (_, s, _n) in enumerate(contigs):
            if s is None:
                continue
            sb = s if isinstance(s, bytes) else s.encode()
            assert lib.mmc_ref_add(self.ctx, tid, sb, len(sb)) == 0, lib.mmc_strerror(self.ctx)
            assert self.O.oracle_ref_add(self.octx, tid, sb, len(sb)) == 0
        assert lib.mmc_ref_commit(self.ctx) == 0
        self.batch = C.POINTER(N.MmcBatch)()
        assert lib.mmc_batch_acquire(self.ctx, C.byref(self.batch)) == 0

    def close(self):
        self.lib.mmc_batch_release(self.ctx, self.batch)
        self.lib.mmc_destroy(self.ctx)
        self.O.oracle_destroy(self.octx)

    def run_device(self):
        rc = self.lib.mmc_batch_submit(self.ctx, self.batch)
        if rc == 0:
            rc = self.lib.mmc_batch_wait(self.ctx, self.batch)
        return rc, (self.lib.mmc_strerror(self.ctx) or b"").decode()

    def run_oracle(self):
        rc = self.O.oracle_process_batch(self.octx, self.batch)
        return rc, self.O.oracle_strerror(self.octx).decode()

    def device_freq(self):
        recs, n = C.POINTER(N.MmcFreqRec)(), C.c_uint64()
        assert self.lib.mmc_freq_finalize(self.ctx, C.byref(recs), C.byref(n)) == 0, self.lib.mmc_strerror(self.ctx)
        a = np.zeros(0, dtype=N.FREQ_DTYPE)
        if n.value:
            a = np.frombuffer((N.MmcFreqRec * n.value).from_address(C.addressof(recs.contents)), dtype=N.FREQ_DTYPE).copy()
        return canon_freq(a, [self.lib.mmc_code_name(self.ctx, i) for i in range(256)])

    def oracle_freq(self):
        recs, n = C.POINTER(N.MmcFreqRec)(), C.c_uint64()
        self.O.oracle_freq_records(self.octx, C.byref(recs), C.byref(n))
        a = np.zeros(0, dtype=N.FREQ_DTYPE)
        if n.value:
            a = np.frombuffer((N.MmcFreqRec * n.value).from_address(C.addressof(recs.contents)), dtype=N.FREQ_DTYPE).copy()
        return canon_freq(a, [self.O.oracle_code_name(self.octx, i) for i in range(256)])

    def device_view(self):
        recs, n = C.POINTER(N.MmcViewRec)(), C.c_uint64()
        assert self.lib.mmc_view_fetch(self.ctx, self.batch, C.byref(recs), C.byref(n)) == 0, self.lib.mmc_strerror(self.ctx)
        a = np.zeros(0, dtype=N.VIEW_DTYPE)
        if n.value:
            a = np.frombuffer((N.MmcViewRec * n.value).from_address(C.addressof(recs.contents)), dtype=N.VIEW_DTYPE).copy()
        return canon_view(a, [self.lib.mmc_code_name(self.ctx, i) for i in range(256)])

    def oracle_view(self):
        recs, n = C.POINTER(N.MmcViewRec)(), C.c_uint64()
        self.O.oracle_view_records(self.octx, C.byref(recs), C.byref(n))
        a = np.zeros(0, dtype=N.VIEW_DTYPE)
        if n.value:
            a = np.frombuffer((N.MmcViewRec * n.value).from_address(C.addressof(recs.contents)), dtype=N.VIEW_DTYPE).copy()
        return canon_view(a, [self.O.oracle_code_name(self.octx, i) for i in range(256)])


def canon_freq(a, names):
    """Sorted list of tuples keyed by code *string* (the two sides number codes independently)."""
    return sorted((int(r["tid"]), int(r["pos"]), int(r["strand"]), (names[r["code"]] or b"").decode(), int(r["ins_offset"]),
                   int(r["hap"]), int(r["n_called"]), int(r["n_mod"])) for r in a)


def canon_view(a, names):
    return sorted((int(r["read"]), int(r["ref_pos"]), int(r["read_pos"]), (names[r["code"]] or b"").decode(), int(r["ins_offset"]),
                   int(r["mod_prob"]), int(r["strand"]), int(r["hp"])) for r in a)
