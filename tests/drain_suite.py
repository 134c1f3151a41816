"""mmc_freq_drain(): rows of finished positions leave while later batches are in flight.  The drained rows followed by
what mmc_freq_finalize() still holds must be the table of one finalize at the end -- byte for byte, in the same order.
Shared by the CPU suite (SIMT-emulation build) and the GPU suite."""
import ctypes as C

import numpy as np

from minimod_b200 import _native as N
from minimod_b200.synth import CONFIG_ARGS, Synth
from parity import make_mods


class Job:
    """One context over a two-contig synthetic job, the reads split into `chunks` coordinate-ordered batches."""

    def __init__(self, lib, config, contig_len=120000, coverage=3.0, chunks=5, **opts):
        self.lib = lib
        self.s = s = Synth(config, contigs=(("chrS", contig_len), ("chrT", contig_len // 2)), coverage=coverage)
        ca = CONFIG_ARGS[config]
        self.mods, n_mods = make_mods(ca["mod_codes"], ca.get("mod_thresh"), N.MMC_FREQ)
        names = (C.c_char_p * 2)(*s.names)
        lens = (C.c_uint32 * 2)(*s.lens)
        o = N.MmcOpts()
        o.struct_size = C.sizeof(N.MmcOpts)
        o.subtool, o.n_mods, o.mods = N.MMC_FREQ, n_mods, self.mods
        o.insertions, o.haplotypes, o.n_slots = int(bool(ca.get("insertions"))), int(bool(ca.get("haplotypes"))), chunks
        per = (s.n_reads + chunks - 1) // chunks
        o.max_reads, o.max_bytes = per + 8, max(8 << 20, per * 60000 * 2)
        for k, v in opts.items():
            setattr(o, k, v)
        self.ctx = C.c_void_p()
        assert lib.mmc_create(C.byref(self.ctx), C.byref(o), 2, names, lens) == 0, lib.mmc_strerror(None)
        for tid in range(2):
            p, n = s.ref(tid)
            assert lib.mmc_ref_add(self.ctx, tid, C.cast(p, C.c_char_p), n) == 0, lib.mmc_strerror(self.ctx)
        assert lib.mmc_ref_commit(self.ctx) == 0
        self.held = []
        for k in range(chunks):
            b = C.POINTER(N.MmcBatch)()
            assert lib.mmc_batch_acquire(self.ctx, C.byref(b)) == 0
            cnt = max(0, min(per, s.n_reads - k * per))
            got, _ = s.fill(b, k * per, cnt, 2)
            assert got == cnt
            self.held.append(b)

    def chk(self, rc):
        assert rc == 0, (rc, self.lib.mmc_strerror(self.ctx))

    def first(self, k):
        b = self.held[k].contents
        return int(b.tid[0]), int(b.pos[0])

    def rows(self, fn, *a):
        recs, n = C.POINTER(N.MmcFreqRec)(), C.c_uint64()
        rc = fn(self.ctx, *a, C.byref(recs), C.byref(n))
        if rc != 0:
            return rc, None
        if not n.value:
            return 0, np.zeros(0, dtype=N.FREQ_DTYPE)
        return 0, np.frombuffer((N.MmcFreqRec * n.value).from_address(C.addressof(recs.contents)), dtype=N.FREQ_DTYPE).copy()

    def table(self, order=None):
        """every batch, one finalize at the end"""
        self.chk(self.lib.mmc_freq_reset(self.ctx))
        for k in (order or range(len(self.held))):
            self.chk(self.lib.mmc_batch_submit(self.ctx, self.held[k]))
        rc, t = self.rows(self.lib.mmc_freq_finalize)
        self.chk(rc)
        return t

    def close(self):
        for b in self.held:
            self.lib.mmc_batch_release(self.ctx, b)
        self.lib.mmc_destroy(self.ctx)
        self.s.close()


def check_drain_equals_finalize(lib, config, lag=1, expect_early=True, **kw):
    j = Job(lib, config, **kw)
    try:
        want = j.table()
        assert len(want) > 100
        j.chk(lib.mmc_freq_reset(j.ctx))
        parts = []
        for k in range(len(j.held)):
            j.chk(lib.mmc_batch_submit(j.ctx, j.held[k]))
            if k >= lag and j.held[k - lag + 1].contents.n_reads:
                rc, part = j.rows(lib.mmc_freq_drain, *j.first(k - lag + 1))     # batches < k-lag+1 are all that start before it
                j.chk(rc)
                parts.append(part)
        rc, rest = j.rows(lib.mmc_freq_finalize)
        j.chk(rc)
        early = sum(len(p) for p in parts)
        got = np.concatenate(parts + [rest])
        assert got.tobytes() == want.tobytes(), (len(got), len(want))
        if expect_early:
            assert early > len(want) // 3, (early, len(want))        # the drains really carried most of the table
        else:
            assert early == 0
        # finalize is repeatable (the remainder again); after a reset the job can run again
        rc, again = j.rows(lib.mmc_freq_finalize)
        j.chk(rc)
        assert again.tobytes() == rest.tobytes()
        assert j.table().tobytes() == want.tobytes()
        return len(want), early
    finally:
        j.close()


def check_drain_order_violation(lib, config=2, **kw):
    """A batch that starts before the drained watermark: finalize reports MMC_EORDER, and after mmc_freq_undrain() the
    complete table comes back (drains never clear counts)."""
    j = Job(lib, config, **kw)
    try:
        want = j.table()
        j.chk(lib.mmc_freq_reset(j.ctx))
        n = len(j.held)
        j.chk(lib.mmc_batch_submit(j.ctx, j.held[1]))
        rc, part = j.rows(lib.mmc_freq_drain, *j.first(2))
        j.chk(rc)
        assert len(part) > 0
        j.chk(lib.mmc_batch_submit(j.ctx, j.held[0]))                  # breaks the promise
        for k in range(2, n):
            j.chk(lib.mmc_batch_submit(j.ctx, j.held[k]))
        rc, nothing = j.rows(lib.mmc_freq_drain, *j.first(n - 1))      # no more early rows once the order is broken
        assert rc == 0 and len(nothing) == 0
        rc, _ = j.rows(lib.mmc_freq_finalize)
        assert rc == N.MMC_EORDER, rc
        j.chk(lib.mmc_freq_undrain(j.ctx))
        rc, full = j.rows(lib.mmc_freq_finalize)
        j.chk(rc)
        assert full.tobytes() == want.tobytes()
    finally:
        j.close()
