"""Python mirror of the reference's batch operator interface, over the C ABI.

Reference (warp9seq/minimod v0.5.0)            here
  init_core()            src/minimod.c:51        Core(...)
  load_ref()+load_ref_contexts() src/ref.c:46    Core.load_ref(fasta)
  init_db()+load_db()    src/minimod.c:164,235   Core.load_db()      -> batch handle (pinned SoA)
  process_db()+merge_db() src/minimod.c:344,373  Core.process_db(b)  -> async H2D + kernels
  output_core()          src/minimod.c:388       Core.output_core()  -> freq text / records
  output_db()            src/minimod.c:354       Core.output_db(b)   -> view text of one batch
Errors the reference reports with ERROR()+exit(1) are raised as MinimodError.
"""
import ctypes as C
import os
import tempfile

import numpy as np

from . import _native as N


class MinimodError(RuntimeError):
    pass


class Core:
    def __init__(self, subtool, bam, mod_codes="m", mod_thresh=None, insertions=False, haplotypes=False,
                 batch_size=512, max_bytes=20 * 1000 * 1000, allow_secondary=False, skip_supplementary=False,
                 device=0, n_slots=3, dense_haps=0, dense_codes=0, sparse_capacity=0, seq_packing=2, cigar_packing=8, lib=None):
        self.lib = lib if lib is not None else N.load_cuda()
        self.host = N.load_host()
        self.subtool = N.MMC_FREQ if subtool in ("freq", N.MMC_FREQ) else N.MMC_VIEW
        self.insertions, self.haplotypes = int(bool(insertions)), int(bool(haplotypes))
        err = C.create_string_buffer(1024)
        self.mods = (N.MmcMod * N.MMC_MAX_MODS)()
        n = self.host.mmh_parse_mods((mod_codes or "m").encode(), (mod_thresh or "").encode(), self.subtool,
                                     self.mods, N.MMC_MAX_MODS, err, 1024)
        if n < 0:
            raise MinimodError(err.value.decode())
        self.n_mods = n
        self.bam = self.host.mmh_bam_open(os.fsencode(bam), err, 1024)
        if not self.bam:
            raise MinimodError(err.value.decode())
        nt = self.host.mmh_bam_n_targets(self.bam)
        self.contig_names = [self.host.mmh_bam_target_name(self.bam, i) for i in range(nt)]
        self.contig_lens = [self.host.mmh_bam_target_len(self.bam, i) for i in range(nt)]
        self.loader = self.host.mmh_loader_new(self.bam, batch_size, int(max_bytes), int(allow_secondary),
                                               int(skip_supplementary), int(self.subtool == N.MMC_VIEW))
        o = N.MmcOpts()
        o.struct_size = C.sizeof(N.MmcOpts)
        o.subtool, o.n_mods, o.mods = self.subtool, n, self.mods
        o.insertions, o.haplotypes, o.device, o.n_slots = self.insertions, self.haplotypes, device, n_slots
        o.max_reads, o.max_bytes = batch_size, int(max_bytes)
        o.dense_haps, o.dense_codes, o.sparse_capacity = dense_haps, dense_codes, sparse_capacity
        o.seq_packing = seq_packing            # SEQ crosses PCIe at 2 bits per base (include/minimod_cuda.h); 4 = BAM nibbles
        o.cigar_packing = cigar_packing        # CIGARs cross PCIe at a byte per op + escapes; 32 = BAM words
        names = (C.c_char_p * max(nt, 1))(*self.contig_names)
        lens = (C.c_uint32 * max(nt, 1))(*self.contig_lens)
        ctx = C.c_void_p()
        if self.lib.mmc_create(C.byref(ctx), C.byref(o), nt, names, lens) != N.MMC_OK:
            raise MinimodError(self.lib.mmc_strerror(None).decode())
        self.ctx = ctx
        self.stats = dict(total_reads=0, total_bytes=0, processed_reads=0, processed_bytes=0, ml_entries=0, bases=0)
        self._more = True

    # ---- helpers
    def _check(self, rc):
        if rc != N.MMC_OK:
            raise MinimodError(self.lib.mmc_strerror(self.ctx).decode().split("\x1f")[0])

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.mmc_destroy(self.ctx)
            self.ctx = None
        if getattr(self, "loader", None):
            self.host.mmh_loader_free(self.loader)
            self.loader = None
        if getattr(self, "bam", None):
            self.host.mmh_bam_close(self.bam)
            self.bam = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- reference
    def load_ref(self, fasta):
        err = C.create_string_buffer(1024)
        fa = self.host.mmh_fasta_load(os.fsencode(fasta), err, 1024)
        if not fa:
            raise MinimodError(err.value.decode())
        try:
            index = {n: i for i, n in enumerate(self.contig_names)}
            for i in range(self.host.mmh_fasta_n(fa)):
                tid = index.get(self.host.mmh_fasta_name(fa, i))
                if tid is None:
                    continue
                seq = C.cast(self.host.mmh_fasta_seq(fa, i), C.c_char_p)
                self.lib.mmc_ref_add(self.ctx, tid, seq, self.host.mmh_fasta_len(fa, i))  # mismatch => contig stays unloaded
            self._check(self.lib.mmc_ref_commit(self.ctx))
        finally:
            self.host.mmh_fasta_free(fa)

    # ---- batches
    def load_db(self):
        """Next batch (None at end of file)."""
        if not self._more:
            return None
        b = C.POINTER(N.MmcBatch)()
        self._check(self.lib.mmc_batch_acquire(self.ctx, C.byref(b)))
        st, err = N.MmhStats(), C.create_string_buffer(1024)
        rc = self.host.mmh_loader_fill(self.loader, b, C.byref(st), err, 1024)
        if rc < 0:
            raise MinimodError(err.value.decode())
        self._more = rc > 0
        s = self.stats
        s["total_reads"] += st.total_reads; s["total_bytes"] += st.total_bytes
        s["processed_reads"] += st.n_recs; s["processed_bytes"] += st.processed_bytes
        s["ml_entries"] += st.ml_entries; s["bases"] += st.bases
        return b

    def process_db(self, b):
        self._check(self.lib.mmc_batch_submit(self.ctx, b))

    def wait_db(self, b):
        self._check(self.lib.mmc_batch_wait(self.ctx, b))

    def free_db(self, b):
        self._check(self.lib.mmc_batch_release(self.ctx, b))

    def code_names(self):
        return [self.lib.mmc_code_name(self.ctx, i) or b"" for i in range(256)]

    # ---- results
    def freq_records(self):
        recs, n = C.POINTER(N.MmcFreqRec)(), C.c_uint64()
        self._check(self.lib.mmc_freq_finalize(self.ctx, C.byref(recs), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=N.FREQ_DTYPE)
        buf = (N.MmcFreqRec * n.value).from_address(C.addressof(recs.contents))
        return np.frombuffer(buf, dtype=N.FREQ_DTYPE).copy()

    def drain_records(self, tid, pos):
        """Rows before the coordinate watermark (tid, pos) that no earlier drain returned (mmc_freq_drain): the caller
        submits no read starting before it from now on.  Copies them out of the library's alternating buffers."""
        recs, n = C.POINTER(N.MmcFreqRec)(), C.c_uint64()
        self._check(self.lib.mmc_freq_drain(self.ctx, int(tid), int(pos), C.byref(recs), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=N.FREQ_DTYPE)
        buf = (N.MmcFreqRec * n.value).from_address(C.addressof(recs.contents))
        return np.frombuffer(buf, dtype=N.FREQ_DTYPE).copy()

    def remaining_records(self, drained):
        """The table after drains: `drained` (list of arrays from drain_records) + what mmc_freq_finalize() still holds.
        If a batch broke the coordinate order after a drain (MMC_EORDER), the drained rows are dropped and the
        complete table is read back instead (counts are never cleared by a drain)."""
        recs, n = C.POINTER(N.MmcFreqRec)(), C.c_uint64()
        rc = self.lib.mmc_freq_finalize(self.ctx, C.byref(recs), C.byref(n))
        if rc == N.MMC_EORDER:
            self._check(self.lib.mmc_freq_undrain(self.ctx))
            return self.freq_records()
        self._check(rc)
        rest = np.zeros(0, dtype=N.FREQ_DTYPE)
        if n.value:
            rest = np.frombuffer((N.MmcFreqRec * n.value).from_address(C.addressof(recs.contents)), dtype=N.FREQ_DTYPE).copy()
        return np.concatenate(list(drained) + [rest]) if drained else rest

    def output_core(self, bedmethyl=False, records=None):
        """freq text exactly as print_freq_header()+print_freq_output() write it."""
        recs, n = C.POINTER(N.MmcFreqRec)(), C.c_uint64()
        if records is not None:
            records = np.ascontiguousarray(records)
            recs, n = C.cast(records.ctypes.data, C.POINTER(N.MmcFreqRec)), C.c_uint64(len(records))
        else:
            self._check(self.lib.mmc_freq_finalize(self.ctx, C.byref(recs), C.byref(n)))
        names = (C.c_char_p * max(1, len(self.contig_names)))(*self.contig_names)
        codes = (C.c_char_p * 256)(*self.code_names())
        with tempfile.NamedTemporaryFile(suffix=".tsv", delete=False) as tf:
            path = tf.name
        try:
            if self.host.mmh_write_freq(path.encode(), int(bedmethyl), self.insertions, self.haplotypes,
                                        len(self.contig_names), names, recs, n.value, 256, codes) != 0:
                raise MinimodError("cannot write " + path)
            with open(path, "rb") as fh:
                return fh.read()
        finally:
            os.unlink(path)

    def view_records(self, b):
        recs, n = C.POINTER(N.MmcViewRec)(), C.c_uint64()
        self._check(self.lib.mmc_view_fetch(self.ctx, b, C.byref(recs), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=N.VIEW_DTYPE)
        buf = (N.MmcViewRec * n.value).from_address(C.addressof(recs.contents))
        return np.frombuffer(buf, dtype=N.VIEW_DTYPE).copy()

    def output_db(self, b, path, append):
        recs, n = C.POINTER(N.MmcViewRec)(), C.c_uint64()
        self._check(self.lib.mmc_view_fetch(self.ctx, b, C.byref(recs), C.byref(n)))
        names = (C.c_char_p * max(1, len(self.contig_names)))(*self.contig_names)
        codes = (C.c_char_p * 256)(*self.code_names())
        if self.host.mmh_write_view(path.encode(), int(append), self.insertions, self.haplotypes, len(self.contig_names),
                                    names, b, self.loader, recs, n.value, 256, codes) != 0:
            raise MinimodError("cannot write " + path)

    def timers(self):
        t = N.MmcTimers()
        self.lib.mmc_get_timers(self.ctx, C.byref(t))
        return {k: getattr(t, k) for k, _ in N.MmcTimers._fields_}


def freq(ref_fa, bam, mod_codes="m", mod_thresh=None, bedmethyl=False, drain=False, **kw):
    """`minimod freq` through the C ABI; returns the stdout bytes the reference would print.
    drain: rows of finished positions leave while later batches are in flight (mmc_freq_drain; coordinate-sorted input,
    anything else falls back to the single read-back at the end)."""
    with Core("freq", bam, mod_codes, mod_thresh, **kw) as core:
        core.load_ref(ref_fa)
        held, drained = [], []
        while True:
            b = core.load_db()
            if b is None:
                break
            first = (int(b.contents.tid[0]), int(b.contents.pos[0])) if b.contents.n_reads else None
            core.process_db(b)
            if drain and first is not None and first[0] >= 0:
                drained.append(core.drain_records(*first))     # every later read starts at or after this batch's first
            held.append(b)
            if len(held) >= 3:
                core.free_db(held.pop(0))
        for b in held:
            core.free_db(b)
        if drain:
            return core.output_core(bedmethyl=bedmethyl, records=core.remaining_records(drained))
        return core.output_core(bedmethyl=bedmethyl)


def view(ref_fa, bam, mod_codes="m", **kw):
    """`minimod view` through the C ABI; returns the stdout bytes the reference would print."""
    with Core("view", bam, mod_codes, None, **kw) as core:
        core.load_ref(ref_fa)
        with tempfile.NamedTemporaryFile(suffix=".tsv", delete=False) as tf:
            path = tf.name
        try:
            first = True
            while True:
                b = core.load_db()
                if b is None:
                    break
                core.process_db(b)
                core.output_db(b, path, append=not first)
                core.free_db(b)
                first = False
            with open(path, "rb") as fh:
                return fh.read()
        finally:
            os.unlink(path)
