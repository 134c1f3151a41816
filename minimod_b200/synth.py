"""Synthetic workloads of the BASELINE.json shapes (generator: minimod_b200/host/synth.cpp)."""
import ctypes as C

from . import _native as N

SEED0 = 20261017          # seed = SEED0 + config id
CHR22_LEN = 50818468

CONFIG_ARGS = {           # the reference command line each synthetic config is run with
    2: dict(mod_codes="m[CG]", mod_thresh="0.8", bedmethyl=True),
    3: dict(mod_codes="m[CG],h[CG]", mod_thresh="0.8,0.7", insertions=True),
    4: dict(mod_codes="m[*],a[A]", haplotypes=True),
    5: dict(mod_codes="m[CG]", mod_thresh="0.8"),
    6: dict(mod_codes="m[CG],h[CG]", mod_thresh="0.8,0.7", insertions=True),
    7: dict(mod_codes="m[CG]", mod_thresh="0.8"),       # ultra-long ONT reads: > 65535 CIGAR ops travel in a CG:B,I tag
}


def cli_args(config):
    a = CONFIG_ARGS[config]
    out = ["-c", a["mod_codes"]]
    if a.get("mod_thresh"):
        out += ["-m", a["mod_thresh"]]
    if a.get("bedmethyl"):
        out.append("-b")
    if a.get("insertions"):
        out.append("--insertions")
    if a.get("haplotypes"):
        out.append("--haplotypes")
    return out


class Synth:
    def __init__(self, config, contigs=(("chr22", CHR22_LEN),), coverage=0.0, seed=None, gids=None):
        """gids: job-wide ids of `contigs` (default 0..n-1).  A contig's sequence and reads depend on (seed, gid) only,
        so the union over any split of a contig table into instances is the same job (contig sharding)."""
        self.host = N.load_host()
        self.config = config
        self.names = [n.encode() for n, _ in contigs]
        self.lens = [l for _, l in contigs]
        names = (C.c_char_p * len(contigs))(*self.names)
        lens = (C.c_uint32 * len(contigs))(*self.lens)
        g = (C.c_uint32 * len(contigs))(*(gids if gids is not None else range(len(contigs))))
        self.h = self.host.mmh_synth_new2(config, SEED0 + config if seed is None else seed, len(contigs), names, lens, g, coverage)
        self.n_reads = self.host.mmh_synth_n_reads(self.h)

    def ref(self, tid):
        n = C.c_uint64()
        p = self.host.mmh_synth_ref(self.h, tid, C.byref(n))
        return p, n.value

    def write_fasta(self, path):
        assert self.host.mmh_synth_write_fasta(self.h, path.encode()) == 0

    def write_bam(self, path, first=0, count=None, threads=8):
        st = N.MmhSynthStats()
        count = self.n_reads - first if count is None else count
        assert self.host.mmh_synth_write_bam(self.h, path.encode(), first, count, threads, C.byref(st)) == 0
        return {k: getattr(st, k) for k, _ in N.MmhSynthStats._fields_}

    def fill(self, batch, first, count, threads=8, stats=None):
        st = stats if stats is not None else N.MmhSynthStats()
        n = self.host.mmh_synth_fill(self.h, batch, first, count, threads, C.byref(st))
        return n, st

    def close(self):
        if self.h:
            self.host.mmh_synth_free(self.h)
            self.h = None
