"""Synthetic-workload parity (BASELINE configs 2-4 + the '.'-status variant 6) on small contigs:
device records vs the CPU oracle on the same packed batch."""
import ctypes as C

from minimod_b200 import _native as N
from minimod_b200.synth import CONFIG_ARGS, Synth
from parity import Pair


def run_synth(lib, config, contig_len, coverage, sub="freq", threads=4, **opts):
    s = Synth(config, contigs=(("chrS", contig_len), ("chrT", contig_len // 2)), coverage=coverage)
    ca = CONFIG_ARGS[config]
    contigs = []
    for tid in range(2):
        p, n = s.ref(tid)
        contigs.append((s.names[tid].decode(), C.string_at(p, n)))
    pair = Pair(lib, sub, contigs, ca["mod_codes"], ca.get("mod_thresh") if sub == "freq" else None,
                bool(ca.get("insertions")), bool(ca.get("haplotypes")), max_reads=s.n_reads + 8,
                max_bytes=max(8 << 20, int(s.n_reads * 60000 * 2)), **opts)
    try:
        got, st = s.fill(pair.batch, 0, s.n_reads, threads)
        assert got == s.n_reads
        rc, msg = pair.run_device()
        assert rc == 0, msg
        orc, omsg = pair.run_oracle()
        assert orc == 0, omsg
        if sub == "freq":
            d, o = pair.device_freq(), pair.oracle_freq()
        else:
            d, o = pair.device_view(), pair.oracle_view()
        assert len(d) == len(o) and d == o
        return len(d), st
    finally:
        pair.close()
        s.close()
