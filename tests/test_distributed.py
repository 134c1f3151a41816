"""N>1 host logic on CPU: two gloo ranks, each driving the (emulated) kernels on the reads that start in its
half of the contig, then the halo exchange of minimod_b200.shard -- the union of the ranks' rows must equal
the single-process result.  Also the contig partitioner."""
import os
import subprocess
import sys

from helpers import ROOT
from minimod_b200 import shard

GRCH38_LIKE = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
               135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
               46709983, 50818468, 156040895, 57227415, 16569] + [40000 + 1000 * i for i in range(170)]


def test_lpt_partition_balances_grch38():
    for g in (1, 2, 4, 8):
        bins, load = shard.lpt_partition(GRCH38_LIKE, g)
        assert sorted(t for b in bins for t in b) == list(range(len(GRCH38_LIKE)))
        assert max(load) - min(load) <= max(GRCH38_LIKE)
        assert max(load) <= sum(GRCH38_LIKE) / g * 1.25 or g == 1


def test_region_bounds_cover_contig():
    b = shard.region_bounds(50818468, 8)
    assert b[0][0] == 0 and b[-1][1] == 50818468
    assert all(b[i][1] == b[i + 1][0] for i in range(7))
    assert shard.owner_of(0, b) == 0 and shard.owner_of(50818467, b) == 7


WORKER = r'''
import ctypes as C, os, sys, pickle
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from minimod_b200 import _native as N, shard
from minimod_b200.synth import Synth, CONFIG_ARGS
from parity import Pair
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
lib = N.load_cuda(os.path.join({root!r}, "tests/kernel_emul/_build/libminimod_emul.so"))
config, clen, insertions = {config!r}, {clen!r}, {insertions!r}
s = Synth(config, contigs=(("chrS", clen),), coverage=3.0)
ca = CONFIG_ARGS[config]
p, n = s.ref(0)
ref = C.string_at(p, n)
bounds = shard.slice_bounds(clen, world)
kw = dict(insertions=insertions, haplotypes=bool(ca.get("haplotypes")), max_reads=s.n_reads + 8, max_bytes=64 << 20)
pair = Pair(lib, "freq", [("chrS", ref)], ca["mod_codes"], ca.get("mod_thresh"), **kw)
# pack everything, then keep only the reads whose start this rank owns (reads are coordinate sorted)
full = Pair(lib, "freq", [("chrS", ref)], ca["mod_codes"], ca.get("mod_thresh"), **kw)
s.fill(full.batch, 0, s.n_reads, 2)
starts = [full.batch.contents.pos[i] for i in range(full.batch.contents.n_reads)]
mine = [i for i, st in enumerate(starts) if min(world - 1, st * world // clen) == rank]
if mine:
    assert mine == list(range(mine[0], mine[-1] + 1))
    s.fill(pair.batch, mine[0], len(mine), 2)
    rc, msg = pair.run_device(); assert rc == 0, msg
halo = shard.exchange_halos(lib, pair.ctx, 0, bounds, rank, dist, cuda=False)
rows = pair.device_freq() if mine else []
gathered = [None] * world
dist.all_gather_object(gathered, (rows, halo))
if rank == 0:
    rc, msg = full.run_device(); assert rc == 0, msg
    single = full.device_freq()
    merged = shard.merge_rows([part for part, _ in gathered])
    assert max(h for _, h in gathered) > {min_halo!r}, gathered[0][1]     # reads do run across the boundaries
    assert merged == single and len(single) > 1000, (len(merged), len(single))
    print("OK", len(single), "rows; halo", max(h for _, h in gathered))
dist.destroy_process_group()
'''


def _run_region(tmp_path, world, port, **fmt):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, **fmt))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
             for r in range(world)]
    outs = [p.communicate(timeout=900) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e.decode()[-3000:]
    assert b"OK" in outs[0][0]


def test_region_sharding_three_ranks_gloo(emul_lib, tmp_path):
    # (three slices: the synthetic contig has an N run in its middle that no read crosses)
    _run_region(tmp_path, 3, 29617, config=3, clen=150000, insertions=False, min_halo=1000)


def test_region_sharding_insertions_three_ranks_gloo(emul_lib, tmp_path):
    """--insertions: sparse rows sit on both sides of a boundary and must be added, not dropped."""
    _run_region(tmp_path, 3, 29637, config=3, clen=150000, insertions=True, min_halo=1000)


def test_region_sharding_halo_wider_than_slice_gloo(emul_lib, tmp_path):
    """50 kb reads over 30 kb slices: a read's counts cross two boundaries; every cell must be summed exactly once."""
    _run_region(tmp_path, 4, 29647, config=4, clen=120000, insertions=False, min_halo=20000)


# ---- contig sharding (BASELINE config 5 shape): GRCh38-like contig table scaled down, contigs dealt to ranks by LPT,
# every rank loads the reference and allocates counts for ITS contigs only and sees only their reads; no collective.
CONTIG_WORKER = r'''
import ctypes as C, os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch.distributed as dist
from minimod_b200 import _native as N, shard
from minimod_b200.synth import Synth
from parity import Pair
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
lib = N.load_cuda(os.path.join({root!r}, "tests/kernel_emul/_build/libminimod_emul.so"))
lens = {lens!r}
contigs = tuple(("ctg%d" % i, l) for i, l in enumerate(lens))
s = Synth(5, contigs=contigs, coverage=1.5)
refs = []
for tid in range(len(contigs)):
    p, n = s.ref(tid); refs.append(C.string_at(p, n))
bins, load = shard.lpt_partition(lens, world)
mine = set(bins[rank])
kw = dict(max_reads=s.n_reads + 8, max_bytes=64 << 20)
full = Pair(lib, "freq", [(contigs[t][0], refs[t]) for t in range(len(contigs))], "m[CG]", "0.8", **kw)
s.fill(full.batch, 0, s.n_reads, 2)
tids = [full.batch.contents.tid[i] for i in range(full.batch.contents.n_reads)]
pair = Pair(lib, "freq", [(contigs[t][0], refs[t] if t in mine else None, lens[t]) for t in range(len(contigs))], "m[CG]", "0.8", **kw)
n_mine = 0
for t in sorted(mine):                                     # reads are sorted by contig: one contiguous range each
    idx = [i for i, x in enumerate(tids) if x == t]
    if not idx: continue
    assert idx == list(range(idx[0], idx[-1] + 1))
    b = pair.batch.contents                                # the packer appends: start every batch empty
    b.n_reads = 0; b.cigar_used = 0; b.seq_used = 0; b.mm_used = 0; b.ml_used = 0; b.seq_exc_used = 0; b.cig8_used = 0
    got, _ = s.fill(pair.batch, idx[0], len(idx), 2); assert got == len(idx)
    rc, msg = pair.run_device(); assert rc == 0, msg
    n_mine += len(idx)
rows = pair.device_freq()
assert all(r[0] in mine for r in rows)
gathered = [None] * world
dist.all_gather_object(gathered, (rows, n_mine))
if rank == 0:
    rc, msg = full.run_device(); assert rc == 0, msg
    single = full.device_freq()
    merged = sorted(r for part, _ in gathered for r in part)
    assert sum(n for _, n in gathered) == s.n_reads
    assert merged == single and len(single) > 500, (len(merged), len(single))
    print("OK", len(single), "rows over", len(lens), "contigs; reads per rank", [n for _, n in gathered])
dist.destroy_process_group()
'''


def test_contig_sharding_two_ranks_gloo(emul_lib, tmp_path):
    lens = [max(3000, l // 3000) for l in GRCH38_LIKE[:25]] + [3000 + 40 * i for i in range(20)]
    script = tmp_path / "contig_worker.py"
    script.write_text(CONTIG_WORKER.format(root=ROOT, lens=lens))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29627", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
             for r in range(2)]
    outs = [p.communicate(timeout=900) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e.decode()[-3000:]
    assert b"OK" in outs[0][0]
