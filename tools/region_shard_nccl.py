#!/usr/bin/env python3
"""Region sharding of one contig across N GPUs with the NCCL halo exchange (minimod_b200.shard), checked against a
single-GPU run of the same reads.  Launch with torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/region_shard_nccl.py
The CPU-side twin of this check (gloo + SIMT emulator) is tests/test_distributed.py."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from minimod_b200 import _native as N, shard
from minimod_b200.synth import Synth
from parity import Pair

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = N.load_cuda()
clen = 6000000
s = Synth(3, contigs=(("chrS", clen),), coverage=10.0)
p, n = s.ref(0)
ref = C.string_at(p, n)
# uneven slices: the synthetic contig has an N run in the middle that no read crosses
edges = [0] + [clen * (2 * r + 1) // (2 * world + 1) for r in range(1, world)] + [clen]
bounds = [(edges[r], edges[r + 1]) for r in range(world)]
kw = dict(max_reads=s.n_reads + 8, max_bytes=int(s.n_reads * 30000), device=local)
pair = Pair(lib, "freq", [("chrS", ref)], "m[CG],h[CG]", "0.8,0.7", **kw)
full = Pair(lib, "freq", [("chrS", ref)], "m[CG],h[CG]", "0.8,0.7", **kw)
s.fill(full.batch, 0, s.n_reads, 4)
starts = [full.batch.contents.pos[i] for i in range(full.batch.contents.n_reads)]
mine = [i for i, st in enumerate(starts) if shard.owner_of(st, bounds) == rank]
assert mine and mine == list(range(mine[0], mine[-1] + 1))
s.fill(pair.batch, mine[0], len(mine), 4)
rc, msg = pair.run_device(); assert rc == 0, msg
torch.cuda.synchronize(); t0 = time.time()
halo = shard.exchange_halos(lib, pair.ctx, 0, bounds, rank, dist, cuda=True)
torch.cuda.synchronize(); t1 = time.time()
rows = pair.device_freq()
own = [r for r in rows if bounds[rank][0] <= r[1] < bounds[rank][1]]
gathered = [None] * world
dist.all_gather_object(gathered, own)
if rank == 0:
    rc, msg = full.run_device(); assert rc == 0, msg
    single = full.device_freq()
    merged = sorted(r for part in gathered for r in part)
    assert halo > 1000, halo
    assert merged == single, (len(merged), len(single))
    print(f"OK region sharding over {world} GPUs (NCCL): {len(single)} rows identical to the single-GPU run; "
          f"{s.n_reads} reads, halo {halo} positions, exchange {1e3 * (t1 - t0):.2f} ms")
dist.destroy_process_group()
