"""Multi-GPU partitioning of the freq path (one process per GPU, torch.distributed for plumbing).

Reads are independent and counts are additive, so the path shards with no data-path collective:
  * contig sharding  -- whole contigs are dealt to ranks by longest-processing-time bin packing; a rank
    only loads the reference / allocates dense counts for its own contigs and only receives their reads.
  * region sharding  -- for a job dominated by one contig, rank r owns the reads that START in its slice
    [start_r, end_r).  A read may run past end_r, so its counts land in rank r's copy of the next slice's
    first positions (the halo).  One exchange at the end -- a sum all-reduce of each boundary's halo cells
    (uint64 pairs reinterpreted as int64; n_called/n_mod never carry into each other) -- moves them to the
    owner.  This is the only collective on the path (NCCL over NVLink on GPUs, gloo in the CPU tests).
There is no reference equivalent: the reference is a single process ("collective inventory").
"""
import ctypes as C

import numpy as np

from . import _native as N


def lpt_partition(lengths, n_ranks):
    """Longest-processing-time bin packing: contig index lists per rank, loads balanced within the largest item."""
    bins = [[] for _ in range(n_ranks)]
    load = [0] * n_ranks
    for tid in sorted(range(len(lengths)), key=lambda i: (-lengths[i], i)):
        r = min(range(n_ranks), key=lambda k: (load[k], k))
        bins[r].append(tid)
        load[r] += lengths[tid]
    return [sorted(b) for b in bins], load


def region_bounds(contig_len, n_ranks):
    """[start, end) of every rank's slice of one contig."""
    edges = [contig_len * r // n_ranks for r in range(n_ranks + 1)]
    return [(edges[r], edges[r + 1]) for r in range(n_ranks)]


def owner_of(pos, bounds):
    for r, (s, e) in enumerate(bounds):
        if s <= pos < e:
            return r
    return len(bounds) - 1


class _CudaView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def dense_tensor(lib, ctx, tid, start, end, cuda):
    """A torch int64 view (no copy) of the dense count cells of [start,end) on contig tid."""
    import torch
    ptr, n = C.c_void_p(), C.c_uint64()
    if lib.mmc_dense_slice(ctx, tid, start, end, C.byref(ptr), C.byref(n)) != 0:
        raise RuntimeError(lib.mmc_strerror(ctx).decode())
    if n.value == 0:
        return torch.zeros(0, dtype=torch.int64, device="cuda" if cuda else "cpu")
    if cuda:
        return torch.as_tensor(_CudaView(ptr.value, n.value), device="cuda")
    arr = np.ctypeslib.as_array((C.c_int64 * n.value).from_address(ptr.value))
    return torch.from_numpy(arr)


def slice_bounds(contig_len, n_ranks):
    """[start, end) of the positions rank k owns under the rule owner(p) = p * n_ranks // contig_len (the rule
    mmc_region_reduce() and `minimod --shard-regions` use)."""
    starts = [(k * contig_len + n_ranks - 1) // n_ranks for k in range(n_ranks)] + [contig_len]
    return [(starts[k], starts[k + 1]) for k in range(n_ranks)]


def exchange_halos(lib, ctx, tid, bounds, rank, dist, cuda, stats=None):
    """After all batches: move every boundary's counts to the rank that owns the positions.  Rank k holds the reads that
    start in bounds[k]; a read may run past its slice (even across a whole slice), so for every owner j the cells of
    [start_j, min(end_j, reach_j)) -- reach_j = how far the reads of ranks < j ran -- are summed over all ranks (one
    all-reduce per boundary: NCCL on GPUs, gloo in the CPU tests), kept by rank j and zeroed everywhere else.  The regions
    of different owners are disjoint, so no cell is ever summed twice.  Sparse rows (insertions, exotic haplotypes) of a
    halo stay on the rank that counted them: merge_rows() adds rows with equal keys when the tables are put together.
    Returns the widest halo (positions)."""
    import torch
    world = dist.get_world_size()
    lo, hi = C.c_uint32(), C.c_uint32()
    if lib.mmc_touched_range(ctx, tid, C.byref(lo), C.byref(hi)) != 0:
        raise RuntimeError(lib.mmc_strerror(ctx).decode())
    mine = torch.tensor([int(lo.value), int(hi.value)], dtype=torch.int64, device="cuda" if cuda else "cpu")
    spans = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(spans, mine)
    spans = [(int(t[0]), int(t[1])) for t in spans]
    if bounds is None:                                   # reads were dealt by index, not by position: rank j owns from its first read on
        starts = [0] + [spans[j][0] if spans[j][1] > spans[j][0] else None for j in range(1, world)]
        for j in range(world - 1, 0, -1):                # a rank without reads owns nothing: its slice collapses onto the next start
            if starts[j] is None:
                starts[j] = starts[j + 1] if j + 1 < world else max(sp[1] for sp in spans)
        bounds = [(starts[j], starts[j + 1] if j + 1 < world else 1 << 62) for j in range(world)]
    widest, reach = 0, 0
    for j in range(1, world):
        if spans[j - 1][1] > spans[j - 1][0]:
            reach = max(reach, spans[j - 1][1])
        s, e = bounds[j][0], min(bounds[j][1], reach)
        if e <= s:
            continue
        cells = dense_tensor(lib, ctx, tid, s, e, cuda)
        dist.all_reduce(cells, op=dist.ReduceOp.SUM)
        if rank == j:
            lib.mmc_dense_touch(ctx, tid, s, e)
        else:
            cells.zero_()
        widest = max(widest, e - s)
        if stats is not None:
            stats["bytes"] = stats.get("bytes", 0) + cells.numel() * 8
    return widest


def merge_rows(parts):
    """Union of the ranks' freq tables (lists of canonical row tuples (tid, pos, strand, code, ins, hap, n_called, n_mod)):
    rows with equal keys -- sparse rows counted on both sides of a boundary -- are added."""
    acc = {}
    for part in parts:
        for r in part:
            k = tuple(r[:6])
            if k in acc:
                acc[k] = (acc[k][0] + r[6], acc[k][1] + r[7])
            else:
                acc[k] = (r[6], r[7])
    return sorted(k + v for k, v in acc.items())
