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


def exchange_halos(lib, ctx, tid, bounds, rank, dist, cuda):
    """After all batches: sum every boundary's halo cells across ranks so that the slice owner holds the
    complete counts.  Returns the halo width used (positions)."""
    import torch
    world = len(bounds)
    lo, hi = C.c_uint32(), C.c_uint32()
    if lib.mmc_touched_range(ctx, tid, C.byref(lo), C.byref(hi)) != 0:
        raise RuntimeError(lib.mmc_strerror(ctx).decode())
    over = max(0, int(hi.value) - bounds[rank][1])
    t = torch.tensor([over], dtype=torch.int64, device="cuda" if cuda else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    halo = int(t.item())
    if halo == 0:
        return 0
    contig_len = bounds[-1][1]
    for k in range(world - 1):                       # boundary between rank k and k+1
        s = bounds[k][1]
        e = min(contig_len, s + halo)
        cells = dense_tensor(lib, ctx, tid, s, e, cuda)
        dist.all_reduce(cells, op=dist.ReduceOp.SUM)
        if rank != k and e > s:                      # everyone but the left neighbour now holds the full halo;
            lib.mmc_dense_touch(ctx, tid, s, e)      # the owner (k+1, or further right for very long reads) emits it
    return halo


def owned_rows(rows, tid, bounds, rank):
    """Rows of the final table this rank is responsible for printing."""
    s, e = bounds[rank]
    keep = (rows["tid"] != tid) | ((rows["pos"] >= s) & (rows["pos"] < e))
    return rows[keep]
