// mmc_sparse.cuh -- finalize of the sparse side buffer on the device (sm_100a).
//
// Counts that have no dense cell (ins_offset > 0 under --insertions, haplotype / code ids beyond the
// dense slots) are appended by the decode kernels as SparseRec {a, b, w}.  With --insertions on
// ONT reads they are a quarter of all output rows (3.8 M records per 149 k-read pass of BASELINE
// config 3), and sorting them on the host took 240 ms of a 380 ms pass.  So the collect + sort of
// print_freq_output() (src/mod.c:644-664) for these rows is done here:
//
//   k_sparse_keys     record -> 25-bit minor key (ins_offset, haplotype order) + its index
//   [radix sort by the minor key, then a stable radix sort by the 64-bit major key a:
//    cub::DeviceRadixSort, library code like the prefix sum below]
//   k_sparse_gather   major key of the records in minor-key order
//   k_sparse_heads    first record of every (a, b) run
//   k_sparse_emit     one output row per run: n_called / n_mod summed over the run
//   k_merge_rows      dense rows (already position ordered) and sparse rows -> one ordered array,
//                     each row placed by a binary search in the other list (both fit in L2)
//
// Row order is the reference's (contig, pos, strand, code, ins_offset, haplotype with '*' first).
#ifndef MMC_SPARSE_CUH
#define MMC_SPARSE_CUH

#include "mmc_device.cuh"
#include "mmc_decode_warp.cuh"

namespace mmc {

constexpr int kSpThreads = 256;

__device__ __forceinline__ uint32_t sparse_minor(uint32_t b) {             // (ins_offset, hap) in row order: '*' (256) first
    const uint32_t h9 = b >> 16;
    return ((b & 0xffffu) << 9) | (h9 == 256u ? 0u : h9 + 1u);
}

// A pass works on the records whose major key lies in [lo, hi): what a drain behind a coordinate watermark takes (lo = the
// previous watermark, hi = this one) or what the final pass still owes (hi = all ones, which also excludes the sentinels).
// Everything else -- records an earlier drain returned, records of batches still in flight (never below a watermark; they may
// even be half written: the slots past the fill level and the unused slots of a warp's chunk hold sentinels) -- reads as a sentinel.
__device__ __forceinline__ unsigned long long sparse_major(const SparseRec &r, unsigned long long lo, unsigned long long hi) {
    return (r.a >= lo && r.a < hi) ? r.a : kSpSentinel;
}

__global__ void __launch_bounds__(kSpThreads) k_sparse_keys(const SparseRec *raw, uint32_t n, uint32_t *minor, uint32_t *idx) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        minor[i] = sparse_minor(raw[i].b);
        idx[i] = i;
    }
}

__global__ void __launch_bounds__(kSpThreads) k_sparse_gather(const SparseRec *raw, const uint32_t *idx, uint32_t n, unsigned long long *major,
                                                              unsigned long long lo, unsigned long long hi) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) major[i] = sparse_major(raw[idx[i]], lo, hi);
}

// flag[i] = 1 when sorted record i starts a run of equal (a, b); sentinels (sorted last) never do
__global__ void __launch_bounds__(kSpThreads) k_sparse_heads(const SparseRec *raw, const uint32_t *idx, uint32_t n, uint32_t *flag,
                                                             unsigned long long lo, unsigned long long hi) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const SparseRec r = raw[idx[i]];
        const unsigned long long ra = sparse_major(r, lo, hi);
        uint32_t f = ra != kSpSentinel;
        if (f && i > 0) { const SparseRec p = raw[idx[i - 1u]]; f = sparse_major(p, lo, hi) != ra || p.b != r.b; }
        flag[i] = f;
    }
}

__global__ void __launch_bounds__(kSpThreads) k_sparse_emit(const SparseRec *raw, const uint32_t *idx, const uint32_t *flag, const uint32_t *off,
                                                            uint32_t n, FreqRecDev *rows, unsigned long long *n_rows,
                                                            unsigned long long lo, unsigned long long hi) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i == n - 1u) *n_rows = (unsigned long long)off[i] + flag[i];
        if (!flag[i]) continue;
        const SparseRec r = raw[idx[i]];
        unsigned long long called = r.w & 0xffffu, mod = r.w >> 16;
        for (uint32_t j = i + 1u; j < n && !flag[j]; ++j) {                // the rest of the run (short: one record per read)
            const SparseRec q = raw[idx[j]];
            if (sparse_major(q, lo, hi) == kSpSentinel) break;
            called += q.w & 0xffffu; mod += q.w >> 16;
        }
        FreqRecDev o;
        o.tid = (int32_t)(r.a >> 41); o.pos = (int32_t)((r.a >> 9) & 0xffffffffull);
        o.n_called = (uint32_t)called; o.n_mod = (uint32_t)mod;
        o.ins_offset = (uint16_t)(r.b & 0xffffu);
        const uint32_t h9 = r.b >> 16;
        o.hap = h9 == 256u ? (int16_t)-1 : (int16_t)h9;
        o.strand = (uint8_t)((r.a >> 8) & 1u); o.code = (uint8_t)(r.a & 0xffu); o.reserved = 0;
        rows[off[i]] = o;
    }
}

struct RowKey { unsigned long long hi; uint32_t lo; };
__device__ __forceinline__ RowKey row_key(const FreqRecDev &r) {
    RowKey k;
    k.hi = ((unsigned long long)(uint32_t)r.tid << 41) | ((unsigned long long)(uint32_t)r.pos << 9) | ((unsigned long long)r.strand << 8) | r.code;
    k.lo = ((uint32_t)r.ins_offset << 9) | (uint32_t)((int32_t)r.hap + 1);
    return k;
}
__device__ __forceinline__ bool key_less(const RowKey &x, const RowKey &y) { return x.hi != y.hi ? x.hi < y.hi : x.lo < y.lo; }

// number of rows of v[lo..hi) (+ lo) whose key is < k  (no key occurs in both lists)
__device__ __forceinline__ unsigned long long rows_below(const FreqRecDev *v, unsigned long long lo, unsigned long long hi, const RowKey &k) {
    while (lo < hi) {
        const unsigned long long mid = (lo + hi) >> 1;
        if (key_less(row_key(v[mid]), k)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Both lists are ordered, so the rows a CTA takes in one step (blockDim.x consecutive rows of one list) fall between the places
// of the step's first and last row in the other list: two full-length searches per step, every other thread searches that
// window only (a few steps instead of log2 of millions).
__global__ void __launch_bounds__(kSpThreads) k_merge_rows(const FreqRecDev *dense, unsigned long long n_dense, const FreqRecDev *sparse, unsigned long long n_sparse,
                                                           FreqRecDev *out) {
    __shared__ unsigned long long win[2];
    const unsigned long long total = n_dense + n_sparse, step = blockDim.x;
    const unsigned long long steps_d = (n_dense + step - 1) / step, steps = steps_d + (n_sparse + step - 1) / step;
    (void)total;
    for (unsigned long long st = blockIdx.x; st < steps; st += gridDim.x) {
        const bool from_dense = st < steps_d;
        const FreqRecDev *mine = from_dense ? dense : sparse, *other = from_dense ? sparse : dense;
        const unsigned long long n_mine = from_dense ? n_dense : n_sparse, n_other = from_dense ? n_sparse : n_dense;
        const unsigned long long i0 = (from_dense ? st : st - steps_d) * step, i1 = i0 + step < n_mine ? i0 + step : n_mine;
        if (threadIdx.x < 2u) {
            const FreqRecDev r = mine[threadIdx.x == 0u ? i0 : i1 - 1u];
            win[threadIdx.x] = rows_below(other, 0, n_other, row_key(r));
        }
        __syncthreads();
        const unsigned long long i = i0 + threadIdx.x;
        if (i < i1) {
            const FreqRecDev r = mine[i];
            out[i + rows_below(other, win[0], win[1], row_key(r))] = r;
        }
        __syncthreads();
    }
}

}  // namespace mmc

#endif  // MMC_SPARSE_CUH
