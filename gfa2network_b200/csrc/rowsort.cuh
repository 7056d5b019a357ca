// rowsort.cuh -- K4: COO triplets -> indptr / indices / data by row bucketing.
//
// Adjacency rows of sequence graphs are tiny (a handful of links per segment), so instead of a global
// radix sort of (row, col) keys the build is a counting sort by row followed by an in-row sort:
//   k_rows_count     one atomic per triplet into cnt[major]                       (histogram)
//   exclusive scan   cnt -> rowptr                                                (common.cuh)
//   k_rows_scatter   entry (minor, dir, emission index) -> atomicAdd(cursor[major]) (any order inside a row)
//   k_rows_big       rows longer than RS_SMALL are sorted in place by a whole CTA (bitonic; rare)
//   k_rows_sort      one lane per row: insertion sort by (minor, dir, emission index) in shared memory,
//                    count of the entries the row will store (duplicates summed, zeros of max() dropped)
//   exclusive scan   counts -> indptr
//   k_rows_write     one lane per row: left-to-right duplicate sum, optional max(S, S^T), output
// No kernel waits on another CTA; the sorted order inside a row is total (the emission
// index breaks ties), so the result is deterministic and duplicate weights are summed in emission
// order exactly like SciPy does for rows of <= 16 stored entries.
//
// Replaces the SciPy C++ the reference reaches through
//   builders.py:283   out_mat.maximum(out_mat.T)  -> coo_tocsr, csr_sort_indices, csr_sum_duplicates, csr_maximum_csr
//   utils.py:55       A.asformat("csr" | "csc")   -> coo_tocsr + sum_duplicates
#pragma once
#include "ids.cuh"

namespace g2n {

#define RS_SMALL 64      // rows up to this many entries are sorted by one lane
#define RS_GROUP_CAP 1024  // entries of a 32-row group staged in shared memory per warp
#define RS_WARPS 4

// entry = minor << 33 | dir << 32 | emission index of the triplet
__device__ __forceinline__ u64 rs_entry(u32 minor, u32 dir, u32 t) { return ((u64)minor << 33) | ((u64)dir << 32) | t; }
__device__ __forceinline__ u32 rs_minor(u64 e) { return (u32)(e >> 33); }
__device__ __forceinline__ u32 rs_dir(u64 e) { return (u32)(e >> 32) & 1u; }
__device__ __forceinline__ u32 rs_t(u64 e) { return (u32)e; }

// Entries of one edge record:  g(major, minor, dir, t)
//   sym == 0: one entry per triplet, major = row (CSR) or col (CSC)
//   sym == 1: two entries per triplet: (row, col, dir 0) and (col, row, dir 1)   [max(S, S^T)]
template <class G>
__device__ __forceinline__ void record_entries(const u32 (&id)[4], int tpe, u32 t0, int sym, int csc, G g)
{
    for (int k = 0; k < tpe; k++) {
        u32 r, c;
        record_triplet(id, k, r, c);
        if (sym) { g(r, c, 0u, t0 + k); g(c, r, 1u, t0 + k); }
        else if (csc) g(c, r, 0u, t0 + k);
        else g(r, c, 0u, t0 + k);
    }
}

// histogram of majors; translates edge_slots to node IDs in place and lays the weights out in emission
// order (w_emit[t]) when there are any
__global__ void __launch_bounds__(256) k_rows_count(const EmitParams E, int sym, int csc, u32* __restrict__ cnt, double* __restrict__ w_emit)
{
    for_each_edge(E, [&](u32 stored, u32 t0, const u32 (&id)[4]) {
        record_entries(id, E.tpe, t0, sym, csc, [&](u32 major, u32, u32, u32) { atomicAdd(&cnt[major], 1u); });
        if (w_emit) {
            const double w = E.edge_w[stored];
            for (int k = 0; k < E.tpe; k++) w_emit[t0 + k] = w;
        }
    });
}

// cursor[major] starts at rowptr[major]; one atomicAdd hands out the entry's position inside the row's
// range (order inside the range is arbitrary; the in-row sort restores a total order)
__global__ void __launch_bounds__(256) k_rows_scatter(const EmitParams E, int sym, int csc, u32* __restrict__ cursor, u64* __restrict__ entries)
{
    for_each_edge(E, [&](u32, u32 t0, const u32 (&id)[4]) {
        record_entries(id, E.tpe, t0, sym, csc, [&](u32 major, u32 minor, u32 dir, u32 t) {
            entries[atomicAdd(&cursor[major], 1u)] = rs_entry(minor, dir, t);
        });
    });
}

// same two steps for caller-provided COO arrays (g2n_coo_to_compressed)
__global__ void __launch_bounds__(256) k_coo_count(const int32_t* __restrict__ row, const int32_t* __restrict__ col, u64 nnz, int csc, u32* __restrict__ cnt)
{
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (u64)gridDim.x * blockDim.x)
        atomicAdd(&cnt[(u32)(csc ? col[t] : row[t])], 1u);
}
__global__ void __launch_bounds__(256) k_coo_scatter(const int32_t* __restrict__ row, const int32_t* __restrict__ col, u64 nnz, int csc,
                                                      u32* __restrict__ cursor, u64* __restrict__ entries)
{
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += (u64)gridDim.x * blockDim.x) {
        const u32 major = (u32)(csc ? col[t] : row[t]), minor = (u32)(csc ? row[t] : col[t]);
        entries[atomicAdd(&cursor[major], 1u)] = rs_entry(minor, 0u, (u32)t);
    }
}

// ---------------------------------------------------------------- long rows (rare)
__global__ void __launch_bounds__(256) k_rows_find_big(const u32* __restrict__ rowptr, const u32* __restrict__ n_dev, u32* __restrict__ biglist, u32* __restrict__ bigcount)
{
    const u32 n = *n_dev;
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
        if (rowptr[r + 1] - rowptr[r] > RS_SMALL) biglist[atomicAdd(bigcount, 1u)] = r;
}

#define RS_BIG_SMEM 4096
// One CTA sorts one long row with a normalized bitonic network (every comparison ascending) over the
// next power of two; indices past the end act as +infinity and are never touched.  Rows that fit are
// sorted in shared memory, longer ones in place in global memory.
__global__ void __launch_bounds__(256) k_rows_big(const u32* __restrict__ rowptr, const u32* __restrict__ biglist,
                                                   const u32* __restrict__ bigcount, u64* __restrict__ entries)
{
    __shared__ u64 s_big[RS_BIG_SMEM];
    const u32 nbig = *bigcount;
    for (u32 b = blockIdx.x; b < nbig; b += gridDim.x) {
        const u32 r = biglist[b];
        const u32 lo = rowptr[r], len = rowptr[r + 1] - lo;
        u64* a = entries + lo;
        const bool in_smem = len <= RS_BIG_SMEM;
        if (in_smem) {
            for (u32 i = threadIdx.x; i < len; i += blockDim.x) s_big[i] = a[i];
            a = s_big;
        }
        __syncthreads();
        u32 p2 = 1;
        while (p2 < len) p2 <<= 1;
        for (u32 k = 2; k <= p2; k <<= 1) {
            for (u32 j = k >> 1; j > 0; j >>= 1) {
                const bool mirror = (j == (k >> 1));
                for (u32 i = threadIdx.x; i < p2; i += blockDim.x) {
                    const u32 l = mirror ? (i ^ (k - 1)) : (i ^ j);
                    if (l > i && l < len) {
                        const u64 x = a[i], y = a[l];
                        if (x > y) { a[i] = y; a[l] = x; }
                    }
                }
                __syncthreads();
            }
        }
        if (in_smem)
            for (u32 i = threadIdx.x; i < len; i += blockDim.x) entries[lo + i] = s_big[i];
        __syncthreads();
    }
}

// ---------------------------------------------------------------- finalize
template <typename T>
struct RowAcc {
    // accumulates one (major, minor) group: a = sum of dir-0 weights, b = sum of dir-1 weights
    T a, b;
    bool has_a, has_b;
    __device__ __forceinline__ void reset() { a = zero_t<T>(); b = zero_t<T>(); has_a = has_b = false; }
    __device__ __forceinline__ void add(u32 dir, T x)
    {
        if (dir) { b = has_b ? add_t<T>(b, x) : x; has_b = true; }
        else { a = has_a ? add_t<T>(a, x) : x; has_a = true; }
    }
    // sym == 0: csr_sum_duplicates keeps explicit zeros; sym == 1: csr_maximum_csr drops results == 0
    __device__ __forceinline__ bool result(int sym, T& v) const
    {
        if (sym) { v = lt_t<T>(a, b) ? b : a; return nz_t<T>(v); }
        v = a;
        return true;
    }
};

template <typename T>
__device__ __forceinline__ T entry_weight(u64 e, const double* __restrict__ w_emit, const T* __restrict__ w_typed)
{
    if (w_typed) return w_typed[rs_t(e)];
    if (w_emit) return cast_weight<T>(w_emit[rs_t(e)]);
    return cast_weight<T>(1.0);
}

// Walks one sorted row; emit(minor, value) is called for every stored result.  Returns their count.
template <typename T, class Emit>
__device__ __forceinline__ u32 walk_row(const u64* a, u32 len, int sym, const double* w_emit, const T* w_typed, Emit emit)
{
    u32 out = 0;
    u32 i = 0;
    while (i < len) {
        const u32 minor = rs_minor(a[i]);
        RowAcc<T> acc;
        acc.reset();
        while (i < len && rs_minor(a[i]) == minor) {
            acc.add(rs_dir(a[i]), entry_weight<T>(a[i], w_emit, w_typed));
            i++;
        }
        T v;
        if (acc.result(sym, v)) { emit(out, minor, v); out++; }
    }
    return out;
}

__device__ __forceinline__ void insertion_sort(u64* a, u32 len)
{
    for (u32 i = 1; i < len; i++) {
        const u64 x = a[i];
        u32 j = i;
        while (j > 0 && a[j - 1] > x) { a[j] = a[j - 1]; j--; }
        a[j] = x;
    }
}

// Pass A: one warp per group of 32 consecutive rows, one lane per row: sort the row (staged in shared
// memory when the group fits), write it back, count the entries it will store.
template <typename T>
__global__ void __launch_bounds__(RS_WARPS * 32) k_rows_sort(const u32* __restrict__ rowptr, u64* __restrict__ entries, const u32* __restrict__ n_dev, int sym,
                                                              const double* __restrict__ w_emit, const T* __restrict__ w_typed,
                                                              u32* __restrict__ ucnt)
{
    __shared__ u64 s_ent[RS_WARPS][RS_GROUP_CAP];
    const u32 n = *n_dev;
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u64* sm = s_ent[wid];
    const u32 n_groups = (n + 31) / 32;
    const u32 warp = blockIdx.x * RS_WARPS + wid, n_warps = gridDim.x * RS_WARPS;
    for (u32 g = warp; g < n_groups; g += n_warps) {
        const u32 r = g * 32 + lane;
        const bool live = r < n;
        const u32 lo = live ? rowptr[r] : 0, hi = live ? rowptr[r + 1] : 0;
        const u32 len = hi - lo;
        const u32 g_lo = __shfl_sync(0xffffffffu, lo, 0);
        const u32 g_hi = rowptr[min(g * 32 + 32, n)];
        const u32 g_len = g_hi - g_lo;
        const bool all_small = __all_sync(0xffffffffu, len <= RS_SMALL);
        const bool staged = all_small && g_len <= RS_GROUP_CAP;
        __syncwarp();
        u64* a;
        if (staged) {
            for (u32 i = lane; i < g_len; i += 32) sm[i] = entries[g_lo + i];
            __syncwarp();
            a = sm + (lo - g_lo);
        } else {
            a = entries + lo;  // in place in global memory (rows > RS_SMALL were sorted by k_rows_big)
        }
        if (len > 1 && len <= RS_SMALL) insertion_sort(a, len);
        const u32 mine = walk_row<T>(a, len, sym, w_emit, w_typed, [](u32, u32, T) {});
        if (live) ucnt[r] = mine;
        if (staged) {
            __syncwarp();
            for (u32 i = lane; i < g_len; i += 32) entries[g_lo + i] = sm[i];
        }
    }
}

// Pass B: indptr is known; one lane per row walks its sorted entries and writes indices / data.
template <typename T>
__global__ void __launch_bounds__(256) k_rows_write(const u32* __restrict__ rowptr, const u64* __restrict__ entries, const u32* __restrict__ n_dev, int sym,
                                                     const double* __restrict__ w_emit, const T* __restrict__ w_typed,
                                                     const int32_t* __restrict__ indptr, int32_t* __restrict__ indices, T* __restrict__ data,
                                                     u32* __restrict__ nnz_out)
{
    const u32 n = *n_dev;
    if (blockIdx.x == 0 && threadIdx.x == 0) *nnz_out = (u32)indptr[n];
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const u32 lo = rowptr[r], len = rowptr[r + 1] - lo;
        const u32 out0 = (u32)indptr[r];
        walk_row<T>(entries + lo, len, sym, w_emit, w_typed, [&](u32 k, u32 minor, T v) {
            indices[out0 + k] = (int32_t)minor;
            data[out0 + k] = v;
        });
    }
}

}  // namespace g2n
