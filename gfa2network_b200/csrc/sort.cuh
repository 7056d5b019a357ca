// sort.cuh -- K4: stable LSD radix sort of (major, minor) keys + segmented duplicate sum into
// indptr / indices / data, including the reference's max(S, S^T) symmetrisation.
//
// Replaces the SciPy C++ the reference reaches through
//   builders.py:283   out_mat.maximum(out_mat.T)      -> coo_tocsr, csr_sort_indices, csr_sum_duplicates,
//                                                       csr_maximum_csr (scipy/sparse/sparsetools/csr.h)
//   utils.py:55       A.asformat("csr" | "csc")       -> coo_tocsr + sum_duplicates
// Stability + left-to-right summation reproduce SciPy's emission-order duplicate sums (exact for
// rows of <= 16 stored entries; otherwise within the stated float tolerance).
#pragma once
#include "ids.cuh"

namespace g2n {

#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)  // 4096 keys per tile
#define RS_RADIX 256

// Per-warp digit ranks of one tile.  Keys are taken in memory order: warp w owns RS_ITEMS rounds of 32
// consecutive keys.  On return warp_hist[w][d] = number of keys of warp w with digit d, and
// rank[r] = number of earlier keys *of the same warp* with the same digit.
__device__ __forceinline__ void tile_ranks(const u64* __restrict__ keys, u64 n, u64 tile_base, int shift, u32 mask,
                                           u32 (*warp_hist)[RS_RADIX], u64 (&key)[RS_ITEMS], u32 (&rank)[RS_ITEMS])
{
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (u32 i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const u64 wbase = tile_base + (u64)wid * (RS_ITEMS * 32);
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const u64 idx = wbase + (u64)r * 32 + lane;
        const bool valid = idx < n;
        key[r] = valid ? keys[idx] : ~0ull;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const u64 idx = wbase + (u64)r * 32 + lane;
        const bool valid = idx < n;
        const u32 d = valid ? (u32)((key[r] >> shift) & mask) : 0xFFFFFFFFu;
        const u32 peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        u32 base = 0;
        if (valid && (int)lane == leader) {
            base = warp_hist[wid][d];
            warp_hist[wid][d] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        rank[r] = base + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
}

// tile_hist[d * n_tiles + tile] = number of keys of `tile` with digit d
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const u64* __restrict__ keys, u64 n, int shift, u32 mask,
                                                            u32* __restrict__ tile_hist, u32 n_tiles)
{
    __shared__ u32 warp_hist[RS_WARPS][RS_RADIX];
    u64 key[RS_ITEMS];
    u32 rank[RS_ITEMS];
    for (u32 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        tile_ranks(keys, n, (u64)tile * RS_TILE, shift, mask, warp_hist, key, rank);
        u32 s = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) s += warp_hist[w][threadIdx.x];
        tile_hist[(u64)threadIdx.x * n_tiles + tile] = s;
        __syncthreads();
    }
}

template <bool HAS_PAYLOAD>
__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const u64* __restrict__ keys_in, const u32* __restrict__ pay_in,
                                                               u64* __restrict__ keys_out, u32* __restrict__ pay_out, u64 n, int shift,
                                                               u32 mask, const u32* __restrict__ tile_offs, u32 n_tiles)
{
    __shared__ u32 warp_hist[RS_WARPS][RS_RADIX];
    u64 key[RS_ITEMS];
    u32 rank[RS_ITEMS];
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (u32 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const u64 tile_base = (u64)tile * RS_TILE;
        tile_ranks(keys_in, n, tile_base, shift, mask, warp_hist, key, rank);
        {
            // digit d = threadIdx.x: exclusive scan over warps, seeded with the tile's global offset
            const u32 d = threadIdx.x;
            u32 run = tile_offs[(u64)d * n_tiles + tile];
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) {
                const u32 t = warp_hist[w][d];
                warp_hist[w][d] = run;
                run += t;
            }
        }
        __syncthreads();
        const u64 wbase = tile_base + (u64)wid * (RS_ITEMS * 32);
#pragma unroll
        for (int r = 0; r < RS_ITEMS; r++) {
            const u64 idx = wbase + (u64)r * 32 + lane;
            if (idx < n) {
                const u32 d = (u32)((key[r] >> shift) & mask);
                const u32 pos = warp_hist[wid][d] + rank[r];
                keys_out[pos] = key[r];
                if (HAS_PAYLOAD) pay_out[pos] = pay_in[idx];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- segmented reduce
// Sorted keys: ((major << mbits | minor) << 1) | dir.  A group = equal (major, minor).
//   sym == 0: value = left-to-right sum of the group's weights (csr_sum_duplicates); zeros kept.
//   sym == 1: a = sum of dir-0 weights, b = sum of dir-1 weights (0 if absent);
//             value = (a < b) ? b : a  (csr_maximum_csr); kept only if value != 0.
// Writes per head position: val[i], keep flag[i] (0 for non-heads).
template <typename T>
__global__ void __launch_bounds__(256) k_group_reduce(const u64* __restrict__ keys, const u32* __restrict__ payload,
                                                       const double* __restrict__ w_f64, const T* __restrict__ w_typed, u64 n,
                                                       int sym, T* __restrict__ val, u32* __restrict__ flag)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 g = keys[i] >> 1;
        if (i > 0 && (keys[i - 1] >> 1) == g) { flag[i] = 0; continue; }
        T a = zero_t<T>(), b = zero_t<T>();
        bool has_a = false, has_b = false;
        for (u64 j = i; j < n; j++) {
            const u64 kj = keys[j];
            if ((kj >> 1) != g) break;
            T x;
            if (w_typed) x = w_typed[payload[j]];
            else if (w_f64) x = cast_weight<T>(w_f64[payload[j]]);
            else x = cast_weight<T>(1.0);
            if (kj & 1) { b = has_b ? add_t<T>(b, x) : x; has_b = true; }
            else { a = has_a ? add_t<T>(a, x) : x; has_a = true; }
        }
        T v;
        bool keep;
        if (sym) { v = lt_t<T>(a, b) ? b : a; keep = nz_t<T>(v); }
        else { v = a; keep = true; }
        val[i] = v;
        flag[i] = keep ? 1u : 0u;
    }
}

// Compaction: kept heads -> indices/data at pos[i]; per-major counts for indptr.
template <typename T>
__global__ void __launch_bounds__(256) k_compact(const u64* __restrict__ keys, const T* __restrict__ val, const u32* __restrict__ flag,
                                                  const u32* __restrict__ pos, u64 n, int mbits, int32_t* __restrict__ indices,
                                                  T* __restrict__ data, u32* __restrict__ major_count)
{
    const u64 mmask = (1ull << mbits) - 1;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        if (!flag[i]) continue;
        const u64 g = keys[i] >> 1;
        const u32 p = pos[i];
        indices[p] = (int32_t)(g & mmask);
        data[p] = val[i];
        atomicAdd(&major_count[(u32)(g >> mbits)], 1u);
    }
}

}  // namespace g2n
