// common.cuh -- shared device helpers, counters and the single-pass prefix scan.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <assert.h>

namespace g2n {

#define G2N_SM_COUNT 148

// -DG2N_CHECKED: bounds assertions on every computed index of the hot kernels (ring positions, table slots, edge
// records, row cursors, IDs).  compute-sanitizer is closed on the B200 pool, so tools/sanitize_run.py is run against a
// checked build instead (profiles/r2_checked_build.md); the product build compiles them away.
#ifdef G2N_CHECKED
#define G2N_CHECK(cond) assert(cond)
#else
#define G2N_CHECK(cond) ((void)0)
#endif

typedef unsigned long long u64;
typedef unsigned int u32;

// ---------------------------------------------------------------- cache-bypassing loads
__device__ __forceinline__ u64 ld_volatile_u64(const u64* p)
{
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u32 ld_volatile_u32(const u32* p)
{
    u32 v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(u64* p, u64 v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_nc_v4(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// L2 eviction-priority policies: streamed input must not push the hash table out of the 126 MB L2
__device__ __forceinline__ u64 policy_evict_first()
{
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ u64 policy_evict_last()
{
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ld_stream_v4(const void* p, u64 pol)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}

// ---------------------------------------------------------------- warp / block scans
__device__ __forceinline__ u32 warp_incl_scan(u32 v)
{
    const u32 lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (u32)d) v += t;
    }
    return v;
}
__device__ __forceinline__ u64 warp_incl_scan64(u64 v)
{
    const u32 lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u64 t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (u32)d) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one u64 per thread; returns exclusive prefix, *total = block sum.
// `sm` must hold (blockDim.x / 32 + 1) u64.  Ends with a __syncthreads().
__device__ __forceinline__ u64 block_excl_scan64(u64 v, u64* sm, u64* total)
{
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    u64 inc = warp_incl_scan64(v);
    if (lane == 31) sm[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        u64 w = lane < nw ? sm[lane] : 0;
        u64 winc = warp_incl_scan64(w);
        if (lane < nw) sm[lane] = winc - w;
        if (lane == nw - 1) sm[nw] = winc;
    }
    __syncthreads();
    u64 res = sm[wid] + inc - v;
    *total = sm[nw];
    __syncthreads();
    return res;
}

// ---------------------------------------------------------------- decoupled look-back
// state word: [63:62] flag (0 = empty, 1 = aggregate only, 2 = inclusive prefix) | [61:0] value
#define LB_AGG (1ull << 62)
#define LB_INC (2ull << 62)
#define LB_VAL(x) ((x) & ((1ull << 62) - 1))

__device__ __forceinline__ u64 warp_sum64(u64 v)
{
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Called by all 32 lanes of ONE warp of the block that owns `tile` (tiles are handed out by an atomic
// ticket, so every predecessor is already running).  Publishes `agg`, resolves and returns the
// exclusive prefix (in every lane).  32 predecessors are inspected per round.
__device__ __forceinline__ u64 lookback_exclusive(u64* state, u32 tile, u64 agg)
{
    const u32 lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) st_volatile_u64(&state[0], LB_INC | agg);
        return 0;
    }
    if (lane == 0) st_volatile_u64(&state[tile], LB_AGG | agg);
    u64 excl = 0;
    long long base = (long long)tile - 1;
    while (true) {
        const long long idx = base - (long long)lane;
        const u64 s = idx >= 0 ? ld_volatile_u64(&state[idx]) : LB_INC;  // virtual tile -1: inclusive 0
        const u64 flag = s >> 62;
        const u32 ready = __ballot_sync(0xffffffffu, flag != 0);
        const u32 inc = __ballot_sync(0xffffffffu, flag == 2);
        if (inc) {
            const int first = __ffs(inc) - 1;  // nearest predecessor holding an inclusive prefix
            const u32 need = first == 31 ? 0xffffffffu : ((2u << first) - 1u);
            if ((ready & need) != need) continue;  // a nearer tile has not published yet
            excl += warp_sum64((int)lane <= first ? LB_VAL(s) : 0ull);
            break;
        }
        if (ready != 0xffffffffu) continue;
        excl += warp_sum64(LB_VAL(s));
        base -= 32;
    }
    if (lane == 0) {
        __threadfence();
        st_volatile_u64(&state[tile], LB_INC | (excl + agg));
    }
    return excl;
}

// ---------------------------------------------------------------- generic exclusive scan
// out[i] = sum_{j<i} load(j) for i in [0, n]; out has n+1 entries (out[n] = total).  `out2`, if given,
// receives a second copy (the row cursors of the scatter pass).  n = *n_dev when n_dev is given (sizes that
// only the device knows, see DevSizes), else n_host.
// grid: any size >= 1 (persistent, ticketed tiles of SCAN_TILE items); block: 256 threads.
// `state` (one word per tile) and `ticket` must be zero at launch.
#define SCAN_ITEMS 8
#define SCAN_TILE (256 * SCAN_ITEMS)

template <typename Tout, class LoadOp>
__global__ void __launch_bounds__(256) k_scan_exclusive(LoadOp load, Tout* __restrict__ out, Tout* __restrict__ out2, u64 n_host,
                                                         const u32* __restrict__ n_dev, u64* __restrict__ state, u32* __restrict__ ticket)
{
    __shared__ u64 sm[10];
    __shared__ u32 s_tile;
    __shared__ u64 s_base;
    const u64 n = n_dev ? (u64)*n_dev : n_host;
    const u64 n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= n_tiles) break;
        const u64 base = (u64)tile * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
        u64 v[SCAN_ITEMS];
        u64 sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            v[k] = (base + k < n) ? (u64)load(base + k) : 0;
            sum += v[k];
        }
        u64 total;
        u64 excl = block_excl_scan64(sum, sm, &total);
        if (threadIdx.x < 32) {
            const u64 e = lookback_exclusive(state, tile, total);
            if (threadIdx.x == 0) s_base = e;
        }
        __syncthreads();
        u64 run = s_base + excl;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (base + k < n) {
                out[base + k] = (Tout)run;
                if (out2) out2[base + k] = (Tout)run;
            }
            run += v[k];
        }
        if (tile == n_tiles - 1 && threadIdx.x == 255) {
            out[n] = (Tout)run;
            if (out2) out2[n] = (Tout)run;
        }
        __syncthreads();
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        out[0] = (Tout)0;
        if (out2) out2[0] = (Tout)0;
    }
}

// ---------------------------------------------------------------- gang scan
// Same contract as k_scan_exclusive, for launches whose CTAs are all co-resident (cooperative launch,
// grid <= a few CTAs per SM): every CTA owns one contiguous range, reduces it, publishes the sum, meets
// the others at a grid-wide barrier, adds up its predecessors' sums and scans its range.  Three memory
// round trips in a row instead of a look-back chain: about half the latency for the 10^4 .. 10^7 element
// scans of this pipeline.  load.peek(i) must be free of side effects; load(i) runs once per element.
// `state`: gridDim.x + 2 words, zero at launch ([0] arrival counter, [1 + b] sum of CTA b).
template <typename Tout, class LoadOp>
__global__ void __launch_bounds__(256) k_scan_gang(LoadOp load, Tout* __restrict__ out, Tout* __restrict__ out2, u64 n_host,
                                                    const u32* __restrict__ n_dev, u64* __restrict__ state)
{
    __shared__ u64 sm[10];
    const u64 n = n_dev ? (u64)*n_dev : n_host;
    const u32 G = gridDim.x, b = blockIdx.x;
    u64 chunk = (n + G - 1) / G;
    chunk = (chunk + SCAN_TILE - 1) / SCAN_TILE * SCAN_TILE;
    const u64 lo = min(n, (u64)b * chunk), hi = min(n, lo + chunk);
    // pass 1: my range's sum (coalesced, four loads in flight)
    u64 s = 0;
    for (u64 i = lo + threadIdx.x; i < hi; i += 4 * 256) {
        u64 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (i + k * 256 < hi) ? (u64)load.peek(i + k * 256) : 0ull;
        s += v[0] + v[1] + v[2] + v[3];
    }
    u64 total;
    block_excl_scan64(s, sm, &total);
    if (threadIdx.x == 0) {
        st_volatile_u64(&state[1 + b], total);
        __threadfence();
        atomicAdd((unsigned long long*)&state[0], 1ull);
        while (ld_volatile_u64(&state[0]) < (u64)G) {}
        __threadfence();
    }
    __syncthreads();
    // sums of my predecessors
    u64 p = 0;
    for (u32 j = threadIdx.x; j < b; j += 256) p += ld_volatile_u64(&state[1 + j]);
    u64 base;
    block_excl_scan64(p, sm, &base);
    // pass 2: scan my range tile by tile
    for (u64 t0 = lo; t0 < hi; t0 += SCAN_TILE) {
        const u64 i0 = t0 + (u64)threadIdx.x * SCAN_ITEMS;
        u64 v[SCAN_ITEMS];
        u64 sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            v[k] = (i0 + k < hi) ? (u64)load(i0 + k) : 0ull;
            sum += v[k];
        }
        u64 tile_total;
        const u64 excl = block_excl_scan64(sum, sm, &tile_total);
        u64 run = base + excl;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (i0 + k < hi) {
                out[i0 + k] = (Tout)run;
                if (out2) out2[i0 + k] = (Tout)run;
            }
            run += v[k];
        }
        base += tile_total;
    }
    // whoever owns the end of the input writes the total (CTA 0 when the input is empty)
    if (threadIdx.x == 0 && ((hi == n && lo < n) || (n == 0 && b == 0))) {
        out[n] = (Tout)base;
        if (out2) out2[n] = (Tout)base;
    }
}

template <typename T>
struct LoadArray {
    const T* p;
    __device__ __forceinline__ u64 operator()(u64 i) const { return (u64)p[i]; }
    __device__ __forceinline__ u64 peek(u64 i) const { return (u64)p[i]; }
};
// The first-appearance bitmap is ranked in groups of 8 words (256 bits = one 32-byte sector): the scan runs over
// groups (8x fewer elements and prefix words), a bit's rank adds the popcounts inside its group.  (A one-CTA scan for
// inputs this short was measured and dropped: 81 us per launch against 17 us for the gang scan -- one SM's load latency
// per 4096-element chunk, in sequence.)
#define BM_GROUP 8
struct LoadPopc8 {
    const u32* p;  // padded to whole groups
    __device__ __forceinline__ u64 operator()(u64 g) const
    {
        const uint4 a = reinterpret_cast<const uint4*>(p)[2 * g], b = reinterpret_cast<const uint4*>(p)[2 * g + 1];
        return (u64)(__popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w));
    }
    __device__ __forceinline__ u64 peek(u64 g) const { return (*this)(g); }
};
// number of set bits below bit `ob` of the bitmap, given the exclusive prefix over groups
__device__ __forceinline__ u32 bitmap_rank(const u32* __restrict__ bitmap, const u32* __restrict__ gprefix, u32 ob)
{
    const u32 wd = ob >> 5, g = wd / BM_GROUP, k = wd % BM_GROUP;
    const uint4 a = reinterpret_cast<const uint4*>(bitmap)[2 * g], b = reinterpret_cast<const uint4*>(bitmap)[2 * g + 1];
    const u32 w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    u32 r = gprefix[g];
#pragma unroll
    for (u32 j = 0; j < 8; j++) r += j < k ? __popc(w[j]) : (j == k ? __popc(w[j] & ((1u << (ob & 31)) - 1u)) : 0u);
    return r;
}

}  // namespace g2n
