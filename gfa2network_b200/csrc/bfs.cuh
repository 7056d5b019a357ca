// bfs.cuh -- SURVEY 8(f) row 4: distances on the device-resident CSR.
//
// The reference computes path-to-path distances with one multi_source_dijkstra_path_length per path on
// the NetworkX graph parse_gfa(build_graph=True) builds without a weight tag (analysis.py:219-222, 236-240):
// every edge counts 1, so the lengths are hop counts of a multi-source BFS over the out-neighbours -- the
// rows of the CSR that the matrix path leaves in HBM (asymmetric build for the default DiGraph, undirected
// build for nx.Graph).  One co-resident gang of CTAs (cooperative launch) runs all levels of a search:
// expand the frontier queue, grid barrier, swap queues -- no launch and no host round trip per level,
// which matters because sequence graphs are near-linear and searches from few sources are thousands of
// levels deep.
#pragma once
#include "common.cuh"

namespace g2n {

struct BfsCtl {
    u64 arrived;   // monotone arrival counter of the grid barrier
    u32 size[2];   // entries in the two frontier queues
    u32 depth;     // levels expanded (diagnostic)
    u32 pad;
};

__device__ __forceinline__ void bfs_barrier(BfsCtl* c, u64& phase)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        phase += gridDim.x;
        __threadfence();
        atomicAdd((unsigned long long*)&c->arrived, 1ull);
        while (ld_volatile_u64(&c->arrived) < phase) {}
        __threadfence();
    }
    __syncthreads();
}

// level[] = -1 everywhere (memset), ctl zeroed; sources get level 0 and form the first frontier
__global__ void __launch_bounds__(256) k_bfs_seed(const int32_t* __restrict__ sources, u64 n_src, u32 n, int32_t* __restrict__ level,
                                                   u32* __restrict__ q0, BfsCtl* __restrict__ ctl)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += (u64)gridDim.x * blockDim.x) {
        const u32 s = (u32)sources[i];
        if (s < n && atomicCAS(&level[s], -1, 0) == -1) q0[atomicAdd(&ctl->size[0], 1u)] = s;
    }
}

__global__ void __launch_bounds__(256) k_bfs_gang(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int32_t* __restrict__ level,
                                                   u32* __restrict__ q0, u32* __restrict__ q1, BfsCtl* __restrict__ ctl)
{
    const u32 gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    u64 phase = 0;
    u32 cur = 0;
    int32_t depth = 0;
    while (true) {
        const u32 nf = ld_volatile_u32(&ctl->size[cur]);
        if (nf == 0) break;
        const u32* q = cur ? q1 : q0;
        u32* qn = cur ? q0 : q1;
        for (u32 i = gtid; i < nf; i += gthreads) {
            const u32 u = q[i];
            const int32_t a = indptr[u], b = indptr[u + 1];
            for (int32_t e = a; e < b; e++) {
                const int32_t v = indices[e];
                if (level[v] == -1 && atomicCAS(&level[v], -1, depth + 1) == -1) qn[atomicAdd(&ctl->size[cur ^ 1], 1u)] = (u32)v;
            }
        }
        bfs_barrier(ctl, phase);            // the next frontier is complete
        if (gtid == 0) { ctl->size[cur] = 0; ctl->depth = (u32)depth + 1; }  // this queue is filled again two levels on
        bfs_barrier(ctl, phase);            // ... and nobody pushes into it before the reset is visible
        cur ^= 1;
        depth++;
    }
}

// {min level, sum of levels, reachable count} over a node list (duplicates count as often as they occur)
__global__ void __launch_bounds__(256) k_levels_reduce(const int32_t* __restrict__ level, const int32_t* __restrict__ nodes, u64 n_nodes, u32 n,
                                                        long long* __restrict__ out)
{
    long long mn = 0x7fffffffffffffffLL, sum = 0, cnt = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += (u64)gridDim.x * blockDim.x) {
        const u32 v = (u32)nodes[i];
        if (v >= n) continue;
        const int32_t l = level[v];
        if (l < 0) continue;
        mn = l < mn ? l : mn;
        sum += l;
        cnt++;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const long long o = __shfl_xor_sync(0xffffffffu, mn, d);
        mn = o < mn ? o : mn;
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicMin(&out[0], mn);
        atomicAdd((unsigned long long*)&out[1], (unsigned long long)sum);
        atomicAdd((unsigned long long*)&out[2], (unsigned long long)cnt);
    }
}

}  // namespace g2n
