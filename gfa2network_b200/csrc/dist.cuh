// dist.cuh -- device side of the multi-GPU build (one process per GPU; the host moves the exchange
// buffers between ranks with NCCL, see gfa2network_b200/dist.py and SURVEY.md 8(e)).
//
//   phase 1  every rank tokenizes its own byte range (k_tokenize, local table, local `order`)
//   phase 2  k_dist_export: local distinct keys + their first local order  -> all-gather ->
//            k_dist_insert: every rank builds the same global table keyed by the same bytes, keeping the
//            global first appearance (records before the rank + records before the tile + index in tile);
//            IDs by the same bitmap ranking as on one GPU; k_dist_localmap: local slot -> global ID
//   phase 3  k_dist_dest_count / k_dist_dest_scatter: row entries bucketed by owner(row) -> all-to-all ->
//            k_pairs_count / k_pairs_scatter + the single-GPU row sort: each rank builds its CSR slab
// Restricted to inline (<= 15 byte) keys and unweighted builds; the host mirror refuses anything else.
#pragma once
#include "rowsort.cuh"

namespace g2n {

struct DistKey {
    u64 k0, k1;
    u64 order;  // local order (export) -- rewritten to the global order by the receiver
    u64 pad;
};

// local distinct keys (any order)
__global__ void __launch_bounds__(256) k_dist_export(const TKey* __restrict__ tkeys, const u64* __restrict__ tfirst, u32 cap,
                                                      DistKey* __restrict__ out, u32* __restrict__ counter)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        const TKey k = tkeys[i];
        if (k.x == 0 && k.y == 0) continue;
        const u32 j = atomicAdd(counter, 1u);
        DistKey d;
        d.k0 = k.x; d.k1 = k.y; d.order = ~tfirst[i]; d.pad = 0;
        out[j] = d;
    }
}

struct DistMergeParams {
    const uint8_t* keys;       // rank s: DistKey[n_keys[s]] at keys + s * key_stride (bytes)
    const uint8_t* tile_base;  // rank s: u64[n_tiles + 1] at tile_base + s * tile_stride (bytes): exclusive scan of (n_rec << 32 | n_edge)
    u64 key_stride, tile_stride;
    u64 n_keys[8];
    u64 rec_base[8];         // records before each rank
    int world;
};

// insert every rank's keys with their GLOBAL order = global record ordinal << 2 | sub-rank
__global__ void __launch_bounds__(256) k_dist_insert(const ScanParams P, const DistMergeParams D)
{
    u32 claimed = 0;
    for (int s = 0; s < D.world; s++) {
        const DistKey* keys = reinterpret_cast<const DistKey*>(D.keys + (u64)s * D.key_stride);
        const u64* tb = reinterpret_cast<const u64*>(D.tile_base + (u64)s * D.tile_stride);
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < D.n_keys[s]; i += (u64)gridDim.x * blockDim.x) {
            const DistKey d = keys[i];
            const u64 tile = d.order >> 12;
            const u64 rec = D.rec_base[s] + (tb[tile] >> 32) + ((d.order >> 2) & 1023u);
            table_probe(P, d.k0, d.k1, (rec << 2) | (d.order & 3u), claimed);
        }
    }
    if (claimed) atomicAdd(&P.cnt->n_keys, claimed);
}

// global table: order already is the bit index
__global__ void __launch_bounds__(256) k_dist_mark(const TKey* __restrict__ tkeys, const u64* __restrict__ tfirst, u32 cap, u32* __restrict__ bitmap)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        const TKey k = tkeys[i];
        if (k.x == 0 && k.y == 0) continue;
        const u64 bit = ~tfirst[i];
        atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));
    }
}

__global__ void __launch_bounds__(256) k_dist_assign(const TKey* __restrict__ tkeys, const u64* __restrict__ tfirst, u32 cap,
                                                      const u32* __restrict__ bitmap, const u32* __restrict__ wprefix,
                                                      u32* __restrict__ slot_id, u32* __restrict__ id2slot, u32* __restrict__ name_len)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        const TKey k = tkeys[i];
        if (k.x == 0 && k.y == 0) continue;
        const u64 ob = ~tfirst[i];
        const u32 wd = (u32)(ob >> 5), bit = (u32)(ob & 31);
        const u32 id = wprefix[wd] + __popc(bitmap[wd] & ((1u << bit) - 1u));
        slot_id[i] = id;
        id2slot[id] = i;
        name_len[id] = slot_key_len(k.y);
    }
}

// local slot -> global node ID (lookup of the local key in the global table; it is always present)
__global__ void __launch_bounds__(256) k_dist_localmap(const TKey* __restrict__ ltkeys, u32 lcap, const TKey* __restrict__ gtkeys, u32 gmask,
                                                        const u32* __restrict__ gslot_id, u32* __restrict__ lslot_id)
{
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < lcap; i += gridDim.x * blockDim.x) {
        const TKey k = ltkeys[i];
        if (k.x == 0 && k.y == 0) continue;
        ProbeSeq q = probe_seq(k.x, k.y, gmask, 0);  // the sequence probe_issue / probe_finish walk (the merge inserts with bidirected = 0)
        u32 found = 0xFFFFFFFFu, visited = 0;
        while (true) {
            const u32 j = q.slot();
            const TKey a = gtkeys[j], b = gtkeys[j + 1];
            if (a.x == k.x && a.y == k.y) { found = j; break; }
            if (b.x == k.x && b.y == k.y) { found = j + 1; break; }
            if ((a.x == 0 && a.y == 0) || (b.x == 0 && b.y == 0)) break;
            if (!q.next(gmask, visited)) break;
        }
        lslot_id[i] = found == 0xFFFFFFFFu ? 0xFFFFFFFFu : gslot_id[found];
    }
}

// ---------------------------------------------------------------- phase 3
struct DistPair {  // multi-GPU builds are unweighted: no emission index travels (rowsort.cuh: Ent32)
    u32 entry;  // minor << 1 | dir
    u32 major;
};

__global__ void __launch_bounds__(256) k_dist_dest_count(const EmitParams E, int sym, int csc, u32 rows_per, u32* __restrict__ dest_cnt)
{
    __shared__ u32 s_cnt[8];
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for_each_edge(E, [&](u32, u32 t0, const u32 (&id)[4]) {
        record_entries(id, E.tpe, t0, sym, csc, [&](u32 major, u32, u32, u32) { atomicAdd(&s_cnt[major / rows_per], 1u); });
    });
    __syncthreads();
    if (threadIdx.x < 8 && s_cnt[threadIdx.x]) atomicAdd(&dest_cnt[threadIdx.x], s_cnt[threadIdx.x]);
}

// dest_off[d] = first send-buffer index of destination d; dest_cur[d] = entries reserved so far.
// One warp per tile: count the tile's entries per destination in shared memory, reserve the ranges
// with one global atomic per destination, then write (requires edge_slots to hold node IDs already).
__global__ void __launch_bounds__(256) k_dist_dest_scatter(const EmitParams E, int sym, int csc, u32 rows_per, int world,
                                                            const u32* __restrict__ dest_off, u32* __restrict__ dest_cur,
                                                            DistPair* __restrict__ send)
{
    __shared__ u32 s_cnt[8][8];
    __shared__ u32 s_base[8][8];
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    if (!E.ds->ok) return;
    for (u32 tile = warp; tile < E.n_tiles; tile += n_warps) {
        const TileInfo ti = E.tile_info[tile];
        if (ti.n_edge == 0) continue;
        const u32 e0 = (u32)E.tile_base[tile];
        if (lane < 8) s_cnt[wid][lane] = 0;
        __syncwarp();
        for (int pass = 0; pass < 2; pass++) {
            for (u32 j = lane; j < ti.n_edge; j += 32) {
                const u32* sl = E.edge_slots + (u64)(ti.edge_alloc + j) * E.slots_per_edge;
                u32 id[4] = {sl[0], sl[1], 0, 0};
                if (E.slots_per_edge == 4) { id[2] = sl[2]; id[3] = sl[3]; }
                record_entries(id, E.tpe, (e0 + j) * (u32)E.tpe, sym, csc, [&](u32 major, u32 minor, u32 dir, u32 t) {
                    const u32 d = major / rows_per;
                    const u32 k = atomicAdd(&s_cnt[wid][d], 1u);
                    if (pass) {
                        DistPair p;
                        p.entry = Ent32::make(minor, dir, 0u);
                        p.major = major;
                        send[s_base[wid][d] + k] = p;
                    }
                });
            }
            __syncwarp();
            if (pass == 0 && lane < (u32)world) {
                const u32 c = s_cnt[wid][lane];
                s_base[wid][lane] = dest_off[lane] + (c ? atomicAdd(&dest_cur[lane], c) : 0u);
                s_cnt[wid][lane] = 0;
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(256) k_pairs_count(const DistPair* __restrict__ pairs, u64 n, u32 row0, u32* __restrict__ cnt)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        atomicAdd(&cnt[pairs[i].major - row0], 1u);
}

// multi-GPU builds are unweighted: the slab keeps 32-bit entries (minor << 1 | dir), see Ent32
__global__ void __launch_bounds__(256) k_pairs_scatter(const DistPair* __restrict__ pairs, u64 n, u32 row0, u32* __restrict__ cursor,
                                                        u32* __restrict__ entries)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const DistPair p = pairs[i];
        entries[atomicAdd(&cursor[p.major - row0], 1u)] = p.entry;
    }
}

}  // namespace g2n
