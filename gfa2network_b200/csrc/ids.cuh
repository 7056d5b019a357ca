// ids.cuh -- K2b: first-appearance ranking -> node IDs and the node-name table; K3: COO emission.
//
// Node IDs reproduce the reference's `node2idx[n] = len(node2idx)` (builders.py:194-198, 219-221):
// every key's slot holds the minimum order (tile, record index in tile, sub-rank) over all its
// mentions; with the exclusive scan of the per-tile record counts that is a global record ordinal, one
// bit per (record, sub-rank) marks the positions that are first appearances, and a key's ID is the
// number of marked bits below its own.  Triplet order follows add_mat_edge (builders.py:218-234).
#pragma once
#include "tokenize.cuh"
#include "tokenize_slow.cuh"

namespace g2n {

// tile_base[t] = (records before tile t) << 32 | (edge records before tile t)
__device__ __forceinline__ u32 order_bit(const u64* __restrict__ tile_base, u64 order)
{
    const u32 tile = (u32)(order >> 12);
    const u32 rec = (u32)(tile_base[tile] >> 32) + (u32)((order >> 2) & 1023u);
    return (rec << 2) | (u32)(order & 3u);
}

// Capacities the host sized this build's buffers for (exact after a host round trip, learnt from the
// previous build of the same shape otherwise).
struct SizeCaps {
    u32 n_cap, E_cap, R_cap;
    u32 n_tiles;
    int tpe, sym;
    int slow_ran;  // k_tokenize_slow was launched (deferred lines are only legal if it was)
};

// One thread: counters -> DevSizes.  ok = 0 (and every size 0) if anything needs the host: an error
// record, a full table / list, a hash collision, or more nodes / edges / records than the buffers hold.
__global__ void k_sizes(const Counters* __restrict__ cnt, const u64* __restrict__ tile_base, const SizeCaps c, DevSizes* __restrict__ ds)
{
    if (blockIdx.x || threadIdx.x) return;
    const u64 tot = c.n_tiles ? tile_base[c.n_tiles] : 0ull;  // records << 32 | edge records
    const u64 R = tot >> 32, E = tot & 0xFFFFFFFFull, n = cnt->n_keys;
    const u64 T = E * (u64)c.tpe, M = c.sym ? 2 * T : T;
    const u32 fatal = CF_TABLE_FULL | CF_EDGE_FULL | CF_LONG_FULL | CF_DEFER_FULL;
    bool ok = !(cnt->flags & fatal) && !cnt->collision && cnt->first_error_inv == 0;
    ok = ok && (cnt->n_defer == 0 || c.slow_ran) && cnt->edge_alloc <= c.E_cap;
    ok = ok && n <= c.n_cap && E <= c.E_cap && R <= c.R_cap && n <= 0x7FFFFFFFull && R < (1ull << 30) - 1 && M < 0xFFFFFFF0ull;
    DevSizes s;
    memset(&s, 0, sizeof(s));
    if (ok) {
        s.n = (u32)n; s.E = (u32)E; s.R = (u32)R; s.words = (u32)((4 * R + 31) / 32 + 1); s.wgroups = (s.words + BM_GROUP - 1) / BM_GROUP;
        s.T = (u32)T; s.M = (u32)M; s.ok = 1; s.nnz = (u32)T; s.rows = (u32)n;
    }
    *ds = s;
}

// host-known sizes (slab builds, caller-provided COO): rows and entries only
__global__ void k_set_sizes(DevSizes* __restrict__ ds, u32 rows, u32 M)
{
    if (blockIdx.x || threadIdx.x) return;
    DevSizes s;
    memset(&s, 0, sizeof(s));
    s.n = rows; s.rows = rows; s.M = M; s.T = M; s.ok = 1;
    *ds = s;
}

// bit (global record ordinal << 2 | sub-rank) of `bitmap` set for every occupied slot
// streaming read of one slot: {k0, k1, first, rep}
__device__ __forceinline__ void ld_slot_stream(const Slot* s, u64 (&v)[4])
{
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(s));
}

__global__ void __launch_bounds__(256) k_mark_first(const Slot* __restrict__ slots, u32 cap,
                                                     const u64* __restrict__ tile_base, u32* __restrict__ bitmap, const DevSizes* __restrict__ ds)
{
    if (!ds->ok) return;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        u64 v[4];
        ld_slot_stream(&slots[i], v);
        if (v[0] == 0 && v[1] == 0) continue;
        const u32 bit = order_bit(tile_base, ~v[2]);
        G2N_CHECK((bit >> 5) < ds->words);
        atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));
    }
}

__device__ __forceinline__ u32 slot_key_len(u64 k1)
{
    const u32 top = (u32)(k1 >> 56);
    return top == 0xFF ? (u32)((k1 >> 32) & 0xFFFFFF) : top - 1;
}

// length of node id's name, from its table slot: only the node-name paths need it (g2n_names_bytes / g2n_fetch_names, the node
// map), so it is not materialised by the build -- one random 4-byte store per node less in k_assign_ids / k_dx_send_rank
struct LoadNameLen {
    const Slot* slots;
    const u32* id2slot;
    __device__ __forceinline__ u64 operator()(u64 i) const { return (u64)slot_key_len(slots[id2slot[i]].k1); }
    __device__ __forceinline__ u64 peek(u64 i) const { return (*this)(i); }
};

// slot -> id ; id -> slot ; id -> entries of the node's row (when the tokenizer counted them)
__global__ void __launch_bounds__(256) k_assign_ids(const Slot* __restrict__ slots, u32 cap,
                                                     const u64* __restrict__ tile_base, const u32* __restrict__ bitmap,
                                                     const u32* __restrict__ wprefix, u32* __restrict__ slot_id,
                                                     u32* __restrict__ id2slot, const DevSizes* __restrict__ ds,
                                                     const u32* __restrict__ slot_cnt, u32* __restrict__ rowcnt)
{
    if (!ds->ok) return;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        u64 v[4];
        ld_slot_stream(&slots[i], v);
        if (v[0] == 0 && v[1] == 0) continue;
        const u32 ob = order_bit(tile_base, ~v[2]);
        const u32 id = bitmap_rank(bitmap, wprefix, ob);
        G2N_CHECK((ob >> 5) < ds->words && id < ds->n);
        slot_id[i] = id;
        id2slot[id] = i;
        if (slot_cnt) rowcnt[id] = slot_cnt[i];
    }
}

// names[name_off[id] ...] = key bytes of node id (builders.py:284-288 node list, raw bytes)
__global__ void __launch_bounds__(256) k_gather_names(const Slot* __restrict__ slots,
                                                       const u32* __restrict__ id2slot, const u64* __restrict__ name_off, u32 n,
                                                       const uint8_t* __restrict__ text, const LongDesc* __restrict__ longs,
                                                       uint8_t* __restrict__ names)
{
    for (u32 id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
        const u32 slot = id2slot[id];
        const Slot k = slots[slot];
        uint8_t* dst = names + name_off[id];
        const u32 top = (u32)(k.k1 >> 56);
        if (top != 0xFF) {
            const u32 L = top - 1;
            for (u32 j = 0; j < L; j++) dst[j] = (uint8_t)((j < 8 ? k.k0 >> (8 * j) : k.k1 >> (8 * (j - 8))) & 0xFF);
        } else {
            const LongDesc d = longs[k.rep - 1];
            const u32 L = d.base_len + (d.has_ori ? 1 + d.ori_len : 0);
            for (u32 j = 0; j < L; j++) dst[j] = long_byte(text, d, j);
        }
    }
}

// ---------------------------------------------------------------- node map as text (utils.py:108-114)
// "<index>\t<name>\n" per node, in ID order: the file save_node_map writes next to the matrix.
__device__ __forceinline__ u32 dec_digits(u32 v)
{
    u32 d = 1;
    while (v >= 10) { v /= 10; d++; }
    return d;
}
struct LoadTsvLen {  // bytes of node i's line
    LoadNameLen name_len;
    __device__ __forceinline__ u64 operator()(u64 i) const { return (u64)dec_digits((u32)i) + name_len(i) + 2; }
    __device__ __forceinline__ u64 peek(u64 i) const { return (*this)(i); }
};
__global__ void __launch_bounds__(256) k_tsv_write(const Slot* __restrict__ slots,
                                                    const u32* __restrict__ id2slot, const u64* __restrict__ line_off, u32 n,
                                                    const uint8_t* __restrict__ text, const LongDesc* __restrict__ longs,
                                                    uint8_t* __restrict__ out)
{
    for (u32 id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
        uint8_t* dst = out + line_off[id];
        const u32 nd = dec_digits(id);
        u32 v = id;
        for (u32 k = nd; k-- > 0;) { dst[k] = (uint8_t)('0' + v % 10); v /= 10; }
        dst += nd;
        *dst++ = '\t';
        const u32 slot = id2slot[id];
        const Slot k = slots[slot];
        const u32 top = (u32)(k.k1 >> 56);
        u32 L;
        if (top != 0xFF) {
            L = top - 1;
            for (u32 j = 0; j < L; j++) dst[j] = (uint8_t)((j < 8 ? k.k0 >> (8 * j) : k.k1 >> (8 * (j - 8))) & 0xFF);
        } else {
            const LongDesc d = longs[k.rep - 1];
            L = d.base_len + (d.has_ori ? 1 + d.ori_len : 0);
            for (u32 j = 0; j < L; j++) dst[j] = long_byte(text, d, j);
        }
        dst[L] = '\n';
    }
}

// ---------------------------------------------------------------- dtype helpers
// The reference casts the float64 weight list to `dtype` when the COO matrix is built
// (builders.py:281 -> scipy/_coo.py), i.e. BEFORE duplicates are summed.
struct BoolT { uint8_t v; };

template <typename T> __device__ __forceinline__ T cast_weight(double w) { return (T)w; }
template <> __device__ __forceinline__ int8_t cast_weight<int8_t>(double w) { return (int8_t)(int)w; }
template <> __device__ __forceinline__ BoolT cast_weight<BoolT>(double w) { BoolT b; b.v = (w != 0.0) ? 1 : 0; return b; }

template <typename T> __device__ __forceinline__ T add_t(T a, T b) { return a + b; }
template <> __device__ __forceinline__ int8_t add_t<int8_t>(int8_t a, int8_t b) { return (int8_t)((int)a + (int)b); }
template <> __device__ __forceinline__ BoolT add_t<BoolT>(BoolT a, BoolT b) { BoolT r; r.v = a.v | b.v; return r; }  // npy_bool_wrapper +
template <typename T> __device__ __forceinline__ bool lt_t(T a, T b) { return a < b; }
template <> __device__ __forceinline__ bool lt_t<BoolT>(BoolT a, BoolT b) { return a.v < b.v; }
template <typename T> __device__ __forceinline__ bool nz_t(T a) { return a != (T)0; }
template <> __device__ __forceinline__ bool nz_t<BoolT>(BoolT a) { return a.v != 0; }
template <typename T> __device__ __forceinline__ T zero_t() { return (T)0; }
template <> __device__ __forceinline__ BoolT zero_t<BoolT>() { BoolT b; b.v = 0; return b; }

// ---------------------------------------------------------------- K3: emission
// Edge records live in edge_slots in per-tile ranges (tile_info[t].edge_alloc, claimed in arrival
// order); file order is tile order, so the emission index of tile t's j-th edge record is
// (edges before tile t) + j.  One warp walks one tile.
struct EmitParams {
    u32* edge_slots;       // per stored edge record: table slots of its endpoints, or -- once ids_ready -- node IDs
    const double* edge_w;  // NULL when no weight tag: every weight is 1.0
    const u32* slot_id;
    const TileInfo* tile_info;
    const u64* tile_base;  // exclusive scan of (n_rec << 32 | n_edge) over tiles
    u32 n_tiles;
    int slots_per_edge;    // 2 | 4
    int tpe;               // triplets per edge record: 1 (graph_directed) | 2 | 4
    int ids_ready;         // edge_slots already hold node IDs (translated in place by an earlier pass)
    int write_ids;         // translate in place during this pass
    const DevSizes* ds;    // ds->ok == 0: the build was abandoned on the device, touch nothing
};

#define EM_UNROLL 4

// Walks the edge records in emission order, EM_UNROLL records per lane in flight, and calls
// f(stored, t0, id[4]) where t0 is the emission index of the record's first triplet.  The triplets of a
// record are (builders.py:222-234): (a,b) [, (b,a)] [, (c,d), (d,c)] with a = id[0], b = id[1],
// c = id[2] = v:flip(ot), d = id[3] = u:flip(of).
template <class F>
__device__ __forceinline__ void for_each_edge(const EmitParams& E, F f)
{
    const u32 lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    if (!E.ds->ok) return;
    for (u32 tile = warp; tile < E.n_tiles; tile += n_warps) {
        const TileInfo ti = E.tile_info[tile];
        if (ti.n_edge == 0) continue;
        const u32 e0 = (u32)E.tile_base[tile];  // edge records before this tile
        for (u32 j0 = 0; j0 < ti.n_edge; j0 += 32 * EM_UNROLL) {
            u32 id[EM_UNROLL][4];
#pragma unroll
            for (int u = 0; u < EM_UNROLL; u++) {
                const u32 j = j0 + u * 32 + lane;
                if (j < ti.n_edge) {
                    const u32* s = E.edge_slots + (u64)(ti.edge_alloc + j) * E.slots_per_edge;
                    if (E.slots_per_edge == 4) {
                        const uint4 q = *reinterpret_cast<const uint4*>(s);
                        id[u][0] = q.x; id[u][1] = q.y; id[u][2] = q.z; id[u][3] = q.w;
                    } else {
                        const uint2 q = *reinterpret_cast<const uint2*>(s);
                        id[u][0] = q.x; id[u][1] = q.y; id[u][2] = 0; id[u][3] = 0;
                    }
                }
            }
            if (!E.ids_ready) {
#pragma unroll
                for (int u = 0; u < EM_UNROLL; u++) {
                    const u32 j = j0 + u * 32 + lane;
                    if (j < ti.n_edge) {
                        id[u][0] = E.slot_id[id[u][0]];
                        id[u][1] = E.slot_id[id[u][1]];
                        if (E.slots_per_edge == 4) { id[u][2] = E.slot_id[id[u][2]]; id[u][3] = E.slot_id[id[u][3]]; }
                    }
                }
                if (E.write_ids) {
#pragma unroll
                    for (int u = 0; u < EM_UNROLL; u++) {
                        const u32 j = j0 + u * 32 + lane;
                        if (j < ti.n_edge) {
                            u32* s = E.edge_slots + (u64)(ti.edge_alloc + j) * E.slots_per_edge;
                            if (E.slots_per_edge == 4) *reinterpret_cast<uint4*>(s) = make_uint4(id[u][0], id[u][1], id[u][2], id[u][3]);
                            else *reinterpret_cast<uint2*>(s) = make_uint2(id[u][0], id[u][1]);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < EM_UNROLL; u++) {
                const u32 j = j0 + u * 32 + lane;
                if (j < ti.n_edge) f(ti.edge_alloc + j, (e0 + j) * (u32)E.tpe, id[u]);
            }
        }
    }
}

// k-th triplet of a record
__device__ __forceinline__ void record_triplet(const u32 (&id)[4], int k, u32& r, u32& c)
{
    const u32 a = id[k & 2], b = id[(k & 2) + 1];
    if (k & 1) { r = b; c = a; } else { r = a; c = b; }
}

// ---------------------------------------------------------------- edge list as text (cli.py:267-281)
// `export --format edge-list`: one line "<u>\t<v>\n" per L / E / C record in file order, where u and v are the
// record's endpoint keys -- exactly the node keys of a build with the same --bidirected flag (u:of, v:ot).
// Two passes over the edge records (line lengths, exclusive scan, bytes), same emission order as the COO.
struct NameSrc {
    const Slot* slots;
    const LongDesc* longs;
    const uint8_t* text;
    const u32* id2slot;  // non-NULL: edge_slots already hold node IDs
    __device__ __forceinline__ u32 slot_of(u32 v) const { return id2slot ? id2slot[v] : v; }
    __device__ __forceinline__ u32 len(u32 slot) const
    {
        const Slot k = slots[slot];
        if ((u32)(k.k1 >> 56) != 0xFF) return slot_key_len(k.k1);
        const LongDesc d = longs[k.rep - 1];
        return d.base_len + (d.has_ori ? 1 + d.ori_len : 0);
    }
    __device__ __forceinline__ u32 copy(u32 slot, uint8_t* dst) const
    {
        const Slot k = slots[slot];
        if ((u32)(k.k1 >> 56) != 0xFF) {
            const u32 L = slot_key_len(k.k1);
            for (u32 j = 0; j < L; j++) dst[j] = (uint8_t)((j < 8 ? k.k0 >> (8 * j) : k.k1 >> (8 * (j - 8))) & 0xFF);
            return L;
        }
        const LongDesc d = longs[k.rep - 1];
        const u32 L = d.base_len + (d.has_ori ? 1 + d.ori_len : 0);
        for (u32 j = 0; j < L; j++) dst[j] = long_byte(text, d, j);
        return L;
    }
};

__global__ void __launch_bounds__(256) k_edge_line_len(const EmitParams E, const NameSrc N, u32* __restrict__ line_len)
{
    for_each_edge(E, [&](u32, u32 t0, const u32 (&id)[4]) {
        line_len[t0 / (u32)E.tpe] = N.len(N.slot_of(id[0])) + N.len(N.slot_of(id[1])) + 2u;
    });
}

__global__ void __launch_bounds__(256) k_edge_line_write(const EmitParams E, const NameSrc N, const u64* __restrict__ line_off, uint8_t* __restrict__ out)
{
    for_each_edge(E, [&](u32, u32 t0, const u32 (&id)[4]) {
        uint8_t* dst = out + line_off[t0 / (u32)E.tpe];
        dst += N.copy(N.slot_of(id[0]), dst);
        *dst++ = '\t';
        dst += N.copy(N.slot_of(id[1]), dst);
        *dst = '\n';
    });
}

// raw COO in emission order: row, col (int32) and data (dtype)
template <typename T>
__global__ void __launch_bounds__(256) k_emit_coo(const EmitParams E, int32_t* __restrict__ row, int32_t* __restrict__ col, T* __restrict__ data)
{
    for_each_edge(E, [&](u32 stored, u32 t0, const u32 (&id)[4]) {
        const T w = cast_weight<T>(E.edge_w ? E.edge_w[stored] : 1.0);
        for (int k = 0; k < E.tpe; k++) {
            u32 r, c;
            record_triplet(id, k, r, c);
            row[t0 + k] = (int32_t)r;
            col[t0 + k] = (int32_t)c;
            data[t0 + k] = w;
        }
    });
}

}  // namespace g2n
