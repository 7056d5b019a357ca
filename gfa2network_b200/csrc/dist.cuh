// dist.cuh -- device side of the multi-GPU build: one process and one handle per GPU.
//
// Nothing here goes through the host or through a collective library: every exchange is a kernel that
// WRITES STRAIGHT INTO THE PEER'S MEMORY over NVLink (peer pointers from CUDA IPC, or plain pointers when
// several logical ranks share one device); the last CTA of that kernel to finish publishes a per-source
// header and an epoch flag in the peer's control block (release, system scope), and the consuming kernel
// of the peer spins on those flags (acquire, system scope) before it touches the data.  A whole build is queued on the stream
// without a host round trip; the host synchronises once, at the end.
//
//   stage 0  k_tokenize (this rank's byte range: local table, local first-appearance order)
//            k_dx_export     distinct local keys + first local order  --> owner = hash(key) % world     [x1]
//   stage 1  k_dx_insert     owner table: min over (source rank, local order) = global first appearance
//            k_dx_reply_first  owner -> every source: "your mention is the global first"               [x2]
//   stage 2  k_dx_mark + scan  bitmap of this shard's global firsts -> rank among them (builders.py:194-198:
//                            ID = number of earlier first appearances)
//            k_dx_send_rank  first source -> owner: rank inside the shard; count of firsts in the header [x3]
//   stage 3  k_dx_reply_ids  owner -> every source: ID = (firsts of lower shards) + rank               [x4]
//   stage 4  k_dx_localmap   local slot -> global node ID
//            k_dx_entries    row entries --> owner(row) = row / ceil(n / world)                         [x5]
//            (with a weight tag: k_dxw_count / scan / k_dxw_scatter, entries in emission order + weight)
//   stage 5  k_dx_slab_sizes, k_pairs_count / scan / k_pairs_scatter + the single-GPU row kernels:
//            duplicate sum / max(S, S^T) -> this rank's CSR slab (builders.py:279-283, utils.py:55)
//            status word to every rank                                                                  [x6]
//   stage 6  k_dx_final      the build's verdict, the same on every rank: a capacity miss anywhere repeats the
//                            build everywhere
// Per-rank work is proportional to the rank's own shard (keys are hash-partitioned, nobody holds the
// whole dictionary).  Node names stay with the shard of their first appearance: shard s names the
// consecutive IDs [idbase[s], idbase[s] + firsts[s]).
// Keys longer than 15 bytes travel as
// their 128-bit tagged hash (table.cuh: make_key, same seed on every rank); their bytes stay with the shards.
#pragma once
#include "rowsort.cuh"

namespace g2n {

#define DX_MAXW 8
#define DX_NEX 6  // exchanges per build

// `bad` bits: anything set on any rank repeats the build with a host-planned (non-speculative) pass
#define DXB_TOK 1u       // the speculative tokenizer pass did not fit its buffers / needs the host
#define DXB_KEYS 2u      // a (source, owner) key segment overflowed
#define DXB_PAIRS 4u     // a (source, owner) entry segment overflowed
#define DXB_RECV 8u      // more entries received than the slab buffers hold
#define DXB_ROWS 16u     // more rows than the slab buffers hold
#define DXB_GTABLE 32u   // owner table full
#define DXB_TIMEOUT 64u  // a peer's signal did not arrive
#define DXB_RANGE 128u   // an entry for a row outside this slab (only after another failure)
#define DXB_SHAPE 256u   // host: this rank's input does not match the remembered plan

struct DxHdr {  // written by the source into the receiver's control block, before the epoch flag
    u64 count;  // x1: keys in the segment; x5: entries in the segment
    u64 bad;    // everything bad the source knew when it signalled
    u64 aux;    // x3: global firsts of the source's shard
    u64 pad;
};

struct DxCtl {  // device memory of one rank, written by its peers
    u32 sig[DX_NEX][DX_MAXW];  // epoch of the latest signal of exchange e from source s (monotone)
    DxHdr hdr[DX_NEX][DX_MAXW];
};

struct DxLocal {  // zeroed at the start of every build; never written by peers
    u32 bad;
    u32 cur_keys[DX_MAXW];   // keys sent to each owner so far
    u32 cur_pairs[DX_MAXW];  // entries sent to each row owner so far
    u32 idbase[DX_MAXW + 1];
    u32 n_first, n_global, rows_per, row0, n_rows, n_recv;
    u32 seg_off[DX_MAXW + 1];  // received entries: exclusive prefix over sources
    u32 n_keys, n_records, n_edges;  // this shard (copied from DevSizes before the slab overwrites it)
    u32 final_bad;
    u32 done[DX_NEX];  // CTAs of the producing kernel of exchange e that have finished
    u32 pad[2];
};

struct DxLayout {  // the exchange arena of one rank (same layout on every rank)
    u64 kcap, pcap;  // elements per (source, destination) segment
    u64 off_key, off_ord, off_first, off_rank, off_id, off_pair, bytes;
    u64 pair_bytes;  // 8 (DistPair) | 16 (DistPairW, weighted builds)
};

struct DxPeers {
    uint8_t* arena[DX_MAXW];
    DxCtl* ctl[DX_MAXW];
    int rank, world;
    u32 epoch;
};

struct DistPair {  // unweighted builds: no emission index travels (rowsort.cuh: Ent32)
    u32 entry;  // minor << 1 | dir
    u32 major;
};
struct DistPairW {  // weighted builds: the POSITION in the (source, owner) segment is the emission order
    u32 entry;  // minor << 1 | dir
    u32 major;
    double w;
};

__device__ __forceinline__ u32 ld_acquire_sys_u32(const u32* p)
{
    u32 v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(u32* p, u32 v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ u32 dx_owner(u64 k0, u64 k1, int world)
{
    return (u32)((mix64(k0 ^ mix64(k1 + 0x9e3779b97f4a7c15ULL)) >> 33) % (u32)world);
}

// Every CTA of a consuming kernel: wait until all sources have signalled exchange `e` of this epoch.
// Bounded: a peer that never signals raises DXB_TIMEOUT instead of hanging the device.
__device__ __forceinline__ void dx_wait(const DxCtl* my, int e, const DxPeers& X, DxLocal* loc)
{
    if (threadIdx.x < (u32)X.world) {
        const u32* f = &my->sig[e][threadIdx.x];
        u32 spins = 0;
        while ((int)(ld_acquire_sys_u32(f) - X.epoch) < 0) {
            __nanosleep(200);
            if (++spins > 30000000u) { atomicOr(&loc->bad, DXB_TIMEOUT); break; }  // >= 6 s: a peer that is merely slow (first-touch allocations) is waited for
        }
    }
    __syncthreads();
}

// Publishes this rank's header of exchange `e` and the epoch flag in every peer's control block.
// Called by the first `world` threads of ONE CTA once all of this rank's writes of the exchange are ordered
// before it (dx_tail_signal).  st.release.sys: the header (and, cumulatively, everything ordered before)
// is visible to whoever acquires the flag.
__device__ __forceinline__ void dx_publish(const DxPeers& X, int e, DxLocal* loc, const u32* counts, u64 cap, u64 aux)
{
    const u32 d = threadIdx.x;
    if (d >= (u32)X.world) return;
    DxHdr h;
    u64 c = counts ? *(const volatile u32*)&counts[d] : 0u;
    if (c > cap) c = cap;
    h.count = c;
    h.bad = *(const volatile u32*)&loc->bad;
    h.aux = aux;
    h.pad = 0;
    DxCtl* pc = X.ctl[d];
    pc->hdr[e][X.rank] = h;
    st_release_sys_u32(&pc->sig[e][X.rank], X.epoch);
}

// End of a producing kernel (every thread of every CTA gets here): the LAST CTA to arrive signals.
// Each CTA orders its writes into peer memory before its ticket (bar.sync + release fence at system
// scope), the last CTA acquires the tickets before it publishes: no separate signal launch, and the
// peers' consumers can start while this kernel's other CTAs are already gone.
__device__ __forceinline__ void dx_tail_signal(const DxPeers& X, int e, DxLocal* loc, const u32* counts, u64 cap, u64 aux)
{
    __shared__ u32 s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        const u32 ticket = atomicAdd(&loc->done[e], 1u);
        s_last = ticket == gridDim.x - 1;
        if (s_last) asm volatile("fence.acq_rel.sys;" ::: "memory");
    }
    __syncthreads();
    if (s_last) dx_publish(X, e, loc, counts, cap, aux);
}

// ---------------------------------------------------------------- x1: local keys -> owners
// klist[owner][position in the owner's segment of this rank] = local order bit << 32 | table slot: the later passes
// of this rank (mark, send_rank, localmap) walk these compact lists next to the owners' replies, which sit at
// the same positions -- sequential reads instead of three more scans of the (half empty) table
// A CTA handles DXK_SLOTS table slots per round: its keys are grouped by owner in shared memory, the CTA's range
// in every owner's segment is reserved with ONE global atomic per owner, and consecutive threads write
// consecutive elements -- the peers receive whole 128-byte-and-larger pieces instead of one key per lane group.
#define DXK_PER 4
#define DXK_SLOTS (256 * DXK_PER)
__global__ void __launch_bounds__(256) k_dx_export(const Slot* __restrict__ slots, u32 cap,
                                                    const u64* __restrict__ tile_base, const DevSizes* __restrict__ ds,
                                                    const DxPeers X, const DxLayout L, DxLocal* __restrict__ loc, u64* __restrict__ klist)
{
    __shared__ TKey s_key[DXK_SLOTS];
    __shared__ u64 s_meta[DXK_SLOTS];  // order bit << 32 | slot
    __shared__ u32 s_cnt[DX_MAXW], s_off[DX_MAXW + 1], s_base[DX_MAXW];
    if (!ds->ok && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&loc->bad, DXB_TOK);
    const u32 lane = threadIdx.x & 31;
    for (u32 i0 = blockIdx.x * DXK_SLOTS; ds->ok && i0 < cap; i0 += gridDim.x * DXK_SLOTS) {  // cap is a multiple of 1024
        if (threadIdx.x < DX_MAXW) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        TKey k[DXK_PER];
        u64 fst[DXK_PER];
        u32 d[DXK_PER], r[DXK_PER];
#pragma unroll
        for (int u = 0; u < DXK_PER; u++) {
            const u32 i = i0 + u * 256 + threadIdx.x;
            u64 v[4];
            ld_slot_stream(&slots[i], v);
            k[u] = make_ulonglong2(v[0], v[1]);
            fst[u] = v[2];
            const bool occ = !(k[u].x == 0 && k[u].y == 0);
            d[u] = occ ? dx_owner(k[u].x, k[u].y, X.world) : 0xFFu;
            const u32 m = __match_any_sync(0xffffffffu, d[u]);
            r[u] = 0;
            if (occ) {
                const int leader = __ffs(m) - 1;
                u32 base = 0;
                if ((int)lane == leader) base = atomicAdd(&s_cnt[d[u]], (u32)__popc(m));
                r[u] = __shfl_sync(m, base, leader) + __popc(m & ((1u << lane) - 1u));
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 run = 0;
            for (int w = 0; w < X.world; w++) { s_off[w] = run; run += s_cnt[w]; }
            s_off[X.world] = run;
        }
        if (threadIdx.x < (u32)X.world) {
            const u32 c = s_cnt[threadIdx.x];
            const u32 b = c ? atomicAdd(&loc->cur_keys[threadIdx.x], c) : 0u;
            if ((u64)b + c > L.kcap) atomicOr(&loc->bad, DXB_KEYS);
            s_base[threadIdx.x] = b;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < DXK_PER; u++) {
            if (d[u] == 0xFFu) continue;
            const u32 i = i0 + u * 256 + threadIdx.x;
            const u32 q = s_off[d[u]] + r[u];
            s_key[q] = k[u];
            s_meta[q] = ((u64)order_bit(tile_base, ~fst[u]) << 32) | i;
        }
        __syncthreads();
        const u32 total = s_off[X.world];
        for (u32 q = threadIdx.x; q < total; q += 256) {
            u32 w = 0;
            while (q >= s_off[w + 1]) w++;  // staged elements are grouped by owner
            const u32 pos = s_base[w] + (q - s_off[w]);
            if (pos >= L.kcap) continue;
            uint8_t* a = X.arena[w];
            const u64 jj = (u64)X.rank * L.kcap + pos;
            const u64 meta = s_meta[q];
            reinterpret_cast<TKey*>(a + L.off_key)[jj] = s_key[q];
            reinterpret_cast<u32*>(a + L.off_ord)[jj] = (u32)(meta >> 32);
            klist[(u64)w * L.kcap + pos] = meta;
        }
        __syncthreads();
    }
    dx_tail_signal(X, 0, loc, loc->cur_keys, L.kcap, 0);
}

struct DxOwner {  // the owner's table over the keys it is responsible for
    Slot* gslots;  // first = ~min(source rank << 32 | local order)
    u32 gmask;
    u32* gslot;   // [source][position]: table slot of the received key
    u32* gpos;    // [table slot]: FIRST source of the key << 29 | its position in that source's segment
};

// headers are read after dx_wait(); volatile: never through the read-only path
__device__ __forceinline__ u32 dx_count(const DxCtl* my, int e, u32 s, u64 cap)
{
    const u64 c = *(const volatile u64*)&my->hdr[e][s].count;
    return (u32)(c < cap ? c : cap);
}
__device__ __forceinline__ u32 dx_bad(const DxCtl* my, int e, u32 s) { return (u32) * (const volatile u64*)&my->hdr[e][s].bad; }
__device__ __forceinline__ u32 dx_aux(const DxCtl* my, int e, u32 s) { return (u32) * (const volatile u64*)&my->hdr[e][s].aux; }

// The per-peer segments of an exchange as ONE index space: a kernel walks q in [0, off[world]) grid-strided and finds
// (segment, position) by a scan of <= 8 shared words -- one pass with every thread busy instead of `world` short loops in
// a row, each with its own tail of dependent memory latencies.
struct DxFlat {
    u32 off[DX_MAXW + 1];
    __device__ __forceinline__ u32 seg(u32 q) const
    {
        u32 s = 0;
        while (q >= off[s + 1]) s++;
        return s;
    }
};
// counts -> prefix (thread 0), visible to the CTA after the barrier; `unit` elements per index (rounded up per segment)
template <class CountOf>
__device__ __forceinline__ void dx_flat_init(DxFlat& F, int world, u32 unit, CountOf count_of)
{
    if (threadIdx.x == 0) {
        u32 run = 0;
        for (int s = 0; s < world; s++) {
            F.off[s] = run;
            run += (count_of((u32)s) + unit - 1) / unit;
        }
        for (int s = world; s <= DX_MAXW; s++) F.off[s] = run;
    }
    __syncthreads();
}

// ---------------------------------------------------------------- x1 consumer: owner table
__global__ void __launch_bounds__(256) k_dx_insert(const DxPeers X, const DxLayout L, const DxCtl* my, DxLocal* __restrict__ loc,
                                                    const DxOwner G, Counters* __restrict__ cnt)
{
    dx_wait(my, 0, X, loc);
    if (blockIdx.x == 0 && threadIdx.x < (u32)X.world) {
        const u32 b = dx_bad(my, 0, threadIdx.x);
        if (b) atomicOr(&loc->bad, b);
    }
    ScanParams P;
    P.slots = G.gslots;
    P.slot_cnt = nullptr;
    P.table_mask = G.gmask;
    P.cnt = cnt;
    P.bidirected = 0;
    const uint8_t* a = X.arena[X.rank];
    u32 claimed = 0;
    __shared__ DxFlat F;
    dx_flat_init(F, X.world, 1u, [&](u32 s) { return dx_count(my, 0, s, L.kcap); });
    const u32 total = F.off[X.world];
    for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const u32 s = F.seg(q), j = q - F.off[s];
        const u64 jj = (u64)s * L.kcap + j;
        const TKey k = reinterpret_cast<const TKey*>(a + L.off_key)[jj];
        const u32 o = reinterpret_cast<const u32*>(a + L.off_ord)[jj];
        const u32 slot = table_probe(P, k.x, k.y, ((u64)s << 32) | o, false, claimed);
        if (slot == 0xFFFFFFFFu) atomicOr(&loc->bad, DXB_GTABLE);
        G.gslot[jj] = slot;
    }
}

// ---------------------------------------------------------------- x2: owner -> sources, one byte per key
// four keys per thread: one 32-bit store into the peer's memory
__global__ void __launch_bounds__(256) k_dx_reply_first(const DxPeers X, const DxLayout L, const DxCtl* my, DxLocal* __restrict__ loc, const DxOwner G)
{
    __shared__ DxFlat F;
    dx_flat_init(F, X.world, 4u, [&](u32 s) { return dx_count(my, 0, s, L.kcap); });
    const u32 total = F.off[X.world];
    for (u32 qq = blockIdx.x * blockDim.x + threadIdx.x; qq < total; qq += gridDim.x * blockDim.x) {
        const u32 s = F.seg(qq), q = qq - F.off[s];
        const u32 n = dx_count(my, 0, s, L.kcap);
        const u32* gs = G.gslot + (u64)s * L.kcap;
        u32* out = reinterpret_cast<u32*>(X.arena[s] + L.off_first + (u64)X.rank * L.kcap);  // kcap is a multiple of 4
        u32 g[4];
        u64 f[4];
#pragma unroll
        for (u32 t = 0; t < 4; t++) g[t] = q * 4 + t < n ? gs[q * 4 + t] : 0xFFFFFFFFu;
#pragma unroll
        for (u32 t = 0; t < 4; t++) f[t] = g[t] != 0xFFFFFFFFu ? ~G.gslots[g[t]].first : ~0ull;
        u32 word = 0;
#pragma unroll
        for (u32 t = 0; t < 4; t++) {
            if (g[t] != 0xFFFFFFFFu && (u32)(f[t] >> 32) == s) {
                word |= 1u << (8 * t);
                G.gpos[g[t]] = (s << 29) | (q * 4 + t);  // kcap < 2^29 (g2n_dist_plan)
            }
        }
        out[q] = word;
    }
    dx_tail_signal(X, 1, loc, nullptr, 0, 0);
}

// ---------------------------------------------------------------- x2 consumer: bitmap of this shard's global firsts
__device__ __forceinline__ u32 dx_sent_count(const DxLocal* loc, u32 d, u64 kcap)
{
    const u32 c = loc->cur_keys[d];
    return (u32)(c < kcap ? c : kcap);
}

__global__ void __launch_bounds__(256) k_dx_mark(const DevSizes* __restrict__ ds, const DxPeers X, const DxLayout L, const DxCtl* my,
                                                  DxLocal* __restrict__ loc, const u64* __restrict__ klist, u32* __restrict__ bitmap)
{
    dx_wait(my, 1, X, loc);
    if (blockIdx.x == 0 && threadIdx.x < (u32)X.world) {
        const u32 b = dx_bad(my, 1, threadIdx.x);
        if (b) atomicOr(&loc->bad, b);
    }
    if (!ds->ok) return;
    const uint8_t* first = X.arena[X.rank] + L.off_first;
    __shared__ DxFlat F;
    dx_flat_init(F, X.world, 1u, [&](u32 d) { return dx_sent_count(loc, d, L.kcap); });
    const u32 total = F.off[X.world];
    for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const u32 d = F.seg(q), pos = q - F.off[d];
        if (!first[(u64)d * L.kcap + pos]) continue;
        const u32 bit = (u32)(klist[(u64)d * L.kcap + pos] >> 32);
        atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));
    }
}

// ---------------------------------------------------------------- x3: first source -> owner, rank inside the shard
// also this shard's name table: id2slot is indexed by that rank (name lengths: ids.cuh LoadNameLen, on demand)
__global__ void __launch_bounds__(256) k_dx_send_rank(const Slot* __restrict__ slots, const DevSizes* __restrict__ ds, const DxPeers X,
                                                       const DxLayout L, const u64* __restrict__ klist, const u32* __restrict__ bitmap,
                                                       const u32* __restrict__ wprefix, u32* __restrict__ id2slot,
                                                       DxLocal* __restrict__ loc)
{
    const bool ok = ds->ok != 0;
    const uint8_t* first = X.arena[X.rank] + L.off_first;
    __shared__ DxFlat F;
    dx_flat_init(F, X.world, 1u, [&](u32 d) { return ok ? dx_sent_count(loc, d, L.kcap) : 0u; });
    const u32 total = F.off[X.world];
    for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const u32 d = F.seg(q), pos = q - F.off[d];
        if (!first[(u64)d * L.kcap + pos]) continue;
        const u64 e = klist[(u64)d * L.kcap + pos];
        const u32 ob = (u32)(e >> 32), slot = (u32)e;
        const u32 r = bitmap_rank(bitmap, wprefix, ob);
        (reinterpret_cast<u32*>(X.arena[d] + L.off_rank) + (u64)X.rank * L.kcap)[pos] = r;
        id2slot[r] = slot;
    }
    // the popcount prefix ends with the number of marked bits = this shard's global firsts
    dx_tail_signal(X, 2, loc, nullptr, 0, ok ? (u64)wprefix[ds->wgroups] : 0ull);
}

// ---------------------------------------------------------------- x4: owner -> sources, the node ID of every key
__global__ void __launch_bounds__(256) k_dx_reply_ids(const DxPeers X, const DxLayout L, const DxCtl* my, DxLocal* __restrict__ loc,
                                                       const DxOwner G)
{
    __shared__ u32 s_base[DX_MAXW + 1];
    dx_wait(my, 2, X, loc);
    if (threadIdx.x == 0) {
        u32 run = 0, bad = 0;
        for (int s = 0; s < X.world; s++) {
            s_base[s] = run;
            run += dx_aux(my, 2, (u32)s);
            bad |= dx_bad(my, 2, (u32)s);
        }
        s_base[X.world] = run;
        if (blockIdx.x == 0) {
            if (bad) atomicOr(&loc->bad, bad);
            for (int s = 0; s <= X.world; s++) loc->idbase[s] = s_base[s];
            const u32 n = run, rpr = n ? (n + (u32)X.world - 1) / (u32)X.world : 1u;
            const u32 r0 = min(n, (u32)X.rank * rpr);
            loc->n_first = s_base[X.rank + 1] - s_base[X.rank];
            loc->n_global = n;
            loc->rows_per = rpr;
            loc->row0 = r0;
            loc->n_rows = min(n, r0 + rpr) - r0;
        }
    }
    __syncthreads();
    const u32* ranks = reinterpret_cast<const u32*>(X.arena[X.rank] + L.off_rank);
    __shared__ DxFlat F;
    dx_flat_init(F, X.world, 1u, [&](u32 s) { return dx_count(my, 0, s, L.kcap); });
    const u32 total = F.off[X.world];
    for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const u32 s = F.seg(q), j = q - F.off[s];
        const u32 g = G.gslot[(u64)s * L.kcap + j];
        u32 id = 0;
        if (g != 0xFFFFFFFFu) {
            const u32 v = G.gpos[g];  // first source and the key's position in its segment, in one word
            const u32 fs = v >> 29, p = v & ((1u << 29) - 1u);
            if (fs < (u32)X.world && p < L.kcap) id = s_base[fs] + ranks[(u64)fs * L.kcap + p];
        }
        (reinterpret_cast<u32*>(X.arena[s] + L.off_id) + (u64)X.rank * L.kcap)[j] = id;
    }
    dx_tail_signal(X, 3, loc, nullptr, 0, 0);
}

// ---------------------------------------------------------------- x4 consumer: local slot -> global node ID
__global__ void __launch_bounds__(256) k_dx_localmap(const DevSizes* __restrict__ ds, const DxPeers X, const DxLayout L, const DxCtl* my,
                                                      DxLocal* __restrict__ loc, const u64* __restrict__ klist, u32* __restrict__ slot_id, u32 rows_cap)
{
    dx_wait(my, 3, X, loc);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        u32 bad = 0;
        for (int s = 0; s < X.world; s++) bad |= dx_bad(my, 3, (u32)s);
        if (loc->n_rows > rows_cap) bad |= DXB_ROWS;
        if (bad) atomicOr(&loc->bad, bad);
        loc->n_keys = ds->n; loc->n_records = ds->R; loc->n_edges = ds->E;
    }
    if (!ds->ok) return;
    const u32* ids = reinterpret_cast<const u32*>(X.arena[X.rank] + L.off_id);
    __shared__ DxFlat F;
    dx_flat_init(F, X.world, 1u, [&](u32 d) { return dx_sent_count(loc, d, L.kcap); });
    const u32 total = F.off[X.world];
    for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const u32 d = F.seg(q), pos = q - F.off[d];
        slot_id[(u32)klist[(u64)d * L.kcap + pos]] = ids[(u64)d * L.kcap + pos];
    }
}

// ---------------------------------------------------------------- x5: row entries -> owner of their row block
// Flat over the stored edge records (unweighted: nothing depends on the emission index).  A CTA handles
// 256 x DXE_BATCH records per round: count per destination in shared memory, reserve the CTA's range in
// every destination segment with ONE global atomic per destination, then write the entries straight
// into the destination rank's memory.
#define DXE_BATCH 4
template <int TPE>
__global__ void __launch_bounds__(256) k_dx_entries(u32* __restrict__ edge_slots, const u32* __restrict__ slot_id, const DevSizes* __restrict__ ds,
                                                     int sym, int csc, const DxPeers X, const DxLayout L, DxLocal* __restrict__ loc)
{
    constexpr int SPE = TPE == 4 ? 4 : 2;
    __shared__ u32 s_cnt[DX_MAXW], s_base[DX_MAXW], s_go;
    if (threadIdx.x == 0) s_go = ds->ok && !loc->bad;  // one decision per CTA (another CTA may raise `bad` meanwhile)
    __syncthreads();
    const u32 E = s_go ? ds->E : 0u;
    const u32 rpr = loc->rows_per;
    const u32 lane = threadIdx.x & 31;
    const u32 per_round = 256 * DXE_BATCH;
    for (u32 r0 = blockIdx.x * per_round; r0 < E; r0 += gridDim.x * per_round) {  // uniform per CTA
        if (threadIdx.x < DX_MAXW) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        u32 id[DXE_BATCH][4];
#pragma unroll
        for (int u = 0; u < DXE_BATCH; u++) {
            const u32 e = r0 + u * 256 + threadIdx.x;
            id[u][0] = id[u][1] = id[u][2] = id[u][3] = 0;
            if (e < E) {
                if (SPE == 4) {
                    const uint4 q = reinterpret_cast<const uint4*>(edge_slots)[e];
                    id[u][0] = q.x; id[u][1] = q.y; id[u][2] = q.z; id[u][3] = q.w;
                } else {
                    const uint2 q = reinterpret_cast<const uint2*>(edge_slots)[e];
                    id[u][0] = q.x; id[u][1] = q.y;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < DXE_BATCH; u++) {
            if (r0 + u * 256 + threadIdx.x < E) {
#pragma unroll
                for (int k = 0; k < SPE; k++) id[u][k] = slot_id[id[u][k]];
            }
        }
        // the in-segment position of an entry: (entries of the same destination in earlier (u, k, half) steps of this
        // CTA) + rank among the lanes of the same step; pass 0 counts, pass 1 writes
        for (int pass = 0; pass < 2; pass++) {
#pragma unroll
            for (int u = 0; u < DXE_BATCH; u++) {
                const bool valid = r0 + u * 256 + threadIdx.x < E;
                record_entries_t<TPE>(id[u], 0u, sym, csc, [&](u32 major, u32 minor, u32 dir, u32) {
                    u32 d = valid ? major / rpr : 0xFFu;
                    if (valid && d >= (u32)X.world) d = (u32)X.world - 1;  // never: IDs are below n
                    const u32 m = __match_any_sync(0xffffffffu, d);
                    if (!valid) return;
                    const int leader = __ffs(m) - 1;
                    u32 base = 0;
                    if ((int)lane == leader) base = atomicAdd(&s_cnt[d], (u32)__popc(m));
                    base = __shfl_sync(m, base, leader);
                    if (pass) {
                        const u32 pos = s_base[d] + base + __popc(m & ((1u << lane) - 1u));
                        if (pos < L.pcap) {
                            DistPair p;
                            p.entry = Ent32::make(minor, dir, 0u);
                            p.major = major;
                            reinterpret_cast<DistPair*>(X.arena[d] + L.off_pair)[(u64)X.rank * L.pcap + pos] = p;
                        }
                    }
                });
            }
            __syncthreads();
            if (pass == 0) {
                if (threadIdx.x < (u32)X.world) {
                    const u32 c = s_cnt[threadIdx.x];
                    const u32 b = c ? atomicAdd(&loc->cur_pairs[threadIdx.x], c) : 0u;
                    if ((u64)b + c > L.pcap) atomicOr(&loc->bad, DXB_PAIRS);
                    s_base[threadIdx.x] = b;
                    s_cnt[threadIdx.x] = 0;
                }
                __syncthreads();
            }
        }
    }
    dx_tail_signal(X, 4, loc, loc->cur_pairs, L.pcap, 0);
}

// ---------------------------------------------------------------- x5 consumer: this rank's slab
// one warp: received counts -> DevSizes of the slab build (rows, entries)
__global__ void k_dx_slab_sizes(const DxPeers X, const DxLayout L, const DxCtl* my, DxLocal* __restrict__ loc,
                                DevSizes* __restrict__ ds, u32 recv_cap, u32 rows_cap)
{
    dx_wait(my, 4, X, loc);
    if (threadIdx.x == 0) {
    u32 bad = loc->bad, run = 0;
    for (int s = 0; s < X.world; s++) {
        bad |= dx_bad(my, 4, (u32)s);
        loc->seg_off[s] = run;
        run += dx_count(my, 4, (u32)s, L.pcap);
    }
    loc->seg_off[X.world] = run;
    loc->n_recv = run;
    if (run > recv_cap) bad |= DXB_RECV;
    if (loc->n_rows > rows_cap) bad |= DXB_ROWS;
    loc->bad = bad;
    DevSizes s;
    memset(&s, 0, sizeof(s));
    if (!bad) { s.n = loc->n_rows; s.rows = loc->n_rows; s.M = run; s.T = run; s.ok = 1; }
    *ds = s;
    }
    __syncthreads();  // one CTA: thread 0's status precedes the headers
    dx_publish(X, 5, loc, nullptr, 0, 0);
}

__global__ void __launch_bounds__(256) k_pairs_count(const DxPeers X, const DxLayout L, const DxLocal* __restrict__ loc, const DevSizes* __restrict__ ds,
                                                      u32* __restrict__ cnt, u32* __restrict__ bad_out, const RowRange rr)
{
    if (!ds->ok) return;
    const u32 row0 = loc->row0, n_rows = loc->n_rows;
    const DistPair* base = reinterpret_cast<const DistPair*>(X.arena[X.rank] + L.off_pair);
    __shared__ DxFlat F;
    dx_flat_init(F, X.world, 1u, [&](u32 s) { return loc->seg_off[s + 1] - loc->seg_off[s]; });
    const u32 total = F.off[X.world];
    for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const u32 s = F.seg(q);
        const u32 r = base[(u64)s * L.pcap + (q - F.off[s])].major - row0;
        if (r >= n_rows) atomicOr(bad_out, DXB_RANGE);
        else if (rr.has(r)) atomicAdd(&cnt[r], 1u);
    }
}

// multi-GPU builds are unweighted: the slab keeps 32-bit entries (minor << 1 | dir), see Ent32
__global__ void __launch_bounds__(256) k_pairs_scatter(const DxPeers X, const DxLayout L, const DxLocal* __restrict__ loc, const DevSizes* __restrict__ ds,
                                                        u32* __restrict__ cursor, u32* __restrict__ entries, const RowRange rr)
{
    if (!ds->ok) return;
    const u32 row0 = loc->row0, n_rows = loc->n_rows;
    const DistPair* base = reinterpret_cast<const DistPair*>(X.arena[X.rank] + L.off_pair);
    __shared__ DxFlat F;
    dx_flat_init(F, X.world, 1u, [&](u32 s) { return loc->seg_off[s + 1] - loc->seg_off[s]; });
    const u32 total = F.off[X.world];
    for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const u32 s = F.seg(q);
        const DistPair p = base[(u64)s * L.pcap + (q - F.off[s])];
        const u32 r = p.major - row0;
        if (r < n_rows && rr.has(r)) entries[atomicAdd(&cursor[r], 1u)] = p.entry;
    }
}

// ---------------------------------------------------------------- weighted builds: x5 in emission order
// Duplicate weights are summed in emission order (SciPy's order for rows of <= 16 stored entries, SURVEY 8a
// row 13), so the order of the entries inside a (source, owner) segment must be the file order -- the
// receiver then numbers them (segments in rank order = file order) and uses that number as the emission
// index of the single-GPU row kernels.  Two passes, one warp per tile, no atomics on positions:
//   k_dxw_count    entries per (tile, owner)            -> exclusive scan over [owner][tile]
//   k_dxw_scatter  position = scan value + rank inside the tile (records in order, entries of a record in
//                  the order of builders.py:222-234), written straight into the owner's segment
// entry counts of one lane's record per owner: 8 x 16-bit fields in two 64-bit words
__device__ __forceinline__ void dxw_add(u64& lo, u64& hi, u32 d) { if (d < 4) lo += 1ull << (16 * d); else hi += 1ull << (16 * (d - 4)); }
__device__ __forceinline__ u32 dxw_get(u64 lo, u64 hi, u32 d) { return (u32)((d < 4 ? lo >> (16 * d) : hi >> (16 * (d - 4))) & 0xFFFFu); }

template <bool WRITE>
__device__ __forceinline__ void dxw_tiles(const EmitParams& E, int sym, int csc, const DxPeers& X, const DxLayout& L, const DxLocal* loc,
                                          u32* __restrict__ tile_cnt, const u32* __restrict__ tile_off, u32* bad)
{
    const u32 lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const u32 rpr = loc->rows_per, W = (u32)X.world, nt = E.n_tiles;
    for (u32 tile = warp; tile < nt; tile += n_warps) {
        const TileInfo ti = E.tile_info[tile];
        u64 run_lo = 0, run_hi = 0;  // entries of this tile per owner so far (same in every lane)
        for (u32 j0 = 0; j0 < ti.n_edge; j0 += 32) {  // uniform trip count
            const u32 j = j0 + lane;
            const bool valid = j < ti.n_edge;
            u32 id[4] = {0, 0, 0, 0};
            double w = 1.0;
            if (valid) {
                const u32 stored = ti.edge_alloc + j;
                const u32* sl = E.edge_slots + (u64)stored * E.slots_per_edge;
                id[0] = sl[0]; id[1] = sl[1];
                if (E.slots_per_edge == 4) { id[2] = sl[2]; id[3] = sl[3]; }
                if (!E.ids_ready) {  // table slots -> node IDs
                    id[0] = E.slot_id[id[0]]; id[1] = E.slot_id[id[1]];
                    if (E.slots_per_edge == 4) { id[2] = E.slot_id[id[2]]; id[3] = E.slot_id[id[3]]; }
                }
                if (WRITE && E.edge_w) w = E.edge_w[stored];
            }
            // my record's entries per owner
            u64 c_lo = 0, c_hi = 0;
            if (valid) record_entries(id, E.tpe, 0u, sym, csc, [&](u32 major, u32, u32, u32) { dxw_add(c_lo, c_hi, min(major / rpr, W - 1)); });
            const u64 i_lo = warp_incl_scan64(c_lo), i_hi = warp_incl_scan64(c_hi);  // 16-bit fields: <= 32 x 8 entries
            if (WRITE && valid) {
                u64 m_lo = run_lo + i_lo - c_lo, m_hi = run_hi + i_hi - c_hi;  // entries before mine, per owner
                record_entries(id, E.tpe, 0u, sym, csc, [&](u32 major, u32 minor, u32 dir, u32) {
                    const u32 d = min(major / rpr, W - 1);
                    const u32 pos = tile_off[(u64)d * nt + tile] - tile_off[(u64)d * nt] + dxw_get(m_lo, m_hi, d);
                    dxw_add(m_lo, m_hi, d);
                    if (pos < L.pcap) {
                        DistPairW p;
                        p.entry = Ent32::make(minor, dir, 0u);
                        p.major = major;
                        p.w = w;
                        reinterpret_cast<DistPairW*>(X.arena[d] + L.off_pair)[(u64)X.rank * L.pcap + pos] = p;
                    } else {
                        atomicOr(bad, DXB_PAIRS);
                    }
                });
            }
            run_lo += __shfl_sync(0xffffffffu, i_lo, 31);
            run_hi += __shfl_sync(0xffffffffu, i_hi, 31);
        }
        if (!WRITE && lane < W) tile_cnt[(u64)lane * nt + tile] = dxw_get(run_lo, run_hi, lane);
    }
}

__global__ void __launch_bounds__(256) k_dxw_count(const EmitParams E, int sym, int csc, const DxPeers X, const DxLayout L, DxLocal* __restrict__ loc,
                                                    u32* __restrict__ tile_cnt)
{
    __shared__ u32 s_go;
    if (threadIdx.x == 0) s_go = E.ds->ok && !loc->bad;
    __syncthreads();
    if (s_go) dxw_tiles<false>(E, sym, csc, X, L, loc, tile_cnt, nullptr, &loc->bad);
    else {
        // the scan that follows must see defined counts
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < (u64)X.world * E.n_tiles; i += (u64)gridDim.x * blockDim.x) tile_cnt[i] = 0;
    }
}

__global__ void __launch_bounds__(256) k_dxw_scatter(const EmitParams E, int sym, int csc, const DxPeers X, const DxLayout L, DxLocal* __restrict__ loc,
                                                      const u32* __restrict__ tile_off)
{
    __shared__ u32 s_go;
    if (threadIdx.x == 0) s_go = E.ds->ok && !loc->bad;
    __syncthreads();
    if (s_go) {
        dxw_tiles<true>(E, sym, csc, X, L, loc, nullptr, tile_off, &loc->bad);
        if (blockIdx.x == 0 && threadIdx.x < (u32)X.world) {  // entries per owner: the header's count
            const u64 nt = E.n_tiles, d = threadIdx.x;
            loc->cur_pairs[d] = tile_off[(d + 1) * nt] - tile_off[d * nt];
        }
    }
    dx_tail_signal(X, 4, loc, loc->cur_pairs, L.pcap, 0);
}

// receiver: row histogram, then entries numbered in arrival = emission order, weights laid out by that number
__global__ void __launch_bounds__(256) k_pairsw_count(const DxPeers X, const DxLayout L, const DxLocal* __restrict__ loc, const DevSizes* __restrict__ ds,
                                                       u32* __restrict__ cnt, u32* __restrict__ bad_out, const RowRange rr)
{
    if (!ds->ok) return;
    const u32 row0 = loc->row0, n_rows = loc->n_rows;
    const DistPairW* base = reinterpret_cast<const DistPairW*>(X.arena[X.rank] + L.off_pair);
    __shared__ DxFlat F;
    dx_flat_init(F, X.world, 1u, [&](u32 s) { return loc->seg_off[s + 1] - loc->seg_off[s]; });
    const u32 total = F.off[X.world];
    for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const u32 s = F.seg(q);
        const u32 r = base[(u64)s * L.pcap + (q - F.off[s])].major - row0;
        if (r >= n_rows) atomicOr(bad_out, DXB_RANGE);
        else if (rr.has(r)) atomicAdd(&cnt[r], 1u);
    }
}

__global__ void __launch_bounds__(256) k_pairsw_scatter(const DxPeers X, const DxLayout L, const DxLocal* __restrict__ loc, const DevSizes* __restrict__ ds,
                                                         u32* __restrict__ cursor, u64* __restrict__ entries, double* __restrict__ w_emit, const RowRange rr)
{
    if (!ds->ok) return;
    const u32 row0 = loc->row0, n_rows = loc->n_rows;
    const DistPairW* base = reinterpret_cast<const DistPairW*>(X.arena[X.rank] + L.off_pair);
    __shared__ DxFlat F;  // = loc->seg_off: the flat index IS the arrival (= emission) number
    dx_flat_init(F, X.world, 1u, [&](u32 s) { return loc->seg_off[s + 1] - loc->seg_off[s]; });
    const u32 total = F.off[X.world];
    for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const u32 s = F.seg(q);
        const DistPairW p = base[(u64)s * L.pcap + (q - F.off[s])];
        const u32 r = p.major - row0;
        if (r < n_rows && rr.has(r)) {
            w_emit[q] = p.w;
            entries[atomicAdd(&cursor[r], 1u)] = Ent64::make(Ent32::minor(p.entry), Ent32::dir(p.entry), q);
        }
    }
}

__device__ __forceinline__ void dx_pair_weight(const DistPair&, double*, u32) {}
__device__ __forceinline__ void dx_pair_weight(const DistPairW& p, double* w_emit, u32 q) { w_emit[q] = p.w; }

// ---------------------------------------------------------------- slabs far larger than L2: bucketed receive
// Same idea as the single-GPU bucketed build (rowsort.cuh: RowBuckets): the received entries are first partitioned by
// row bucket of the slab into a bucket-major pair list (local row, entry), and k_bucket_rows_count /
// k_bucket_rows_scatter stream over that list.  PAIR = DistPair -> Ent32, DistPairW -> Ent64 (the flat arrival
// number is the emission index, the weight is laid out by it on the way).
template <class PAIR>
__global__ void __launch_bounds__(256) k_pairs_bucket_count(const DxPeers X, const DxLayout L, const DxLocal* __restrict__ loc, const DevSizes* __restrict__ ds,
                                                             const RowBuckets rb, BucketCtl* __restrict__ ctl, u32* __restrict__ bad_out)
{
    __shared__ u32 s_cnt[8][RB_MAX];
    __shared__ DxFlat F;
    for (u32 i = threadIdx.x; i < 8 * RB_MAX; i += 256) (&s_cnt[0][0])[i] = 0;
    if (!ds->ok) return;
    dx_flat_init(F, X.world, 1u, [&](u32 s) { return loc->seg_off[s + 1] - loc->seg_off[s]; });
    const u32 row0 = loc->row0, n_rows = loc->n_rows, wid = threadIdx.x >> 5;
    const PAIR* base = reinterpret_cast<const PAIR*>(X.arena[X.rank] + L.off_pair);
    const u32 total = F.off[X.world];
    for (u32 q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const u32 s = F.seg(q);
        const u32 r = base[(u64)s * L.pcap + (q - F.off[s])].major - row0;
        if (r >= n_rows) atomicOr(bad_out, DXB_RANGE);
        else atomicAdd(&s_cnt[wid][rb.of(r)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < rb.count) {
        u32 c = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) c += s_cnt[w][threadIdx.x];
        if (c) atomicAdd(&ctl->cnt[threadIdx.x], c);
    }
}

template <class PAIR, class ENT>
__global__ void __launch_bounds__(256) k_pairs_bucket_scatter(const DxPeers X, const DxLayout L, const DxLocal* __restrict__ loc, const DevSizes* __restrict__ ds,
                                                               const RowBuckets rb, BucketCtl* __restrict__ ctl, u32* __restrict__ pair_major,
                                                               typename ENT::type* __restrict__ pair_ent, double* __restrict__ w_emit)
{
    typedef typename ENT::type EV;
    constexpr int PER = RB_ROUND / 256;
    __shared__ u32 s_off[RB_MAX], s_cnt[RB_MAX], s_lo[RB_MAX + 1], s_base[RB_MAX], s_ws[8];
    __shared__ u32 s_major[RB_ROUND];
    __shared__ EV s_ent[RB_ROUND];
    __shared__ DxFlat F;
    if (!ds->ok) return;
    dx_flat_init(F, X.world, 1u, [&](u32 s) { return loc->seg_off[s + 1] - loc->seg_off[s]; });
    if (threadIdx.x == 0) {
        u32 run = 0;
        for (u32 b = 0; b < rb.count; b++) { s_off[b] = run; run += ctl->cnt[b]; }
    }
    const u32 row0 = loc->row0, n_rows = loc->n_rows;
    const PAIR* base = reinterpret_cast<const PAIR*>(X.arena[X.rank] + L.off_pair);
    const u64 total = F.off[X.world];
    for (u64 q0 = (u64)blockIdx.x * RB_ROUND; q0 < total; q0 += (u64)gridDim.x * RB_ROUND) {  // uniform per CTA
        if (threadIdx.x < RB_MAX) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        u32 row[PER];
        EV ent[PER];
        unsigned short rk[PER];
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const u64 q = q0 + u * 256 + threadIdx.x;
            row[u] = 0xFFFFFFFFu;
            ent[u] = (EV)0;
            if (q < total) {
                const u32 s = F.seg((u32)q);
                const PAIR p = base[(u64)s * L.pcap + ((u32)q - F.off[s])];
                const u32 r = p.major - row0;
                if (r < n_rows) {
                    row[u] = r;
                    ent[u] = ENT::make(Ent32::minor(p.entry), Ent32::dir(p.entry), (u32)q);
                    dx_pair_weight(p, w_emit, (u32)q);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PER; u++)
            if (row[u] != 0xFFFFFFFFu) rk[u] = (unsigned short)atomicAdd(&s_cnt[rb.of(row[u])], 1u);
        rb_prefix256(s_cnt, rb.count, s_lo, s_ws);
        if (threadIdx.x < rb.count) {
            const u32 c = s_cnt[threadIdx.x];
            s_base[threadIdx.x] = s_off[threadIdx.x] + (c ? atomicAdd(&ctl->cur[threadIdx.x], c) : 0u);
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PER; u++) {
            if (row[u] != 0xFFFFFFFFu) {
                const u32 p = s_lo[rb.of(row[u])] + rk[u];
                s_major[p] = row[u];
                s_ent[p] = ent[u];
            }
        }
        __syncthreads();
        const u32 n = s_lo[rb.count];
        for (u32 p = threadIdx.x; p < n; p += 256) {
            const u32 r = s_major[p];
            const u32 b = rb.of(r);
            const u64 pos = (u64)s_base[b] + (p - s_lo[b]);
            pair_major[pos] = r;
            pair_ent[pos] = s_ent[p];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- x6 consumer: the build's verdict, same on every rank
__global__ void k_dx_final(const DxPeers X, const DxCtl* my, DxLocal* __restrict__ loc)
{
    dx_wait(my, 5, X, loc);
    if (threadIdx.x) return;
    u32 bad = loc->bad;
    for (int s = 0; s < X.world; s++) bad |= dx_bad(my, 5, (u32)s);
    loc->final_bad = bad;
}

}  // namespace g2n
