// table.cuh -- shared types of the tokenizer kernels and the GPU open-addressing hash of node keys.
//
// The table maps node keys (segment names, optionally with a ":+" / ":-" orientation suffix) to slots
// and keeps, per key, the minimum `order` over all its mentions -- the reference's first-appearance
// order (builders.py:194-198, 219-221).  IDs are derived from those minima in ids.cuh.
#pragma once
#include "common.cuh"
#include "numparse.cuh"
#include "../../include/g2n.h"

namespace g2n {

// Tokenizer geometry: one warp owns one 2 KiB tile (64 bytes per lane) plus a look-ahead window.
#define WT_TILE 2048
#define WT_PRE 32
#ifndef WT_LOOK
#define WT_LOOK 480  // look-ahead past the tile: what a line that starts in the tile may use; longer lines go to the generic parser
#endif
#define WT_WIN (WT_PRE + WT_TILE + WT_LOOK)  // 2560 bytes staged per tile
#define WT_WORDS (WT_WIN / 32)               // 32-bit mask words covering the window
#ifndef WT_WARPS
#define WT_WARPS 8                           // warps (tiles in flight) per CTA
#endif
#ifndef WT_LIST
#define WT_LIST 256                          // record lines per batch of the compacted list
#endif
#define TK_NF 0xFFFFFFFFu

// `order` of a node mention: [63:12] tile, [11:2] record index within the tile (a 2 KiB tile holds at
// most 1024 records), [1:0] sub-rank inside the record (builders.py:230-234: u, v, v', u').
// File order == increasing order, and no tile needs to know anything about any other tile.
__device__ __forceinline__ unsigned long long make_order(unsigned tile, unsigned rec_idx, unsigned sub)
{
    return ((unsigned long long)tile << 12) | ((unsigned long long)rec_idx << 2) | sub;
}

struct TileInfo {
    unsigned n_rec;       // records yielded by this tile (S L E C P O)
    unsigned n_edge;      // edge records (L E C)
    unsigned edge_alloc;  // first edge_slots index of this tile's edge records
    unsigned pad;
};

// Hash table: open addressing over 32-byte slots -- ONE slot = ONE L2 / DRAM sector, fetched with one
// 256-bit load, so a mention of a known key costs exactly one sector however large the table is (the
// first-appearance word and the row counter ride in the sector the key compare needs anyway):
//   k0, k1  16-byte key: <= 15 inline bytes + (len+1) in the top byte, or a 0xFF-tagged hash for long keys
//   first   ~min(order) over the key's mentions (atomicMax; skipped when the loaded value already covers the mention)
//   rep     long keys only: 1 + index of a LongDesc holding the key's bytes
// Measured on B200 (tools/ubench/randmem.cu, profiles/r2_randmem.txt): random 32-byte reads run at 216 G/s from L2 and
// 43 G/s from HBM (1 GiB working set), 64-byte reads at half that -- DRAM traffic is sector-granular, so the slot must
// not straddle sectors and nothing a plain mention needs may live in a second array.  A read followed by an atomic
// on the SAME sector runs at 62 G/s (L2) / 16 G/s (HBM): that is why the per-node row counter is NOT in the slot but
// in a compact side array (slot_cnt, 4 bytes per slot, only when that array stays L2-resident).
struct alignas(32) Slot {
    u64 k0, k1;
    u64 first;
    u32 rep;
    u32 spare;
};
typedef ulonglong2 TKey;  // a bare key (exchange arenas of the multi-GPU build)

struct DeferEnt {
    u64 off;  // global offset of the first byte of the line
    u32 tile;
    unsigned short rec_idx, edge_idx;  // within the tile
};

struct LongDesc {
    u64 base_off;
    u64 ori_off;
    u32 base_len;
    u32 ori_len;   // bytes of the orientation string (may be 0)
    u32 ori_char;  // used when ori_len == 1
    u32 has_ori;   // 1: key is base + ':' + ori
};

struct Counters {  // all-zero at the start of a build (one memset)
    u64 first_error_inv;    // max ~(line_offset << 8 | kind): the FIRST offending line; 0 if none
    u64 first_unknown_inv;  // max ~(line_offset << 8 | first byte); 0 if none
    u32 n_records;
    u32 n_edges;
    u32 n_long;
    u32 flags;
    u32 scan_ticket;
    u32 collision;
    u32 n_defer;
    u32 pad0;
    u64 nnz;
    u64 aux[4];
    u64 pad1[5];
    // the two words every tile of the tokenizer bumps sit in 128-byte lines of their own: the L2 serialises atomics per
    // line, and the tile's `flags` poll must not queue behind them
    u32 edge_alloc;  // edge_slots entries handed out so far (one atomicAdd per tile)
    u32 pad2[31];
    u32 n_keys;
    u32 pad3[31];
};
static_assert(sizeof(Counters) == 128 * 3, "Counters layout");
// Sizes of the current build, resident in device memory so that no kernel after the tokenizer needs a
// host round trip: k_sizes (ids.cuh) derives them from the counters, later kernels read what they need.
struct DevSizes {
    u32 n;      // nodes
    u32 E;      // edge records
    u32 R;      // records (S L E C P O)
    u32 words;  // 32-bit words of the first-appearance bitmap (4 bits per record)
    u32 T;      // triplets
    u32 M;      // row entries (T, or 2T for max(S, S^T))
    u32 ok;     // 0: a capacity was exceeded or the host has to look at the input: later kernels do nothing
    u32 nnz;    // stored entries of the result (written by the last kernel of the build)
    u32 rows;   // rows of the matrix this GPU builds (n, or the rows of its slab)
    u32 wgroups;  // groups of BM_GROUP bitmap words (what the rank scan runs over)
    u32 pad[6];
};
struct Ctl {
    Counters c;
    DevSizes s;
};

// slot_cnt = entries of the node's row (K4 major): every endpoint of every edge record when both (r, c) and
// (c, r) are stored (undirected triplets, max(S, S^T)); the source only for a directed CSR, the target for a
// directed CSC; nothing when the result is the raw COO
#define CM_NONE 0
#define CM_ALL 1
#define CM_SRC 2
#define CM_DST 3
__host__ __device__ __forceinline__ bool cm_counts(int mode, unsigned sub) { return mode == CM_ALL || (mode == CM_SRC && sub == 0) || (mode == CM_DST && sub == 1); }

#define CF_TABLE_FULL 1u
#define CF_EDGE_FULL 2u
#define CF_LONG_FULL 4u
#define CF_CAST_OVERFLOW 8u
#define CF_DEFER_FULL 16u

struct ScanParams {
    const uint8_t* text;
    u64 nbytes;
    Slot* slots;
    u32* slot_cnt;   // per slot: row entries the node will own as a K4 major (NULL: counted by a pass over the edge records)
    u32 table_mask;  // slots - 1 (slots is a power of two >= TG_SLOTS)
    u32 table_max_keys;
    u32* edge_slots;
    double* edge_w;
    u32 edge_cap;
    LongDesc* longs;
    u32 long_cap;
    struct DeferEnt* defer;  // lines the hot kernel hands to k_tokenize_slow
    u32 defer_cap;
    TileInfo* tile_info;
    Counters* cnt;
    u32 n_tiles;     // tiles of the whole text
    u32 tile_begin;  // this launch handles tiles [tile_begin, tile_end): the text may still be arriving
    u32 tile_end;    // (host -> device copy in pieces, one launch per piece)
    int bidirected;
    int slots_per_edge;  // 2, or 4 for bidirected without keep_directed_bidir
    int strip_orientation;
    int wt_len;
    int dtype;  // G2N_DTYPE_* the weights will be cast to (only used to flag float32 overflow)
    int count_mode;  // which mentions of an edge record bump slot_cnt: CM_* below
    u64 seed;
    uint8_t wt[64];
};

// ---------------------------------------------------------------- byte window
struct Win {
    const uint8_t* sm;  // shared-memory copy of [base, base + wlen), if any
    const uint8_t* g;
    u64 base;
    u64 n;
    u64 wlen;
    __device__ __forceinline__ uint8_t operator()(u64 p) const
    {
        if (p >= n) return '\n';
        const u64 d = p - base;
        if (d < wlen) return sm[d];
        return g[p];
    }
};

struct Span {
    u64 off;
    u32 len;
};

struct SpanSrc {
    const Win& w;
    u64 off;
    __device__ __forceinline__ uint8_t operator()(int64_t i) const { return w(off + (u64)i); }
};

// A node key: base bytes, optionally followed by ':' + orientation (builders.py:193, 211-212, 234)
struct KeyDesc {
    u64 base_off;
    u64 ori_off;
    u32 base_len;
    u32 ori_len;   // bytes of the orientation string (may be 0: key ends with ':')
    u32 ori_char;  // literal when ori_len == 1
    u32 has_ori;   // 0: plain key (not bidirected)
    __device__ __forceinline__ u32 total_len() const { return base_len + (has_ori ? 1 + ori_len : 0); }
    __device__ __forceinline__ uint8_t byte(const Win& w, u32 i) const
    {
        if (i < base_len) return w(base_off + i);
        if (i == base_len) return ':';
        if (ori_len == 1) return (uint8_t)ori_char;
        return w(ori_off + (i - base_len - 1));
    }
};

__device__ __forceinline__ u64 mix64(u64 x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

__device__ __forceinline__ void cas128(Slot* s, u64 n0, u64 n1, u64& o0, u64& o1)
{
    asm volatile(
        "{\n\t"
        ".reg .b128 c, v, o;\n\t"
        "mov.b128 c, {%2, %3};\n\t"
        "mov.b128 v, {%4, %5};\n\t"
        "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n\t"
        "mov.b128 {%0, %1}, o;\n\t"
        "}"
        : "=l"(o0), "=l"(o1)
        : "l"(0ull), "l"(0ull), "l"(n0), "l"(n1), "l"(s)
        : "memory");
}

// one slot (one 32-byte sector) per 256-bit load; table sectors are kept in L2 (evict_last) while the
// text streams through (evict_first)
__device__ __forceinline__ u64 table_policy()
{
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// v = {k0, k1, first, rep}
__device__ __forceinline__ void ld_slot(const Slot* s, u64 (&v)[4], u64 pol)
{
    asm volatile("ld.global.cg.L2::cache_hint.v4.u64 {%0, %1, %2, %3}, [%4], %5;"
                 : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3])
                 : "l"(s), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void prefetch_slot(const Slot* s)
{
    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(s));
}

// Build the 128-bit table key for a node key.
__device__ __forceinline__ void make_key(const Win& w, const KeyDesc& kd, u64 seed, u64& k0, u64& k1, bool& is_long)
{
    const u32 L = kd.total_len();
    if (L <= 15) {
        u64 a = 0, b = 0;
        for (u32 i = 0; i < L; i++) {
            const u64 c = kd.byte(w, i);
            if (i < 8) a |= c << (8 * i); else b |= c << (8 * (i - 8));
        }
        k0 = a;
        k1 = b | ((u64)(L + 1) << 56);
        is_long = false;
    } else {
        u64 h1 = seed ^ 0x9e3779b97f4a7c15ULL, h2 = ~seed * 0xd6e8feb86659fd93ULL;
        for (u32 i = 0; i < L; i++) {
            const u64 c = kd.byte(w, i);
            h1 = (h1 ^ c) * 0x100000001b3ULL;
            h2 = (h2 + c + 1) * 0xc2b2ae3d27d4eb4fULL;
            h2 ^= h2 >> 29;
        }
        k0 = mix64(h1 ^ (h2 << 1));
        k1 = (0xFFull << 56) | ((u64)(L & 0xFFFFFF) << 32) | (mix64(h2 + h1) & 0xFFFFFFFFull);
        is_long = true;
    }
}

// bytes of a stored long key, read from global text only (any thread, any time)
__device__ __forceinline__ uint8_t long_byte(const uint8_t* text, const LongDesc& d, u32 i)
{
    if (i < d.base_len) return text[d.base_off + i];
    if (i == d.base_len) return ':';
    if (d.ori_len == 1) return (uint8_t)d.ori_char;
    return text[d.ori_off + (i - d.base_len - 1)];
}

// Probe sequence.  Sequence graphs name their segments in runs (s1041, s1042, ... / 1041+, 1041-), and
// links join neighbours, so keys that differ only in their last character are used together.  The
// table therefore hashes the key WITHOUT its last name character to a group of TG_SLOTS consecutive
// slots (1 KiB: one DRAM page burst / eight L2 lines) and lets that character (and the orientation of a
// bidirected key) pick the slot inside the group; the sequence continues with the same slot of
// other groups (double hashing over groups).  Any key set works -- clustering only shortens the
// distance between keys that are likely to be touched together; per-slot load is that of plain
// double hashing because the slot offset is rotated by the hash.
#define TG_SLOTS 32  // slots per group
// group stride of a key's probe sequence, in slots: an odd number of groups; a function of the whole key so that
// the hot kernel can compute it lazily (only a mention whose home slot holds another key needs it)
__device__ __forceinline__ u32 probe_step(u64 k0, u64 k1, u32 mask)
{
    u32 s = ((u32)k0 * 0x9E3779B1u) ^ ((u32)(k0 >> 32) * 0x85EBCA77u) ^ ((u32)k1 * 0xC2B2AE3Du) ^ ((u32)(k1 >> 32) * 0x27D4EB2Fu);
    s ^= s >> 15;
    return (((s << 1) | 1u) * TG_SLOTS) & mask;
}

struct ProbeSeq {
    u32 g;     // first slot of the current group
    u32 boff;  // offset of the key's slot inside every group
    u32 step;  // probe_step of the key
    __device__ __forceinline__ u32 slot() const { return g + boff; }
    // moves to the next slot; false when every group has been visited
    __device__ __forceinline__ bool next(u32 mask, u32& visited)
    {
        visited += TG_SLOTS;
        if (visited > mask) return false;
        g = (g + step) & mask;
        return true;
    }
};

// home slot.  (h0, h1): the key with its cluster byte (and, for the keys of a bidirected build, its '+' / '-'
// orientation byte) zeroed; c: the cluster byte; pair: bidirected build; ori: 1 for the '-' twin.
// A miss in L2 costs a whole 128-byte line of HBM traffic whatever the load's size (tools/ubench/randmem.cu under ncu:
// 133 bytes of DRAM reads per random 32-byte load), i.e. four slots: the two orientations of a segment -- registered
// together by every S line and mentioned together by every edge record of a bidirected build (builders.py:190-198,
// 230-234) -- therefore sit in ADJACENT slots of one line.
__device__ __forceinline__ u32 probe_home(u64 h0, u64 h1, u32 c, u32 pair, u32 ori, u32 mask)
{
    // 32-bit multiply-xorshift mix of the four key words (the quality only matters for speed)
    u32 a = (u32)h0 ^ ((u32)(h0 >> 32) * 0x9E3779B1u) ^ ((u32)h1 * 0x85EBCA77u) ^ ((u32)(h1 >> 32) * 0xC2B2AE3Du);
    a ^= a >> 16; a *= 0x21F0AAADu; a ^= a >> 15; a *= 0x735A2D97u; a ^= a >> 15;
    const u32 b = (a ^ (u32)(h0 >> 32) ^ (u32)h1) * 0x9E3779B1u;  // further bits: slot rotation and group stride
    const u32 gmask = mask & ~(u32)(TG_SLOTS - 1);  // tables are at least TG_SLOTS slots
    // ten digits -> ten of the thirty-two slots (ten of sixteen slot pairs), rotated per group of keys
    const u32 r = c + (b >> 27);
    return (a & gmask) + ((pair ? ((r << 1) | ori) : r) & (TG_SLOTS - 1));
}

// The cluster byte of a key is positional: the last byte, or -- for the keys of a bidirected build, which
// end in ':' + orientation -- the third byte from the end (the last character of the segment name when
// the orientation is one character, as it is in every well-formed file).
__device__ __forceinline__ ProbeSeq probe_seq(u64 k0, u64 k1, u32 mask, int bidir)
{
    const u32 top = (u32)(k1 >> 56);
    u32 c = 0, ori = 0;
    u64 h0 = k0, h1 = k1;
    if (top != 0xFF && top >= 2) {  // inline key of >= 1 byte (0xFF: hashed long key -- no structure to exploit)
        const u32 L = top - 1;
        auto byte_at = [&](u32 p) { return (u32)((p < 8 ? k0 >> (8 * p) : k1 >> (8 * (p - 8))) & 0xFF); };
        u32 pos = L - 1;
        if (bidir) {
            const u32 ob = byte_at(L - 1);
            if (ob == '+' || ob == '-') {  // the twins differ in this byte only: same group, adjacent slots
                ori = ob == '-';
                if (L - 1 < 8) h0 &= ~(0xFFull << (8 * (L - 1))); else h1 &= ~(0xFFull << (8 * (L - 1 - 8)));
            }
            if (L >= 3) pos = L - 3;
        }
        c = byte_at(pos);
        if (pos < 8) h0 &= ~(0xFFull << (8 * pos)); else h1 &= ~(0xFFull << (8 * (pos - 8)));
    }
    // (a hashed long key has no twin structure to exploit: it may sit in any slot of its group)
    const u32 home = probe_home(h0, h1, c, (bidir && top != 0xFF && top >= 2) ? 1u : 0u, ori, mask);
    ProbeSeq q;
    q.g = home & ~(u32)(TG_SLOTS - 1);
    q.boff = home & (TG_SLOTS - 1);
    q.step = probe_step(k0, k1, mask);
    return q;
}

// Examines slot i, whose 32 bytes were loaded as v: claims it if it is empty.  True: the slot holds the key now
// -- min(order) is recorded (atomicMax(~order), skipped when the value that came with the key already covers this
// mention: files are read roughly in order, so all but the first mention of a key take that exit and never dirty
// the sector for it) and the row counter is bumped for a counted mention.  False: another key lives here.
__device__ __forceinline__ bool slot_try(const ScanParams& P, u32 i, u64 k0, u64 k1, const u64 (&v)[4], u64 order, bool count, u32& claimed)
{
    u64 s0 = v[0], s1 = v[1], seen = v[2];
    if (s0 == 0 && s1 == 0) {
        cas128(&P.slots[i], k0, k1, s0, s1);
        if (s0 == 0 && s1 == 0) { claimed++; s0 = k0; s1 = k1; }
        seen = 0;  // whoever owns the slot now: its first-appearance word was not loaded with the key
    }
    if (s0 != k0 || s1 != k1) return false;
#ifndef TK_DBG_NOMAX
    if (~order > seen) atomicMax(&P.slots[i].first, ~order);
#endif
    if (count && P.slot_cnt) atomicAdd(&P.slot_cnt[i], 1u);
    return true;
}

// Lookup-or-insert of a ready 128-bit key; returns its slot (0xFFFFFFFF if the table is full).
__device__ __forceinline__ u32 table_probe(const ScanParams& P, u64 k0, u64 k1, u64 order, bool count, u32& claimed)
{
    const u64 pol = table_policy();
    ProbeSeq q = probe_seq(k0, k1, P.table_mask, P.bidirected);
    u32 visited = 0;
    while (true) {
        const u32 i = q.slot();
        u64 v[4];
        ld_slot(&P.slots[i], v, pol);
        if (slot_try(P, i, k0, k1, v, order, count, claimed)) return i;
        if (!q.next(P.table_mask, visited) || visited > 4096u * TG_SLOTS) {
            atomicOr(&P.cnt->flags, CF_TABLE_FULL);
            return 0xFFFFFFFFu;
        }
    }
}

// Generic lookup-or-insert from a key descriptor (any length, any orientation string).
__device__ __forceinline__ u32 table_insert(const ScanParams& P, const Win& w, const KeyDesc& kd, u64 order, bool count, u32& claimed)
{
    u64 k0, k1;
    bool is_long;
    make_key(w, kd, P.seed, k0, k1, is_long);
    const u32 i = table_probe(P, k0, k1, order, count, claimed);
    if (i == 0xFFFFFFFFu) return i;
    if (is_long) {
        // keep / verify the bytes behind a hashed key: every arrival is compared with some earlier
        // arrival, so all mentions that share the slot are byte-equal unless `collision` is raised
        u32 r = ld_volatile_u32(&P.slots[i].rep);
        if (r == 0) {
            const u32 idx = atomicAdd(&P.cnt->n_long, 1u);
            if (idx >= P.long_cap) { atomicOr(&P.cnt->flags, CF_LONG_FULL); return i; }
            LongDesc d;
            d.base_off = kd.base_off; d.ori_off = kd.ori_off; d.base_len = kd.base_len;
            d.ori_len = kd.ori_len; d.ori_char = kd.ori_char; d.has_ori = kd.has_ori;
            P.longs[idx] = d;
            __threadfence();
            r = atomicExch(&P.slots[i].rep, idx + 1);
        }
        if (r != 0) {
            __threadfence();
            const volatile LongDesc* vd = &P.longs[r - 1];
            LongDesc d;
            d.base_off = vd->base_off; d.ori_off = vd->ori_off; d.base_len = vd->base_len;
            d.ori_len = vd->ori_len; d.ori_char = vd->ori_char; d.has_ori = vd->has_ori;
            const u32 L = kd.total_len();
            bool same = (d.base_len + (d.has_ori ? 1 + d.ori_len : 0)) == L;
            for (u32 j = 0; same && j < L; j++) same = long_byte(P.text, d, j) == kd.byte(w, j);
            if (!same) atomicExch(&P.cnt->collision, 1u);
        }
    }
    return i;
}

}  // namespace g2n
