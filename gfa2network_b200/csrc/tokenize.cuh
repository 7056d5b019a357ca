// tokenize.cuh -- K1+K2 hot kernel: fused line scan, field split, node-key hashing and edge-record
// emission for the record shapes real GFA files are made of.
//
// One pass over the text, no dependency between tiles.  Each WARP owns 2 KiB tiles (64 bytes per
// lane): the tile plus a look-ahead window is fetched into shared memory by the TMA unit (one bulk copy
// per window; the next window is requested as soon as the last line of this one has been parsed and arrives
// while the tile's node mentions are still being looked up), the warp classifies '\n' and '\t' 16 bytes at a
// time into bitmasks, compacts the starts of its record lines with warp shuffles, and parses one line per
// lane per round from the separator bitmask (no byte loops, no block barriers).
// Parsing and hashing are decoupled: a parsed line only APPENDS its node mentions -- the packed 128-bit
// key (<= 15 name bytes inline), the key's home slot and where the result goes -- to a per-warp queue in
// shared memory; the queue is drained 32 mentions at a time, one per lane, so the table code runs with
// every lane busy and the same instructions whatever mix of S / L / E lines the tile holds (a line-per-
// lane probe ran at 9 of 32 lanes, profiles/r1_ncu_full.md).  A mention of a known key is ONE 32-byte
// sector (table.cuh: Slot); new keys are claimed with a 128-bit CAS.
// The table keeps, per key, the minimum `order` = (tile, record index in tile, sub-rank) -- file
// order, i.e. the reference's first-appearance order (builders.py:194-198, 219-221) -- so no prefix
// over earlier tiles is needed while parsing.  Edge records are written to a per-tile range of
// edge_slots claimed with one atomicAdd; per-tile counts go to tile_info and are scanned afterwards.
//
// Handled here: S (any), P/O (field count), L in GFA-1 form with one-byte orientations
// (parser.py:210-216), E in the reference's coord form (parser.py:254-288), weight tags whose value is
// a plain decimal (parser.py:179-204, builders.py:205-209).  Every other line -- compact L,
// orientation-only E, C, lenient numbers, long keys, errors, lines longer than the window -- is
// appended to the deferred list and parsed by k_tokenize_slow (tokenize_slow.cuh) with the generic
// byte-wise parser; since every mention carries its order the result does not depend on which kernel
// handled a line.
#pragma once
#include "table.cuh"

namespace g2n {

struct Tile {
    const uint8_t* win;  // shared-memory window, WT_WIN bytes (+ 32 bytes of slack)
    const u32* nlm;      // bit o: window byte o is '\n'
    const u32* spm;      // bit o: window byte o is '\t' or '\n'
    u64 wbase;           // global offset of window byte 0 (wraps for tile 0)
};

// first separator (TAB or newline) at or after window offset pos; TK_NF if none inside the window
__device__ __forceinline__ u32 find_sep(const Tile& t, u32 pos)
{
    u32 w = pos >> 5;
    u32 m = t.spm[w] & (0xFFFFFFFFu << (pos & 31));
    while (m == 0) {
        if (++w >= WT_WORDS) return TK_NF;
        m = t.spm[w];
    }
    return (w << 5) + (u32)__ffs(m) - 1u;
}

// inline table key of the <= 15 window bytes [off, off + len), optionally followed by ':' + ori
__device__ __forceinline__ void key_inline(const Tile& t, u32 off, u32 len, bool has_ori, u32 ori, u64& k0, u64& k1)
{
    const u32* wp = reinterpret_cast<const u32*>(t.win + (off & ~3u));
    const u32 sh = (off & 3u) * 8u;
    const u32 a0 = wp[0], a1 = wp[1], a2 = wp[2], a3 = wp[3], a4 = wp[4];
    const u32 b0 = __funnelshift_r(a0, a1, sh), b1 = __funnelshift_r(a1, a2, sh);
    const u32 b2 = __funnelshift_r(a2, a3, sh), b3 = __funnelshift_r(a3, a4, sh);
    u64 lo = (u64)b0 | ((u64)b1 << 32), hi = (u64)b2 | ((u64)b3 << 32);
    if (len < 8) { lo &= (1ull << (8 * len)) - 1ull; hi = 0; }
    else hi &= (1ull << (8 * (len - 8))) - 1ull;
    u32 L = len;
    if (has_ori) {
        const u64 suf = (u64)':' | ((u64)ori << 8);  // two bytes at positions len, len + 1
        if (len < 7) lo |= suf << (8 * len);
        else if (len == 7) { lo |= (u64)':' << 56; hi |= (u64)ori; }
        else hi |= suf << (8 * (len - 8));
        L += 2;
    }
    k0 = lo;
    k1 = hi | ((u64)(L + 1) << 56);
}

__device__ __forceinline__ u32 rstrip_pm_win(const Tile& t, u32 off, u32 len)
{
    while (len > 0) {
        const uint8_t c = t.win[off + len - 1];
        if (c != '+' && c != '-') break;
        len--;
    }
    return len;
}

__device__ const double G2N_P10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};

// Plain decimal: [+-] digits [. digits] (float only), <= 15 digit characters.  Exact: the mantissa is
// below 2^53 and the power of ten is exactly representable, so one IEEE division is correctly rounded.
// Anything else (spaces, underscores, exponents, inf/nan, long digit strings) -> false: generic parser.
__device__ __forceinline__ bool simple_number(const Tile& t, u32 a, u32 b, bool is_float, double& out)
{
    if (a >= b) return false;
    bool neg = false;
    uint8_t c = t.win[a];
    if (c == '+' || c == '-') { neg = c == '-'; a++; }
    u64 m = 0;
    u32 nd = 0, frac = 0;
    bool point = false;
    for (u32 q = a; q < b; q++) {
        c = t.win[q];
        const u32 d = (u32)c - '0';
        if (d < 10) { m = m * 10 + d; nd++; if (point) frac++; }
        else if (c == '.' && is_float && !point) point = true;
        else return false;
    }
    if (nd == 0 || nd > 15) return false;
    double v = (double)m;
    if (frac) v = v / G2N_P10[frac];
    // float("-0") is -0.0 but float(int("-0")) is 0.0
    out = (neg && (is_float || m != 0)) ? -v : v;
    return true;
}

// Weight from the tag fields that follow separator `q` (parser.py:179-204 restricted to the key
// builders.py:206 reads).  Returns false if any field needs the generic parser.
__device__ __forceinline__ bool fast_weight(const ScanParams& P, const Tile& t, u32 q, double& w_out)
{
    bool has = false;
    double w = 1.0;
    const u32 wl = (u32)P.wt_len;
    while (t.win[q] == '\t') {
        const u32 a = q + 1;
        const u32 en = find_sep(t, a);
        if (en == TK_NF) return false;
        q = en;
        if (en - a < wl + 1 || t.win[a + wl] != ':') continue;  // tag name differs (it cannot hold ':')
        bool match = true;
        for (u32 k = 0; k < wl; k++) match = match && (t.win[a + k] == P.wt[k]);
        if (!match) continue;
        // second colon; a field with fewer than three parts is skipped (parser.py:184-186)
        u32 c2 = TK_NF;
        bool ascii = true;
        for (u32 k = a + wl + 1; k < en; k++) {
            const uint8_t ch = t.win[k];
            if (ch == ':' && c2 == TK_NF) c2 = k;
            ascii = ascii && ch < 0x80;
        }
        if (!ascii) return false;  // UTF-8 validity decides whether the field counts: generic parser
        if (c2 == TK_NF) continue;
        const uint8_t typ = (c2 == a + wl + 2) ? t.win[a + wl + 1] : 0;
        if (typ == 'i' || typ == 'f') {
            if (!simple_number(t, c2 + 1, en, typ == 'f', w)) return false;
            has = true;
        } else {
            has = false;  // str / list value: builders.py:208 falls back to 1.0
        }
    }
    w_out = has ? w : 1.0;
    return true;
}

// MODE bits select the specialisation of the hot kernel (dead paths compile away: smaller code,
// fewer registers): 1 = bidirected keys, 2 = four slots per edge record, 4 = weight tag present
#define TM_BIDIR 1
#define TM_FOUR 2
#define TM_WEIGHT 4

// the first separators of a short line from one 64-bit slice of the separator mask starting at `pos`
__device__ __forceinline__ u64 mask_slice(const u32* m, u32 pos)
{
    const u32 w = pos >> 5, sh = pos & 31;
    const u32 a = m[w], b = m[w + 1], c = m[w + 2];  // both masks have two words of slack
    const u32 lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);
    return (u64)lo | ((u64)hi << 32);
}
__device__ __forceinline__ u64 sep_slice(const Tile& t, u32 pos) { return mask_slice(t.spm, pos); }

// the <= 8 window bytes [off, off + len) as a little-endian u64, bytes past len = '0'
__device__ __forceinline__ u64 load8_digits(const Tile& t, u32 off, u32 len)
{
    const u32* wp = reinterpret_cast<const u32*>(t.win + (off & ~3u));
    const u32 sh = (off & 3u) * 8u;
    const u32 a0 = wp[0], a1 = wp[1], a2 = wp[2];
    const u64 x = (u64)__funnelshift_r(a0, a1, sh) | ((u64)__funnelshift_r(a1, a2, sh) << 32);
    const u64 keep = len >= 8 ? ~0ull : ((1ull << (8 * len)) - 1ull);
    return (x & keep) | (0x3030303030303030ull & ~keep);
}
// every byte of x is an ASCII digit (borrows / carries only travel upwards from a byte that is already invalid)
__device__ __forceinline__ bool all_digits8(u64 x)
{
    return (((x + 0x4646464646464646ull) | (x - 0x3030303030303030ull) | x) & 0x8080808080808080ull) == 0;
}

__device__ __forceinline__ void defer_line(const ScanParams& P, u64 off, u32 tile, u32 rec_idx, u32 edge_idx)
{
    const u32 idx = atomicAdd(&P.cnt->n_defer, 1u);
    if (idx < P.defer_cap) {
        DeferEnt d;
        d.off = off; d.tile = tile; d.rec_idx = (unsigned short)rec_idx; d.edge_idx = (unsigned short)edge_idx;
        P.defer[idx] = d;
    } else {
        atomicOr(&P.cnt->flags, CF_DEFER_FULL);
    }
}

// What a parsed line leaves behind: where its node names sit in the window and how many node mentions it
// makes (builders.py:190-198 for S, :218-234 for edge records).
struct LineOut {
    u32 uo, ul, vo, vl;  // window offset / length of the two names (an S line: both describe its id)
    u32 ocu, ocv;        // orientation characters (bidirected keys)
    u32 nm;              // mentions: 0 (P / O), 1 or 2 (S), 2 or 4 (edge record)
    bool edge;
    double w;
};

// Common record shapes, parsed from the separator bitmask.  Returns false when the line must go to the
// generic parser (rare shapes, errors, long keys, fields running past the window).
template <int MODE>
__device__ __forceinline__ bool parse_line_fast(const ScanParams& P, const Tile& t, u32 s, LineOut& lo)
{
    constexpr bool BIDIR = (MODE & TM_BIDIR) != 0;
    constexpr bool FOUR = (MODE & TM_FOUR) != 0;
    constexpr bool want_w = (MODE & TM_WEIGHT) != 0;
    const uint8_t c0 = t.win[s];
    lo.nm = 0; lo.edge = false; lo.w = 1.0;
    if (t.win[s + 1] != '\t') return false;  // record with no fields at all: error paths
    const u32 p1 = s + 2;
    const u32 e1 = find_sep(t, p1);
    if (e1 == TK_NF) return false;
    const u32 maxlen = BIDIR ? 13u : 15u;  // longest base that still fits the inline key
    if (c0 == 'S') {
        // parser.py:135-163 -> builders.py:190-198: only fields[1] matters; bidirected: id:+ then id:-
        const u32 len = e1 - p1;
        if (len > maxlen) return false;
        lo.uo = lo.vo = p1; lo.ul = lo.vl = len; lo.ocu = '+'; lo.ocv = '-';
        lo.nm = BIDIR ? 2u : 1u;
        return true;
    }
    if (c0 == 'P' || c0 == 'O') return t.win[e1] == '\t';  // >= 3 fields; otherwise the generic parser raises
    if (t.win[e1] != '\t') return false;
    u32 uo, ul, vo, vl, oc_u, oc_v, tag_from;
    if (c0 == 'L') {
        // GFA-1 form with one-byte orientations (parser.py:210-216): L u o v o [ovl [tags]]
        // the four separators after fields[1] come from one 64-bit slice of the separator mask
        u64 m = sep_slice(t, e1 + 1);
        if (__popcll(m) < 3) return false;  // line longer than the slice (or malformed): generic parser
        const u32 e2 = e1 + 1 + (u32)__ffsll((long long)m) - 1; m &= m - 1;
        const u32 e3 = e1 + 1 + (u32)__ffsll((long long)m) - 1; m &= m - 1;
        const u32 e4 = e1 + 1 + (u32)__ffsll((long long)m) - 1; m &= m - 1;
        if (e2 != e1 + 2 || e4 != e3 + 2 || t.win[e2] != '\t' || t.win[e3] != '\t') return false;
        oc_u = t.win[e1 + 1];
        if (oc_u != '+' && oc_u != '-') return false;
        oc_v = t.win[e3 + 1];
        if (oc_v >= 0x80) return false;
        uo = p1; ul = e1 - p1; vo = e2 + 1; vl = e3 - e2 - 1;
        tag_from = e4;  // separator that ends fields[4]; fields[5] (overlap) is not a tag
        if (want_w && t.win[tag_from] == '\t') {
            const u32 e5 = find_sep(t, tag_from + 1);
            if (e5 == TK_NF) return false;
            tag_from = e5;
        }
    } else if (c0 == 'E') {
        // coord form (parser.py:254-288): E id u+- s e v+- s e cigar tags...
        u32 e[9];
        e[1] = e1;
        // the seven separators after fields[1] from one 64-bit slice of the separator mask when the nine fields are that
        // short (they are: ids and coordinates); none of the first six may be the line's end
        u64 m = sep_slice(t, e1 + 1);
        if (__popcll(m) >= 7) {
            const u64 nlm64 = mask_slice(t.nlm, e1 + 1);
            u64 first6 = 0;
#pragma unroll
            for (int k = 2; k <= 8; k++) {
                const u64 low = m & (0 - m);
                e[k] = e1 + 1 + (u32)__ffsll((long long)m) - 1;
                if (k <= 7) first6 |= low;
                m &= m - 1;
            }
            if (nlm64 & first6) return false;
        } else {
#pragma unroll
            for (int k = 2; k <= 8; k++) {
                if (t.win[e[k - 1]] != '\t') return false;
                e[k] = find_sep(t, e[k - 1] + 1);
                if (e[k] == TK_NF) return false;
            }
        }
        // int() probes on fields 3, 4, 6, 7: plain ASCII digits only on the fast path
#pragma unroll
        for (int k = 3; k <= 7; k++) {
            if (k == 5) continue;
            const u32 a = e[k - 1] + 1, b = e[k];
            if (b == a || b - a > 18) return false;
            if (b - a <= 8) {
                if (!all_digits8(load8_digits(t, a, b - a))) return false;
            } else {
                for (u32 q = a; q < b; q++)
                    if ((uint8_t)(t.win[q] - '0') > 9) return false;
            }
        }
        uo = e[1] + 1; ul = e[2] - uo; vo = e[4] + 1; vl = e[5] - vo;
        if (ul == 0 || vl == 0) return false;
        oc_u = t.win[uo + ul - 1] == '-' ? '-' : '+';
        oc_v = t.win[vo + vl - 1] == '-' ? '-' : '+';
        ul = rstrip_pm_win(t, uo, ul);
        vl = rstrip_pm_win(t, vo, vl);
        tag_from = e[8];
    } else {
        return false;  // C records: generic parser
    }
    if (P.strip_orientation) { ul = rstrip_pm_win(t, uo, ul); vl = rstrip_pm_win(t, vo, vl); }
    if (ul > maxlen || vl > maxlen) return false;
    if (want_w && !fast_weight(P, t, tag_from, lo.w)) return false;
    lo.uo = uo; lo.ul = ul; lo.vo = vo; lo.vl = vl; lo.ocu = oc_u; lo.ocv = oc_v;
    lo.nm = FOUR ? 4u : 2u;
    lo.edge = true;
    return true;
}

// Sparse tiles (a handful of records per 2 KiB: segments with their sequences): no separator mask is built at all,
// the few fields that matter are found byte by byte.  Only S / P / O lines get here (a tile with an edge record gets
// the separator mask after all); same results as parse_line_fast.
template <int MODE>
__device__ __forceinline__ bool parse_line_sparse(const Tile& t, u32 s, LineOut& lo)
{
    constexpr bool BIDIR = (MODE & TM_BIDIR) != 0;
    const uint8_t c0 = t.win[s];
    lo.nm = 0; lo.edge = false; lo.w = 1.0;
    if (t.win[s + 1] != '\t') return false;
    const u32 p1 = s + 2;
    const u32 maxlen = BIDIR ? 13u : 15u;
    if (c0 == 'S') {
        u32 len = 0;
        while (len <= maxlen) {
            const uint8_t c = t.win[p1 + len];
            if (c == '\t' || c == '\n') break;
            len++;
        }
        if (len > maxlen) return false;  // long key: generic parser
        lo.uo = lo.vo = p1; lo.ul = lo.vl = len; lo.ocu = '+'; lo.ocv = '-';
        lo.nm = BIDIR ? 2u : 1u;
        return true;
    }
    if (c0 == 'P' || c0 == 'O') {  // >= 3 fields <=> the name field ends with a TAB; a long name: generic parser
        for (u32 q = p1; q < p1 + 64u && q < WT_WIN; q++) {
            const uint8_t c = t.win[q];
            if (c == '\t') return true;
            if (c == '\n') return false;
        }
        return false;
    }
    return false;
}

// ---------------------------------------------------------------- the mention queue
// A ring per warp that lives ACROSS tiles.  One entry per node mention: the packed key and
//   x.x  the slot to look at next (home slot first: probe_home)        x.z  tile (high bits of the mention's order)
//   x.w  where the result goes: index of the edge record in edge_slots (absolute)
//   x.y  meta: [11:0] low bits of the order (record index in the tile << 2 | sub-rank = position in the record's slot tuple)
//              [12] the mention belongs to an edge record (its slot is stored)   [13] it bumps the node's row counter
//              [31:24] probes made so far
// The queue is drained in groups of 32, one mention per lane, and a group does ONE probe step: a mention whose slot
// holds another key goes back to the tail with its next slot (double hashing over groups, probe_step) and is looked at
// again with a later group.  So every lane is busy in the table code whatever the chain lengths are (looping inside the
// group ran at 10 - 14 of 32 lanes, profiles/r2_ncu_tokenize.md), the line mix of a tile does not matter, and tiles with
// a handful of records (segments with long sequences) share groups with their successors.  (Measured and dropped:
// prefetching the home slot into L2 at enqueue time with the drain lagging a round behind -- C5 shape at 5 %: 2.91 ms
// against 2.69 ms without; the extra L2 request per mention costs more than the shorter wait saves.)
#define QM_EDGE (1u << 12)
#define QM_CNT (1u << 13)
// ring capacity: one group stays in flight across a round of lines (< 64 mentions queued) + 32 lines x 2 mentions; or < 32
// mentions left over + 32 lines x 4 mentions for the records of a bidirected build (which carries no group across rounds:
// a 192-entry ring at two CTAs per SM was measured slower, C3 shape at 25 %: 1.85 ms against 1.59 ms)
template <int MODE> struct QCap { static constexpr u32 value = (MODE & TM_FOUR) ? 160u : 128u; };

// per-warp shared memory: the text window, the two bitmasks, the compacted line list, the mention queue, one mbarrier
template <u32 QCAP>
struct alignas(128) WarpSmemT {
    static constexpr u32 kCap = QCAP;
    uint8_t win[WT_WIN + 32];  // + 32 bytes of '\n' slack read by key_inline
    ulonglong2 qk[QCAP];
    uint4 qx[QCAP];
    u32 list[WT_LIST];  // [15:0] window offset of the line, [31:16] edge index within the tile
    u32 nl[WT_WORDS + 2];  // + 2 words of slack for mask_slice
    u32 sp[WT_WORDS + 2];
    u64 bar;
    static __device__ __forceinline__ u32 wrap(u32 pos) { return (QCAP & (QCAP - 1)) == 0 ? (pos & (QCAP - 1)) : (pos >= QCAP ? pos - QCAP : pos); }
};
template <int MODE> using WarpSmem = WarpSmemT<QCap<MODE>::value>;
template <int MODE> constexpr size_t tk_smem_bytes() { return WT_WARPS * sizeof(WarpSmem<MODE>); }
template <int MODE> constexpr int tk_min_blocks() { return 3; }

template <int MODE>
__device__ __forceinline__ void enqueue_mention(const ScanParams& P, const Tile& t, WarpSmem<MODE>& S, u32 idx, u32 off, u32 len, u32 ori, u32 meta, u32 tile, u32 edge_ord)
{
    u64 k0, k1;
    key_inline(t, off, len, (MODE & TM_BIDIR) != 0, ori, k0, k1);
    u32 home;
    if (len == 0) {
        const ProbeSeq q = probe_seq(k0, k1, P.table_mask, P.bidirected);
        home = q.slot();
    } else {
        // the home slot from what the parser already knows (same result as probe_seq on the packed key):
        // the cluster byte is the last character of the name, at key position len - 1
        const u32 cp = len - 1;
        u64 h0 = k0, h1 = k1;
        if (cp < 8) h0 &= ~(0xFFull << (8 * cp)); else h1 &= ~(0xFFull << (8 * (cp - 8)));
        if ((MODE & TM_BIDIR) && (ori == '+' || ori == '-')) {  // the orientation byte sits at key position len + 1
            const u32 op = len + 1;
            if (op < 8) h0 &= ~(0xFFull << (8 * op)); else h1 &= ~(0xFFull << (8 * (op - 8)));
        }
        home = probe_home(h0, h1, t.win[off + cp], (MODE & TM_BIDIR) ? 1u : 0u, ((MODE & TM_BIDIR) && ori == '-') ? 1u : 0u, P.table_mask);
    }
    G2N_CHECK(idx < WarpSmem<MODE>::kCap && home <= P.table_mask && off + len <= WT_WIN);
    S.qk[idx] = make_ulonglong2(k0, k1);
    S.qx[idx] = make_uint4(home, meta, tile, edge_ord);
}

// the slot sector of the mention at ring position qh + ahead + lane (the lookup's only long-latency access)
template <class SM>
__device__ __forceinline__ void group_load(const ScanParams& P, const SM& S, u32 qh, u32 qn, u32 ahead, u64 pol, u64 (&v)[4])
{
    const u32 lane = threadIdx.x & 31;
#ifndef TK_DBG_NOPROBE
    if (ahead + lane < qn) {
        const u32 i = S.qx[SM::wrap(qh + ahead + lane)].x;
        G2N_CHECK(i <= P.table_mask && qn <= SM::kCap);
        ld_slot(&P.slots[i], v, pol);
    }
#endif
}

// One group of <= 32 queued mentions from the head of the ring, one per lane, ONE probe step each (the slot was
// loaded into v by group_load): lookup-or-insert, first-appearance order, row counter, the slot into the edge record;
// unresolved mentions go back to the tail.
template <int MODE>
__device__ __forceinline__ void drain_group(const ScanParams& P, WarpSmem<MODE>& S, u32& qh, u32& qn, const u64 (&v)[4], u32& claimed)
{
    typedef WarpSmem<MODE> SM;
    constexpr u32 SPE = (MODE & TM_FOUR) ? 4u : 2u;
    const u32 lane = threadIdx.x & 31;
    const u32 take = qn < 32u ? qn : 32u;
    const bool live = lane < take;
    bool miss = false;
    ulonglong2 k = make_ulonglong2(0, 0);
    uint4 x = make_uint4(0, 0, 0, 0);
    if (live) {
        const u32 j = SM::wrap(qh + lane);
        k = S.qk[j];
        x = S.qx[j];
        u32 i = x.x;
#ifndef TK_DBG_NOPROBE
        const u64 order = ((u64)x.z << 12) | (x.y & 0xFFFu);
        if (!slot_try(P, i, k.x, k.y, v, order, (x.y & QM_CNT) != 0, claimed)) {
            const u32 probes = (x.y >> 24) + 1u;
            if (probes >= 255u || probes * TG_SLOTS > P.table_mask) {
                atomicOr(&P.cnt->flags, CF_TABLE_FULL);  // the host repeats the pass with a larger table
                i = 0xFFFFFFFFu;
            } else {
                miss = true;
                x.x = ((((i & ~(u32)(TG_SLOTS - 1)) + probe_step(k.x, k.y, P.table_mask)) & P.table_mask) | (i & (TG_SLOTS - 1)));
                x.y += 1u << 24;
            }
        }
#endif
        G2N_CHECK(miss ? x.x <= P.table_mask : (i <= P.table_mask || i == 0xFFFFFFFFu));
        if (!miss && (x.y & QM_EDGE) && x.w < P.edge_cap) P.edge_slots[(u64)x.w * SPE + (x.y & 3u)] = i;  // (edge_alloc > edge_cap: the host retries with room)
    }
    const u32 mb = __ballot_sync(0xffffffffu, miss);  // every lane has read its entry: the group's places may be reused
    qh = SM::wrap(qh + take);
    qn -= take;
    if (miss) {
        const u32 j = SM::wrap(qh + qn + (u32)__popc(mb & ((1u << lane) - 1u)));
        G2N_CHECK(j < SM::kCap && qn + (u32)__popc(mb) <= SM::kCap);
        S.qk[j] = k;
        S.qx[j] = x;
    }
    qn += (u32)__popc(mb);
    __syncwarp();
}

// Drains the ring's full groups -- all of it when `final`.  Latency hiding: the slot sectors of the group BEHIND the
// current one are requested before the current group is worked on, and the last full group of a call is left IN FLIGHT
// (`have`: its sectors are in v) while the warp parses its next round of lines; it is the first group of the next call.
template <int MODE>
__device__ __forceinline__ void drain(const ScanParams& P, WarpSmem<MODE>& S, u32& qh, u32& qn, bool final, u64 pol, u32& claimed, u64 (&v)[4], bool& have)
{
    if (!have) {
        if (qn == 0 || (!final && qn < 32u)) return;
        group_load(P, S, qh, qn, 0, pol, v);
        have = true;
    }
    constexpr bool CARRY = (MODE & TM_FOUR) == 0;
    while (true) {
        const bool more = qn >= 64u;  // a full group behind the one in flight (what the current group puts back only adds to it)
        if (CARRY && !more && !final) return;
        u64 w[4] = {0, 0, 0, 0};
        if (more) group_load(P, S, qh, qn, 32, pol, w);
        drain_group<MODE>(P, S, qh, qn, v, claimed);
        if (more) {
#pragma unroll
            for (int q = 0; q < 4; q++) v[q] = w[q];
        } else {
            if (qn == 0 || (!final && qn < 32u)) { have = false; return; }
            group_load(P, S, qh, qn, 0, pol, v);  // what the last group put back
        }
    }
}

// ---------------------------------------------------------------- the kernel

__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }

// TMA bulk copy global -> shared, completion counted on an mbarrier (bytes and addresses: multiples of 16)
__device__ __forceinline__ void tma_load(void* dst, const void* src, u32 bytes, u64* bar, u64 pol)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(smem_addr(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TK_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra TK_WAIT_%=;\n\t"
        "}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// 0x80 in every byte of w that equals the byte replicated in c (exact, no cross-byte carries)
__device__ __forceinline__ u32 eq_bytes(u32 w, u32 c)
{
    const u32 x = w ^ c;
    const u32 t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | x | 0x7F7F7F7Fu);
}
// the four 0x80 flags of a word -> bits 0..3
__device__ __forceinline__ u32 flags4(u32 f) { return (f * 0x00204081u) >> 28; }

template <int MODE>
__global__ void __launch_bounds__(WT_WARPS * 32, tk_min_blocks<MODE>()) k_tokenize(const __grid_constant__ ScanParams P)
{
    constexpr bool BIDIR = (MODE & TM_BIDIR) != 0;
    extern __shared__ __align__(128) uint8_t s_raw[];
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    typedef WarpSmem<MODE> SM;
    SM& S = reinterpret_cast<SM*>(s_raw)[wid];
    u32* nlm = S.nl;
    u32* spm = S.sp;
    u32* list = S.list;
    uint8_t* win = S.win;
    const u64 pol_text = policy_evict_first();
    const u64 pol_table = table_policy();
    if (lane < 8) reinterpret_cast<u32*>(win + WT_WIN)[lane] = 0x0A0A0A0Au;
    if (lane < 2) { spm[WT_WORDS + lane] = 0; nlm[WT_WORDS + lane] = 0; }
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&S.bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const u32 n_warps = gridDim.x * WT_WARPS;
    // a window the TMA unit can fetch in one piece: inside the text, 16-byte granular (text is 16-byte aligned)
    const u64 n16 = P.nbytes & ~15ull;
    auto tma_ok = [&](u32 tile) { return tile > 0 && (u64)tile * WT_TILE - WT_PRE + WT_WIN <= n16; };
    u32 tile = P.tile_begin + blockIdx.x * WT_WARPS + wid;
    u32 parity = 0;  // phase the next wait on the mbarrier expects
    bool fetched = false;  // the window of `tile` has been requested from the TMA unit
    if (tile < P.tile_end && tma_ok(tile)) {
        if (lane == 0) tma_load(win, P.text + ((u64)tile * WT_TILE - WT_PRE), WT_WIN, &S.bar, pol_text);
        fetched = true;
    }
    u32 aborted = 0;
    bool quick = false;   // this warp's previous tile held at most one record: look for a line start before building masks
    bool sparse = false;  // ... held a handful of records and no edge record: build the newline mask only
    u32 qh = 0, qn = 0;  // the mention ring: head and fill (warp-uniform)
    u32 claimed = 0;
    u64 fly[4] = {0, 0, 0, 0};  // slot sectors of the ring's head group while it is in flight (drain)
    bool have_fly = false;
    for (; tile < P.tile_end && !aborted; tile += n_warps) {
        const u64 t0 = (u64)tile * WT_TILE;
        const u64 wbase = t0 - WT_PRE;  // wraps for tile 0: only ever used as wbase + offset
        const u32 nxt = tile + n_warps;
        const bool nxt_tma = nxt < P.tile_end && tma_ok(nxt);
        // requests the next tile's window; every lane is done with this one (called after a __syncwarp)
        auto fetch_next = [&]() {
            if (lane == 0 && nxt_tma) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the window vs the async write
                tma_load(win, P.text + ((u64)nxt * WT_TILE - WT_PRE), WT_WIN, &S.bar, pol_text);
            }
        };
        if (fetched) {
            mbar_wait(&S.bar, parity);
            parity ^= 1u;
        } else {
            // first and last windows: virtual '\n' before byte 0 and after the last byte
#pragma unroll
            for (int k = 0; k < WT_WIN / 16 / 32; k++) {
                const u32 piece = lane + 32 * k;
                const u64 g = wbase + (u64)piece * 16;  // global offset of this 16-byte piece
                uint4 v;
                if (tile == 0 && piece < WT_PRE / 16) {
                    v = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
                } else if (g + 16 <= P.nbytes) {
                    v = ld_stream_v4(P.text + g, pol_text);
                } else {
                    uint8_t tmp[16];
#pragma unroll
                    for (int b = 0; b < 16; b++) tmp[b] = (g + b < P.nbytes) ? P.text[g + b] : (uint8_t)'\n';
                    v = *reinterpret_cast<uint4*>(tmp);
                }
                reinterpret_cast<uint4*>(win)[piece] = v;
            }
            __syncwarp();
        }
        fetched = nxt_tma;
        // the next window is requested from the TMA unit only when this tile's lines are parsed (one window buffer per warp):
        // ask L2 for its lines now, so that the bulk copy finds them there
#ifndef TK_PF_DIST
#define TK_PF_DIST 1
#endif
        {
            const u64 pft = (u64)tile + (u64)TK_PF_DIST * n_warps;
            if (pft < P.tile_end && lane < (WT_WIN + 127) / 128 + 1) {
                const u64 po = (pft * WT_TILE - WT_PRE) + (u64)lane * 128;
                if (po < P.nbytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.text + po));
            }
        }
        // ---- long-line regime (P / W / sequence lines of megabytes, SURVEY 8a row 9: "skipped at full bandwidth"):
        // when this warp's previous tile held at most one record, first look for a line start at all -- a '\n' in
        // window bytes [WT_PRE - 1, WT_PRE + WT_TILE - 1) -- and leave the tile without building any mask if there is none
        if (quick) {
            u32 f = 0;
#pragma unroll
            for (int k = 0; k < WT_TILE / 16 / 32; k++) {
                const uint4 v = reinterpret_cast<const uint4*>(win)[WT_PRE / 16 + lane + 32 * k];
                f |= eq_bytes(v.x, 0x0A0A0A0Au) | eq_bytes(v.y, 0x0A0A0A0Au) | eq_bytes(v.z, 0x0A0A0A0Au) | eq_bytes(v.w, 0x0A0A0A0Au);
            }
            // byte WT_PRE - 1 starts a line at the tile's first byte (the tile's last byte is looked at too: harmless, the
            // full path decides)
            if (lane == 0 && win[WT_PRE - 1] == '\n') f |= 1u;
            if (!__any_sync(0xffffffffu, f != 0)) {
                if (lane == 0) {
                    TileInfo ti;
                    ti.n_rec = 0; ti.n_edge = 0; ti.edge_alloc = 0; ti.pad = 0;
                    P.tile_info[tile] = ti;
                }
                __syncwarp();
                fetch_next();
                continue;
            }
        }
        // ---- classify '\n' and '\t' 16 bytes at a time into the two bitmasks; a tile that follows a sparse one gets the
        // newline mask only (less than half the work) and the separator mask later, if it turns out to need one
        bool dense = !sparse;
        if (dense) {
#pragma unroll
            for (int k = 0; k < WT_WIN / 16 / 32; k++) {
                const u32 piece = lane + 32 * k;
                const uint4 v = reinterpret_cast<const uint4*>(win)[piece];
                const u32 ww[4] = {v.x, v.y, v.z, v.w};
                u32 mn = 0, mt = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    mn |= flags4(eq_bytes(ww[b], 0x0A0A0A0Au)) << (4 * b);
                    mt |= flags4(eq_bytes(ww[b], 0x09090909u)) << (4 * b);
                }
                reinterpret_cast<unsigned short*>(nlm)[piece] = (unsigned short)mn;
                reinterpret_cast<unsigned short*>(spm)[piece] = (unsigned short)(mn | mt);
            }
        } else {
            // line starts lie in window bytes [WT_PRE - 1, WT_PRE + WT_TILE): mask words 0 .. WT_TILE / 32
#pragma unroll
            for (int k = 0; k < (WT_PRE + WT_TILE) / 16 / 32 + 1; k++) {
                const u32 piece = lane + 32 * k;
                if (piece < (WT_PRE + WT_TILE) / 16) {
                    const uint4 v = reinterpret_cast<const uint4*>(win)[piece];
                    const u32 mn = flags4(eq_bytes(v.x, 0x0A0A0A0Au)) | (flags4(eq_bytes(v.y, 0x0A0A0A0Au)) << 4) |
                                   (flags4(eq_bytes(v.z, 0x0A0A0A0Au)) << 8) | (flags4(eq_bytes(v.w, 0x0A0A0A0Au)) << 12);
                    reinterpret_cast<unsigned short*>(nlm)[piece] = (unsigned short)mn;
                }
            }
        }
        __syncwarp();
        Tile t{win, nlm, spm, wbase};
        // ---- line starts in my 64-byte chunk: a line starts right after every '\n'
        const u32 c = 1 + lane * 2;  // mask word of my first 32 bytes (window offset 32 + 64 lane)
        const u64 nl = (u64)nlm[c] | ((u64)nlm[c + 1] << 32);
        u64 ls = (nl << 1) | (u64)(nlm[c - 1] >> 31);
        const u64 chunk0 = t0 + (u64)lane * 64;
        if (chunk0 >= P.nbytes) ls = 0;
        else if (chunk0 + 64 > P.nbytes) ls &= (1ull << (P.nbytes - chunk0)) - 1;
        const u32 woff0 = WT_PRE + lane * 64;
        // ---- classify my lines; keep only yielded records (bit set in `rec`), edge records in `edg`
        u64 rec = 0, edg = 0;
        for (u64 m = ls; m; m &= m - 1) {
            const int bit = __ffsll((long long)m) - 1;
            const u32 off = woff0 + bit;
            const uint8_t c0 = win[off], c1 = win[off + 1];
            const bool known = (c0 == 'S' || c0 == 'L' || c0 == 'P' || c0 == 'E' || c0 == 'C' || c0 == 'O');
            if (known && (c1 == '\t' || c1 == '\n')) {
                rec |= 1ull << bit;
                if (c0 == 'L' || c0 == 'E' || c0 == 'C') edg |= 1ull << bit;
            } else if (!known && c0 != 'H' && c0 != 'F') {
                const u64 val = ((wbase + off) << 8) | c0;
                if (~val > ld_volatile_u64(&P.cnt->first_unknown_inv)) atomicMax(&P.cnt->first_unknown_inv, ~val);
            }
        }
        // ---- warp scan of (records, edges): index of my first record / edge inside the tile
        const u32 packed = ((u32)__popcll(rec) << 16) | (u32)__popcll(edg);
        const u32 inc = warp_incl_scan(packed);
        const u32 tot = __shfl_sync(0xffffffffu, inc, 31);
        const u32 n_rec_tile = tot >> 16, n_edge_tile = tot & 0xFFFFu;
        const u32 my_rec0 = (inc - packed) >> 16, my_edge0 = (inc - packed) & 0xFFFFu;
        if (!dense && (n_edge_tile != 0 || n_rec_tile > 24u)) {
            // not that sparse after all: the separator mask (and the newline mask of the look-ahead) now
#pragma unroll
            for (int k = 0; k < WT_WIN / 16 / 32; k++) {
                const u32 piece = lane + 32 * k;
                const uint4 v = reinterpret_cast<const uint4*>(win)[piece];
                const u32 ww[4] = {v.x, v.y, v.z, v.w};
                u32 mn = 0, mt = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    mn |= flags4(eq_bytes(ww[b], 0x0A0A0A0Au)) << (4 * b);
                    mt |= flags4(eq_bytes(ww[b], 0x09090909u)) << (4 * b);
                }
                reinterpret_cast<unsigned short*>(nlm)[piece] = (unsigned short)mn;
                reinterpret_cast<unsigned short*>(spm)[piece] = (unsigned short)(mn | mt);
            }
            dense = true;
            __syncwarp();
        }
        // one atomicAdd claims this tile's range of edge_slots; the abort flag rides along.  Both results are
        // consumed only after the first round of lines has been parsed, so their latency is hidden.
        u32 alloc_l0 = 0, flags_l0 = 0;
        if (lane == 0) {
            if (n_edge_tile) alloc_l0 = atomicAdd(&P.cnt->edge_alloc, n_edge_tile);
            flags_l0 = ld_volatile_u32(&P.cnt->flags);
        }
        u32 alloc = 0;
        bool alloc_ready = false;
        // ---- parse and enqueue.  Record lines are compacted into `list` (WT_LIST per batch; one batch
        // unless the tile holds very short lines) and handed out one line per lane per round, so
        // neighbouring lanes parse neighbouring lines; full groups of 32 mentions are drained after every round.
        for (u32 lo = 0; lo < n_rec_tile; lo += WT_LIST) {
            {
                u32 ri = my_rec0, ei = my_edge0;
                for (u64 m = rec; m; m &= m - 1) {
                    const int bit = __ffsll((long long)m) - 1;
                    G2N_CHECK(woff0 + bit < WT_WIN && ei < 0x10000u);
                    if (ri - lo < WT_LIST) list[ri - lo] = (woff0 + bit) | (ei << 16);
                    ri++;
                    ei += (u32)((edg >> bit) & 1);
                }
            }
            __syncwarp();
            const u32 nb = min(n_rec_tile - lo, (u32)WT_LIST);
            for (u32 i0 = 0; i0 < nb; i0 += 32) {  // uniform trip count: the shuffles below need every lane
                const u32 i = i0 + lane;
                LineOut L;
                L.nm = 0; L.edge = false;
                bool ok = true;
                u32 off = 0, eidx = 0;
                if (i < nb) {
                    const u32 ent = list[i];
                    off = ent & 0xFFFFu;
                    eidx = ent >> 16;
                    ok = dense ? parse_line_fast<MODE>(P, t, off, L) : parse_line_sparse<MODE>(t, off, L);
                }
                if (!alloc_ready) {
                    alloc = __shfl_sync(0xffffffffu, alloc_l0, 0);
                    alloc_ready = true;
                }
                if (i < nb && !ok) defer_line(P, wbase + off, tile, lo + i, eidx);
                if (!ok) L.nm = 0;
                // positions of my mentions in the ring
                const u32 ninc = warp_incl_scan(L.nm);
                const u32 pos0 = qh + qn + ninc - L.nm;
                if (L.nm) {
                    u32 meta = (lo + i) << 2;
                    const u32 edge_ord = alloc + eidx;
                    if (L.edge) {
                        meta |= QM_EDGE;
                        if ((MODE & TM_WEIGHT) && edge_ord < P.edge_cap) {
                            P.edge_w[edge_ord] = L.w;
                            if (P.dtype == G2N_DTYPE_F32 && isfinite(L.w) && isinf((float)L.w)) atomicOr(&P.cnt->flags, CF_CAST_OVERFLOW);
                        }
                    }
                    // registration order u:of, v:ot, v:flip(ot), u:flip(of)  (builders.py:230-234); S: id[:+], id:-
                    const u32 c0m = (L.edge && cm_counts(P.count_mode, 0)) ? QM_CNT : 0u;
                    enqueue_mention<MODE>(P, t, S, SM::wrap(pos0), L.uo, L.ul, L.ocu, meta | c0m, tile, edge_ord);
                    if (L.nm >= 2) {
                        const u32 c1m = (L.edge && cm_counts(P.count_mode, 1)) ? QM_CNT : 0u;
                        enqueue_mention<MODE>(P, t, S, SM::wrap(pos0 + 1), L.vo, L.vl, L.ocv, meta | 1u | c1m, tile, edge_ord);
                    }
                    if ((MODE & TM_FOUR) && L.nm == 4) {
                        const u32 cm = cm_counts(P.count_mode, 2) ? QM_CNT : 0u;
                        enqueue_mention<MODE>(P, t, S, SM::wrap(pos0 + 2), L.vo, L.vl, L.ocv == '+' ? '-' : '+', meta | 2u | cm, tile, edge_ord);
                        enqueue_mention<MODE>(P, t, S, SM::wrap(pos0 + 3), L.uo, L.ul, L.ocu == '+' ? '-' : '+', meta | 3u | cm, tile, edge_ord);
                    }
                }
                qn += __shfl_sync(0xffffffffu, ninc, 31);
                G2N_CHECK(qn <= SM::kCap);
                __syncwarp();
                drain<MODE>(P, S, qh, qn, false, pol_table, claimed, fly, have_fly);
            }
        }
        // ---- the window is no longer needed: request the next one
        __syncwarp();
        fetch_next();
        if (lane == 0) {
            TileInfo ti;
            ti.n_rec = n_rec_tile; ti.n_edge = n_edge_tile; ti.edge_alloc = alloc_l0; ti.pad = 0;
            P.tile_info[tile] = ti;
        }
        aborted = __shfl_sync(0xffffffffu, flags_l0, 0) & (CF_TABLE_FULL | CF_DEFER_FULL);
        quick = n_rec_tile <= 1;
        sparse = n_rec_tile <= 12u && n_edge_tile == 0;
        // new keys of this tile: one fire-and-forget atomic per warp (the host checks the load factor)
#pragma unroll
        for (int d = 16; d; d >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, d);
        if (lane == 0 && claimed) atomicAdd(&P.cnt->n_keys, claimed);
        claimed = 0;
        __syncwarp();  // every lane is done with this tile's masks and list
    }
    // an aborted pass (table / defer list full) must not leave a bulk copy in flight towards its shared memory
    if (aborted && fetched) mbar_wait(&S.bar, parity);
    // ---- what is still queued (the pass is repeated anyway after an abort)
    if (!aborted) {
        drain<MODE>(P, S, qh, qn, true, pol_table, claimed, fly, have_fly);
#pragma unroll
        for (int d = 16; d; d >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, d);
        if (lane == 0 && claimed) atomicAdd(&P.cnt->n_keys, claimed);
    }
}

}  // namespace g2n
