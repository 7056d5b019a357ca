// tokenize.cuh -- K1+K2: fused line scan, field split, node-key hashing and edge-record emission.
//
// One pass over the text.  Each CTA takes 16 KiB byte tiles (ticketed, so a decoupled look-back can
// carry the record / edge-record ordinals across tiles), stages the tile plus a look-ahead window in
// shared memory, classifies newlines 16 bytes at a time, and lets each thread parse the lines that
// START in its 64-byte chunk.  Node keys are inserted straight into an open-addressing table
// (inline 15-byte keys, 128-bit CAS); the table keeps, per key, the minimum (record ordinal,
// sub-rank) -- the reference's first-appearance order (builders.py:194-198, 219-221).
//
// Reference semantics implemented here (gfa2network/parser.py unless noted):
//   :114-132  newline-only line split, first-byte filter, one-shot unknown-record warning
//   :133-134  split on TAB; only a 1-byte first field matches a record type
//   :135-163  S -> fields[1]           :206-227  L (GFA-1 and compact forms)
//   :249-295  E (coord / orientation)  :297-341  C            :229-247, 343-361  P / O (field count only)
//   :179-204  tags -> builders.py:205-209 weight
//   builders.py:190-234  node registration order and (in emit.cuh) triplet order
#pragma once
#include "common.cuh"
#include "numparse.cuh"
#include "../../include/g2n.h"

namespace g2n {

#define TK_TILE 16384
#define TK_LOOK 2032
#define TK_PRE 16
#define TK_WIN (TK_PRE + TK_TILE + TK_LOOK)
#define TK_THREADS 256
#define TK_CHUNK (TK_TILE / TK_THREADS)  // 64 bytes per thread

struct __align__(16) Slot {
    u64 k0, k1;      // key: <= 15 inline bytes + (len+1) in the top byte, or 0xFF-tagged hash for long keys
    u64 first_inv;   // ~min(order); 0 = never set.  order = record_ordinal << 2 | sub-rank
    u32 rep;         // long keys: 1 + index of a LongDesc holding the key's bytes
    u32 pad;
};

struct LongDesc {
    u64 base_off;
    u64 ori_off;
    u32 base_len;
    u32 ori_len;   // bytes of the orientation string (may be 0)
    u32 ori_char;  // used when ori_len == 1
    u32 has_ori;   // 1: key is base + ':' + ori
};

struct Counters {
    u64 first_error;    // min (line_offset << 8 | kind); ~0 if none
    u64 first_unknown;  // min (line_offset << 8 | first byte); ~0 if none
    u32 n_records;
    u32 n_edges;
    u32 n_keys;
    u32 n_long;
    u32 flags;
    u32 ticket;
    u32 scan_ticket;
    u32 collision;
    u64 nnz;
    u64 aux[4];
};
#define CF_TABLE_FULL 1u
#define CF_EDGE_FULL 2u
#define CF_LONG_FULL 4u
#define CF_CAST_OVERFLOW 8u

struct ScanParams {
    const uint8_t* text;
    u64 nbytes;
    Slot* table;
    u32 table_mask;
    u32 table_max_keys;
    u32* edge_slots;
    double* edge_w;
    u32 edge_cap;
    LongDesc* longs;
    u32 long_cap;
    u64* tile_state;
    Counters* cnt;
    u32 n_tiles;
    int bidirected;
    int slots_per_edge;  // 2, or 4 for bidirected without keep_directed_bidir
    int strip_orientation;
    int wt_len;
    int dtype;  // G2N_DTYPE_* the weights will be cast to (only used to flag float32 overflow)
    u64 seed;
    uint8_t wt[64];
};

// ---------------------------------------------------------------- byte window
struct Win {
    const uint8_t* sm;  // shared-memory copy of [base, base + TK_WIN)
    const uint8_t* g;
    u64 base;
    u64 n;
    __device__ __forceinline__ uint8_t operator()(u64 p) const
    {
        if (p >= n) return '\n';
        const u64 d = p - base;
        if (d < (u64)TK_WIN) return sm[d];
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

__device__ __forceinline__ void ld_key(const Slot* s, u64& k0, u64& k1)
{
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(k0), "=l"(k1) : "l"(s) : "memory");
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

// Lookup-or-insert; returns the slot index (0xFFFFFFFF if the table is full).
// `claimed` is incremented when this call created the key.
__device__ __forceinline__ u32 table_insert(const ScanParams& P, const Win& w, const KeyDesc& kd, u64 order, u32& claimed)
{
    u64 k0, k1;
    bool is_long;
    make_key(w, kd, P.seed, k0, k1, is_long);
    u32 i = (u32)mix64(k0 ^ (k1 * 0x9e3779b97f4a7c15ULL)) & P.table_mask;
    u32 probes = 0;
    while (true) {
        Slot* s = &P.table[i];
        u64 s0, s1;
        ld_key(s, s0, s1);
        if (s0 == 0 && s1 == 0) {
            cas128(s, k0, k1, s0, s1);
            if (s0 == 0 && s1 == 0) { claimed++; s0 = k0; s1 = k1; }
        }
        if (s0 == k0 && s1 == k1) break;
        i = (i + 1) & P.table_mask;
        if (++probes > 8192u || probes > P.table_mask) { atomicOr(&P.cnt->flags, CF_TABLE_FULL); return 0xFFFFFFFFu; }
    }
    Slot* s = &P.table[i];
    const u64 inv = ~order;
    if (ld_volatile_u64(&s->first_inv) < inv) atomicMax(&s->first_inv, inv);
    if (is_long) {
        // keep / verify the bytes behind a hashed key: every arrival is compared with some earlier
        // arrival, so all mentions that share the slot are byte-equal unless `collision` is raised
        u32 r = ld_volatile_u32(&s->rep);
        if (r == 0) {
            const u32 idx = atomicAdd(&P.cnt->n_long, 1u);
            if (idx >= P.long_cap) { atomicOr(&P.cnt->flags, CF_LONG_FULL); return i; }
            LongDesc d;
            d.base_off = kd.base_off; d.ori_off = kd.ori_off; d.base_len = kd.base_len;
            d.ori_len = kd.ori_len; d.ori_char = kd.ori_char; d.has_ori = kd.has_ori;
            P.longs[idx] = d;
            __threadfence();
            r = atomicExch(&s->rep, idx + 1);
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

// ---------------------------------------------------------------- field cursor
struct Cursor {
    u64 p;     // start of the next field
    bool eol;  // the previous field ended the line
};

// Reads the next TAB-separated field starting at c.p; returns false if the line is exhausted.
__device__ __forceinline__ bool next_field(const Win& w, Cursor& c, Span& f)
{
    if (c.eol) return false;
    u64 q = c.p;
    uint8_t ch;
    while ((ch = w(q)) != '\t' && ch != '\n') q++;
    f.off = c.p;
    f.len = (u32)(q - c.p);
    c.eol = (ch == '\n');
    c.p = q + 1;
    return true;
}

__device__ __forceinline__ Span rstrip_pm(const Win& w, Span s)
{
    while (s.len > 0) {
        const uint8_t c = w(s.off + s.len - 1);
        if (c != '+' && c != '-') break;
        s.len--;
    }
    return s;
}

__device__ __forceinline__ bool utf8_valid(const Win& w, Span s)
{
    u32 i = 0;
    const u32 n = s.len;
    while (i < n) {
        const uint8_t c = w(s.off + i);
        if (c < 0x80) { i++; continue; }
        uint8_t c1 = i + 1 < n ? w(s.off + i + 1) : 0, c2 = i + 2 < n ? w(s.off + i + 2) : 0, c3 = i + 3 < n ? w(s.off + i + 3) : 0;
        if (c >= 0xC2 && c <= 0xDF) {
            if ((c1 & 0xC0) != 0x80) return false;
            i += 2;
        } else if (c >= 0xE0 && c <= 0xEF) {
            if ((c1 & 0xC0) != 0x80 || (c2 & 0xC0) != 0x80) return false;
            if (c == 0xE0 && c1 < 0xA0) return false;
            if (c == 0xED && c1 > 0x9F) return false;
            i += 3;
        } else if (c >= 0xF0 && c <= 0xF4) {
            if ((c1 & 0xC0) != 0x80 || (c2 & 0xC0) != 0x80 || (c3 & 0xC0) != 0x80) return false;
            if (c == 0xF0 && c1 < 0x90) return false;
            if (c == 0xF4 && c1 > 0x8F) return false;
            i += 4;
        } else return false;
    }
    return true;
}

struct WeightState {
    double w;
    bool has;
    bool huge;  // the current value is an int too large for a double (OverflowError if it survives)
    int err;
};

// One tag field -> weight state.  parser.py:179-204 restricted to the key builders.py:206 reads.
__device__ __noinline__ void process_tag(const ScanParams& P, const Win& w, Span f, WeightState& ws)
{
    // f.decode().split(":", 2) must give three parts
    u32 c1 = f.len, c2 = f.len;
    for (u32 i = 0; i < f.len; i++) {
        if (w(f.off + i) == ':') {
            if (c1 == f.len) c1 = i;
            else { c2 = i; break; }
        }
    }
    if (c2 == f.len) return;
    if ((int)c1 != P.wt_len) return;
    for (int i = 0; i < P.wt_len; i++)
        if (w(f.off + i) != P.wt[i]) return;
    if (!utf8_valid(w, f)) return;
    const u32 typlen = c2 - c1 - 1;
    const uint8_t typ = typlen == 1 ? w(f.off + c1 + 1) : 0;
    if (typ == 'i' || typ == 'f') {
        SpanSrc src{w, f.off + c2 + 1};
        double v;
        const int st = parse_py_number(src, (int64_t)(f.len - c2 - 1), typ == 'f', true, &v);
        if (st == NUM_OK) { ws.w = v; ws.has = true; ws.huge = false; }
        else if (st == NUM_OVERFLOW) { ws.has = true; ws.huge = true; }
        else if (st == NUM_NONASCII) ws.err = G2N_PE_UNSUPPORTED_NUM;
        // NUM_BAD: ValueError swallowed, entry left unchanged (parser.py:190-191, 195-196)
    } else {
        ws.has = false;  // str / list value: builders.py:208 falls back to 1.0
        ws.huge = false;
    }
}

__device__ __noinline__ bool int_probe(const Win& w, Span f)
{
    SpanSrc src{w, f.off};
    return py_int_ok(src, (int64_t)f.len);
}

__device__ __forceinline__ void report_error(Counters* cnt, u64 line_off, int kind)
{
    const u64 v = (line_off << 8) | (u64)kind;
    if (v < ld_volatile_u64(&cnt->first_error)) atomicMin(&cnt->first_error, v);
}

struct EdgeParse {
    Span u, v;
    Span of, ot;       // orientation strings as the reference stores them
    u32 ofc, otc;      // literal orientation chars when synthesised (of.len/ot.len == 1 and off unused)
    bool of_lit, ot_lit;
};

__device__ __forceinline__ void set_ori_from_last(const Win& w, Span f, Span& o, u32& oc, bool& lit, bool compact_l)
{
    // compact L (parser.py:220-221): last byte if it is +/- else "+";  E/C coord (parser.py:265-266):
    // "-" if the field ends with "-" else "+"
    uint8_t last = f.len ? w(f.off + f.len - 1) : 0;
    if (compact_l) oc = (last == '+' || last == '-') ? last : '+';
    else oc = (last == '-') ? '-' : '+';
    o.off = 0; o.len = 1; lit = true;
}

__device__ __forceinline__ KeyDesc node_key(const ScanParams& P, const Win& w, Span base, Span o, u32 oc, bool lit)
{
    KeyDesc k;
    k.base_off = base.off; k.base_len = base.len;
    k.has_ori = P.bidirected ? 1u : 0u;
    if (!P.bidirected) { k.ori_len = 0; k.ori_off = 0; k.ori_char = 0; return k; }
    if (lit) { k.ori_len = 1; k.ori_char = oc; k.ori_off = 0; }
    else if (o.len == 1) { k.ori_len = 1; k.ori_char = w(o.off); k.ori_off = 0; }
    else { k.ori_len = o.len; k.ori_off = o.off; k.ori_char = 0; }
    return k;
}

// Handles one line that starts at global offset p.  rec_ord / edge_ord are this line's ordinals.
__device__ __forceinline__ void parse_line(const ScanParams& P, const Win& w, u64 p, u32 rec_ord, u32 edge_ord, u32& claimed)
{
    const uint8_t c0 = w(p);
    Cursor cur{p + 2, w(p + 1) == '\n'};
    Span f1, f2, f3, f4, f5, f6, f7, f8, ft;
    const u64 order0 = (u64)rec_ord << 2;
    if (c0 == 'S') {
        if (!next_field(w, cur, f1)) { report_error(P.cnt, p, G2N_PE_S_NO_ID); return; }
        if (P.bidirected) {
            KeyDesc k = node_key(P, w, f1, f1, '+', true);
            table_insert(P, w, k, order0, claimed);
            k.ori_char = '-';
            table_insert(P, w, k, order0 | 1, claimed);
        } else {
            KeyDesc k = node_key(P, w, f1, f1, 0, true);
            table_insert(P, w, k, order0, claimed);
        }
        return;
    }
    if (c0 == 'P' || c0 == 'O') {
        // >= 3 fields  <=>  the name field is followed by a TAB
        if (!next_field(w, cur, f1) || cur.eol) report_error(P.cnt, p, c0 == 'P' ? G2N_PE_MALFORMED_P : G2N_PE_MALFORMED_O);
        return;
    }
    EdgeParse e;
    e.of_lit = e.ot_lit = false; e.ofc = e.otc = 0;
    WeightState ws;
    ws.w = 1.0; ws.has = false; ws.huge = false; ws.err = 0;
    const bool want_w = P.wt_len > 0;
    if (c0 == 'L') {
        if (!next_field(w, cur, f1) || !next_field(w, cur, f2) || !next_field(w, cur, f3) || !next_field(w, cur, f4)) {
            report_error(P.cnt, p, G2N_PE_MALFORMED_L);
            return;
        }
        const uint8_t o2 = f2.len == 1 ? w(f2.off) : 0;
        if (o2 == '+' || o2 == '-') {
            e.u = f1; e.of = f2; e.v = f3; e.ot = f4;
            if (!utf8_valid(w, f4)) { report_error(P.cnt, p, G2N_PE_ORI_UTF8); return; }
            if (want_w) {
                next_field(w, cur, ft);  // overlap (fields[5])
                while (next_field(w, cur, ft)) process_tag(P, w, ft, ws);
            }
        } else {
            if (f1.len == 0 || f2.len == 0) { report_error(P.cnt, p, G2N_PE_COMPACT_EMPTY); return; }
            set_ori_from_last(w, f1, e.of, e.ofc, e.of_lit, true);
            set_ori_from_last(w, f2, e.ot, e.otc, e.ot_lit, true);
            e.u = rstrip_pm(w, f1);
            e.v = rstrip_pm(w, f2);
            if (want_w) {
                process_tag(P, w, f4, ws);  // tags = fields[4:]
                while (next_field(w, cur, ft)) process_tag(P, w, ft, ws);
            }
        }
    } else {
        // E: fields[2..7] ; C: fields[1..7] share the coord test on fields 3,4,6,7
        const bool isE = c0 == 'E';
        bool ok = next_field(w, cur, f1) && next_field(w, cur, f2) && next_field(w, cur, f3) && next_field(w, cur, f4);
        if (ok && isE) ok = next_field(w, cur, f5);
        if (!ok) { report_error(P.cnt, p, isE ? G2N_PE_MALFORMED_E : G2N_PE_MALFORMED_C); return; }
        bool have5 = isE ? true : next_field(w, cur, f5);
        bool have6 = have5 && next_field(w, cur, f6);
        bool have7 = have6 && next_field(w, cur, f7);
        bool have8 = have7 && next_field(w, cur, f8);
        bool coord = have8 && int_probe(w, f3) && int_probe(w, f4) && int_probe(w, f6) && int_probe(w, f7);
        if (coord) {
            set_ori_from_last(w, f2, e.of, e.ofc, e.of_lit, false);
            set_ori_from_last(w, f5, e.ot, e.otc, e.ot_lit, false);
            e.u = rstrip_pm(w, f2);
            e.v = rstrip_pm(w, f5);
            if (want_w) while (next_field(w, cur, ft)) process_tag(P, w, ft, ws);
        } else {
            if (isE) { e.u = f2; e.of = f3; e.v = f4; e.ot = f5; }
            else { e.u = f1; e.of = f2; e.v = f3; e.ot = f4; }
            if (!utf8_valid(w, e.of) || !utf8_valid(w, e.ot)) { report_error(P.cnt, p, G2N_PE_ORI_UTF8); return; }
            if (want_w) {
                // tags = fields[6:] (E) / fields[5:] (C)
                if (!isE && have5) process_tag(P, w, f5, ws);
                if (have6) process_tag(P, w, f6, ws);
                if (have7) process_tag(P, w, f7, ws);
                if (have8) process_tag(P, w, f8, ws);
                while (next_field(w, cur, ft)) process_tag(P, w, ft, ws);
            }
        }
    }
    if (ws.has && ws.huge) ws.err = G2N_PE_WEIGHT_OVERFLOW;  // builders.py:209 float(val)
    if (ws.err) { report_error(P.cnt, p, ws.err); return; }
    Span u = e.u, v = e.v;
    if (P.strip_orientation) { u = rstrip_pm(w, u); v = rstrip_pm(w, v); }
    // registration order u:of, v:ot, v:flip(ot), u:flip(of)  (builders.py:230-234)
    KeyDesc ku = node_key(P, w, u, e.of, e.ofc, e.of_lit);
    KeyDesc kv = node_key(P, w, v, e.ot, e.otc, e.ot_lit);
    const u32 su = table_insert(P, w, ku, order0, claimed);
    const u32 sv = table_insert(P, w, kv, order0 | 1, claimed);
    u32 sv2 = 0, su2 = 0;
    if (P.slots_per_edge == 4) {
        // rev = "-" if ori == "+" else "+"   (builders.py:232-233)
        const bool of_plus = ku.ori_len == 1 && ku.ori_char == '+';
        const bool ot_plus = kv.ori_len == 1 && kv.ori_char == '+';
        KeyDesc kv2 = kv, ku2 = ku;
        kv2.ori_len = 1; kv2.ori_char = ot_plus ? '-' : '+';
        ku2.ori_len = 1; ku2.ori_char = of_plus ? '-' : '+';
        sv2 = table_insert(P, w, kv2, order0 | 2, claimed);
        su2 = table_insert(P, w, ku2, order0 | 3, claimed);
    }
    if (edge_ord < P.edge_cap) {
        if (P.slots_per_edge == 4) {
            reinterpret_cast<uint4*>(P.edge_slots)[edge_ord] = make_uint4(su, sv, sv2, su2);
        } else {
            reinterpret_cast<uint2*>(P.edge_slots)[edge_ord] = make_uint2(su, sv);
        }
        if (want_w) {
            const double wv = ws.has ? ws.w : 1.0;
            P.edge_w[edge_ord] = wv;
            if (P.dtype == G2N_DTYPE_F32 && isfinite(wv) && isinf((float)wv)) atomicOr(&P.cnt->flags, CF_CAST_OVERFLOW);
        }
    }
}

// ---------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(TK_THREADS) k_tokenize(const ScanParams P)
{
    __shared__ __align__(16) uint8_t s_win[TK_WIN];
    __shared__ u32 s_nl[TK_TILE / 32 + 1];  // bit i: byte i of the tile is '\n' ; word 0 bit 31 of entry [-1] kept separately
    __shared__ u64 s_scan[TK_THREADS / 32 + 2];
    __shared__ u32 s_tile;
    __shared__ u64 s_base;
    __shared__ u32 s_prev_nl;

    const u32 tid = threadIdx.x;
    u32 claimed = 0;
    while (true) {
        if (tid == 0) s_tile = atomicAdd(&P.cnt->ticket, 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= P.n_tiles) break;
        const u64 t0 = (u64)tile * TK_TILE;
        const u64 wbase = t0 - TK_PRE;  // may wrap for tile 0: handled below
        // ---- stage [t0 - 16, t0 + TILE + LOOK) in shared memory, classify newlines of the tile
        for (u32 piece = tid; piece < TK_WIN / 16; piece += TK_THREADS) {
            const u64 g = wbase + (u64)piece * 16;  // global offset of this 16-byte piece
            uint4 v;
            if (tile == 0 && piece == 0) {
                v = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);  // virtual '\n' before byte 0
            } else if (g + 16 <= P.nbytes) {
                v = ld_nc_v4(P.text + g);
            } else {
                uint8_t tmp[16];
#pragma unroll
                for (int k = 0; k < 16; k++) tmp[k] = (g + k < P.nbytes) ? P.text[g + k] : (uint8_t)'\n';
                v = *reinterpret_cast<uint4*>(tmp);
            }
            reinterpret_cast<uint4*>(s_win)[piece] = v;
            if (piece >= 1 && piece <= TK_TILE / 16) {
                // 16-bit newline mask of this piece
                u32 m = 0;
                const u32 ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const u32 eq = __vcmpeq4(ww[k], 0x0A0A0A0Au) & 0x01010101u;  // 1 per matching byte
                    const u32 bits = (eq * 0x01020408u) >> 24;                  // gather to 4 bits (byte k -> bit k)
                    m |= (bits & 0xF) << (4 * k);
                }
                reinterpret_cast<unsigned short*>(s_nl)[piece - 1] = (unsigned short)m;
            }
            if (piece == 0) s_prev_nl = ((v.w >> 24) == 0x0A) ? 1u : 0u;
        }
        __syncthreads();
        Win w{s_win, P.text, wbase, P.nbytes};
        // ---- line starts in my 64-byte chunk: a line starts right after every '\n'
        const u32 c = tid * 2;  // index of my first 32-bit mask word
        const u64 nl = (u64)s_nl[c] | ((u64)s_nl[c + 1] << 32);
        const u32 prev = (tid == 0) ? s_prev_nl : (s_nl[c - 1] >> 31);
        u64 ls = (nl << 1) | prev;
        const u64 chunk0 = t0 + (u64)tid * TK_CHUNK;
        if (chunk0 >= P.nbytes) ls = 0;
        else if (chunk0 + 64 > P.nbytes) ls &= (1ull << (P.nbytes - chunk0)) - 1;
        // ---- pass 1: count yielded records and edge records
        u32 nrec = 0, nedge = 0;
        for (u64 m = ls; m; m &= m - 1) {
            const u32 off = tid * TK_CHUNK + (__ffsll((long long)m) - 1) + TK_PRE;
            const uint8_t c0 = s_win[off];
            const uint8_t c1 = (t0 + off - TK_PRE + 1 < P.nbytes) ? s_win[off + 1] : (uint8_t)'\n';
            const bool rec = (c0 == 'S' || c0 == 'L' || c0 == 'P' || c0 == 'E' || c0 == 'C' || c0 == 'O') && (c1 == '\t' || c1 == '\n');
            if (rec) { nrec++; if (c0 == 'L' || c0 == 'E' || c0 == 'C') nedge++; }
            else if (!(c0 == 'S' || c0 == 'L' || c0 == 'P' || c0 == 'E' || c0 == 'C' || c0 == 'O') && c0 != 'H' && c0 != 'F') {
                const u64 val = ((t0 + off - TK_PRE) << 8) | c0;
                if (val < ld_volatile_u64(&P.cnt->first_unknown)) atomicMin(&P.cnt->first_unknown, val);
            }
        }
        u64 total;
        const u64 packed = ((u64)nrec << 32) | nedge;
        const u64 excl = block_excl_scan64(packed, s_scan, &total);
        if (tid == 0) s_base = lookback_exclusive(P.tile_state, tile, total);
        __syncthreads();
        const u64 base = s_base + excl;
        u32 rec_ord = (u32)(base >> 32), edge_ord = (u32)base;
        // ---- pass 2: parse, hash, emit (skipped once a capacity overflow has been flagged: the host
        // grows the buffers and reruns)
        if (ld_volatile_u32(&P.cnt->flags) & CF_TABLE_FULL) ls = 0;
        for (u64 m = ls; m; m &= m - 1) {
            const u32 off = tid * TK_CHUNK + (__ffsll((long long)m) - 1) + TK_PRE;
            const uint8_t c0 = s_win[off];
            const uint8_t c1 = (t0 + off - TK_PRE + 1 < P.nbytes) ? s_win[off + 1] : (uint8_t)'\n';
            const bool rec = (c0 == 'S' || c0 == 'L' || c0 == 'P' || c0 == 'E' || c0 == 'C' || c0 == 'O') && (c1 == '\t' || c1 == '\n');
            if (!rec) continue;
            const bool is_edge = (c0 == 'L' || c0 == 'E' || c0 == 'C');
            parse_line(P, w, t0 + off - TK_PRE, rec_ord, edge_ord, claimed);
            rec_ord++;
            if (is_edge) edge_ord++;
        }
        if (tile == P.n_tiles - 1 && tid == TK_THREADS - 1) {
            // last thread of the last tile knows the grand totals
            P.cnt->n_records = rec_ord;
            P.cnt->n_edges = edge_ord;
            if (edge_ord > P.edge_cap) atomicOr(&P.cnt->flags, CF_EDGE_FULL);
        }
        // new keys of this tile: one atomic per warp
        {
            u32 wsum = claimed;
#pragma unroll
            for (int d = 16; d; d >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, d);
            if ((tid & 31) == 0 && wsum) {
                const u32 before = atomicAdd(&P.cnt->n_keys, wsum);
                if (before + wsum > P.table_max_keys) atomicOr(&P.cnt->flags, CF_TABLE_FULL);
            }
            claimed = 0;
        }
        __syncthreads();
    }
}

}  // namespace g2n
