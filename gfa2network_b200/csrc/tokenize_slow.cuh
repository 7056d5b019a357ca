// tokenize_slow.cuh -- generic byte-wise record parser: every record shape and every Python-ism the
// reference accepts (compact L, orientation-only E/C, odd orientation strings, lenient int()/float(),
// keys longer than the 15-byte inline form, lines longer than a tile window).  The hot kernel
// (tokenize.cuh) hands such lines over through a deferred-line list; because every node mention
// carries its record ordinal, the order in which lines are processed does not matter.
//
// Reference semantics (gfa2network/parser.py unless noted):
//   :133-134  split on TAB; only a 1-byte first field matches a record type
//   :135-163  S -> fields[1]           :206-227  L (GFA-1 and compact forms)
//   :249-295  E (coord / orientation)  :297-341  C            :229-247, 343-361  P / O (field count only)
//   :179-204  tags -> builders.py:205-209 weight
//   builders.py:190-234  node registration order
#pragma once
#include "numparse.cuh"
#include "table.cuh"

namespace g2n {

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
__device__ __noinline__ void process_tag(const ScanParams& P, const Win w, const Span f, WeightState& ws)
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

__device__ __noinline__ bool int_probe(const Win w, const Span f)
{
    SpanSrc src{w, f.off};
    return py_int_ok(src, (int64_t)f.len);
}

__device__ __forceinline__ void report_error(Counters* cnt, u64 line_off, int kind)
{
    const u64 v = (line_off << 8) | (u64)kind;
    if (~v > ld_volatile_u64(&cnt->first_error_inv)) atomicMax(&cnt->first_error_inv, ~v);
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
__device__ __forceinline__ void parse_line_body(const ScanParams& P, const Win& w, u64 p, u64 order0, u32 edge_ord, u32& claimed)
{
    const uint8_t c0 = w(p);
    Cursor cur{p + 2, w(p + 1) == '\n'};
    Span f1, f2, f3, f4, f5, f6, f7, f8, ft;
    if (c0 == 'S') {
        if (!next_field(w, cur, f1)) { report_error(P.cnt, p, G2N_PE_S_NO_ID); return; }
        if (P.bidirected) {
            KeyDesc k = node_key(P, w, f1, f1, '+', true);
            table_insert(P, w, k, order0, false, claimed);
            k.ori_char = '-';
            table_insert(P, w, k, order0 | 1, false, claimed);
        } else {
            KeyDesc k = node_key(P, w, f1, f1, 0, true);
            table_insert(P, w, k, order0, false, claimed);
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
    const u32 su = table_insert(P, w, ku, order0, cm_counts(P.count_mode, 0), claimed);
    const u32 sv = table_insert(P, w, kv, order0 | 1, cm_counts(P.count_mode, 1), claimed);
    u32 sv2 = 0, su2 = 0;
    if (P.slots_per_edge == 4) {
        // rev = "-" if ori == "+" else "+"   (builders.py:232-233)
        const bool of_plus = ku.ori_len == 1 && ku.ori_char == '+';
        const bool ot_plus = kv.ori_len == 1 && kv.ori_char == '+';
        KeyDesc kv2 = kv, ku2 = ku;
        kv2.ori_len = 1; kv2.ori_char = ot_plus ? '-' : '+';
        ku2.ori_len = 1; ku2.ori_char = of_plus ? '-' : '+';
        sv2 = table_insert(P, w, kv2, order0 | 2, cm_counts(P.count_mode, 2), claimed);
        su2 = table_insert(P, w, ku2, order0 | 3, cm_counts(P.count_mode, 3), claimed);
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

// One deferred line per thread.
__global__ void __launch_bounds__(128) k_tokenize_slow(const __grid_constant__ ScanParams P)
{
    const u32 n_defer = min(P.cnt->n_defer, P.defer_cap);  // nobody appends while this kernel runs
    u32 claimed = 0;
    Win w{nullptr, P.text, 0, P.nbytes, 0};
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_defer; i += gridDim.x * blockDim.x) {
        const DeferEnt d = P.defer[i];
        parse_line_body(P, w, d.off, make_order(d.tile, d.rec_idx, 0), P.tile_info[d.tile].edge_alloc + d.edge_idx, claimed);
    }
    if (claimed) {
        const u32 before = atomicAdd(&P.cnt->n_keys, claimed);
        if (before + claimed > P.table_max_keys) atomicOr(&P.cnt->flags, CF_TABLE_FULL);
    }
}

}  // namespace g2n
