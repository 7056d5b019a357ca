/*
 * oracle/gfa_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * A single-threaded CPU restatement (plain C) of the reference's GFA -> COO
 * triplet path: the line tokenizer of gfa2network/parser.py and the matrix half
 * of parse_gfa() in gfa2network/builders.py.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (gfa2network_b200) never does, and fails loudly without its CUDA library.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement
 * against tests/golden/ *.json vectors that tools/gen_golden.py produced by importing
 * the real reference (/root/reference) in the build container, and -- when the
 * reference is importable -- against the live reference on randomized inputs.
 *
 * What lives here: tokenizer + node-ID assignment + triplet emission + node names.
 * What does not: duplicate summing / maximum(A, A.T) / format conversion.  Those are
 * SciPy (third-party, the same wheel on both sides); oracle/oracle.py calls SciPy
 * exactly where the reference does (builders.py:281-283, utils.py:55).
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/gfa2network/).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORA_OK 0
/* error kinds: numbering shared (by convention only) with include/g2n.h */
#define ORA_ERR_NONE 0
#define ORA_ERR_MALFORMED_L 1      /* parser.py:208-209 */
#define ORA_ERR_MALFORMED_E 2      /* parser.py:251-252 */
#define ORA_ERR_MALFORMED_C 3      /* parser.py:299-300 */
#define ORA_ERR_MALFORMED_P 4      /* parser.py:231-232 */
#define ORA_ERR_MALFORMED_O 5      /* parser.py:345-346 */
#define ORA_ERR_S_NO_ID 6          /* parser.py:163  fields[1] -> IndexError */
#define ORA_ERR_COMPACT_EMPTY 7    /* parser.py:220-221  u_field[-1] on b"" -> IndexError */
#define ORA_ERR_ORI_UTF8 8         /* parser.py:214,291,293,337,339  .decode() -> UnicodeDecodeError */
#define ORA_ERR_WEIGHT_OVERFLOW 9  /* builders.py:209  float(int) -> OverflowError */
#define ORA_ERR_UNSUPPORTED_NUM 10 /* non-ASCII numeric weight value: declared outside parity scope */

typedef struct {
    int32_t directed;
    int32_t bidirected;
    int32_t keep_directed_bidir;
    int32_t strip_orientation;
    const uint8_t *weight_tag; /* UTF-8 bytes of weight_tag, NULL/0 when falsy (builders.py:206) */
    int32_t weight_tag_len;
} ora_params;

typedef struct {
    int64_t n_nodes;
    int64_t n_triplets;
    int32_t *rows, *cols; /* emission order, builders.py:222-228 */
    double *data;
    uint8_t *names;    /* node names concatenated in ID order */
    int64_t *name_off; /* n_nodes + 1 */
    int64_t n_records; /* records yielded by the parser (S L E C P O), builders.py:163 lineno */
    int64_t n_edge_records;
    int32_t err_kind;
    int64_t err_offset; /* byte offset of the start of the offending line */
    int64_t err_aux_off, err_aux_len; /* offending field (ORI_UTF8) */
    int32_t unknown_byte; /* -1 if no unsupported record was seen before the error */
    int64_t unknown_offset;
} ora_result;

/* ------------------------------------------------------------------ dictionary */
typedef struct {
    uint64_t *hash; /* 0 = empty */
    int32_t *id;
    int64_t cap;
    int64_t n;
    uint8_t *arena;
    int64_t arena_len, arena_cap;
    int64_t *off; /* n+1 offsets into arena */
    int64_t off_cap;
} dict_t;

static uint64_t fnv1a(const uint8_t *p, int64_t n)
{
    uint64_t h = 1469598103934665603ULL;
    for (int64_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ULL; h ^= h >> 32;
    return h ? h : 1;
}

static void dict_init(dict_t *d)
{
    d->cap = 1 << 16;
    d->hash = calloc(d->cap, sizeof(uint64_t));
    d->id = malloc(d->cap * sizeof(int32_t));
    d->n = 0;
    d->arena_cap = 1 << 20; d->arena = malloc(d->arena_cap); d->arena_len = 0;
    d->off_cap = 1 << 16; d->off = malloc(d->off_cap * sizeof(int64_t)); d->off[0] = 0;
}

static void dict_grow(dict_t *d)
{
    int64_t ncap = d->cap * 2;
    uint64_t *nh = calloc(ncap, sizeof(uint64_t));
    int32_t *ni = malloc(ncap * sizeof(int32_t));
    for (int64_t i = 0; i < d->cap; i++) if (d->hash[i]) {
        int64_t j = d->hash[i] & (ncap - 1);
        while (nh[j]) j = (j + 1) & (ncap - 1);
        nh[j] = d->hash[i]; ni[j] = d->id[i];
    }
    free(d->hash); free(d->id);
    d->hash = nh; d->id = ni; d->cap = ncap;
}

/* node2idx lookup-or-insert: "if n not in node2idx: node2idx[n] = len(node2idx)"
 * (builders.py:194-198, 219-221).  The key is the concatenation a||b. */
static int32_t dict_get(dict_t *d, const uint8_t *a, int64_t alen, const uint8_t *b, int64_t blen)
{
    uint8_t stackbuf[256];
    uint8_t *key = stackbuf;
    int64_t klen = alen + blen;
    if (klen > (int64_t)sizeof(stackbuf)) key = malloc(klen);
    memcpy(key, a, alen);
    if (blen) memcpy(key + alen, b, blen);
    uint64_t h = fnv1a(key, klen);
    int64_t j = h & (d->cap - 1);
    int32_t res = -1;
    while (d->hash[j]) {
        if (d->hash[j] == h) {
            int32_t id = d->id[j];
            int64_t o = d->off[id], l = d->off[id + 1] - o;
            if (l == klen && memcmp(d->arena + o, key, klen) == 0) { res = id; break; }
        }
        j = (j + 1) & (d->cap - 1);
    }
    if (res < 0) {
        if (d->arena_len + klen > d->arena_cap) {
            while (d->arena_len + klen > d->arena_cap) d->arena_cap *= 2;
            d->arena = realloc(d->arena, d->arena_cap);
        }
        if (d->n + 2 > d->off_cap) { d->off_cap *= 2; d->off = realloc(d->off, d->off_cap * sizeof(int64_t)); }
        memcpy(d->arena + d->arena_len, key, klen);
        d->arena_len += klen;
        res = (int32_t)d->n;
        d->off[d->n + 1] = d->arena_len;
        d->hash[j] = h; d->id[j] = res;
        d->n++;
        if (d->n * 2 > d->cap) dict_grow(d);
    }
    if (key != stackbuf) free(key);
    return res;
}

/* ------------------------------------------------------------------ triplets */
typedef struct { int32_t *r, *c; double *v; int64_t n, cap; } trip_t;
static void trip_push(trip_t *t, int32_t r, int32_t c, double v)
{
    if (t->n == t->cap) {
        t->cap = t->cap ? t->cap * 2 : (1 << 16);
        t->r = realloc(t->r, t->cap * sizeof(int32_t));
        t->c = realloc(t->c, t->cap * sizeof(int32_t));
        t->v = realloc(t->v, t->cap * sizeof(double));
    }
    t->r[t->n] = r; t->c[t->n] = c; t->v[t->n] = v; t->n++;
}

/* ------------------------------------------------------------------ Python-isms */
typedef struct { const uint8_t *p; int64_t n; } span_t;

/* bytes.rstrip(b"+-")  (parser.py:222-223, 267-268; builders.py:203-204) */
static span_t rstrip_pm(span_t s)
{
    while (s.n > 0 && (s.p[s.n - 1] == '+' || s.p[s.n - 1] == '-')) s.n--;
    return s;
}

static int py_isspace(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

/* strict UTF-8 validity == bytes.decode() succeeding (parser.py:184, 214, 291...) */
static int utf8_valid(const uint8_t *s, int64_t n)
{
    int64_t i = 0;
    while (i < n) {
        uint8_t c = s[i];
        if (c < 0x80) { i++; continue; }
        if (c >= 0xC2 && c <= 0xDF) {
            if (i + 1 >= n || (s[i + 1] & 0xC0) != 0x80) return 0;
            i += 2;
        } else if (c >= 0xE0 && c <= 0xEF) {
            if (i + 2 >= n || (s[i + 1] & 0xC0) != 0x80 || (s[i + 2] & 0xC0) != 0x80) return 0;
            if (c == 0xE0 && s[i + 1] < 0xA0) return 0;
            if (c == 0xED && s[i + 1] > 0x9F) return 0;
            i += 3;
        } else if (c >= 0xF0 && c <= 0xF4) {
            if (i + 3 >= n || (s[i + 1] & 0xC0) != 0x80 || (s[i + 2] & 0xC0) != 0x80 || (s[i + 3] & 0xC0) != 0x80) return 0;
            if (c == 0xF0 && s[i + 1] < 0x90) return 0;
            if (c == 0xF4 && s[i + 1] > 0x8F) return 0;
            i += 4;
        } else return 0;
    }
    return 1;
}

/* Python int(x) on ASCII text: strip whitespace, optional sign, decimal digits with single
 * underscores between digits.  Returns 1 and the digit string (underscores removed, sign
 * split off) in buf on success, 0 on ValueError.  (parser.py:189, 256-259) */
static int py_int_syntax(const uint8_t *s, int64_t n, char *buf, int64_t *ndig, int *neg)
{
    while (n > 0 && py_isspace(s[0])) { s++; n--; }
    while (n > 0 && py_isspace(s[n - 1])) n--;
    *neg = 0;
    if (n > 0 && (s[0] == '+' || s[0] == '-')) { *neg = s[0] == '-'; s++; n--; }
    if (n == 0) return 0;
    int64_t k = 0;
    int prev_digit = 0;
    for (int64_t i = 0; i < n; i++) {
        uint8_t c = s[i];
        if (c >= '0' && c <= '9') { if (buf) buf[k] = (char)c; k++; prev_digit = 1; }
        else if (c == '_') {
            if (!prev_digit || i + 1 >= n || s[i + 1] < '0' || s[i + 1] > '9') return 0;
            prev_digit = 0;
        } else return 0;
    }
    if (k > 4300) return 0; /* CPython int-string digit limit -> ValueError */
    if (buf) buf[k] = 0;
    *ndig = k;
    return 1;
}

static int all_ascii(const uint8_t *s, int64_t n)
{
    for (int64_t i = 0; i < n; i++) if (s[i] >= 0x80) return 0;
    return 1;
}

static int ci_eq(const char *s, int64_t n, const char *lit)
{
    int64_t m = (int64_t)strlen(lit);
    if (n != m) return 0;
    for (int64_t i = 0; i < n; i++) {
        char c = s[i];
        if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
        if (c != lit[i]) return 0;
    }
    return 1;
}

/* Python float(str) on ASCII text (parser.py:194): whitespace stripped, underscores only
 * between digits, decimal grammar or inf/infinity/nan, correctly rounded (glibc strtod is). */
static int py_float(const uint8_t *s, int64_t n, double *out)
{
    while (n > 0 && py_isspace(s[0])) { s++; n--; }
    while (n > 0 && py_isspace(s[n - 1])) n--;
    if (n == 0) return 0;
    char *buf = malloc(n + 1);
    int64_t k = 0;
    for (int64_t i = 0; i < n; i++) {
        uint8_t c = s[i];
        if (c == '_') {
            if (i == 0 || i + 1 >= n || s[i - 1] < '0' || s[i - 1] > '9' || s[i + 1] < '0' || s[i + 1] > '9') { free(buf); return 0; }
            continue;
        }
        if (c == 0) { free(buf); return 0; }
        buf[k++] = (char)c;
    }
    buf[k] = 0;
    const char *q = buf;
    int64_t m = k;
    int neg = 0;
    if (m > 0 && (q[0] == '+' || q[0] == '-')) { neg = q[0] == '-'; q++; m--; }
    int ok = 0;
    if (ci_eq(q, m, "inf") || ci_eq(q, m, "infinity")) { *out = neg ? -INFINITY : INFINITY; ok = 1; }
    else if (ci_eq(q, m, "nan")) { *out = neg ? -NAN : NAN; ok = 1; }
    else {
        /* digits [. digits*] | . digits+ ; optional exponent */
        int64_t i = 0, nd = 0;
        while (i < m && q[i] >= '0' && q[i] <= '9') { i++; nd++; }
        if (i < m && q[i] == '.') { i++; while (i < m && q[i] >= '0' && q[i] <= '9') { i++; nd++; } }
        if (nd > 0) {
            if (i < m && (q[i] == 'e' || q[i] == 'E')) {
                i++;
                if (i < m && (q[i] == '+' || q[i] == '-')) i++;
                int64_t ne = 0;
                while (i < m && q[i] >= '0' && q[i] <= '9') { i++; ne++; }
                if (ne == 0) i = -1;
            }
            if (i == m) { *out = strtod(buf, NULL); ok = 1; }
        }
    }
    free(buf);
    return ok;
}

/* ------------------------------------------------------------------ tags -> weight */
/* Follows GFAParser._parse_tags (parser.py:179-204) restricted to the one key the builder
 * reads (builders.py:205-209).  state: has_num / w.  Returns error kind or 0. */
static int weight_from_tags(const span_t *fields, int nf, int first, const ora_params *pr, int *has_w, double *w)
{
    *has_w = 0;
    int huge = 0;
    if (!pr->weight_tag || pr->weight_tag_len == 0) return 0;
    for (int i = first; i < nf; i++) {
        const uint8_t *f = fields[i].p;
        int64_t n = fields[i].n;
        /* f.decode().split(":", 2) must give 3 parts */
        const uint8_t *c1 = memchr(f, ':', n);
        if (!c1) continue;
        const uint8_t *c2 = memchr(c1 + 1, ':', n - (c1 + 1 - f));
        if (!c2) continue;
        if (!utf8_valid(f, n)) continue;
        int64_t taglen = c1 - f;
        if (taglen != pr->weight_tag_len || memcmp(f, pr->weight_tag, taglen) != 0) continue;
        int64_t typlen = c2 - c1 - 1;
        const uint8_t *val = c2 + 1;
        int64_t vlen = n - (val - f);
        if (typlen == 1 && c1[1] == 'i') {
            if (!all_ascii(val, vlen)) return ORA_ERR_UNSUPPORTED_NUM;
            char *buf = malloc(vlen + 2);
            int64_t nd; int neg;
            if (py_int_syntax(val, vlen, buf, &nd, &neg)) {
                double d = strtod(buf, NULL); /* == float(int(value)), round-half-even */
                /* an int too large for a double only raises (builders.py:209) if a later tag
                 * does not overwrite it: remember it as +-inf, checked after the loop */
                huge = isinf(d);
                *w = (neg && d != 0.0) ? -d : d; *has_w = 1; /* float(int("-0")) is +0.0 */
            } /* else: ValueError -> entry left unchanged (parser.py:190-191) */
            free(buf);
        } else if (typlen == 1 && c1[1] == 'f') {
            if (!all_ascii(val, vlen)) return ORA_ERR_UNSUPPORTED_NUM;
            double d;
            if (py_float(val, vlen, &d)) { *w = d; *has_w = 1; huge = 0; }
        } else {
            /* B -> list, anything else -> str: not int/float, so weight falls back to 1.0
             * (builders.py:208) and overrides any earlier numeric value (dict overwrite) */
            *has_w = 0; huge = 0;
        }
    }
    if (*has_w && huge) return ORA_ERR_WEIGHT_OVERFLOW;
    return 0;
}

/* int(bytes) succeeds?  (parser.py:256-259, 303-306) */
static int py_int_ok(span_t s)
{
    int64_t nd; int neg;
    return py_int_syntax(s.p, s.n, NULL, &nd, &neg);
}

/* ------------------------------------------------------------------ main loop */
#define MAXF 64

typedef struct {
    span_t u, v;
    span_t of, ot; /* orientation strings as the reference stores them */
} edge_t;

static const uint8_t PLUS[1] = {'+'}, MINUS[1] = {'-'};

static span_t lit(const uint8_t *p) { span_t s = {p, 1}; return s; }

static void add_mat_edge(dict_t *d, trip_t *t, int graph_directed, double w,
                         span_t a, span_t asuf, span_t b, span_t bsuf)
{
    /* builders.py:218-228 */
    int32_t ia, ib;
    if (asuf.p) {
        /* key = a + b":" + ori */
        uint8_t *tmp = malloc(1 + asuf.n);
        tmp[0] = ':'; memcpy(tmp + 1, asuf.p, asuf.n);
        ia = dict_get(d, a.p, a.n, tmp, 1 + asuf.n);
        free(tmp);
        tmp = malloc(1 + bsuf.n);
        tmp[0] = ':'; memcpy(tmp + 1, bsuf.p, bsuf.n);
        ib = dict_get(d, b.p, b.n, tmp, 1 + bsuf.n);
        free(tmp);
    } else {
        ia = dict_get(d, a.p, a.n, NULL, 0);
        ib = dict_get(d, b.p, b.n, NULL, 0);
    }
    trip_push(t, ia, ib, w);
    if (!graph_directed) trip_push(t, ib, ia, w);
}

int ora_parse(const uint8_t *text, int64_t nbytes, const ora_params *pr, ora_result *res)
{
    dict_t d; dict_init(&d);
    trip_t t = {0};
    memset(res, 0, sizeof(*res));
    res->unknown_byte = -1;
    /* builders.py:143 */
    int graph_directed = pr->keep_directed_bidir || (!pr->bidirected && pr->directed);
    int64_t pos = 0;
    int64_t nrec = 0, nedge = 0;
    int fcap = 1 << 10;
    span_t *fields = malloc(sizeof(span_t) * fcap);
    int err = 0;
    while (pos < nbytes) {
        /* parser.py:114  "for line in fh" -- lines end at '\n' only */
        const uint8_t *nl = memchr(text + pos, '\n', nbytes - pos);
        int64_t lend = nl ? (nl - text) : nbytes; /* exclusive, without '\n' */
        int64_t lstart = pos;
        pos = nl ? lend + 1 : nbytes;
        uint8_t c0 = text[lstart]; /* line is never empty: it holds at least "\n" or a byte */
        /* parser.py:117-132 */
        if (!(c0 == 'S' || c0 == 'L' || c0 == 'P' || c0 == 'E' || c0 == 'C' || c0 == 'O')) {
            if (c0 != 'H' && c0 != 'F' && res->unknown_byte < 0) {
                res->unknown_byte = c0; res->unknown_offset = lstart;
            }
            continue;
        }
        /* parser.py:133  line.rstrip(b"\n").split(b"\t") ; only fields[0] of length 1 matches */
        const uint8_t *lp = text + lstart;
        int64_t ln = lend - lstart;
        if (!(ln == 1 || (ln > 1 && lp[1] == '\t'))) continue; /* e.g. b"Sx": matches no branch */
        int nf = 0;
        {
            int64_t s = 0;
            for (int64_t i = 0; i <= ln; i++) {
                if (i == ln || lp[i] == '\t') {
                    if (nf == fcap) { fcap *= 2; fields = realloc(fields, sizeof(span_t) * fcap); }
                    fields[nf].p = lp + s; fields[nf].n = i - s; nf++;
                    s = i + 1;
                    /* P/O/S lines can be enormous; nothing past a handful of fields matters for
                     * them (the reference splits everything; the result is the same) */
                    if ((c0 == 'P' || c0 == 'O') && nf >= 3) break;
                    if (c0 == 'S' && nf >= 2) break;
                }
            }
        }
        nrec++;
        if (c0 == 'S') {
            /* parser.py:135-163 ; only fields[1] reaches the matrix path */
            if (nf < 2) { err = ORA_ERR_S_NO_ID; res->err_offset = lstart; break; }
            span_t seg = fields[1];
            /* builders.py:190-198 */
            if (pr->bidirected) {
                uint8_t suf[2] = {':', '+'};
                dict_get(&d, seg.p, seg.n, suf, 2);
                suf[1] = '-';
                dict_get(&d, seg.p, seg.n, suf, 2);
            } else dict_get(&d, seg.p, seg.n, NULL, 0);
            continue;
        }
        if (c0 == 'P' || c0 == 'O') {
            /* parser.py:229-247, 343-361 ; yielded then ignored by builders.py:164-199 */
            if (nf < 3) { err = c0 == 'P' ? ORA_ERR_MALFORMED_P : ORA_ERR_MALFORMED_O; res->err_offset = lstart; break; }
            continue;
        }
        edge_t e; int tag_first;
        memset(&e, 0, sizeof(e));
        if (c0 == 'L') {
            /* parser.py:206-227 */
            if (nf < 5) { err = ORA_ERR_MALFORMED_L; res->err_offset = lstart; break; }
            if (fields[2].n == 1 && (fields[2].p[0] == '+' || fields[2].p[0] == '-')) {
                e.u = fields[1]; e.of = fields[2]; e.v = fields[3]; e.ot = fields[4];
                if (!utf8_valid(e.ot.p, e.ot.n)) {
                    err = ORA_ERR_ORI_UTF8; res->err_offset = lstart;
                    res->err_aux_off = e.ot.p - text; res->err_aux_len = e.ot.n; break;
                }
                tag_first = 6;
            } else {
                span_t uf = fields[1], vf = fields[2];
                if (uf.n == 0 || vf.n == 0) { err = ORA_ERR_COMPACT_EMPTY; res->err_offset = lstart; break; }
                uint8_t lu = uf.p[uf.n - 1], lv = vf.p[vf.n - 1];
                e.of = lit(lu == '-' ? MINUS : PLUS);
                e.ot = lit(lv == '-' ? MINUS : PLUS);
                e.u = rstrip_pm(uf); e.v = rstrip_pm(vf);
                tag_first = 4;
            }
        } else {
            /* parser.py:249-295 (E) and 297-341 (C) */
            int minf = c0 == 'E' ? 6 : 5;
            if (nf < minf) { err = c0 == 'E' ? ORA_ERR_MALFORMED_E : ORA_ERR_MALFORMED_C; res->err_offset = lstart; break; }
            if (nf >= 9 && py_int_ok(fields[3]) && py_int_ok(fields[4]) && py_int_ok(fields[6]) && py_int_ok(fields[7])) {
                span_t uf = fields[2], vf = fields[5];
                e.of = lit(uf.n && uf.p[uf.n - 1] == '-' ? MINUS : PLUS);
                e.ot = lit(vf.n && vf.p[vf.n - 1] == '-' ? MINUS : PLUS);
                e.u = rstrip_pm(uf); e.v = rstrip_pm(vf);
                tag_first = 9;
            } else {
                int b = c0 == 'E' ? 2 : 1;
                e.u = fields[b]; e.of = fields[b + 1]; e.v = fields[b + 2]; e.ot = fields[b + 3];
                if (!utf8_valid(e.of.p, e.of.n)) {
                    err = ORA_ERR_ORI_UTF8; res->err_offset = lstart;
                    res->err_aux_off = e.of.p - text; res->err_aux_len = e.of.n; break;
                }
                if (!utf8_valid(e.ot.p, e.ot.n)) {
                    err = ORA_ERR_ORI_UTF8; res->err_offset = lstart;
                    res->err_aux_off = e.ot.p - text; res->err_aux_len = e.ot.n; break;
                }
                tag_first = b + 4;
            }
        }
        nedge++;
        /* builders.py:199-234 */
        int has_w; double w = 1.0;
        int werr = weight_from_tags(fields, nf, tag_first, pr, &has_w, &w);
        if (werr) { err = werr; res->err_offset = lstart; break; }
        if (!has_w) w = 1.0;
        span_t u = e.u, v = e.v;
        if (pr->strip_orientation) { u = rstrip_pm(u); v = rstrip_pm(v); }
        span_t none = {NULL, 0};
        if (pr->bidirected) {
            add_mat_edge(&d, &t, graph_directed, w, u, e.of, v, e.ot);
            if (!pr->keep_directed_bidir) {
                /* builders.py:231-234: rev = "-" if ori == "+" else "+" */
                int of_plus = e.of.n == 1 && e.of.p[0] == '+';
                int ot_plus = e.ot.n == 1 && e.ot.p[0] == '+';
                add_mat_edge(&d, &t, graph_directed, w, v, lit(ot_plus ? MINUS : PLUS), u, lit(of_plus ? MINUS : PLUS));
            }
        } else {
            add_mat_edge(&d, &t, graph_directed, w, u, none, v, none);
        }
    }
    free(fields);
    res->n_records = nrec;
    res->n_edge_records = nedge;
    res->err_kind = err;
    res->n_nodes = d.n;
    res->n_triplets = t.n;
    res->rows = t.r; res->cols = t.c; res->data = t.v;
    res->names = d.arena; res->name_off = d.off;
    free(d.hash); free(d.id);
    return err;
}

void ora_free(ora_result *res)
{
    free(res->rows); free(res->cols); free(res->data);
    free(res->names); free(res->name_off);
    memset(res, 0, sizeof(*res));
}
