/* synth.c -- deterministic synthetic GFA text generator (host, plain C) for the benchmark
 * configurations of BASELINE.json / SURVEY.md 8(d).  Not part of the compute path: it only
 * produces input bytes (both the GPU path and the CPU oracle consume the same buffer).
 *
 *   kind 1: GFA-1   S\ts<i>\t<seq|*>  +  L\ts<u>\t<o>\ts<v>\t<o>\t0M
 *   kind 2: reference E dialect (parser.py:254-273)
 *           S\ts<i>\t<len>\t*  +  E\t*\ts<u><o>\t0\t<len>\ts<v><o>\t0\t<len>\t<len>M\tRC:f:<k/8>
 * endpoints: 90 % local (v = u + d, d in [1,8], clipped), 10 % uniform; 1 % of links repeat an
 * earlier link exactly; optional P / W lines walking every segment (0.1 % random skips);
 * optional interleaving of S and L lines in blocks.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static uint64_t sm64(uint64_t *s)
{
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

static uint8_t *put_u64(uint8_t *p, uint64_t v)
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = (uint8_t)tmp[--n];
    return p;
}

static uint8_t *put_str(uint8_t *p, const char *s)
{
    while (*s) *p++ = (uint8_t)*s++;
    return p;
}

typedef struct {
    uint64_t n_seg, n_link, seed, id_base;
    int32_t kind;          /* 1 | 2 */
    int32_t seq_mean;      /* 0: "*" ; > 0: random ACGT, geometric length with this mean (kind 1) */
    int32_t n_paths;       /* P lines walking all segments */
    int32_t n_walks;       /* W lines walking all segments */
    int32_t interleave;    /* 0: all S then all L ; > 0: alternate blocks of this many S / 3x L lines */
    int32_t header;        /* 1: emit H line */
    uint64_t uni_base, uni_n; /* id range of the 10 % "uniform" link targets (0,0: this shard's own segments);
                                 lets the shards of a multi-GPU run reference each other's segments */
} synth_params;

/* upper bound of the bytes g2n_synth writes */
uint64_t g2n_synth_bound(const synth_params *sp)
{
    uint64_t per_s = 32 + (sp->kind == 1 ? (uint64_t)sp->seq_mean * 24 + 64 : 16);
    uint64_t per_l = sp->kind == 1 ? 64 : 128;
    uint64_t walk = (uint64_t)(sp->n_paths + sp->n_walks) * (sp->n_seg * 24 + 64);
    uint64_t b = 64 + sp->n_seg * 40 + sp->n_link * per_l + walk;
    if (sp->seq_mean > 0) b += sp->n_seg * ((uint64_t)sp->seq_mean * 3) + (1 << 20) + sp->n_seg * per_s / 8;
    return b;
}

static uint8_t *emit_seg(uint8_t *p, const synth_params *sp, uint64_t i, uint64_t *rng, uint8_t *end)
{
    *p++ = 'S'; *p++ = '\t'; *p++ = 's';
    p = put_u64(p, sp->id_base + i);
    *p++ = '\t';
    if (sp->kind == 2) {
        p = put_u64(p, 100 + (sm64(rng) % 900));
        *p++ = '\t'; *p++ = '*';
    } else if (sp->seq_mean > 0) {
        /* geometric length, mean seq_mean, capped so the bound holds */
        uint64_t len = 1;
        const uint64_t thr = UINT64_MAX / (uint64_t)sp->seq_mean;
        while (sm64(rng) > thr && len < (uint64_t)sp->seq_mean * 20) len++;
        if (p + len + 8 > end) len = 1;
        for (uint64_t k = 0; k < len; k += 32) {
            uint64_t r = sm64(rng);
            for (uint64_t j = k; j < len && j < k + 32; j++) { *p++ = (uint8_t)"ACGT"[r & 3]; r >>= 2; }
        }
    } else {
        *p++ = '*';
    }
    *p++ = '\n';
    return p;
}

typedef struct { uint32_t u, v; uint8_t o; uint16_t k; } link_t;

static uint8_t *emit_link(uint8_t *p, const synth_params *sp, const link_t *l)
{
    const char of = (l->o & 1) ? '-' : '+', ot = (l->o & 2) ? '-' : '+';
    if (sp->kind == 1) {
        *p++ = 'L'; *p++ = '\t'; *p++ = 's'; p = put_u64(p, l->u);
        *p++ = '\t'; *p++ = (uint8_t)of; *p++ = '\t'; *p++ = 's'; p = put_u64(p, l->v);
        *p++ = '\t'; *p++ = (uint8_t)ot; *p++ = '\t'; *p++ = '0'; *p++ = 'M'; *p++ = '\n';
    } else {
        const uint64_t len = 10 + (l->k % 90);
        *p++ = 'E'; *p++ = '\t'; *p++ = '*'; *p++ = '\t'; *p++ = 's'; p = put_u64(p, l->u); *p++ = (uint8_t)of;
        *p++ = '\t'; *p++ = '0'; *p++ = '\t'; p = put_u64(p, len);
        *p++ = '\t'; *p++ = 's'; p = put_u64(p, l->v); *p++ = (uint8_t)ot;
        *p++ = '\t'; *p++ = '0'; *p++ = '\t'; p = put_u64(p, len); *p++ = '\t'; p = put_u64(p, len); *p++ = 'M';
        p = put_str(p, "\tRC:f:");
        /* k/8 for k in [1, 8000]: exactly representable, sums are order independent */
        const uint32_t k = 1 + (l->k % 8000);
        p = put_u64(p, k / 8);
        static const char *frac[8] = {"", ".125", ".25", ".375", ".5", ".625", ".75", ".875"};
        if (k % 8) p = put_str(p, frac[k % 8]); else p = put_str(p, ".0");
        *p++ = '\n';
    }
    return p;
}

uint64_t g2n_synth(uint8_t *buf, uint64_t cap, const synth_params *sp)
{
    uint8_t *p = buf, *end = buf + cap;
    uint64_t rng = sp->seed * 0x2545F4914F6CDD1DULL + 1;
    if (sp->header) p = put_str(p, "H\tVN:Z:1.0\n");
    link_t *links = (link_t *)malloc(sizeof(link_t) * (sp->n_link ? sp->n_link : 1));
    if (!links) return 0;
    const uint64_t n = sp->n_seg ? sp->n_seg : 1;
    for (uint64_t j = 0; j < sp->n_link; j++) {
        uint64_t r = sm64(&rng);
        if (j > 16 && (r % 100) == 0) { links[j] = links[sm64(&rng) % j]; continue; }
        link_t l;
        l.u = (uint32_t)(sp->id_base + sm64(&rng) % n);
        if ((r >> 8) % 10 == 0) l.v = (uint32_t)(sp->uni_n ? sp->uni_base + sm64(&rng) % sp->uni_n : sp->id_base + sm64(&rng) % n);
        else { uint64_t v = l.u + 1 + ((r >> 16) % 8); l.v = (uint32_t)(v >= sp->id_base + n ? sp->id_base + n - 1 : v); }
        l.o = (uint8_t)((r >> 24) & 3);
        l.k = (uint16_t)(r >> 32);
        links[j] = l;
    }
    uint64_t si = 0, li = 0;
    const uint64_t sblk = sp->interleave > 0 ? (uint64_t)sp->interleave : sp->n_seg + 1;
    const uint64_t lblk = sp->interleave > 0 ? (uint64_t)sp->interleave * (sp->n_seg ? (sp->n_link + sp->n_seg - 1) / sp->n_seg : 1) : sp->n_link + 1;
    while (si < sp->n_seg || li < sp->n_link) {
        for (uint64_t k = 0; k < sblk && si < sp->n_seg; k++, si++) {
            if (p + 64 > end) goto done;
            p = emit_seg(p, sp, si, &rng, end);
        }
        for (uint64_t k = 0; k < lblk && li < sp->n_link; k++, li++) {
            if (p + 160 > end) goto done;
            p = emit_link(p, sp, &links[li]);
        }
    }
    for (int h = 0; h < sp->n_paths + sp->n_walks; h++) {
        const int is_w = h >= sp->n_paths;
        if (p + 128 > end) goto done;
        if (is_w) {
            p = put_str(p, "W\tsample"); p = put_u64(p, (uint64_t)h); p = put_str(p, "\t1\tchr1\t0\t"); p = put_u64(p, sp->n_seg); *p++ = '\t';
        } else {
            p = put_str(p, "P\thap"); p = put_u64(p, (uint64_t)h); *p++ = '\t';
        }
        int first = 1;
        for (uint64_t i = 0; i < sp->n_seg; i++) {
            if (sm64(&rng) % 1000 == 0) continue;
            if (p + 40 > end) goto done;
            if (is_w) { *p++ = '>'; *p++ = 's'; p = put_u64(p, sp->id_base + i); }
            else { if (!first) *p++ = ','; *p++ = 's'; p = put_u64(p, sp->id_base + i); *p++ = '+'; }
            first = 0;
        }
        if (!is_w) { *p++ = '\t'; *p++ = '*'; }
        *p++ = '\n';
    }
done:
    free(links);
    return (uint64_t)(p - buf);
}
