// numparse.cuh -- Python int()/float() text semantics -> IEEE double, host+device.
//
// Replaces, on device, what the reference gets from CPython in
//   gfa2network/parser.py:189   tags[tag] = int(value)
//   gfa2network/parser.py:194   tags[tag] = float(value)
//   gfa2network/builders.py:209 w = float(val)
//   gfa2network/parser.py:256-259, 303-306   int(fields[k]) probes that select the E/C coord form
// i.e. ASCII-whitespace stripping, optional sign, single underscores between digits,
// inf/infinity/nan spellings, the 4300-digit int limit, and *correctly rounded*
// decimal -> double conversion (round-half-even), including float(int) for arbitrarily long ints.
//
// Conversion: Eisel-Lemire with a 128-bit power-of-five table for <= 19 significant digits
// (exact: Mushtak & Lemire, "Fast number parsing without fallback"); for longer digit strings
// the truncated value w and w+1 are both converted and, if they disagree, an exact big-integer
// comparison against the half-way point decides.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define G2N_HD __host__ __device__ __forceinline__
#define G2N_HDN __host__ __device__
#else
#define G2N_HD inline
#define G2N_HDN
#endif

#include "pow5_table.h"
static const unsigned long long G2N_POW5_HOST[G2N_POW5_COUNT] = {G2N_POW5_VALUES};
#if defined(__CUDACC__)
__device__ const unsigned long long G2N_POW5_DEV[G2N_POW5_COUNT] = {G2N_POW5_VALUES};
#endif
#if defined(__CUDA_ARCH__)
#define G2N_POW5_TAB G2N_POW5_DEV
#else
#define G2N_POW5_TAB G2N_POW5_HOST
#endif

namespace g2n {

enum NumStatus { NUM_OK = 0, NUM_BAD = 1, NUM_OVERFLOW = 2, NUM_NONASCII = 3 };

G2N_HD bool py_isspace(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }
G2N_HD bool is_digit(uint8_t c) { return (uint8_t)(c - '0') < 10; }

G2N_HD void mul64(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo)
{
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    lo = (uint64_t)p;
    hi = (uint64_t)(p >> 64);
#endif
}

G2N_HD int clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

G2N_HD double bits_to_double(uint64_t b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    __builtin_memcpy(&d, &b, 8);
    return d;
#endif
}

// Eisel-Lemire: w * 10^q -> (mantissa bits, biased exponent) packed as IEEE bits (sign excluded).
G2N_HDN inline uint64_t eisel_lemire_bits(uint64_t w, int64_t q)
{
    if (w == 0 || q < G2N_POW5_MIN_Q) return 0;
    if (q > G2N_POW5_MAX_Q) return 0x7FFull << 52;
    int lz = clz64(w);
    w <<= lz;
    const int idx = 2 * (int)(q - G2N_POW5_MIN_Q);
    uint64_t hi, lo;
    mul64(w, G2N_POW5_TAB[idx], hi, lo);
    if ((hi & 0x1FF) == 0x1FF) {
        uint64_t hi2, lo2;
        mul64(w, G2N_POW5_TAB[idx + 1], hi2, lo2);
        lo += hi2;
        if (hi2 > lo) hi++;
    }
    const int upperbit = (int)(hi >> 63);
    uint64_t mant = hi >> (upperbit + 64 - 52 - 3);
    int32_t power2 = (int32_t)((((152170 + 65536) * q) >> 16) + 63) + upperbit - lz + 1023;
    if (power2 <= 0) {
        if (-power2 + 1 >= 64) return 0;
        mant >>= -power2 + 1;
        mant += (mant & 1);
        mant >>= 1;
        power2 = (mant < (1ull << 52)) ? 0 : 1;
        return ((uint64_t)power2 << 52) | (mant & ~(1ull << 52));
    }
    if (lo <= 1 && q >= -4 && q <= 23 && (mant & 3) == 1) {
        if ((mant << (upperbit + 64 - 52 - 3)) == hi) mant &= ~1ull;
    }
    mant += (mant & 1);
    mant >>= 1;
    if (mant >= (2ull << 52)) {
        mant = 1ull << 52;
        power2++;
    }
    mant &= ~(1ull << 52);
    if (power2 >= 0x7FF) return 0x7FFull << 52;
    return ((uint64_t)power2 << 52) | mant;
}

// ---------------------------------------------------------------- exact tie-breaker (rare)
struct BigNat {
    static const int CAP = 140;
    uint32_t v[CAP];
    int n;
    G2N_HDN void set(uint64_t x)
    {
        n = 0;
        while (x) { v[n++] = (uint32_t)x; x >>= 32; }
    }
    G2N_HDN void mul_add(uint32_t m, uint32_t a)
    {
        uint64_t carry = a;
        for (int i = 0; i < n; i++) {
            uint64_t t = (uint64_t)v[i] * m + carry;
            v[i] = (uint32_t)t;
            carry = t >> 32;
        }
        if (carry && n < CAP) v[n++] = (uint32_t)carry;
    }
    G2N_HDN void mul_pow5(int k)
    {
        while (k >= 13) { mul_add(1220703125u, 0); k -= 13; }
        uint32_t m = 1;
        while (k-- > 0) m *= 5;
        if (m != 1) mul_add(m, 0);
    }
    G2N_HDN void shl(int s)
    {
        if (n == 0 || s <= 0) return;
        int ws = s >> 5, bs = s & 31;
        int nn = n + ws + 1;
        if (nn > CAP) nn = CAP;
        for (int i = nn - 1; i >= 0; i--) {
            int src = i - ws;
            uint32_t x = 0;
            if (src >= 0 && src < n) x = v[src] << bs;
            if (bs && src - 1 >= 0 && src - 1 < n) x |= v[src - 1] >> (32 - bs);
            v[i] = x;
        }
        n = nn;
        while (n > 0 && v[n - 1] == 0) n--;
    }
    G2N_HDN int cmp(const BigNat& o) const
    {
        if (n != o.n) return n < o.n ? -1 : 1;
        for (int i = n - 1; i >= 0; i--)
            if (v[i] != o.v[i]) return v[i] < o.v[i] ? -1 : 1;
        return 0;
    }
};

// ---------------------------------------------------------------- the parser
// Src: callable (int64 i) -> uint8_t over the value span [0, n).
template <class Src>
G2N_HDN inline int parse_py_number(const Src& src, int64_t n, bool is_float, bool check_ascii, double* out)
{
    int64_t a = 0, b = n;
    if (check_ascii)
        for (int64_t i = 0; i < n; i++)
            if (src(i) >= 0x80) return NUM_NONASCII;
    while (a < b && py_isspace(src(a))) a++;
    while (b > a && py_isspace(src(b - 1))) b--;
    bool neg = false;
    if (a < b && (src(a) == '+' || src(a) == '-')) { neg = src(a) == '-'; a++; }
    if (a >= b) return NUM_BAD;
    const uint64_t sign = neg ? (1ull << 63) : 0;
    if (is_float) {
        // inf / infinity / nan, case-insensitive, nothing else around
        const int64_t m = b - a;
        if (m == 3 || m == 8) {
            const char* lit_inf = "infinity";
            bool isinf = true;
            for (int64_t i = 0; i < m; i++) {
                uint8_t c = src(a + i);
                if (c >= 'A' && c <= 'Z') c = (uint8_t)(c + 32);
                if (c != (uint8_t)lit_inf[i]) { isinf = false; break; }
            }
            if (isinf) { *out = bits_to_double(sign | (0x7FFull << 52)); return NUM_OK; }
            if (m == 3) {
                uint8_t c0 = src(a) | 32, c1 = src(a + 1) | 32, c2 = src(a + 2) | 32;
                if (c0 == 'n' && c1 == 'a' && c2 == 'n') {
                    // CPython: float("nan") = +qNaN, float("-nan") = -qNaN
                    *out = bits_to_double(sign | 0x7FF8000000000000ull);
                    return NUM_OK;
                }
            }
        }
    }
    // ---- digits: [0-9_]* [. [0-9_]*] [e [+-] [0-9_]+]; '_' only between two digits
    uint64_t w = 0;          // first <= 19 significant digits
    int nsig = 0;            // significant digits consumed into w
    int64_t dropped = 0;     // significant digits after the 19th (integer part + fraction)
    int64_t exp10 = 0;       // decimal exponent applying to w
    int64_t ndig_total = 0;  // digit characters in the mantissa
    int64_t first_sig = -1;  // position of the first significant digit (for the slow path)
    bool seen_point = false;
    int64_t i = a;
    int64_t mant_end = b;
    bool prev_digit = false;
    for (; i < b; i++) {
        const uint8_t c = src(i);
        if (is_digit(c)) {
            ndig_total++;
            if (c != '0' || nsig > 0) {
                if (first_sig < 0) first_sig = i;
                if (nsig < 19) { w = w * 10 + (c - '0'); nsig++; if (seen_point) exp10--; }
                else { dropped++; if (!seen_point) exp10++; }
            } else if (seen_point) exp10--;  // leading zero after the point
            prev_digit = true;
        } else if (c == '_') {
            if (!prev_digit || i + 1 >= b || !is_digit(src(i + 1))) return NUM_BAD;
            prev_digit = false;
        } else if (c == '.' && is_float && !seen_point) {
            seen_point = true;
            prev_digit = false;
        } else {
            break;
        }
    }
    mant_end = i;
    if (ndig_total == 0) return NUM_BAD;
    if (!is_float) {
        if (i != b) return NUM_BAD;
        if (ndig_total > 4300) return NUM_BAD;  // CPython int-string digit limit -> ValueError
    } else if (i < b) {
        const uint8_t c = src(i);
        if (c != 'e' && c != 'E') return NUM_BAD;
        i++;
        bool eneg = false;
        if (i < b && (src(i) == '+' || src(i) == '-')) { eneg = src(i) == '-'; i++; }
        int64_t e = 0, ne = 0;
        prev_digit = false;
        for (; i < b; i++) {
            const uint8_t d = src(i);
            if (is_digit(d)) { if (e < 100000000) e = e * 10 + (d - '0'); ne++; prev_digit = true; }
            else if (d == '_') {
                if (!prev_digit || i + 1 >= b || !is_digit(src(i + 1))) return NUM_BAD;
                prev_digit = false;
            } else return NUM_BAD;
        }
        if (ne == 0) return NUM_BAD;
        exp10 += eneg ? -e : e;
    }
    if (w == 0) { *out = bits_to_double(is_float ? sign : 0); return NUM_OK; }  // float(int("-0")) is +0.0
    // decimal magnitude cut-offs (value = w.xxx * 10^(exp10 + nsig - 1 ...)); keep q in table range
    uint64_t bits;
    if (exp10 + nsig > 310) bits = 0x7FFull << 52;
    else if (exp10 + nsig < -330) bits = 0;
    else {
        bits = eisel_lemire_bits(w, exp10);
        if (dropped > 0) {
            const uint64_t bits_hi = eisel_lemire_bits(w + 1, exp10);
            if (bits_hi != bits) {
                // exact comparison of the full digit string against the half-way point above `bits`
                // value V = D * 10^E ; D = all significant digits (<= 768 kept, rest sticky)
                BigNat D;
                D.n = 0;
                int kept = 0;
                bool sticky = false;
                for (int64_t j = first_sig; j < mant_end; j++) {
                    const uint8_t c = src(j);
                    if (!is_digit(c)) continue;  // '.' and '_'
                    if (kept < 768) { D.mul_add(10, (uint32_t)(c - '0')); kept++; }
                    else if (c != '0') sticky = true;
                }
                // all S = 19 + dropped significant digits form Dfull * 10^(exp10 - dropped);
                // keeping the first `kept` of them: D * 10^E with
                const int64_t E = exp10 - dropped + ((19 + dropped) - kept);
                const uint64_t mant = bits & ((1ull << 52) - 1);
                const int bexp = (int)(bits >> 52);
                uint64_t m;
                int e2;
                if (bexp == 0) { m = mant; e2 = -1074; }
                else { m = mant | (1ull << 52); e2 = bexp - 1075; }
                BigNat P;
                P.set(2 * m + 1);
                const int s = e2 - 1;  // midpoint = P * 2^s
                int c;
                if (E >= 0) {
                    D.mul_pow5((int)E);
                    const int64_t sh = E - s;
                    if (sh >= 0) D.shl((int)sh); else P.shl((int)-sh);
                } else {
                    const int64_t k = -E;
                    P.mul_pow5((int)k);
                    const int64_t sh = s + k;
                    if (sh >= 0) P.shl((int)sh); else D.shl((int)-sh);
                }
                c = D.cmp(P);
                if (c == 0 && sticky) c = 1;
                if (c > 0 || (c == 0 && (m & 1))) bits = bits_hi;
            }
        }
    }
    if (!is_float && bits == (0x7FFull << 52)) return NUM_OVERFLOW;  // float(int) -> OverflowError
    *out = bits_to_double(sign | bits);
    return NUM_OK;
}

// int(bytes) succeeds?  bytes flavour: non-ASCII is a plain ValueError (parser.py:256-259)
template <class Src>
G2N_HDN inline bool py_int_ok(const Src& src, int64_t n)
{
    int64_t a = 0, b = n;
    while (a < b && py_isspace(src(a))) a++;
    while (b > a && py_isspace(src(b - 1))) b--;
    if (a < b && (src(a) == '+' || src(a) == '-')) a++;
    if (a >= b) return false;
    int64_t nd = 0;
    bool prev_digit = false;
    for (int64_t i = a; i < b; i++) {
        const uint8_t c = src(i);
        if (is_digit(c)) { nd++; prev_digit = true; }
        else if (c == '_') {
            if (!prev_digit || i + 1 >= b || !is_digit(src(i + 1))) return false;
            prev_digit = false;
        } else return false;
    }
    return nd <= 4300;
}

}  // namespace g2n
