// Exact number codecs shared by every kernel (and compiled for the host by the
// CPU unit-test harness tests/hostsim.cpp, so the arithmetic is checked in the
// build container as well as on the B200).
//
//   parse_double   == Python float(str)            (sofa_preprocess.py:115,117,161,261,266,...)
//   parse_i64      == Python int(str)              (:150-151,171-173,183-186,306-312)
//   fmt_fixed6     == glibc printf("%.6f")         (to_csv float_format, :1206-1211,1439-1444,1802-1807)
//   json_double    == pandas to_json (ujson, 10 digits) -> json.loads -> float.__repr__  (:343-374)
//   log10_cr/log2_cr == correctly rounded log (np.log10 :136, np.log2 :139)
//   interp1        == np.interp scalar path (:132)
//
// All double arithmetic that must match the reference operation-for-operation goes
// through d_add/d_mul/d_div so that nvcc never contracts it into FMAs.
#pragma once
#include <stdint.h>
#include "log_tables.inc"

#if defined(__CUDACC__)
#define SOFA_HD __host__ __device__ __forceinline__
#define SOFA_HD_COLD __host__ __device__ __noinline__
// number formatters: out of line in translation units that define SOFA_OUTLINE_NUM.  Measured on B200 (round 2): the
// fully inlined JSON emit kernels are 0.8-2.1 MB of SASS, but calling the formatters (plans through local memory)
// made json_emit 59 -> 63 ms and csv_emit 28 -> 33 ms per 100 M events, so nothing defines it
#ifdef SOFA_OUTLINE_NUM
#define SOFA_HD_FN static __host__ __device__ __noinline__
#else
#define SOFA_HD_FN __host__ __device__ __forceinline__
#endif
#else
#define SOFA_HD_FN inline
#define SOFA_HD inline
#define SOFA_HD_COLD __attribute__((noinline))
#endif

namespace sofa {

typedef unsigned __int128 u128;

enum { P_OK = 0, P_INVALID = 1, P_UNSUPPORTED = 2 };

// ----------------------------------------------------------------------------- tables
#define SOFA_P10D_INIT {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, \
                        1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22}
#define SOFA_P10U_INIT {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, \
                        100000000ull, 1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, \
                        10000000000000ull, 100000000000000ull, 1000000000000000ull, 10000000000000000ull, \
                        100000000000000000ull, 1000000000000000000ull, 10000000000000000000ull}
static const double h_P10D[23] = SOFA_P10D_INIT;
static const uint64_t h_P10U[20] = SOFA_P10U_INIT;
static const double h_LOG_INV[128] = SOFA_LOG_INV_INIT;
static const double h_L10_HI[128] = SOFA_L10_HI_INIT;
static const double h_L10_LO[128] = SOFA_L10_LO_INIT;
static const double h_L2_HI[128] = SOFA_L2_HI_INIT;
static const double h_L2_LO[128] = SOFA_L2_LO_INIT;
#if defined(__CUDACC__)
static __device__ const double d_P10D[23] = SOFA_P10D_INIT;
static __device__ const uint64_t d_P10U[20] = SOFA_P10U_INIT;
static __device__ const double d_LOG_INV[128] = SOFA_LOG_INV_INIT;
static __device__ const double d_L10_HI[128] = SOFA_L10_HI_INIT;
static __device__ const double d_L10_LO[128] = SOFA_L10_LO_INIT;
static __device__ const double d_L2_HI[128] = SOFA_L2_HI_INIT;
static __device__ const double d_L2_LO[128] = SOFA_L2_LO_INIT;
#endif
#if defined(__CUDA_ARCH__)
#define SOFA_TAB(name) d_##name
#else
#define SOFA_TAB(name) h_##name
#endif

// ----------------------------------------------------------------------------- IEEE ops
SOFA_HD double d_add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
SOFA_HD double d_sub(double a, double b) { return d_add(a, -b); }
SOFA_HD double d_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
SOFA_HD double d_div(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
SOFA_HD double d_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
SOFA_HD uint64_t d_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; __builtin_memcpy(&u, &x, 8); return u;
#endif
}
SOFA_HD double bits_d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; __builtin_memcpy(&x, &u, 8); return x;
#endif
}
SOFA_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
SOFA_HD int bitlen128(u128 v) {
    uint64_t hi = (uint64_t)(v >> 64), lo = (uint64_t)v;
    return hi ? 128 - clz64(hi) : 64 - clz64(lo);
}
// exact 2^e for -1022 <= e <= 1023
SOFA_HD double pow2d(int e) { return bits_d((uint64_t)(e + 1023) << 52); }

SOFA_HD bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31); }

// Round the integer q * 2^e2 (sticky = "there are non-zero bits below q") to the nearest
// double, ties to even.  Result must be a normal double (callers guarantee the range).
SOFA_HD double round_u128(u128 q, int e2, bool sticky) {
    int bl = bitlen128(q);
    if (bl <= 53) {
        // exact unless sticky (callers only pass sticky with bl >= 64)
        return d_mul((double)(uint64_t)q, pow2d(e2));
    }
    int sh = bl - 53;
    uint64_t mant = (uint64_t)(q >> sh);
    bool rb = (uint64_t)(q >> (sh - 1)) & 1;
    bool st = sticky || ((q & ((((u128)1) << (sh - 1)) - 1)) != 0);
    if (rb && (st || (mant & 1))) mant++;
    return d_mul((double)mant, pow2d(e2 + sh));
}

// ----------------------------------------------------------------------------- decimal -> double
// value = (-1)^neg * w * 10^e10, correctly rounded.  Supported: w < 2^53 with |e10| <= 22
// (one exact IEEE operation) or |e10| <= 19 (128-bit integer arithmetic).
SOFA_HD int dec_to_double(uint64_t w, int e10, bool neg, double* out) {
    double r;
    if (w == 0) {
        r = 0.0;
    } else if (w < (1ull << 53) && e10 >= -22 && e10 <= 22) {
        r = (double)w;
        r = e10 < 0 ? d_div(r, SOFA_TAB(P10D)[-e10]) : d_mul(r, SOFA_TAB(P10D)[e10]);
    } else if (e10 >= 0 && e10 <= 19) {
        r = round_u128((u128)w * SOFA_TAB(P10U)[e10], 0, false);
    } else if (e10 < 0 && e10 >= -19) {
        uint64_t d = SOFA_TAB(P10U)[-e10];
        int s = 64 + clz64(w);                 // w << s has its top bit at bit 127
        u128 num = (u128)w << s;
        u128 q = num / d;
        u128 rem = num - q * d;
        r = round_u128(q, -s, rem != 0);
    } else {
        return P_UNSUPPORTED;
    }
    *out = neg ? -r : r;
    return P_OK;
}

// Python float(s[0:n]).  P_INVALID = Python raises ValueError; P_UNSUPPORTED = Python
// accepts it (underscores, inf/nan, > 19 significant digits, huge exponents) but this
// codec does not: callers must fail loudly.
template <typename S>
SOFA_HD int parse_double(S s, int n, double* out) {
    int i = 0;
    while (i < n && is_space((unsigned char)s[i])) i++;
    while (n > i && is_space((unsigned char)s[n - 1])) n--;
    if (i >= n) return P_INVALID;
    bool neg = false;
    if (s[i] == '+' || s[i] == '-') { neg = (s[i] == '-'); i++; }
    uint64_t w = 0;
    int nd = 0, e10 = 0;
    bool any = false, lost = false;
    for (; i < n; i++) {
        unsigned c = (unsigned char)s[i] - '0';
        if (c > 9) break;
        any = true;
        if (nd < 19) { if (w | c) { w = w * 10 + c; nd++; } }
        else { e10++; lost |= (c != 0); }
    }
    if (i < n && s[i] == '.') {
        i++;
        for (; i < n; i++) {
            unsigned c = (unsigned char)s[i] - '0';
            if (c > 9) break;
            any = true;
            if (nd < 19) { if (w | c) { w = w * 10 + c; nd++; } e10--; }
            else lost |= (c != 0);
        }
    }
    if (any && i < n && (s[i] == 'e' || s[i] == 'E')) {
        int j = i + 1;
        bool eneg = false;
        if (j < n && (s[j] == '+' || s[j] == '-')) { eneg = (s[j] == '-'); j++; }
        int ex = 0, nde = 0;
        for (; j < n; j++) {
            unsigned c = (unsigned char)s[j] - '0';
            if (c > 9) break;
            if (ex < 100000) ex = ex * 10 + (int)c;
            nde++;
        }
        if (nde == 0) return P_INVALID;
        e10 += eneg ? -ex : ex;
        i = j;
    }
    if (!any || i != n) {
        // Python-valid spellings this codec refuses: digit-group underscores, inf / nan.
        for (int k = 0; k < n; k++) {
            char c = s[k];
            if (c == '_' || c == 'i' || c == 'I' || c == 'n' || c == 'N') return P_UNSUPPORTED;
        }
        return P_INVALID;
    }
    if (lost) {
        // > 19 significant digits: the value lies strictly between w and w+1 (times 10^e10);
        // accept when both ends round to the same double, refuse otherwise.
        double a, b;
        int s1 = dec_to_double(w, e10, neg, &a);
        int s2 = dec_to_double(w + 1, e10, neg, &b);
        if (s1 != P_OK || s2 != P_OK || a != b) return P_UNSUPPORTED;
        *out = a;
        return P_OK;
    }
    return dec_to_double(w, e10, neg, out);
}

// Python int(s[0:n]) for plain decimal spellings.
template <typename S>
SOFA_HD int parse_i64(S s, int n, int64_t* out) {
    int i = 0;
    while (i < n && is_space((unsigned char)s[i])) i++;
    while (n > i && is_space((unsigned char)s[n - 1])) n--;
    if (i >= n) return P_INVALID;
    bool neg = false;
    if (s[i] == '+' || s[i] == '-') { neg = (s[i] == '-'); i++; }
    if (i >= n) return P_INVALID;
    uint64_t v = 0;
    for (; i < n; i++) {
        unsigned c = (unsigned char)s[i] - '0';
        if (c > 9) return s[i] == '_' ? P_UNSUPPORTED : P_INVALID;
        if (v > 922337203685477580ull) return P_UNSUPPORTED;
        v = v * 10 + c;
    }
    if (v > 9223372036854775807ull) return P_UNSUPPORTED;
    *out = neg ? -(int64_t)v : (int64_t)v;
    return P_OK;
}

// ----------------------------------------------------------------------------- integer -> text
// 32-bit fast paths: almost every value in a trace row fits 32 bits, and 32-bit
// divide-by-constant is a multiply + shift on the GPU (64-bit is several instructions).
SOFA_HD int ndigits_u32(uint32_t v) {
    if (v < 10u) return 1;
    if (v < 100u) return 2;
    if (v < 1000u) return 3;
    if (v < 10000u) return 4;
    if (v < 100000u) return 5;
    if (v < 1000000u) return 6;
    if (v < 10000000u) return 7;
    if (v < 100000000u) return 8;
    if (v < 1000000000u) return 9;
    return 10;
}
SOFA_HD int ndigits_u64(uint64_t v) {
    if ((v >> 32) == 0) return ndigits_u32((uint32_t)v);
    if (v < 100000000000000ull) {                 // 10..14 digits
        uint32_t hi = (uint32_t)(v / 100000ull);  // < 1e9
        return 5 + ndigits_u32(hi);
    }
    uint64_t hi = v / 1000000000ull;              // >= 1e5
    return 9 + ndigits_u32((uint32_t)(hi > 0xffffffffull ? 0xffffffffu : hi)) + (hi > 9999999999ull ? 1 : 0);
}
template <typename O>
SOFA_HD void fmt_u32_pad(uint32_t v, int width, O out) {       // zero padded, exactly `width` (<= 10)
    for (int i = width - 1; i >= 0; i--) { uint32_t q = v / 10u; out[i] = (char)('0' + (v - q * 10u)); v = q; }
}
template <typename O>
SOFA_HD int fmt_u32(uint32_t v, O out) {
    int n = ndigits_u32(v);
    fmt_u32_pad(v, n, out);
    return n;
}
template <typename O>
SOFA_HD void fmt_u64_pad(uint64_t v, int width, O out) {       // zero padded, exactly `width`
    while (width > 9) {
        uint64_t q = v / 1000000000ull;
        fmt_u32_pad((uint32_t)(v - q * 1000000000ull), 9, out + (width - 9));
        v = q; width -= 9;
    }
    fmt_u32_pad((uint32_t)v, width, out);
}
template <typename O>
SOFA_HD int fmt_u64(uint64_t v, O out) {
    if ((v >> 32) == 0) return fmt_u32((uint32_t)v, out);
    int n = ndigits_u64(v);
    fmt_u64_pad(v, n, out);
    return n;
}
template <typename O>
SOFA_HD_FN int fmt_i64(int64_t v, O out) {
    if (v < 0) { out[0] = '-'; return 1 + fmt_u64((uint64_t)0 - (uint64_t)v, out + 1); }
    return fmt_u64((uint64_t)v, out);
}
SOFA_HD int len_i64(int64_t v) {
    return v < 0 ? 1 + ndigits_u64((uint64_t)0 - (uint64_t)v) : ndigits_u64((uint64_t)v);
}

// ----------------------------------------------------------------------------- "%.6f"
// Splits |x| (finite, < 2^63) into integer part and round-half-even(frac * 10^6) on the
// exact binary value, as glibc does.  Returns false when |x| >= 2^63.
SOFA_HD bool fixed6_parts(double x, bool* neg, uint64_t* ip, uint32_t* fp) {
    uint64_t b = d_bits(x);
    *neg = (b >> 63) != 0;
    int be = (int)((b >> 52) & 0x7ff);
    uint64_t m = b & 0xfffffffffffffull;
    int e;
    if (be == 0) { e = -1074; } else { m |= 1ull << 52; e = be - 1075; }
    if (e >= 11) return false;
    uint64_t I, F = 0;
    if (e >= 0) {
        I = m << e;
    } else {
        int s = -e;
        uint64_t fm;
        if (s >= 64) { I = 0; fm = m; } else { I = m >> s; fm = m & ((1ull << s) - 1); }
        if (s < 74) {
            u128 P = (u128)fm * 1000000u;
            u128 q = P >> s;
            u128 rem = P - (q << s);
            u128 half = ((u128)1) << (s - 1);
            if (rem > half || (rem == half && ((uint64_t)q & 1))) q += 1;
            F = (uint64_t)q;
            if (F == 1000000u) { I++; F = 0; }
        }
    }
    *ip = I; *fp = (uint32_t)F;
    return true;
}
// "%.6f" split into a plan (sign, integer part, 6 fraction digits, exact length) and an emit
// step, so kernels can size a row first and print its digits exactly once afterwards.
struct F6Plan { uint64_t I; uint32_t F; uint8_t neg, ok, len; };
SOFA_HD_FN F6Plan fixed6_plan(double x) {
    F6Plan p; bool neg;
    p.ok = fixed6_parts(x, &neg, &p.I, &p.F) ? 1 : 0;
    p.neg = neg ? 1 : 0;
    p.len = (uint8_t)(p.neg + ndigits_u64(p.I) + 7);
    return p;
}
template <typename O>
SOFA_HD_FN int fixed6_emit(const F6Plan& p, O out) {
    int n = 0;
    if (p.neg) out[n++] = '-';
    n += fmt_u64(p.I, out + n);
    out[n++] = '.';
    fmt_u32_pad(p.F, 6, out + n);
    return n + 6;
}
SOFA_HD int len_fixed6(double x) {          // length of "%.6f" (x finite, |x| < 2^63), -1 otherwise
    F6Plan p = fixed6_plan(x);
    return p.ok ? p.len : -1;
}
template <typename O>
SOFA_HD int fmt_fixed6(double x, O out) {
    F6Plan p = fixed6_plan(x);
    if (!p.ok) return -1;
    return fixed6_emit(p, out);
}

// ----------------------------------------------------------------------------- report.js floats
// pandas to_json prints float64 with its vendored ultrajson: whole part, then
// (value - whole) * 1e10 evaluated in double with its own half rule; json.loads turns that
// decimal string back into a double y; json.dump prints repr(y) (shortest round trip).
// json_plan decides the digits (kind 0 = outside the fixed-notation range of ultrajson,
// |v| > 1e16 or 0 < |v| < 1e-15, which this codec does not cover); json_emit prints them.
enum { JP_BAD = 0, JP_NULL = 1, JP_FIXED = 2, JP_SCI = 3 };
struct JPlan { uint64_t I, dg; uint8_t kind, neg, nd, ex, len; };
SOFA_HD JPlan jp_fixed(bool neg, uint64_t I, uint64_t dg, int nd) {
    JPlan p; p.kind = JP_FIXED; p.neg = neg ? 1 : 0; p.I = I; p.dg = dg; p.nd = (uint8_t)nd; p.ex = 0;
    p.len = (uint8_t)(p.neg + ndigits_u64(I) + 1 + nd);
    return p;
}
// Python repr digits of y for 1e5 <= y < 1e16 (fixed notation): shortest decimal that rounds
// back to y, found by an exact 128-bit test of the rounding interval; closest one on ties of length.
SOFA_HD_FN JPlan jp_shortest_ge1e5(bool neg, double y) {
    JPlan bad; bad.I = 0; bad.dg = 0; bad.neg = 0; bad.nd = 0; bad.ex = 0; bad.len = 0; bad.kind = JP_BAD;
    uint64_t yb = d_bits(y);
    int e = (int)((yb >> 52) & 0x7ff) - 1075;
    uint64_t m = (yb & 0xfffffffffffffull) | (1ull << 52);
    if (e >= 0) return jp_fixed(neg, m << e, 0, 1);        // integer valued (< 1e16): digits + ".0"
    int s = -e;                         // 1..36 because y >= 1e5 > 2^16
    if (s > 36) return bad;
    uint64_t I = m >> s;
    uint64_t r0 = m & ((1ull << s) - 1);
    if (r0 == 0) return jp_fixed(neg, I, 0, 1);
    bool even = (m & 1) == 0;
    bool pow2 = (m == (1ull << 52));    // lower neighbour is half as far
    u128 one = ((u128)1) << s;
    uint64_t p10 = 1;
    for (int j = 1; j <= 12; j++) {
        p10 *= 10;
        u128 F = (u128)r0 * p10;
        u128 fl = F >> s;
        u128 rem = F - (fl << s);
        u128 up = one - rem;
        // candidate below y: distance rem / (2^s 10^j); half gap 2^-(s+1) (2^-(s+2) at a power of two)
        u128 lo_lhs = pow2 ? rem * 4 : rem * 2;
        bool lo_ok = (lo_lhs < (u128)p10) || (even && lo_lhs == (u128)p10);
        u128 hi_lhs = up * 2;
        bool hi_ok = (hi_lhs < (u128)p10) || (even && hi_lhs == (u128)p10);
        if (!(lo_ok || hi_ok)) continue;
        uint64_t dg = (uint64_t)fl;
        bool take_hi;
        if (lo_ok && hi_ok) take_hi = (up < rem) || (up == rem && (dg & 1));
        else take_hi = hi_ok;
        if (take_hi) dg++;
        if (dg == p10) return jp_fixed(neg, I + 1, 0, 1);
        return jp_fixed(neg, I, dg, j);
    }
    return bad;
}
SOFA_HD_FN JPlan json_plan(double x) {
    JPlan p; p.I = 0; p.dg = 0; p.neg = 0; p.nd = 0; p.ex = 0; p.len = 0; p.kind = JP_BAD;
    uint64_t xb = d_bits(x);
    if (((xb >> 52) & 0x7ff) == 0x7ff) { p.kind = JP_NULL; p.len = 4; return p; }
    bool neg = x < 0;
    double v = neg ? -x : x;
    if (v > 1e16 || (v != 0.0 && v < 1e-15)) return p;
    uint64_t whole = (uint64_t)v;
    double tmp = d_mul(d_sub(v, (double)whole), 1e10);
    uint64_t frac = (uint64_t)tmp;
    double diff = d_sub(tmp, (double)frac);
    if (diff > 0.5) ++frac;
    else if (diff == 0.5 && (frac == 0 || (frac & 1))) ++frac;
    if (frac >= 10000000000ull) { frac = 0; ++whole; }
    if (whole < 100000ull) {
        // <= 15 significant digits: the decimal string itself is the shortest repr of y
        if (whole > 0 || frac >= 1000000ull || frac == 0) {
            if (frac == 0) return jp_fixed(neg, whole, 0, 1);
            int nz = 10;
            while (frac % 10 == 0) { frac /= 10; nz--; }
            return jp_fixed(neg, whole, frac, nz);
        }
        // 0 < y < 1e-4: Python repr switches to exponent notation ("2.9e-09", "1e-10")
        int ex = 11 - ndigits_u64(frac);          // 5..10
        while (frac % 10 == 0) frac /= 10;
        int L = ndigits_u64(frac);
        p.kind = JP_SCI; p.neg = neg ? 1 : 0; p.dg = frac; p.nd = (uint8_t)L; p.ex = (uint8_t)ex;
        p.len = (uint8_t)(p.neg + L + (L > 1 ? 1 : 0) + 4);
        return p;
    }
    // ---- whole >= 1e5: y = v when v >= 2^19 (the 10-digit decimal is within half an ulp),
    // else y = (whole*1e10 + frac) / 1e10 (numerator < 2^53, one exact division)
    double y = v;
    if (whole < 524288ull) y = d_div((double)(whole * 10000000000ull + frac), 1e10);
    return jp_shortest_ge1e5(neg, y);
}
template <typename O>
SOFA_HD_FN int json_emit(const JPlan& p, O out) {
    if (p.kind == JP_NULL) { out[0] = 'n'; out[1] = 'u'; out[2] = 'l'; out[3] = 'l'; return 4; }
    int n = 0;
    if (p.neg) out[n++] = '-';
    if (p.kind == JP_FIXED) {
        n += fmt_u64(p.I, out + n);
        out[n++] = '.';
        fmt_u64_pad(p.dg, p.nd, out + n);
        return n + p.nd;
    }
    char ds[8];
    fmt_u32_pad((uint32_t)p.dg, p.nd, ds);
    out[n++] = ds[0];
    if (p.nd > 1) { out[n++] = '.'; for (int k = 1; k < p.nd; k++) out[n++] = ds[k]; }
    out[n++] = 'e'; out[n++] = '-';
    out[n++] = (char)('0' + p.ex / 10); out[n++] = (char)('0' + p.ex % 10);
    return n;
}
template <typename O>
SOFA_HD int json_double(double x, O out) {
    JPlan p = json_plan(x);
    if (p.kind == JP_BAD) return -1;
    return json_emit(p, out);
}

// repr(float("%.6f" % x)): what DataFrame.to_csv prints (no float_format) for a value that was
// read back from a trace CSV - comm.csv (sofa_common.py:160-167).  |x| < 9e9.
SOFA_HD_FN JPlan repr6_plan(double x) {
    JPlan p; p.I = 0; p.dg = 0; p.neg = 0; p.nd = 0; p.ex = 0; p.len = 0; p.kind = JP_BAD;
    bool neg; uint64_t I; uint32_t F;
    if (!fixed6_parts(x, &neg, &I, &F) || I >= 9000000000ull) return p;
    if (I < 100000ull) {                       // <= 11 significant digits: the decimal itself
        if (I > 0 || F >= 100u || F == 0u) {
            if (F == 0u) return jp_fixed(neg, I, 0, 1);
            int nd = 6; uint32_t f = F;
            while (f % 10u == 0u) { f /= 10u; nd--; }
            return jp_fixed(neg, I, f, nd);
        }
        int ex = 7 - ndigits_u32(F);           // 0 < y < 1e-4: "2e-06", "1.5e-05"
        uint32_t f = F;
        while (f % 10u == 0u) f /= 10u;
        int L = ndigits_u32(f);
        p.kind = JP_SCI; p.neg = neg ? 1 : 0; p.dg = f; p.nd = (uint8_t)L; p.ex = (uint8_t)ex;
        p.len = (uint8_t)(p.neg + L + (L > 1 ? 1 : 0) + 4);
        return p;
    }
    return jp_shortest_ge1e5(neg, d_div((double)(I * 1000000ull + F), 1e6));
}

// repr(x) of an arbitrary double - what DataFrame.to_csv prints without float_format
// (usr_sys.csv, netbandwidth.csv: sofa_preprocess.py:545, 1323).  Shortest decimal that rounds
// back to x (closest one, even digit on a tie), found by an exact 128-bit test of the rounding
// interval.  Covers 0, integers below 1e16 and 1e-6 <= |x| < 1e16; anything else is JP_BAD.
// kind JP_SCI here carries up to 17 digits: print with repr_emit, not json_emit.
SOFA_HD JPlan repr_plan(double x) {
    JPlan bad; bad.I = 0; bad.dg = 0; bad.neg = 0; bad.nd = 0; bad.ex = 0; bad.len = 0; bad.kind = JP_BAD;
    uint64_t xb = d_bits(x);
    bool neg = (xb >> 63) != 0;
    int be = (int)((xb >> 52) & 0x7ff);
    if (be == 0x7ff) return bad;
    uint64_t frac = xb & 0xfffffffffffffull;
    if (be == 0) return frac == 0 ? jp_fixed(neg, 0, 0, 1) : bad;          // "0.0" / "-0.0"; subnormals unsupported
    uint64_t m = frac | (1ull << 52);
    int e = be - 1075;
    if (e >= 0) {
        if (e > 1 || (m << e) >= 10000000000000000ull) return bad;         // repr switches to 1e+16 there
        return jp_fixed(neg, m << e, 0, 1);
    }
    int s = -e;
    if (s > 74) return bad;                       // |x| < 2^-22
    uint64_t I = s >= 64 ? 0 : (m >> s);
    uint64_t r0 = s >= 64 ? m : (m & ((1ull << s) - 1));
    if (I >= 10000000000000000ull) return bad;
    if (r0 == 0) return jp_fixed(neg, I, 0, 1);
    bool even = (m & 1) == 0;
    bool pow2 = (frac == 0);                      // lower neighbour is half as far
    u128 one = ((u128)1) << s;
    u128 p10 = 1;
    for (int j = 1; j <= 22; j++) {
        p10 *= 10;
        u128 F = (u128)r0 * p10;                  // < 2^53 * 10^22 < 2^127
        u128 fl = F >> s;
        u128 rem = F - (fl << s);
        u128 up = one - rem;
        u128 lo_lhs = pow2 ? rem * 4 : rem * 2;
        bool lo_ok = (lo_lhs < p10) || (even && lo_lhs == p10);
        u128 hi_lhs = up * 2;
        bool hi_ok = (hi_lhs < p10) || (even && hi_lhs == p10);
        if (!(lo_ok || hi_ok)) continue;
        u128 dg = fl;
        bool take_hi;
        if (lo_ok && hi_ok) take_hi = (up < rem) || (up == rem && ((uint64_t)dg & 1));
        else take_hi = hi_ok;
        if (take_hi) dg++;
        if (dg == p10) return jp_fixed(neg, I + 1, 0, 1);
        if (dg >> 63) return bad;
        uint64_t d = (uint64_t)dg;
        if (I > 0 || j <= 4 || d >= (uint64_t)(p10 / 10000)) {            // |x| >= 1e-4: fixed notation
            while (d % 10 == 0) { d /= 10; j--; }                          // (cannot happen for the shortest j; kept for safety)
            return jp_fixed(neg, I, d, j);
        }
        // 0 < |x| < 1e-4: "d.ddde-XX"
        int L = ndigits_u64(d);
        JPlan p; p.kind = JP_SCI; p.neg = neg ? 1 : 0; p.I = 0; p.dg = d; p.nd = (uint8_t)L;
        p.ex = (uint8_t)(j - L + 1);
        p.len = (uint8_t)(p.neg + L + (L > 1 ? 1 : 0) + 4);
        return p;
    }
    return bad;
}
template <typename O>
SOFA_HD int repr_emit(const JPlan& p, O out) {
    if (p.kind != JP_SCI) return json_emit(p, out);
    int n = 0;
    if (p.neg) out[n++] = '-';
    char ds[20];
    fmt_u64_pad(p.dg, p.nd, ds);
    out[n++] = ds[0];
    if (p.nd > 1) { out[n++] = '.'; for (int k = 1; k < p.nd; k++) out[n++] = ds[k]; }
    out[n++] = 'e'; out[n++] = '-';
    out[n++] = (char)('0' + p.ex / 10); out[n++] = (char)('0' + p.ex % 10);
    return n;
}

// ----------------------------------------------------------------------------- double-double
struct dd { double hi, lo; };
SOFA_HD dd two_sum(double a, double b) {
    double s = d_add(a, b);
    double bb = d_sub(s, a);
    double e = d_add(d_sub(a, d_sub(s, bb)), d_sub(b, bb));
    return dd{s, e};
}
SOFA_HD dd fast_two_sum(double a, double b) {      // |a| >= |b|
    double s = d_add(a, b);
    return dd{s, d_sub(b, d_sub(s, a))};
}
SOFA_HD dd two_prod(double a, double b) {
    double p = d_mul(a, b);
    return dd{p, d_fma(a, b, -p)};
}
SOFA_HD dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo = d_add(s.lo, t.hi);
    s = fast_two_sum(s.hi, s.lo);
    s.lo = d_add(s.lo, t.lo);
    return fast_two_sum(s.hi, s.lo);
}
SOFA_HD dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo = d_add(s.lo, a.lo);
    return fast_two_sum(s.hi, s.lo);
}
SOFA_HD dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo = d_add(p.lo, d_add(d_mul(a.hi, b.lo), d_mul(a.lo, b.hi)));
    return fast_two_sum(p.hi, p.lo);
}
SOFA_HD dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo = d_add(p.lo, d_mul(a.lo, b));
    return fast_two_sum(p.hi, p.lo);
}

// ln(1 + r) as a double-double, |r| <= 2^-7, absolute error < 2^-80.
SOFA_HD dd log1p_dd(double r) {
    dd r2 = two_prod(r, r);
    dd r3 = dd_mul_d(r2, r);
    dd t3 = dd_mul(r3, dd{SOFA_THIRD_HI, SOFA_THIRD_LO});
    // tail = r^4 * (-1/4 + r/5 - r^2/6 + ... + r^9/13)
    double t = 1.0 / 13.0;
    t = d_fma(t, r, -1.0 / 12.0);
    t = d_fma(t, r, 1.0 / 11.0);
    t = d_fma(t, r, -1.0 / 10.0);
    t = d_fma(t, r, 1.0 / 9.0);
    t = d_fma(t, r, -1.0 / 8.0);
    t = d_fma(t, r, 1.0 / 7.0);
    t = d_fma(t, r, -1.0 / 6.0);
    t = d_fma(t, r, 1.0 / 5.0);
    t = d_fma(t, r, -1.0 / 4.0);
    double tail = d_mul(d_mul(r2.hi, r2.hi), t);
    dd q = dd_add(dd{r, 0.0}, dd{-0.5 * r2.hi, -0.5 * r2.lo});
    q = dd_add(q, t3);
    return dd_add_d(q, tail);
}

// x = m * 2^E, m in [1,2); idx = top 7 fraction bits; r = m*inv[idx] - 1 exactly.
SOFA_HD void log_reduce(double x, int* E, int* idx, double* r) {
    uint64_t b = d_bits(x);
    int be = (int)((b >> 52) & 0x7ff);
    *E = be - 1023;
    *idx = (int)((b >> 45) & 127);
    double m = bits_d((b & 0xfffffffffffffull) | (1023ull << 52));
    *r = d_fma(m, SOFA_TAB(LOG_INV)[*idx], -1.0);
}

// log10(x) for normal positive x, rounded from a ~2^-80-accurate double-double.
SOFA_HD double log10_cr(double x) {
    int E, idx; double r;
    log_reduce(x, &E, &idx, &r);
    dd l = dd_mul(log1p_dd(r), dd{SOFA_INV_LN10_HI, SOFA_INV_LN10_LO});
    double fe = (double)E;
    dd acc = two_sum(d_mul(fe, SOFA_LOG10_2_HI), SOFA_TAB(L10_HI)[idx]);      // E*HI is exact
    dd e2 = two_prod(fe, SOFA_LOG10_2_MID);
    e2.lo = d_add(e2.lo, d_mul(fe, SOFA_LOG10_2_LO));
    acc = dd_add(acc, e2);
    acc = dd_add(acc, dd{SOFA_TAB(L10_LO)[idx], 0.0});
    acc = dd_add(acc, l);
    return d_add(acc.hi, acc.lo);
}
SOFA_HD double log2_cr(double x) {
    int E, idx; double r;
    log_reduce(x, &E, &idx, &r);
    dd l = dd_mul(log1p_dd(r), dd{SOFA_INV_LN2_HI, SOFA_INV_LN2_LO});
    dd acc = two_sum((double)E, SOFA_TAB(L2_HI)[idx]);
    acc = dd_add(acc, dd{SOFA_TAB(L2_LO)[idx], 0.0});
    acc = dd_add(acc, l);
    return d_add(acc.hi, acc.lo);
}

// ----------------------------------------------------------------------------- np.interp (scalar x)
// numpy/_core/src/multiarray/compiled_base.c arr_interp, scalar-x path (slopes not cached).
SOFA_HD double interp1(double x, const double* xp, const double* fp, int n) {
    if (x > xp[n - 1]) return fp[n - 1];
    if (x < xp[0]) return fp[0];
    int lo = 0, hi = n;               // largest j with xp[j] <= x
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (xp[mid] <= x) lo = mid; else hi = mid; }
    int j = lo;
    if (j == n - 1) return fp[j];
    if (xp[j] == x) return fp[j];
    double slope = d_div(d_sub(fp[j + 1], fp[j]), d_sub(xp[j + 1], xp[j]));
    double res = d_add(d_mul(slope, d_sub(x, xp[j])), fp[j]);
    if (res != res) {
        res = d_add(d_mul(slope, d_sub(x, xp[j + 1])), fp[j + 1]);
        if (res != res && fp[j] == fp[j + 1]) res = fp[j];
    }
    return res;
}

// event_raw = 1.0 * int("0x01" + hex, 16): (16^h + ip) rounded to double (sofa_preprocess.py:118).
SOFA_HD double ip_event_raw(uint64_t ip, int hexdigits) {
    u128 v = (((u128)1) << (4 * hexdigits)) | (u128)ip;
    return round_u128(v, 0, false);
}

// ----------------------------------------------------------------------------- values as read back from a trace CSV
// sofa_analyze works on pd.read_csv(<src>trace.csv): every float column has been through
// "%.6f" and a correctly rounded strtod (sofa_analyze.py:898-977).
SOFA_HD double round6_value(double x) {
    bool neg; uint64_t I; uint32_t F;
    if (!fixed6_parts(x, &neg, &I, &F)) return x;          // |x| >= 2^63: caller's emit fails loudly first
    double r;
    if (I < 9000000000ull) r = d_div((double)(I * 1000000ull + F), 1e6);
    else {
        u128 N = (u128)I * 1000000u + F;
        if ((N >> 64) != 0) return x;
        dec_to_double((uint64_t)N, -6, false, &r);
    }
    return neg ? -r : r;
}

// ----------------------------------------------------------------------------- numpy pairwise summation
// np.add.reduce over a contiguous float64 array: 0.0 + pairwise_sum (numpy/_core/src/umath/
// loops_utils.h DOUBLE_pairwise_sum): blocks of <= 128 with 8 accumulators, halves rounded
// down to a multiple of 8.  Series.sum(), get_group(k).sum() and aggregate(np.sum) all end here.
template <typename A>
SOFA_HD double pairwise_leaf(A a, uint64_t n) {
    if (n < 8) {
        double r = 0.0;
        for (uint64_t i = 0; i < n; i++) r = d_add(r, a[i]);
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    uint64_t i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] = d_add(r[j], a[i + j]);
    double res = d_add(d_add(d_add(r[0], r[1]), d_add(r[2], r[3])), d_add(d_add(r[4], r[5]), d_add(r[6], r[7])));
    for (; i < n; i++) res = d_add(res, a[i]);
    return res;
}
SOFA_HD uint64_t pairwise_split(uint64_t n) { uint64_t n2 = n / 2; return n2 - (n2 % 8); }

template <typename A>
SOFA_HD double pairwise_sum(A a, uint64_t n) {
    struct Fr { uint64_t s, n; int st; double left; };
    Fr stk[64];
    int sp = 0;
    stk[0].s = 0; stk[0].n = n; stk[0].st = 0; stk[0].left = 0;
    double ret = 0.0;
    while (sp >= 0) {
        Fr& f = stk[sp];
        if (f.n <= 128) { ret = pairwise_leaf(a + f.s, f.n); sp--; continue; }
        uint64_t n2 = pairwise_split(f.n);
        if (f.st == 0) { f.st = 1; stk[sp + 1].s = f.s; stk[sp + 1].n = n2; stk[sp + 1].st = 0; sp++; continue; }
        if (f.st == 1) { f.left = ret; f.st = 2; stk[sp + 1].s = f.s + n2; stk[sp + 1].n = f.n - n2; stk[sp + 1].st = 0; sp++; continue; }
        ret = d_add(f.left, ret);
        sp--;
    }
    return d_add(0.0, ret);
}
// sub-tree `t` (0 <= t < 2^depth) of the pairwise tree over n elements; every node above
// `depth` must be longer than 128 (callers pick depth with n >> depth >= 1024).
SOFA_HD void pairwise_subtree(uint64_t n, int depth, uint32_t t, uint64_t* start, uint64_t* len) {
    uint64_t s = 0, l = n;
    for (int lev = depth - 1; lev >= 0; lev--) {
        uint64_t n2 = pairwise_split(l);
        if ((t >> lev) & 1) { s += n2; l -= n2; } else l = n2;
    }
    *start = s; *len = l;
}
}  // namespace sofa
