// Per-record logic of the SOFA preprocess hot path as __host__ __device__ functions:
// the kernels in kernels.cu are thin tiling/scan wrappers around these, and
// tests/hostsim.cpp compiles the same code for the CPU unit tests.
//
//   parse_perf  == cpu_trace_read   sofa_preprocess.py:110-154
//   parse_gpu   == gpu_trace_read   sofa_preprocess.py:249-341
//   parse_net   == net_trace_read   sofa_preprocess.py:156-201
//   csv_row*    == DataFrame.to_csv(index=False, float_format='%.6f') rows  (:1206,1439,1802)
//   json_row*   == to_json(orient='records') -> json.loads -> json.dump records (:343-374)
#pragma once
#include "num.cuh"

namespace sofa {

enum { SRC_PERF = 0, SRC_GPU = 1, SRC_NET = 2, SRC_SIDE = 3,         // SIDE: a 13-column trace CSV row
       SRC_RAW = 4, SRC_SIDEOUT = 5 };                               // RAW: a recorder side-file line; SIDEOUT: the rows derived from them

// row-level error codes (the reference raises for all of these; we fail loudly with the row)
enum {
    E_OK = 0,
    E_FEW_TOKENS = 1,       // IndexError in the reference
    E_BAD_FLOAT = 2,        // float() ValueError
    E_BAD_INT = 3,          // int() ValueError
    E_UNSUPPORTED_NUM = 4,  // Python accepts the spelling, this codec does not
    E_BAD_HEX = 5,          // int("0x01"+ip, 16) ValueError
    E_LINE_TOO_LONG = 6,
    E_RANGE = 7,            // integer outside the column's storage range
    E_CR = 8,               // '\r' in the text (universal-newline handling not supported)
    E_TOO_MANY_LINES = 9,   // > capacity short lines in one tile (all would be E_FEW_TOKENS)
    E_FORMAT = 10,          // value outside the supported range of a text codec
    E_UTF8 = 11             // invalid UTF-8 in a name
};

struct IngestCfg {
    double time_base;        // cfg.time_base (sofa_preprocess.py:418-422)
    double t_offset;         // perf: unix-uptime (:1796); gpu: t_glb_gpu_base - t_base (:1430); net: 0
    int absolute_timestamp;
    // perf
    int perf_has_cycles;     // cfg.perf_events.find('cycles') != -1 (:138)
    int n_mhz;               // len(cpu_mhz_xp) (:131)
    const double* mhz_xp;
    const double* mhz_fp;
    // nvprof gpu trace
    double ts_rescale, dt_rescale, payload_unit;    // (:1392-1415)
    int n_cudaproc;                                   // pid column (:288)
    int idx_start, idx_dur, idx_size, idx_thr, idx_ctx, idx_stream, idx_name, idx_srcctx, idx_dstctx;
    int api_mode;            // nvprof --print-api-trace rows (cuda_api_trace_read, :203-247)
    int csv_mode;            // trace sources: the text is <src>trace.csv itself (stand-alone `sofa analyze`, :898-977)
    int side_table;          // SRC_SIDE: 0 mpstat.csv (name carries usr|sys|idl|iow|irq), 1 nvsmi_trace.csv, 2 netstat.csv;
                             // SRC_RAW: 0 mpstat.txt, 1 nvsmi_query.txt, 2 netstat.txt
};

struct Tok { int a, b; };

// SWAR byte-class tests on a 64-bit word (bit 7 of each lane set where the test holds)
SOFA_HD uint64_t lanes_eq(uint64_t w, unsigned char c) {       // lane == c  (c < 128)
    const uint64_t K7 = 0x7f7f7f7f7f7f7f7full, K8 = 0x8080808080808080ull;
    uint64_t t = (w & K7) ^ (0x0101010101010101ull * c);
    return ~((t + K7) | w) & K8;
}
SOFA_HD uint64_t lanes_first(int nb) { return nb >= 8 ? ~0ull : ((1ull << (8 * nb)) - 1ull); }

// Default word feeder: byte by byte.  Sinks with cheap whole-word paths define put_word.
template <class K> SOFA_HD void put_word_bytes(K& k, uint64_t w, int nb) {
    for (int b = 0; b < nb; b++) { k.put((char)(w & 0xff)); w >>= 8; }
}

// Feed the bytes p[0..len) to a sink.  On the device, global-memory spans are read as aligned
// 64-bit words (8x fewer load instructions than byte loads; the text buffer is padded, so the
// over-read of the last word is safe) and handed to the sink a word at a time (put_word).
// `wide` is false for shared-memory / local spans.
template <class K>
SOFA_HD void put_span(K& k, const char* p, int len, bool wide) {
#if defined(__CUDA_ARCH__) || defined(SOFA_HOST_WIDE)
#if !defined(__CUDA_ARCH__)
#define __ldg(ptr) (*(ptr))
#endif
    if (wide && len > 0) {
        unsigned long long a = (unsigned long long)p;
        const unsigned long long* w = reinterpret_cast<const unsigned long long*>(a & ~7ull);
        int sh = (int)(a & 7);
        unsigned long long cur = __ldg(w++) >> (8 * sh);
        int avail = 8 - sh;
        if constexpr (K::WORDS) {
            while (true) {
                int take = avail < len ? avail : len;
                len -= take;
                unsigned long long nxt = len > 0 ? __ldg(w++) : 0ull;     // issue the next load early
                k.put_word(cur, take);
                if (len <= 0) break;
                cur = nxt; avail = 8;
            }
        } else {
            for (int i = 0; i < len; i++) {
                if (avail == 0) { cur = __ldg(w++); avail = 8; }
                k.put((char)(cur & 0xff));
                cur >>= 8; avail--;
            }
        }
        return;
    }
#if !defined(__CUDA_ARCH__)
#undef __ldg
#endif
#endif
    (void)wide;
    for (int i = 0; i < len; i++) k.put(p[i]);
}

// ---------------------------------------------------------------------------- SWAR line scanning
// Lines are scanned 8 bytes at a time: aligned 64-bit loads (shared memory on the device),
// byte classes evaluated on all 8 lanes with carry-free 7-bit arithmetic, the 8 flag bits
// gathered with one multiply.  Word j of a line covers positions [8j - sh, 8j - sh + 8), where
// sh is the misalignment of the line's first byte; positions outside [0, n) are masked off.
SOFA_HD uint64_t line_word(const char* s, int n, int sh, int j) {
#if defined(__CUDA_ARCH__)
    (void)n;
    return *reinterpret_cast<const uint64_t*>(s - sh + 8 * j);
#else
    uint64_t w = 0;
    for (int b = 0; b < 8; b++) { int p = 8 * j - sh + b; if (p >= 0 && p < n) w |= (uint64_t)(unsigned char)s[p] << (8 * b); }
    return w;
#endif
}
SOFA_HD int line_misalign(const char* s) {
#if defined(__CUDA_ARCH__)
    return (int)((unsigned long long)s & 7ull);
#else
    return (int)((unsigned long long)s & 7ull);        // exercised by the CPU fuzz tests as well
#endif
}
SOFA_HD uint32_t valid8(int n, int sh, int j) {           // bit b set iff 0 <= 8j - sh + b < n
    int lo = sh - 8 * j; if (lo < 0) lo = 0;
    int hi = n + sh - 8 * j; if (hi > 8) hi = 8;
    if (hi <= lo) return 0;
    return ((1u << hi) - 1u) & ~((1u << lo) - 1u);
}
SOFA_HD uint32_t gather8(uint64_t hb) {                   // bit 7 of byte b -> bit b
    uint64_t x = (hb >> 7) & 0x0101010101010101ull;
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);          // two 32-bit multiplies (1 instruction each)
    return ((lo * 0x01020408u) >> 24) | (((hi * 0x01020408u) >> 24) << 4);
}
SOFA_HD uint32_t eq8(uint64_t w, unsigned char c) {       // lanes equal to c (c < 128)
    const uint64_t K7 = 0x7f7f7f7f7f7f7f7full, K8 = 0x8080808080808080ull;
    uint64_t t = (w & K7) ^ (0x0101010101010101ull * c);
    uint64_t ne = ((t + K7) | w) & K8;                    // low 7 bits differ, or byte >= 128
    return gather8(~ne & K8);
}
SOFA_HD uint32_t nonspace8(uint64_t w) {                  // lanes that are NOT str.split() whitespace
    const uint64_t K7 = 0x7f7f7f7f7f7f7f7full, K8 = 0x8080808080808080ull;
    uint64_t L = w & K7;
    uint64_t eq20 = ~(((L ^ 0x2020202020202020ull) + K7)) & K8;
    uint64_t r1 = (L + 0x7777777777777777ull) & ~(L + 0x7272727272727272ull) & K8;     // 9..13
    uint64_t r2 = (L + 0x6464646464646464ull) & ~(L + 0x6060606060606060ull) & K8;     // 28..31
    uint64_t sp = (eq20 | r1 | r2) & ~w & K8;             // bytes >= 128 are never whitespace
    return gather8(~sp & K8);
}
SOFA_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

SOFA_HD int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
SOFA_HD int popc_lanes(uint64_t m) {                       // number of lanes flagged (bit 7 set)
#if defined(__CUDA_ARCH__)
    return __popcll(m & 0x8080808080808080ull);
#else
    return __builtin_popcountll(m & 0x8080808080808080ull);
#endif
}

// str.split() with no argument (ASCII whitespace); returns number of tokens, capped at maxt.
template <typename S>
SOFA_HD int tokenize(S s, int n, Tok* t, int maxt) {
    int k = 0;
    const int sh = line_misalign(s);
    const int nw = (n + sh + 7) >> 3;
    uint32_t in = 0;                                      // class of the previous byte (1 = token)
    bool open = false;
    int start = 0;
    for (int j = 0; j < nw; j++) {
        uint32_t m = nonspace8(line_word(s, n, sh, j)) & valid8(n, sh, j);
        uint32_t tr = (m ^ ((m << 1) | in)) & 0xffu;      // byte positions where the class changes
        in = (m >> 7) & 1u;
        while (tr) {
            int p = 8 * j - sh + ctz32(tr);
            tr &= tr - 1;
            if (!open) { start = p; open = true; }
            else { t[k].a = start; t[k].b = p; k++; open = false; if (k >= maxt) return k; }
        }
    }
    if (open && k < maxt) { t[k].a = start; t[k].b = n; k++; }
    return k;
}

// first position in [from, to) holding byte c (c < 128), or -1; 8 bytes per step
SOFA_HD int find_char(const char* s, int n, int from, int to, unsigned char c) {
    if (from >= to) return -1;
    if (to - from <= 10) {                         // short token: a byte loop beats the word setup
        for (int i = from; i < to; i++) if ((unsigned char)s[i] == c) return i;
        return -1;
    }
    const int sh = line_misalign(s);
    int j0 = (from + sh) >> 3, j1 = (to - 1 + sh) >> 3;
    for (int j = j0; j <= j1; j++) {
        uint32_t m = eq8(line_word(s, n, sh, j), c) & valid8(to, sh, j);
        int base = 8 * j - sh;
        if (base < from) m &= ~((1u << (from - base)) - 1u);
        if (m) return base + ctz32(m);
    }
    return -1;
}

SOFA_HD int num_status_to_err(int st, bool is_float) {
    if (st == P_OK) return E_OK;
    if (st == P_UNSUPPORTED) return E_UNSUPPORTED_NUM;
    return is_float ? E_BAD_FLOAT : E_BAD_INT;
}

// ============================================================================ perf
struct PerfRow {
    double ts, event, dur;
    int32_t pid, tid;
    uint16_t evt_off, evt_len, sym_off, sym_len, dso_off, dso_len;   // relative to the line start
    // viz name (:1811-1816) = sym[0:va_len) + dso[0:vb_len) in the common case where the event
    // token holds neither ']' nor '+'; va_len == 0xffff selects the generic slicer.
    uint16_t va_len, vb_len;
};

template <typename S>
SOFA_HD int parse_perf(S s, int n, const IngestCfg& c, PerfRow* r) {
    if (n > 65535) return E_LINE_TOO_LONG;
    Tok t[8];
    int nt = tokenize(s, n, t, 8);
    if (nt < 2) return E_FEW_TOKENS;
    int sh = 0;
    {   // re.match('\[\d+\]', fields[1])  (:114)
        int a = t[1].a, b = t[1].b;
        if (b - a >= 3 && s[a] == '[') {
            int i = a + 1;
            while (i < b && (unsigned)((unsigned char)s[i] - '0') <= 9) i++;
            if (i > a + 1 && i < b && s[i] == ']') sh = 1;
        }
    }
    if (nt < 7 + sh) return E_FEW_TOKENS;
    double time, counts;
    {
        int a = t[1 + sh].a, b = t[1 + sh].b;
        int e = find_char(s, n, a, b, ':');                  // .split(':')[0]
        if (e < 0) e = b;
        int st = parse_double(s + a, e - a, &time);
        if (st) return num_status_to_err(st, true);
        st = parse_double(s + t[2 + sh].a, t[2 + sh].b - t[2 + sh].a, &counts);
        if (st) return num_status_to_err(st, true);
    }
    uint64_t ip = 0;
    int hd = t[4 + sh].b - t[4 + sh].a;
    {
        int a = t[4 + sh].a;
        if (hd > 16) return E_UNSUPPORTED_NUM;
        for (int i = 0; i < hd; i++) {
            unsigned char ch = (unsigned char)s[a + i];
            unsigned v;
            if (ch >= '0' && ch <= '9') v = ch - '0';
            else if (ch >= 'a' && ch <= 'f') v = ch - 'a' + 10;
            else if (ch >= 'A' && ch <= 'F') v = ch - 'A' + 10;
            else return ch == '_' ? E_UNSUPPORTED_NUM : E_BAD_HEX;
            ip = (ip << 4) | v;
        }
    }
    double event_raw = ip_event_raw(ip, hd);
    if (!c.absolute_timestamp) time = d_sub(time, c.time_base);
    double t_begin = d_add(time, c.t_offset);
    double dur;
    if (c.n_mhz > 1) dur = d_div(counts, d_mul(interp1(t_begin, c.mhz_xp, c.mhz_fp, c.n_mhz), 1e6));
    else dur = d_div(counts, 3000.0 * 1e6);
    if (!c.perf_has_cycles) dur = log2_cr(d_div(event_raw, 1e14));
    r->ts = t_begin;
    r->event = log10_cr(event_raw);
    r->dur = dur;
    {   // int(fields[0].split('/')[0]), int(fields[0].split('/')[1])
        int a = t[0].a, b = t[0].b;
        int s1 = find_char(s, n, a, b, '/');
        if (s1 < 0) return E_FEW_TOKENS;
        int s2 = find_char(s, n, s1 + 1, b, '/');
        if (s2 < 0) s2 = b;
        int64_t p, q;
        int st = parse_i64(s + a, s1 - a, &p);
        if (st) return num_status_to_err(st, false);
        st = parse_i64(s + s1 + 1, s2 - s1 - 1, &q);
        if (st) return num_status_to_err(st, false);
        if (p < -2147483647 || p > 2147483647 || q < -2147483647 || q > 2147483647) return E_RANGE;
        r->pid = (int32_t)p; r->tid = (int32_t)q;
    }
    r->evt_off = (uint16_t)t[3 + sh].a; r->evt_len = (uint16_t)(t[3 + sh].b - t[3 + sh].a);
    r->sym_off = (uint16_t)t[5 + sh].a; r->sym_len = (uint16_t)(t[5 + sh].b - t[5 + sh].a);
    r->dso_off = (uint16_t)t[6 + sh].a; r->dso_len = (uint16_t)(t[6 + sh].b - t[6 + sh].a);
    {   // precompute the viz slice: name = '[' evt ']' sym dso ; [find(']')+1 : find('+')] cut at '@'
        int eo = r->evt_off, el = r->evt_len;
        bool exotic = find_char(s, n, eo, eo + el, ']') >= 0 || find_char(s, n, eo, eo + el, '+') >= 0;
        if (exotic) { r->va_len = 0xffff; r->vb_len = 0; }
        else {
            int so = r->sym_off, sl = r->sym_len, dof = r->dso_off, dl = r->dso_len;
            int plus = find_char(s, n, so, so + sl, '+');
            if (plus >= 0) plus -= so;
            else { plus = find_char(s, n, dof, dof + dl, '+'); if (plus >= 0) plus = sl + (plus - dof); }
            int b = plus >= 0 ? plus : sl + dl - 1;           // [a:-1] when there is no '+'
            if (b < 0) b = 0;
            int va = b < sl ? b : sl, vb = b > sl ? b - sl : 0;
            int at = find_char(s, n, so, so + va, '@');
            if (at >= 0) { va = at - so; vb = 0; }
            else { at = find_char(s, n, dof, dof + vb, '@'); if (at >= 0) vb = at - dof; }
            r->va_len = (uint16_t)va; r->vb_len = (uint16_t)vb;
        }
    }
    return E_OK;
}

// evt.replace('-', '_') on the fly
template <class K>
struct DashFilter {
    static constexpr bool WORDS = K::WORDS;
    K* k;
    SOFA_HD void put(char ch) { k->put(ch == '-' ? '_' : ch); }
    SOFA_HD void put_word(uint64_t w, int nb) { k->put_word(w ^ ((lanes_eq(w, '-') >> 7) * 0x72ull), nb); }      // '-' ^ '_' = 0x72
};

// A name taken verbatim from the name field of a trace CSV (csv_mode): the field without its enclosing quotes; inside,
// a doubled quote is one quote (csv QUOTE_MINIMAL) and 0x1f stands for a newline that the pre-pass took out of the
// quoted field so that rows are lines (k_csv_mask_newlines).
template <class K>
struct CsvLiteralFilter {
    static constexpr bool WORDS = false;
    K* k; bool q;
    SOFA_HD void put(char ch) {
        if (ch == '"') { if (q) { k->put('"'); q = false; } else q = true; return; }
        q = false;
        k->put(ch == 0x1f ? '\n' : ch);
    }
    SOFA_HD void put_word(uint64_t w, int nb) { put_word_bytes(*this, w, nb); }
};
constexpr uint16_t PERF_LITERAL = 0xffff;       // PerfRow.evt_len of a CSV-ingested row: name = the sym span, verbatim

// name = '[%s]' % evt.replace('-','_') + sym + dso   (:116,121)
template <typename S>
struct PerfName {
    S line; PerfRow r; bool wide;
    template <class K> SOFA_HD void gen(K& k) const {
        if (r.evt_len == PERF_LITERAL) {
            CsvLiteralFilter<K> lf{&k, false};
            put_span(lf, line + r.sym_off, r.sym_len, wide);
            return;
        }
        k.put('[');
        { DashFilter<K> df{&k}; put_span(df, line + r.evt_off, r.evt_len, wide); }
        k.put(']');
        put_span(k, line + r.sym_off, r.sym_len, wide);
        put_span(k, line + r.dso_off, r.dso_len, wide);
    }
};

// viz name = name[name.find(']')+1 : name.find('+')].split('@')[0]  (:1811-1816), before demangling
struct FindSink {
    static constexpr bool WORDS = false;
    int pos, first_rb, first_plus;
    SOFA_HD FindSink() : pos(0), first_rb(-1), first_plus(-1) {}
    SOFA_HD void put(char ch) {
        if (ch == ']' && first_rb < 0) first_rb = pos;
        if (ch == '+' && first_plus < 0) first_plus = pos;
        pos++;
    }
    SOFA_HD void put_word(uint64_t w, int nb) { put_word_bytes(*this, w, nb); }
};
template <class K>
struct SliceSink {          // forwards characters [a, b) until the first '@'
    static constexpr bool WORDS = false;
    K* k; int pos, a, b; bool stop;
    SOFA_HD void put(char ch) {
        if (pos >= a && pos < b && !stop) { if (ch == '@') stop = true; else k->put(ch); }
        pos++;
    }
    SOFA_HD void put_word(uint64_t w, int nb) { put_word_bytes(*this, w, nb); }
};
template <typename S>
struct PerfVizName {
    PerfName<S> full;
    template <class K> SOFA_HD void gen(K& k) const {
        if (full.r.va_len != 0xffff) {
            put_span(k, full.line + full.r.sym_off, full.r.va_len, full.wide);
            put_span(k, full.line + full.r.dso_off, full.r.vb_len, full.wide);
            return;
        }
        FindSink f; full.gen(f);
        int len = f.pos;
        int a = f.first_rb + 1;                         // find()==-1 -> 0
        int b = f.first_plus >= 0 ? f.first_plus : len - 1;   // [a:-1]
        if (b < 0) b = 0;
        SliceSink<K> sl{&k, 0, a, b, false};
        full.gen(sl);
    }
};

// ============================================================================ nvprof gpu trace
struct GpuRow {
    double ts, dur, bw;
    int64_t payload;
    int32_t dev, tid, src, dst;
    uint8_t kind, bw_is_float;
    uint16_t name_off, name_len;      // raw span of the Name field (quotes still inside)
};

// parse a numeric field whose '"' characters have been removed (record.replace('"',''))
template <typename S>
SOFA_HD int parse_double_q(S s, int a, int b, double* out) {
    {   // common case: the field is a bare or a fully quoted number - parse it in place (a quote anywhere
        // else makes parse_double fail, and the buffer path below then decides)
        int a2 = a, b2 = b;
        if (a2 < b2 && s[a2] == '"') a2++;
        if (b2 > a2 && s[b2 - 1] == '"') b2--;
        if (parse_double(s + a2, b2 - a2, out) == P_OK) return P_OK;
    }
    char buf[48];
    int m = 0;
    for (int i = a; i < b; i++) { char ch = s[i]; if (ch != '"') { if (m >= 47) return P_UNSUPPORTED; buf[m++] = ch; } }
    return parse_double((const char*)buf, m, out);
}
template <typename S>
SOFA_HD int parse_i64_q(S s, int a, int b, int64_t* out) {
    char buf[32];
    int m = 0;
    for (int i = a; i < b; i++) { char ch = s[i]; if (ch != '"') { if (m >= 31) return P_UNSUPPORTED; buf[m++] = ch; } }
    return parse_i64((const char*)buf, m, out);
}
// int(float(x)) inside try/except -> (ok, value)
SOFA_HD bool float_to_int_trunc(double v, int64_t* out, int* err) {
    uint64_t b = d_bits(v);
    if (((b >> 52) & 0x7ff) == 0x7ff) return false;            // inf/nan -> exception -> default
    if (v >= 9223372036854775808.0 || v <= -9223372036854775808.0) { *err = E_RANGE; return false; }
    *out = (int64_t)v;
    return true;
}

// `n` includes the trailing '\n' when the line has one (the Name field keeps it, :258).
template <typename S>
SOFA_HD int parse_gpu(S s, int n, const IngestCfg& c, GpuRow* r) {
    if (n > 65535) return E_LINE_TOO_LONG;
    // field spans for the wanted columns (split(',') after the quotes are dropped)
    int fa[9], fb[9];
    int want[9] = {c.idx_start, c.idx_dur, c.idx_size, c.idx_thr, c.idx_ctx, c.idx_stream,
                   c.idx_name, c.idx_srcctx, c.idx_dstctx};
    uint32_t wantmask = 0;
    for (int k = 0; k < 9; k++) { fa[k] = -1; fb[k] = -1; if (want[k] >= 0 && want[k] < 32) wantmask |= 1u << want[k]; }
    int f = 0, a = 0;
    {
        const int sh = line_misalign(s);
        const int nw = (n + sh + 7) >> 3;
        for (int j = 0; j < nw; j++) {
            uint32_t m = eq8(line_word(s, n, sh, j), ',') & valid8(n, sh, j);
            while (m) {
                int i = 8 * j - sh + ctz32(m);
                m &= m - 1;
                if (f >= 32 || ((wantmask >> f) & 1u)) {            // most columns are not wanted: skip the slot search
                    for (int k = 0; k < 9; k++) if (want[k] == f) { fa[k] = a; fb[k] = i; }
                }
                f++; a = i + 1;
            }
        }
        for (int k = 0; k < 9; k++) if (want[k] == f) { fa[k] = a; fb[k] = n; }
    }
    if (fa[6] < 0 || fa[0] < 0 || fa[1] < 0) return E_FEW_TOKENS;
    double start, dur;
    int st = parse_double_q(s, fa[0], fb[0], &start);
    if (st) return num_status_to_err(st, true);
    double time = d_add(d_div(start, c.ts_rescale), c.t_offset);
    if (!c.absolute_timestamp) time = d_sub(time, c.time_base);
    st = parse_double_q(s, fa[1], fb[1], &dur);
    if (st) return num_status_to_err(st, true);
    r->ts = time;
    r->dur = d_div(dur, c.dt_rescale);
    if (c.api_mode) {        // cuda_api_trace_read: constants besides time, duration and the name
        r->payload = 0; r->bw = 0.0; r->bw_is_float = 0; r->dev = -1; r->tid = -1; r->src = 0; r->dst = 0; r->kind = 0;
        r->name_off = (uint16_t)fa[6]; r->name_len = (uint16_t)(fb[6] - fa[6]);
        return E_OK;
    }
    int err = E_OK;
    double v; int64_t iv;
    r->payload = 0;
    if (fa[2] >= 0) {
        st = parse_double_q(s, fa[2], fb[2], &v);
        if (st == P_UNSUPPORTED) return E_UNSUPPORTED_NUM;
        if (st == P_OK && float_to_int_trunc(d_mul(v, c.payload_unit), &iv, &err)) r->payload = iv;
    }
    r->bw = 0.0; r->bw_is_float = 0;
    if (fa[3] >= 0) {
        st = parse_double_q(s, fa[3], fb[3], &v);
        if (st == P_UNSUPPORTED) return E_UNSUPPORTED_NUM;
        if (st == P_OK) { r->bw = v; r->bw_is_float = 1; }
    }
    int64_t dev = -1, tid = -1;
    if (fa[4] >= 0) {
        st = parse_double_q(s, fa[4], fb[4], &v);
        if (st == P_UNSUPPORTED) return E_UNSUPPORTED_NUM;
        if (st == P_OK && float_to_int_trunc(v, &iv, &err)) dev = iv;
    }
    if (fa[5] >= 0) {
        st = parse_double_q(s, fa[5], fb[5], &v);
        if (st == P_UNSUPPORTED) return E_UNSUPPORTED_NUM;
        if (st == P_OK && float_to_int_trunc(v, &iv, &err)) tid = iv;
    }
    if (err) return err;
    if (dev < -2147483647 || dev > 2147483647 || tid < -2147483647 || tid > 2147483647) return E_RANGE;
    // substring tags on the quote-free name (:293-323).  Word pass: every tag is [HDP] 't' 'o' [HDP], so only
    // the positions where a 't' is directly followed by an 'o' are looked at (one or two per name at most);
    // a quote anywhere but at the ends of the field falls back to the byte loop over the quote-free text.
    bool h2d = false, d2h = false, d2d = false, p2p = false;
    {
        const int na = fa[6], nb_ = fb[6];
        if (nb_ > na) {
            const uint64_t K8 = 0x8080808080808080ull;
            const int sh = line_misalign(s);
            const int j0 = (na + sh) >> 3, j1 = (nb_ - 1 + sh) >> 3;
            uint64_t prev_t = 0;
            int quotes = 0;
            for (int j = j0; j <= j1; j++) {
                const uint64_t w = line_word(s, n, sh, j);
                const int base = 8 * j - sh;
                const int lo = na > base ? na - base : 0, hi = nb_ - base < 8 ? nb_ - base : 8;
                const uint64_t vm = (hi >= 8 ? ~0ull : ((1ull << (8 * hi)) - 1ull)) & ~((1ull << (8 * lo)) - 1ull) & K8;
                const uint64_t t = lanes_eq(w, 't') & vm, o = lanes_eq(w, 'o') & vm;
                uint64_t hit = ((t << 8) & o) | (prev_t & o);          // flagged lane = the 'o'
                prev_t = t >> 56;
                quotes += popc_lanes(lanes_eq(w, '"') & vm);
                while (hit) {
                    const int po = base + (ctz64(hit) >> 3);           // position of the 'o'
                    hit &= hit - 1;
                    if (po - 2 >= na && po + 1 < nb_) {
                        const char x = s[po - 2], y = s[po + 1];
                        h2d |= (x == 'H' && y == 'D'); d2h |= (x == 'D' && y == 'H');
                        d2d |= (x == 'D' && y == 'D'); p2p |= (x == 'P' && y == 'P');
                    }
                }
            }
            quotes -= (s[na] == '"') ? 1 : 0;
            if (nb_ - 1 > na) quotes -= (s[nb_ - 1] == '"') ? 1 : 0;
            if (nb_ - 2 > na) quotes -= (s[nb_ - 2] == '"') ? 1 : 0;
            if (quotes != 0) {                                         // quotes inside the name: exact scan of the quote-free text
                h2d = d2h = d2d = p2p = false;
                uint32_t w = 0;
                for (int i = na; i < nb_; i++) {
                    char ch = s[i];
                    if (ch == '"') continue;
                    w = (w << 8) | (unsigned char)ch;
                    h2d |= (w == 0x48746f44u);   // "HtoD"
                    d2h |= (w == 0x44746f48u);   // "DtoH"
                    d2d |= (w == 0x44746f44u);   // "DtoD"
                    p2p |= (w == 0x50746f50u);   // "PtoP"
                }
            }
        }
    }
    int64_t src = 0, dst = 0;
    int kind = 0;
    if (h2d) { kind = 1; src = 0; dst = dev; }
    else if (d2h) { kind = 2; src = dev; dst = 0; }
    else if (d2d) { kind = 8; src = dev; dst = dev; }
    else if (p2p) {
        kind = 10;
        if (fa[7] >= 0) { st = parse_i64_q(s, fa[7], fb[7], &iv); if (st == P_UNSUPPORTED) return E_UNSUPPORTED_NUM; if (st == P_OK) src = iv; }
        if (fa[8] >= 0) { st = parse_i64_q(s, fa[8], fb[8], &iv); if (st == P_UNSUPPORTED) return E_UNSUPPORTED_NUM; if (st == P_OK) dst = iv; }
        if (src < -2147483647 || src > 2147483647 || dst < -2147483647 || dst > 2147483647) return E_RANGE;
    }
    r->dev = (int32_t)dev; r->tid = (int32_t)tid; r->src = (int32_t)src; r->dst = (int32_t)dst;
    r->kind = (uint8_t)kind;
    r->name_off = (uint16_t)fa[6]; r->name_len = (uint16_t)(fb[6] - fa[6]);
    return E_OK;
}

template <class K> SOFA_HD void put_cstr(K& k, const char* lit) { for (int i = 0; lit[i]; i++) k.put(lit[i]); }
// string literal: sinks with word paths get it as compile-time constant 8-byte words (a per-byte feed
// costs a class test per character and, in a warp whose lanes print different kinds of names, runs with
// a handful of lanes active)
template <class K, int N> SOFA_HD void put_lit(K& k, const char (&lit)[N]) {
    if constexpr (K::WORDS) {
#pragma unroll
        for (int o = 0; o < N - 1; o += 8) {
            uint64_t w = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) if (o + b < N - 1) w |= (uint64_t)(unsigned char)lit[o + b] << (8 * b);
            k.put_word(w, (N - 1 - o) < 8 ? (N - 1 - o) : 8);
        }
    } else {
        for (int i = 0; i < N - 1; i++) k.put(lit[i]);
    }
}

// decimal digits of v straight into a sink: digits are produced least-significant first into
// byte stacks held in registers (no local-memory buffer), then popped most-significant first
template <class K> SOFA_HD void put_i64(K& k, int64_t v) {
    uint64_t u = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    if (v < 0) k.put('-');
    if ((u >> 32) == 0) {
        uint32_t x = (uint32_t)u;
        if (x < 10u) { k.put((char)('0' + x)); return; }
        uint64_t r0 = 0, r1 = 0; int n = 0;
        do { uint32_t q = x / 10u; uint64_t d = '0' + (x - q * 10u); if (n < 8) r0 = (r0 << 8) | d; else r1 = (r1 << 8) | d; x = q; n++; } while (x);
        if (n > 8) k.put_word(r1, n - 8);               // the stacks already hold the digits in output order
        k.put_word(r0, n < 8 ? n : 8);
        return;
    }
    uint64_t r0 = 0, r1 = 0, r2 = 0; int n = 0;
    do {
        uint64_t q = u / 10u; uint64_t d = '0' + (u - q * 10u);
        if (n < 8) r0 = (r0 << 8) | d; else if (n < 16) r1 = (r1 << 8) | d; else r2 = (r2 << 8) | d;
        u = q; n++;
    } while (u);
    if (n > 16) k.put_word(r2, n - 16);
    if (n > 8) k.put_word(r1, n < 16 ? n - 8 : 8);
    k.put_word(r0, n < 8 ? n : 8);
}

template <class K>
struct QuoteFilter {
    static constexpr bool WORDS = K::WORDS;
    K* k;
    SOFA_HD void put(char ch) { if (ch != '"') k->put(ch); }
    SOFA_HD void put_word(uint64_t w, int nb) {
        uint64_t q = lanes_eq(w, '"') & lanes_first(nb);
        while (q) {                                     // squeeze the quote bytes out of the word (usually one, at an end)
            const int p = ctz64(q) >> 3;
            uint64_t lo = p ? (w & ((1ull << (8 * p)) - 1ull)) : 0ull;
            uint64_t hi = p < 7 ? ((w >> (8 * p + 8)) << (8 * p)) : 0ull;
            w = lo | hi; nb--;
            uint64_t ql = p ? (q & ((1ull << (8 * p)) - 1ull)) : 0ull;
            uint64_t qh = p < 7 ? ((q >> (8 * p + 8)) << (8 * p)) : 0ull;
            q = ql | qh;
        }
        if (nb > 0) k->put_word(w, nb);
    }
};
template <typename S>
struct GpuName {
    S line; GpuRow r; bool wide; bool api;
    template <class K> SOFA_HD void gen(K& k) const {
        if (r.bw_is_float & 2) {                                                       // CSV-ingested row: the name field, verbatim
            CsvLiteralFilter<K> lf{&k, false};
            put_span(lf, line + r.name_off, r.name_len, wide);
            return;
        }
        if (api) put_lit(k, "[CUDA_API]");                                             // :213
        if (r.dev != -1) { put_lit(k, "[gpu"); put_i64(k, r.dev); k.put(']'); }        // :325-326
        if (r.kind == 0) {
            QuoteFilter<K> q{&k};
            put_span(q, line + r.name_off, r.name_len, wide);
        } else if (r.kind == 10) {
            put_lit(k, "[CUDA_COPY_P2P]from_gpu"); put_i64(k, r.src); put_lit(k, "_to_gpu"); put_i64(k, r.dst);
            k.put('_'); put_i64(k, r.payload); k.put('B');
        } else {
            put_lit(k, "CUDA_COPY_");
            k.put_word(r.kind == 1 ? 0x5f443248ull : (r.kind == 2 ? 0x5f483244ull : 0x5f443244ull), 4);      // "H2D_" / "D2H_" / "D2D_"
            put_i64(k, r.payload); k.put('B');
        }
    }
};

// ============================================================================ tcpdump
struct NetRow { double ts; int64_t payload, src, dst; uint8_t proto; };   // proto 0 = [] row, 1 = UDP, 2 = tcp

template <typename S>
SOFA_HD int dotted4(S s, int a, int b, int64_t* out) {           // sum int(part_i) * 1000^(3-i)
    int64_t acc = 0;
    int p = a;
    for (int i = 0; i < 4; i++) {
        if (p > b) return E_FEW_TOKENS;
        int e = p;
        while (e < b && s[e] != '.') e++;
        int64_t v;
        int st = parse_i64(s + p, e - p, &v);
        if (st) return num_status_to_err(st, false);
        if (v < -1000000 || v > 1000000) return E_RANGE;
        acc = acc * 1000 + v;
        p = e + 1;           // next part starts after the dot; p > b means "no more parts"
        if (e >= b) p = b + 1;
    }
    *out = acc;
    return E_OK;
}

template <typename S>
SOFA_HD int parse_net(S s, int n, const IngestCfg& c, NetRow* r) {
    Tok t[8];
    int nt = tokenize(s, n, t, 8);
    if (nt < 1) return E_FEW_TOKENS;
    double time;
    int st = parse_double(s + t[0].a, t[0].b - t[0].a, &time);
    if (st) return num_status_to_err(st, true);
    if (!c.absolute_timestamp) time = d_sub(time, c.time_base);
    r->ts = d_add(time, c.t_offset);
    r->proto = 0; r->payload = 0; r->src = 0; r->dst = 0;
    if (nt < 2) return E_FEW_TOKENS;
    if (!(t[1].b - t[1].a == 2 && s[t[1].a] == 'I' && s[t[1].a + 1] == 'P')) return E_OK;     // [] row
    if (nt < 6) return E_FEW_TOKENS;
    int a = t[5].a, l = t[5].b - t[5].a;
    int ptok;
    if (l == 4 && s[a] == 'U' && s[a + 1] == 'D' && s[a + 2] == 'P' && s[a + 3] == ',') { r->proto = 1; ptok = 7; }
    else if (l == 3 && s[a] == 't' && s[a + 1] == 'c' && s[a + 2] == 'p') { r->proto = 2; ptok = 6; }
    else return E_OK;                                                                            // [] row
    if (nt <= ptok) return E_FEW_TOKENS;
    int64_t pl;
    st = parse_i64(s + t[ptok].a, t[ptok].b - t[ptok].a, &pl);
    if (st) return num_status_to_err(st, false);
    if (pl > 90000000000000000ll || pl < -90000000000000000ll) return E_RANGE;
    r->payload = pl;
    int e = dotted4(s, t[2].a, t[2].b, &r->src);
    if (e) return e;
    e = dotted4(s, t[4].a, t[4].b, &r->dst);
    if (e) return e;
    return E_OK;
}

struct NetName {
    NetRow r;
    template <class K> SOFA_HD void gen(K& k) const {     // "network:%s:%d_to_%d_with_%d" (:198)
        if (r.proto == 1) put_lit(k, "network:UDP,:"); else put_lit(k, "network:tcp:");
        put_i64(k, r.src); put_lit(k, "_to_"); put_i64(k, r.dst); put_lit(k, "_with_"); put_i64(k, r.payload);
    }
};

// ============================================================================ side-table CSV rows
// pd.read_csv row of mpstat.csv / nvsmi_trace.csv / netstat.csv as concurrency_breakdown uses it
// (sofa_analyze.py:75-243): timestamp, event, duration, deviceId, bandwidth and, for mpstat, the five
// integers of name.split(':')[1].split('|')[1:6] (:136-147).  Quoted fields are refused.
struct SideRow { double ts, event, dur, dev, bw; int32_t r[5]; };

template <typename S>
SOFA_HD int parse_side(S s, int n, const IngestCfg& c, SideRow* r) {
    if (n > 65535) return E_LINE_TOO_LONG;
    int fa[13], fb[13];
    int f = 0, a = 0;
    {
        const int sh = line_misalign(s);
        const int nw = (n + sh + 7) >> 3;
        for (int j = 0; j < nw; j++) {
            uint64_t w = line_word(s, n, sh, j);
            if (eq8(w, '"') & valid8(n, sh, j)) return E_UNSUPPORTED_NUM;
            uint32_t m = eq8(w, ',') & valid8(n, sh, j);
            while (m) {
                int i = 8 * j - sh + ctz32(m);
                m &= m - 1;
                if (f < 13) { fa[f] = a; fb[f] = i; }
                f++; a = i + 1;
            }
        }
        if (f < 13) { fa[f] = a; fb[f] = n; }
        f++;
    }
    if (f != 13) return E_FEW_TOKENS;
    double* dst[5] = {&r->ts, &r->event, &r->dur, &r->dev, &r->bw};
    const int col[5] = {0, 1, 2, 3, 6};
    for (int k = 0; k < 5; k++) {
        int st = parse_double(s + fa[col[k]], fb[col[k]] - fa[col[k]], dst[k]);
        if (st) return num_status_to_err(st, true);
    }
    for (int k = 0; k < 5; k++) r->r[k] = 0;
    if (c.side_table == 0) {
        int p = find_char(s, n, fa[11], fb[11], ':');                 // .split(':')[1]
        if (p < 0) return E_FEW_TOKENS;
        int e2 = find_char(s, n, p + 1, fb[11], ':');
        int end = e2 < 0 ? fb[11] : e2;
        int q = find_char(s, n, p + 1, end, '|');                     // .split('|')[0] is dropped
        for (int k = 0; k < 5; k++) {
            if (q < 0) return E_FEW_TOKENS;
            int q2 = find_char(s, n, q + 1, end, '|');
            int stop = q2 < 0 ? end : q2;
            int64_t v;
            int st = parse_i64(s + q + 1, stop - q - 1, &v);
            if (st) return num_status_to_err(st, false);
            if (v < -1000000 || v > 1000000) return E_RANGE;
            r->r[k] = (int32_t)v;
            q = q2;
        }
    } else if (c.side_table == 3) {
        // vmstat.csv (vmstat_profile, sofa_analyze.py:712-733): name = 'r=..|b=..|...|si=..|so=..|bi=..|bo=..|in=..|cs=..|...';
        // pieces 8..11 of name.split('|') with their first three characters dropped: bi, bo, in, cs
        int q = fa[11] - 1;                                            // "separator" in front of piece 0
        for (int piece = 0; piece < 12; piece++) {
            int q2 = find_char(s, n, q + 1, fb[11], '|');
            int stop = q2 < 0 ? fb[11] : q2;
            if (piece >= 8) {
                int a0 = q + 1 + 3;
                if (a0 > stop) return E_BAD_FLOAT;
                int64_t v;
                int st = parse_i64(s + a0, stop - a0, &v);
                if (st) return num_status_to_err(st, true);
                if (v < -2000000000 || v > 2000000000) return E_RANGE;
                r->r[piece - 8] = (int32_t)v;
            }
            if (q2 < 0 && piece < 11) return E_FEW_TOKENS;
            q = q2;
        }
    }
    return E_OK;
}

// ============================================================================ trace CSV rows (re-ingest)
// One row of cputrace.csv / gputrace.csv / nettrace.csv as pd.read_csv sees it (sofa_analyze.py:898-977 works on these
// files).  The first eleven fields are numeric, the last one (category) too; the name is everything between the 11th
// comma and the last one, quoted iff it holds ',' '"' or a newline (newlines inside quotes were replaced by 0x1f).
template <typename S>
SOFA_HD int csv_row_fields(S s, int n, int* fa, int* fb) {       // 13 spans; returns E_OK / E_FEW_TOKENS
    if (n > 65535) return E_LINE_TOO_LONG;
    int f = 0, a = 0;
    for (int i = 0; i < n && f < 11; i++) if (s[i] == ',') { fa[f] = a; fb[f] = i; f++; a = i + 1; }
    if (f < 11) return E_FEW_TOKENS;
    int last = -1;
    for (int i = n - 1; i >= a; i--) if (s[i] == ',') { last = i; break; }
    if (last < 0) return E_FEW_TOKENS;
    fa[11] = a; fb[11] = last; fa[12] = last + 1; fb[12] = n;
    if (fb[11] - fa[11] >= 2 && s[fa[11]] == '"' && s[fb[11] - 1] == '"') { fa[11]++; fb[11]--; }
    return E_OK;
}
template <typename S>
SOFA_HD int csv_num(S s, int a, int b, double* out, bool* is_float) {   // empty field = NaN (the [] packet rows)
    if (b <= a) { *out = bits_d(0x7ff8000000000000ull); *is_float = true; return E_OK; }
    *is_float = false;
    for (int i = a; i < b; i++) if (s[i] == '.' || s[i] == 'e' || s[i] == 'n' || s[i] == 'i') *is_float = true;
    int st = parse_double(s + a, b - a, out);
    return st ? num_status_to_err(st, true) : E_OK;
}
SOFA_HD bool csv_to_i64(double v, int64_t* out) {
    if (!(v >= -9.2e18 && v <= 9.2e18)) return false;
    *out = (int64_t)v;
    return (double)*out == v;
}
template <typename S>
SOFA_HD int parse_csv_perf(S s, int n, PerfRow* r) {
    int fa[13], fb[13];
    int e = csv_row_fields(s, n, fa, fb); if (e) return e;
    double v[11]; bool fl;
    const int need[5] = {0, 1, 2, 9, 10};
    for (int q = 0; q < 5; q++) { e = csv_num(s, fa[need[q]], fb[need[q]], &v[need[q]], &fl); if (e) return e; }
    int64_t pid, tid;
    if (!csv_to_i64(v[9], &pid) || !csv_to_i64(v[10], &tid) || pid < -2147483647 || pid > 2147483647 || tid < -2147483647 || tid > 2147483647)
        return E_RANGE;
    r->ts = v[0]; r->event = v[1]; r->dur = v[2]; r->pid = (int32_t)pid; r->tid = (int32_t)tid;
    r->evt_off = 0; r->evt_len = PERF_LITERAL;
    r->sym_off = (uint16_t)fa[11]; r->sym_len = (uint16_t)(fb[11] - fa[11]);
    r->dso_off = 0; r->dso_len = 0; r->va_len = PERF_LITERAL; r->vb_len = 0;
    return E_OK;
}
template <typename S>
SOFA_HD int parse_csv_gpu(S s, int n, GpuRow* r) {
    int fa[13], fb[13];
    int e = csv_row_fields(s, n, fa, fb); if (e) return e;
    double v[11]; bool fl[11];
    for (int q = 0; q < 11; q++) { if (q == 1 || q == 9) continue; e = csv_num(s, fa[q], fb[q], &v[q], &fl[q]); if (e) return e; }
    int64_t dev, kind, payload, src, dst, tid;
    if (!csv_to_i64(v[3], &dev) || !csv_to_i64(v[4], &kind) || !csv_to_i64(v[5], &payload) || !csv_to_i64(v[7], &src) ||
        !csv_to_i64(v[8], &dst) || !csv_to_i64(v[10], &tid)) return E_BAD_INT;
    if (dev < -2147483647 || dev > 2147483647 || tid < -2147483647 || tid > 2147483647 || src < -2147483647 || src > 2147483647 ||
        dst < -2147483647 || dst > 2147483647 || kind < 0 || kind > 255) return E_RANGE;
    r->ts = v[0]; r->dur = v[2]; r->bw = v[6]; r->payload = payload;
    r->dev = (int32_t)dev; r->tid = (int32_t)tid; r->src = (int32_t)src; r->dst = (int32_t)dst;
    r->kind = (uint8_t)kind; r->bw_is_float = (uint8_t)((fl[6] ? 1 : 0) | 2);
    r->name_off = (uint16_t)fa[11]; r->name_len = (uint16_t)(fb[11] - fa[11]);
    return E_OK;
}
template <typename S>
SOFA_HD int parse_csv_net(S s, int n, NetRow* r) {
    int fa[13], fb[13];
    int e = csv_row_fields(s, n, fa, fb); if (e) return e;
    double ts, pl, src, dst; bool fl;
    e = csv_num(s, fa[0], fb[0], &ts, &fl); if (e) return e;
    r->ts = ts; r->proto = 0; r->payload = 0; r->src = 0; r->dst = 0;
    if (fb[5] <= fa[5]) return E_OK;                                  // the all-NaN row of a non-IP packet
    e = csv_num(s, fa[5], fb[5], &pl, &fl); if (e) return e;
    e = csv_num(s, fa[7], fb[7], &src, &fl); if (e) return e;
    e = csv_num(s, fa[8], fb[8], &dst, &fl); if (e) return e;
    if (!csv_to_i64(pl, &r->payload) || !csv_to_i64(src, &r->src) || !csv_to_i64(dst, &r->dst)) return E_BAD_INT;
    // 'network:tcp:...' or 'network:UDP,:...' (:198)
    const int a = fa[11], l = fb[11] - fa[11];
    if (l >= 12 && s[a + 8] == 't' && s[a + 9] == 'c' && s[a + 10] == 'p') r->proto = 2;
    else if (l >= 13 && s[a + 8] == 'U' && s[a + 9] == 'D' && s[a + 10] == 'P') r->proto = 1;
    else return E_FORMAT;
    return E_OK;
}

// ============================================================================ 13-column row values
struct RowVals {
    bool valid;              // false: the all-NaN row a `[]` record becomes
    double ts, dur;          // columns 0, 2 (float; the recorder side tables print an int duration: dur_i)
    bool dur_is_f; int64_t dur_i;
    bool ts_is_f; int64_t ts_i;          // vmstat / nvidia-smi dmon rows count seconds: an int column unless absolute
    int64_t cat;             // column 12 (0 for the trace sources)
    bool ev_is_f; double ev_f; int64_t ev_i;       // column 1
    int64_t dev, kind, payload;                     // 3, 4, 5
    bool bw_is_f; double bw_f; int64_t bw_i;        // 6
    int64_t src, dst, pid, tid;                     // 7..10
    // 11 = name (generator), 12 = category (always 0)
};
SOFA_HD RowVals vals_perf(const PerfRow& r) {
    RowVals v; v.valid = true; v.ts = r.ts; v.dur = r.dur; v.dur_is_f = true; v.dur_i = 0; v.ts_is_f = true; v.ts_i = 0; v.cat = 0; v.ev_is_f = true; v.ev_f = r.event; v.ev_i = 0;
    v.dev = -1; v.kind = -1; v.payload = 0; v.bw_is_f = false; v.bw_f = 0; v.bw_i = 0;
    v.src = -1; v.dst = -1; v.pid = r.pid; v.tid = r.tid; return v;
}
SOFA_HD RowVals vals_gpu(const GpuRow& r, int n_cudaproc) {
    RowVals v; v.valid = true; v.ts = r.ts; v.dur = r.dur; v.dur_is_f = true; v.dur_i = 0; v.ts_is_f = true; v.ts_i = 0; v.cat = 0; v.ev_is_f = false; v.ev_f = 0;
    v.ev_i = r.payload * 100 + 17;                                   // :329
    v.dev = r.dev; v.kind = r.kind; v.payload = r.payload; v.bw_is_f = (r.bw_is_float & 1) != 0; v.bw_f = r.bw; v.bw_i = 0;
    v.src = r.src; v.dst = r.dst; v.pid = n_cudaproc; v.tid = r.tid; return v;
}
SOFA_HD RowVals vals_net(const NetRow& r) {
    RowVals v; v.valid = r.proto != 0; v.ts = r.ts; v.dur_is_f = true; v.dur_i = 0; v.ts_is_f = true; v.ts_i = 0; v.cat = 0;
    v.dur = d_div((double)r.payload, 128.0e6);                       // :178
    v.ev_is_f = false; v.ev_f = 0; v.ev_i = r.payload * 100 + 17;
    v.dev = -1; v.kind = -1; v.payload = r.payload; v.bw_is_f = true; v.bw_f = 128.0e6; v.bw_i = 0;
    v.src = r.src; v.dst = r.dst; v.pid = -1; v.tid = -1; return v;
}

// column c of the row as (is_float, f, i) under the file's float-column mask
SOFA_HD void col_value(const RowVals& v, int c, uint32_t fmask, bool* isf, double* f, int64_t* i) {
    bool nat_f = false; double nf = 0; int64_t ni = 0;
    switch (c) {
        case 0: nat_f = v.ts_is_f; nf = v.ts; ni = v.ts_i; break;
        case 1: nat_f = v.ev_is_f; nf = v.ev_f; ni = v.ev_i; break;
        case 2: nat_f = v.dur_is_f; nf = v.dur; ni = v.dur_i; break;
        case 3: ni = v.dev; break;
        case 4: ni = v.kind; break;
        case 5: ni = v.payload; break;
        case 6: nat_f = v.bw_is_f; nf = v.bw_f; ni = v.bw_i; break;
        case 7: ni = v.src; break;
        case 8: ni = v.dst; break;
        case 9: ni = v.pid; break;
        case 10: ni = v.tid; break;
        default: ni = v.cat; break;    // 12: category
    }
    bool colf = ((fmask >> c) & 1) != 0 || nat_f;
    *isf = colf;
    *f = nat_f ? nf : (double)ni;
    *i = ni;
}

// ============================================================================ CSV (QUOTE_MINIMAL)
struct CsvCount {
    static constexpr bool WORDS = true;
    int len; bool special;
    SOFA_HD CsvCount() : len(0), special(false) {}
    SOFA_HD void put(char ch) {
        len++;
        if (ch == '"') { len++; special = true; }
        else if (ch == ',' || ch == '\n' || ch == '\r') special = true;
    }
    SOFA_HD void put_word(uint64_t w, int nb) {               // no per-byte loop: quotes are counted with a popcount
        const uint64_t first = lanes_first(nb);
        const uint64_t qt = lanes_eq(w, '"') & first;
        const uint64_t sp = (qt | lanes_eq(w, ',') | lanes_eq(w, '\n') | lanes_eq(w, '\r')) & first;
        len += nb + popc_lanes(qt);
        special |= (sp != 0);
    }
};
template <typename O>
struct CsvWrite {
    static constexpr bool WORDS = true;
    O out; int n;
    SOFA_HD void put(char ch) { if (ch == '"') out[n++] = '"'; out[n++] = ch; }
    SOFA_HD void put_word(uint64_t w, int nb) {
        if ((lanes_eq(w, '"') & lanes_first(nb)) == 0) {
            // eight predicated stores whatever nb is: lanes of a warp sit at different word boundaries of
            // their names, and a separate loop for partial words serialises them (measured 4-7 of 32 lanes)
#pragma unroll
            for (int b = 0; b < 8; b++) if (b < nb) out[n + b] = (char)((w >> (8 * b)) & 0xff);
            n += nb;
        } else put_word_bytes(*this, w, nb);
    }
};

// Float columns that can hold non-integral values (timestamp, event, duration, bandwidth) get
// a cached digit plan; other columns are integers even when the file prints them as floats.
SOFA_HD int fslot(int c) { return c <= 2 ? c : (c == 6 ? 3 : -1); }

struct CsvRowPlan { int len; bool special; F6Plan f[4]; };

// exact row length (incl. '\n') and the digit plans; len = -1 on an unformattable value
template <class NG>
SOFA_HD CsvRowPlan csv_row_plan(const RowVals& v, uint32_t fmask, const NG& ng) {
    CsvRowPlan p; p.special = false;
    if (!v.valid) { p.len = 13; return p; }          // twelve commas + '\n'
    int n = 13;                                      // separators
    bool bad = false;
#pragma unroll
    for (int c = 0; c < 13; c++) {
        if (c == 11) {
            CsvCount k; ng.gen(k);
            n += k.len + (k.special ? 2 : 0);
            p.special = k.special;
        } else {
            bool isf; double f; int64_t i;
            col_value(v, c, fmask, &isf, &f, &i);
            if (isf) {
                F6Plan q = fixed6_plan(f);
                if (!q.ok) bad = true;
                n += q.len;
                if (fslot(c) >= 0) p.f[fslot(c)] = q;
            } else n += len_i64(i);
        }
    }
    p.len = bad ? -1 : n;
    return p;
}
template <class NG, typename O>
SOFA_HD int csv_row_emit(const RowVals& v, uint32_t fmask, const NG& ng, const CsvRowPlan& p, O out) {
    if (!v.valid) { for (int i = 0; i < 12; i++) out[i] = ','; out[12] = '\n'; return 13; }
    int n = 0;
#pragma unroll
    for (int c = 0; c < 13; c++) {
        if (c == 11) {
            if (p.special) out[n++] = '"';
            CsvWrite<O> k{out + n, 0}; ng.gen(k); n += k.n;
            if (p.special) out[n++] = '"';
        } else {
            bool isf; double f; int64_t i;
            col_value(v, c, fmask, &isf, &f, &i);
            if (isf) n += (fslot(c) >= 0) ? fixed6_emit(p.f[fslot(c)], out + n) : fmt_fixed6(f, out + n);
            else n += fmt_i64(i, out + n);
        }
        out[n++] = (c == 12) ? '\n' : ',';
    }
    return n;
}
// returns total row length (incl. '\n'), or -1 on an unformattable value; *name_special out
template <class NG>
SOFA_HD int csv_row_len(const RowVals& v, uint32_t fmask, const NG& ng, bool* name_special) {
    CsvRowPlan p = csv_row_plan(v, fmask, ng);
    *name_special = p.special;
    return p.len;
}
template <class NG, typename O>
SOFA_HD int csv_row_write(const RowVals& v, uint32_t fmask, const NG& ng, bool name_special, O out) {
    CsvRowPlan p = csv_row_plan(v, fmask, ng);
    (void)name_special;
    return csv_row_emit(v, fmask, ng, p, out);
}

// ============================================================================ JSON records
// Final escaping is json.dump's (ensure_ascii=True): \" \\ \n \r \t \b \f, other chars
// outside ' '..'~' as \uXXXX (UTF-8 decoded, surrogate pairs above U+FFFF).
struct JsonEsc {
    uint32_t cp; int need; bool bad;
    SOFA_HD JsonEsc() : cp(0), need(0), bad(false) {}
    // feeds one byte; returns number of output chars into buf (<= 12)
    SOFA_HD int feed(unsigned char b, char* buf) {
        const char* hexd = "0123456789abcdef";
        if (need > 0) {
            if ((b & 0xC0) != 0x80) { bad = true; need = 0; return 0; }
            cp = (cp << 6) | (b & 0x3F);
            if (--need > 0) return 0;
            int n = 0;
            uint32_t u = cp;
            if (u >= 0x10000) {
                u -= 0x10000;
                uint32_t hi = 0xD800 + (u >> 10), lo = 0xDC00 + (u & 0x3FF);
                buf[n++] = '\\'; buf[n++] = 'u';
                buf[n++] = hexd[(hi >> 12) & 15]; buf[n++] = hexd[(hi >> 8) & 15]; buf[n++] = hexd[(hi >> 4) & 15]; buf[n++] = hexd[hi & 15];
                u = lo;
            }
            buf[n++] = '\\'; buf[n++] = 'u';
            buf[n++] = hexd[(u >> 12) & 15]; buf[n++] = hexd[(u >> 8) & 15]; buf[n++] = hexd[(u >> 4) & 15]; buf[n++] = hexd[u & 15];
            return n;
        }
        if (b < 0x80) {
            if (b == '"') { buf[0] = '\\'; buf[1] = '"'; return 2; }
            if (b == '\\') { buf[0] = '\\'; buf[1] = '\\'; return 2; }
            if (b >= 0x20 && b <= 0x7e) { buf[0] = (char)b; return 1; }
            buf[0] = '\\';
            switch (b) {
                case '\n': buf[1] = 'n'; return 2;
                case '\r': buf[1] = 'r'; return 2;
                case '\t': buf[1] = 't'; return 2;
                case '\b': buf[1] = 'b'; return 2;
                case '\f': buf[1] = 'f'; return 2;
                default: break;
            }
            buf[1] = 'u'; buf[2] = '0'; buf[3] = '0'; buf[4] = hexd[b >> 4]; buf[5] = hexd[b & 15];
            return 6;
        }
        if ((b & 0xE0) == 0xC0) { cp = b & 0x1F; need = 1; return 0; }
        if ((b & 0xF0) == 0xE0) { cp = b & 0x0F; need = 2; return 0; }
        if ((b & 0xF8) == 0xF0) { cp = b & 0x07; need = 3; return 0; }
        bad = true;
        return 0;
    }
};
SOFA_HD bool json_plain(unsigned char b) { return b >= 0x20 && b < 0x7f && b != '"' && b != '\\'; }
// lanes that json.dump would escape: < 0x20, >= 0x7f, '"', '\\'
SOFA_HD uint64_t json_special_lanes(uint64_t w) {
    const uint64_t K7 = 0x7f7f7f7f7f7f7f7full, K8 = 0x8080808080808080ull;
    uint64_t L = w & K7;
    uint64_t lt20 = ~(L + 0x6060606060606060ull) & K8;           // low 7 bits < 0x20
    uint64_t eq7f = (L + 0x0101010101010101ull) & K8;            // low 7 bits == 0x7f
    return (lt20 | eq7f | (w & K8) | lanes_eq(w, '"') | lanes_eq(w, '\\')) & K8;
}
struct JsonCount {
    static constexpr bool WORDS = false;
    JsonEsc e; int len;
    SOFA_HD JsonCount() : len(0) {}
    SOFA_HD void put(char ch) {
        if (e.need == 0 && json_plain((unsigned char)ch)) { len++; return; }
        char b[12]; len += e.feed((unsigned char)ch, b);
    }
    // (a whole-word fast path measured slower here: short viz names, higher register pressure)
    SOFA_HD void put_word(uint64_t w, int nb) { put_word_bytes(*this, w, nb); }
};
template <typename O>
struct JsonWrite {
    static constexpr bool WORDS = false;
    O out; int n; JsonEsc e;
    SOFA_HD void put(char ch) {
        if (e.need == 0 && json_plain((unsigned char)ch)) { out[n++] = ch; return; }
        char b[12]; int l = e.feed((unsigned char)ch, b); for (int i = 0; i < l; i++) out[n++] = b[i];
    }
    SOFA_HD void put_word(uint64_t w, int nb) { put_word_bytes(*this, w, nb); }
};

// Common-case name sinks: a name made only of printable ASCII plus the characters json.dump
// writes as two-byte escapes (\" \\ \n \r \t \b \f) needs no UTF-8 / \uXXXX state machine.
SOFA_HD uint64_t lanes_two_byte_escape(uint64_t w) {
    return lanes_eq(w, '"') | lanes_eq(w, '\\') | lanes_eq(w, '\n') | lanes_eq(w, '\r') | lanes_eq(w, '\t') |
           lanes_eq(w, '\b') | lanes_eq(w, '\f');
}
struct JsonScan {             // raw length, number of two-byte escapes, "needs the full encoder"
    static constexpr bool WORDS = true;
    int raw, two; bool complex;
    SOFA_HD JsonScan() : raw(0), two(0), complex(false) {}
    SOFA_HD void put(char ch) {
        unsigned char b = (unsigned char)ch;
        raw++;
        if (json_plain(b)) return;
        if (b == '"' || b == '\\' || b == '\n' || b == '\r' || b == '\t' || b == '\b' || b == '\f') two++;
        else complex = true;
    }
    SOFA_HD void put_word(uint64_t w, int nb) {               // mask arithmetic only, no per-byte loop
        const uint64_t first = lanes_first(nb);
        const uint64_t sp = json_special_lanes(w) & first;
        raw += nb;
        if (sp == 0) return;
        const uint64_t tw = lanes_two_byte_escape(w) & first;
        two += popc_lanes(tw);
        complex |= ((sp & ~tw) != 0);
    }
};
template <typename O>
struct JsonSimpleWrite {      // valid only when JsonScan found complex == false
    static constexpr bool WORDS = true;
    O out; int n;
    SOFA_HD void put(char ch) {
        unsigned char b = (unsigned char)ch;
        if (json_plain(b)) { out[n++] = ch; return; }
        out[n++] = '\\';
        char e = ch;
        if (b == '\n') e = 'n'; else if (b == '\r') e = 'r'; else if (b == '\t') e = 't';
        else if (b == '\b') e = 'b'; else if (b == '\f') e = 'f';
        out[n++] = e;
    }
    SOFA_HD void put_word(uint64_t w, int nb) {
        if ((json_special_lanes(w) & lanes_first(nb)) == 0) {
#pragma unroll
            for (int b = 0; b < 8; b++) if (b < nb) out[n + b] = (char)((w >> (8 * b)) & 0xff);     // predicated, see CsvWrite
            n += nb;
            return;
        }
        // a word holding a character that needs a two-byte escape (every nvprof kernel name ends in '\n'):
        // straight-line code, no loop, so lanes that reach it at different times spend little on it
        const uint64_t sp = json_special_lanes(w);
#pragma unroll
        for (int b = 0; b < 8; b++) {
            if (b < nb) {
                char ch = (char)((w >> (8 * b)) & 0xff);
                if ((sp >> (8 * b + 7)) & 1) {
                    out[n++] = '\\';
                    ch = ch == '\n' ? 'n' : (ch == '\r' ? 'r' : (ch == '\t' ? 't' : (ch == '\b' ? 'b' : (ch == '\f' ? 'f' : ch))));
                }
                out[n++] = ch;
            }
        }
    }
};

// Record text: {"x": V, "event": V, "y": V, "deviceId": V, ... "name": "S", "category": V}
// (x=timestamp, y=duration renamed, :351-355).  Keys are written with fully unrolled
// constant stores (one store per byte, no loads).  A record is emitted as ", " + {...}; the
// series writer drops the first two bytes.
#define SOFA_JKEY(lit) do { const char k_[] = lit; _Pragma("unroll") for (int i_ = 0; i_ < (int)sizeof(k_) - 1; i_++) out[n + i_] = k_[i_]; n += (int)sizeof(k_) - 1; } while (0)
#define SOFA_JNUM(c) do { \
        if (!v.valid) { out[n] = 'n'; out[n + 1] = 'u'; out[n + 2] = 'l'; out[n + 3] = 'l'; n += 4; } \
        else { bool isf; double f; int64_t i; col_value(v, c, fmask, &isf, &f, &i); \
               if (isf) n += (fslot(c) >= 0) ? json_emit(p.f[fslot(c) < 0 ? 0 : fslot(c)], out + n) : json_double(f, out + n); \
               else n += fmt_i64(i, out + n); } \
    } while (0)
// total bytes of the 13 key texts + ", " + "{" ... "}" framing
#define SOFA_JSON_FIXED (2 + 6 + 11 + 7 + 14 + 14 + 13 + 15 + 13 + 13 + 9 + 9 + 10 + 14 + 1)

// rare path (control characters / non-ASCII in a name): full json.dump escaping, kept out of line
// so that its state machine does not inflate the registers of the common path
template <class NG>
SOFA_HD_COLD int json_name_full_len(const NG& ng) {
    JsonCount k; ng.gen(k);
    if (k.e.bad || k.e.need) return -2;
    return k.len;
}
template <class NG, typename O>
SOFA_HD_COLD int json_name_full_write(const NG& ng, O out) {
    JsonWrite<O> w{out, 0, JsonEsc()}; ng.gen(w);
    return w.n;
}

struct JsonRowPlan { int len; bool simple_name; JPlan f[4]; };     // len: -1 unformattable number, -2 invalid UTF-8

// Series whose y_field is 'bandwidth' (tx_bandwidth / rx_bandwidth, sofa_preprocess.py:2051-2067):
// the record keeps the key "duration" and renames "bandwidth" to "y".  Only the side-table name
// generator specialises this; for the trace sources it folds to false at compile time.
template <class NG> struct JsonYIsBw { static SOFA_HD bool get(const NG&) { return false; } };

template <class NG>
SOFA_HD JsonRowPlan json_row_plan(const RowVals& v, uint32_t fmask, const NG& ng) {
    JsonRowPlan p;
    p.simple_name = true;
    int n = SOFA_JSON_FIXED - (JsonYIsBw<NG>::get(ng) ? 1 : 0);
    if (!v.valid) { p.len = n + 13 * 4; return p; }
    int bad = 0;
#pragma unroll
    for (int c = 0; c < 13; c++) {
        if (c == 11) {
            JsonScan sc; ng.gen(sc);
            p.simple_name = !sc.complex;
            if (!sc.complex) n += sc.raw + sc.two + 2;
            else {
                int l = json_name_full_len(ng);
                if (l < 0) bad = -2; else n += l + 2;
            }
        } else {
            bool isf; double f; int64_t i;
            col_value(v, c, fmask, &isf, &f, &i);
            if (isf) {
                JPlan q = json_plan(f);
                if (q.kind == JP_BAD && !bad) bad = -1;
                n += q.len;
                if (fslot(c) >= 0) p.f[fslot(c)] = q;
            } else n += len_i64(i);
        }
    }
    p.len = bad ? bad : n;
    return p;
}
template <class NG, typename O>
SOFA_HD int json_row_emit(const RowVals& v, uint32_t fmask, const NG& ng, const JsonRowPlan& p, O out) {
    int n = 0;
    SOFA_JKEY(", {\"x\": "); SOFA_JNUM(0);
    SOFA_JKEY(", \"event\": "); SOFA_JNUM(1);
    const bool ybw = JsonYIsBw<NG>::get(ng);
    if (ybw) SOFA_JKEY(", \"duration\": "); else SOFA_JKEY(", \"y\": ");
    SOFA_JNUM(2);
    SOFA_JKEY(", \"deviceId\": "); SOFA_JNUM(3);
    SOFA_JKEY(", \"copyKind\": "); SOFA_JNUM(4);
    SOFA_JKEY(", \"payload\": "); SOFA_JNUM(5);
    if (ybw) SOFA_JKEY(", \"y\": "); else SOFA_JKEY(", \"bandwidth\": ");
    SOFA_JNUM(6);
    SOFA_JKEY(", \"pkt_src\": "); SOFA_JNUM(7);
    SOFA_JKEY(", \"pkt_dst\": "); SOFA_JNUM(8);
    SOFA_JKEY(", \"pid\": "); SOFA_JNUM(9);
    SOFA_JKEY(", \"tid\": "); SOFA_JNUM(10);
    SOFA_JKEY(", \"name\": ");
    if (!v.valid) { out[n] = 'n'; out[n + 1] = 'u'; out[n + 2] = 'l'; out[n + 3] = 'l'; n += 4; }
    else {
        out[n++] = '"';
        if (p.simple_name) { JsonSimpleWrite<O> w{out + n, 0}; ng.gen(w); n += w.n; }
        else n += json_name_full_write(ng, out + n);
        out[n++] = '"';
    }
    SOFA_JKEY(", \"category\": "); SOFA_JNUM(12);
    out[n++] = '}';
    return n;
}
// returns length, or -1 (unformattable number) / -2 (invalid UTF-8)
template <class NG, typename O>
SOFA_HD int json_row_write(const RowVals& v, uint32_t fmask, const NG& ng, O out) {
    JsonRowPlan p = json_row_plan(v, fmask, ng);
    if (p.len < 0) return p.len;
    return json_row_emit(v, fmask, ng, p, out);
}
template <class NG>
SOFA_HD int json_row_len(const RowVals& v, uint32_t fmask, const NG& ng) {
    return json_row_plan(v, fmask, ng).len;
}

// ============================================================================ comm.csv rows
// df_gpu.to_csv(columns=[timestamp, pkt_src, pkt_dst, payload, bandwidth]) with the index and
// pandas' default float text = repr of the value read back from gputrace.csv (sofa_common.py:160-167)
struct CommRowPlan { int len; JPlan ts, bw; };
SOFA_HD CommRowPlan comm_row_plan(uint64_t index, const RowVals& v, uint32_t fmask) {
    CommRowPlan p;
    p.ts = repr6_plan(v.ts);
    bool bwf = ((fmask >> 6) & 1) != 0;
    int n = len_i64((int64_t)index) + len_i64(v.src) + len_i64(v.dst) + len_i64(v.payload) + 6 + p.ts.len;
    bool bad = p.ts.kind == JP_BAD;
    if (bwf) { p.bw = repr6_plan(v.bw_is_f ? v.bw_f : 0.0); bad |= p.bw.kind == JP_BAD; n += p.bw.len; }
    else n += len_i64(v.bw_i);
    p.len = bad ? -1 : n;
    return p;
}
template <typename O>
SOFA_HD int comm_row_emit(uint64_t index, const RowVals& v, uint32_t fmask, const CommRowPlan& p, O out) {
    int n = fmt_i64((int64_t)index, out);
    out[n++] = ',';
    n += json_emit(p.ts, out + n); out[n++] = ',';
    n += fmt_i64(v.src, out + n); out[n++] = ',';
    n += fmt_i64(v.dst, out + n); out[n++] = ',';
    n += fmt_i64(v.payload, out + n); out[n++] = ',';
    if ((fmask >> 6) & 1) n += json_emit(p.bw, out + n); else n += fmt_i64(v.bw_i, out + n);
    out[n++] = '\n';
    return n;
}

// ============================================================================ keyword filters
// Series.str.contains(keyword) for metacharacter-free keywords == byte-substring search
// (SURVEY.md A.4).  Shift-and matcher over up to 8 keywords of <= 32 bytes.
struct FilterTable {
    uint32_t mask[8][256];
    uint32_t final_bit[8];
    int nk;
    // packed variant: all keywords laid side by side in one 64-bit shift-and state (used when the
    // summed keyword length is <= 64): one table load + 4 ALU ops per byte for every keyword at once
    uint64_t pmask[256];
    uint64_t pstart, pfinal;
    uint8_t pfinal_pos[8];
    int packed;
};
struct MatchSink {
    static constexpr bool WORDS = true;
    const FilterTable* t; uint32_t st[8]; uint32_t hit; uint64_t ps, phit;
    SOFA_HD MatchSink(const FilterTable* tt) : t(tt), hit(0), ps(0), phit(0) { for (int k = 0; k < 8; k++) st[k] = 0; }
    SOFA_HD void put(char ch) {
        unsigned char b = (unsigned char)ch;
        if (t->packed) {
            ps = ((ps << 1) | t->pstart) & t->pmask[b];
            phit |= ps & t->pfinal;
            return;
        }
        for (int k = 0; k < t->nk; k++) {
            st[k] = ((st[k] << 1) | 1u) & t->mask[k][b];
            if (st[k] & t->final_bit[k]) hit |= 1u << k;
        }
    }
    SOFA_HD void put_word(uint64_t w, int nb) {
        if (!t->packed) { put_word_bytes(*this, w, nb); return; }
#pragma unroll
        for (int b = 0; b < 8; b++) {                    // straight-line shift-and steps, predicated on the byte count
            if (b < nb) {
                ps = ((ps << 1) | t->pstart) & t->pmask[(w >> (8 * b)) & 0xff];
                phit |= ps & t->pfinal;
            }
        }
    }
    SOFA_HD uint32_t hits() const {
        if (!t->packed) return hit;
        uint32_t h = 0;
        for (int k = 0; k < t->nk; k++) if ((phit >> t->pfinal_pos[k]) & 1) h |= 1u << k;
        return h;
    }
};
// host-side table construction (keywords: 1..32 bytes each, at most 8)
inline void filter_table_build(FilterTable* ft, int nkw, const char* const* kws, const size_t* lens) {
    for (int i = 0; i < 8; i++) { ft->final_bit[i] = 0; ft->pfinal_pos[i] = 0; for (int c = 0; c < 256; c++) ft->mask[i][c] = 0; }
    for (int c = 0; c < 256; c++) ft->pmask[c] = 0;
    ft->nk = nkw; ft->pstart = 0; ft->pfinal = 0;
    size_t tot = 0;
    for (int i = 0; i < nkw; i++) tot += lens[i];
    ft->packed = (nkw > 0 && tot <= 64) ? 1 : 0;
    size_t pos = 0;
    for (int i = 0; i < nkw; i++) {
        for (size_t j = 0; j < lens[i]; j++) {
            ft->mask[i][(unsigned char)kws[i][j]] |= 1u << j;
            if (ft->packed) ft->pmask[(unsigned char)kws[i][j]] |= 1ull << (pos + j);
        }
        ft->final_bit[i] = 1u << (lens[i] - 1);
        if (ft->packed) { ft->pstart |= 1ull << pos; ft->pfinal |= 1ull << (pos + lens[i] - 1); ft->pfinal_pos[i] = (uint8_t)(pos + lens[i] - 1); }
        pos += lens[i];
    }
}

// the packed variant alone (2 KB instead of 10): what the report.js main-series kernel stages in shared memory to
// match the keywords while it sizes the name
struct PackedFilter {
    uint64_t pmask[256];
    uint64_t pstart, pfinal;
    uint8_t pfinal_pos[8];
    int nk, pad;
};
struct PackedMatchSink {
    static constexpr bool WORDS = true;
    const PackedFilter* t; uint64_t ps, phit;
    SOFA_HD PackedMatchSink(const PackedFilter* tt) : t(tt), ps(0), phit(0) {}
    SOFA_HD void put(char ch) {
        ps = ((ps << 1) | t->pstart) & t->pmask[(unsigned char)ch];
        phit |= ps & t->pfinal;
    }
    SOFA_HD void put_word(uint64_t w, int nb) {
#pragma unroll
        for (int b = 0; b < 8; b++) {
            if (b < nb) {
                ps = ((ps << 1) | t->pstart) & t->pmask[(w >> (8 * b)) & 0xff];
                phit |= ps & t->pfinal;
            }
        }
    }
    SOFA_HD uint32_t hits() const {
        uint32_t h = 0;
        for (int k = 0; k < t->nk; k++) if ((phit >> t->pfinal_pos[k]) & 1) h |= 1u << k;
        return h;
    }
};
inline bool packed_filter_from(const FilterTable& ft, PackedFilter* pf) {
    if (!ft.packed) return false;
    for (int c = 0; c < 256; c++) pf->pmask[c] = ft.pmask[c];
    pf->pstart = ft.pstart; pf->pfinal = ft.pfinal;
    for (int i = 0; i < 8; i++) pf->pfinal_pos[i] = ft.pfinal_pos[i];
    pf->nk = ft.nk; pf->pad = 0;
    return true;
}

// FNV-1a 64 over generated characters (name dictionary keys)
struct HashSink {
    static constexpr bool WORDS = true;
    uint64_t h; int len; char c0, c1;
    SOFA_HD HashSink() : h(1469598103934665603ull), len(0), c0(0), c1(0) {}
    SOFA_HD void put(char ch) {
        if (len == 0) c0 = ch; else if (len == 1) c1 = ch;
        h = (h ^ (unsigned char)ch) * 1099511628211ull; len++;
    }
    SOFA_HD void put_word(uint64_t w, int nb) {                // FNV-1a steps, straight-line
        if (len == 0) { c0 = (char)(w & 0xff); if (nb > 1) c1 = (char)((w >> 8) & 0xff); }
        else if (len == 1) c1 = (char)(w & 0xff);
#pragma unroll
        for (int b = 0; b < 8; b++) if (b < nb) h = (h ^ ((w >> (8 * b)) & 0xff)) * 1099511628211ull;
        len += nb;
    }
};
template <typename O>
struct RawWrite {
    static constexpr bool WORDS = false;
    O out; int n;
    SOFA_HD void put(char ch) { out[n++] = ch; }
    SOFA_HD void put_word(uint64_t w, int nb) { put_word_bytes(*this, w, nb); }
};
struct LenSink {
    static constexpr bool WORDS = true;
    int len;
    SOFA_HD LenSink() : len(0) {}
    SOFA_HD void put(char) { len++; }
    SOFA_HD void put_word(uint64_t, int nb) { len += nb; }
};

}  // namespace sofa
