// Recorder side files, second group: diskstat.txt, blktrace.txt, strace.txt, pystacks.txt as
// sofa_preprocess.py:548-673, 688-781, 1618-1704, 1706-1761 turn them into diskstat.csv / diskstat_vector.csv /
// diskstat_vector_ui.csv, blktrace.csv, strace.csv, pystacks.csv and the diskstat, blktrace_starting_block,
// blktrace_time, strace, pystacks series of report.js.
//
// Row-level functions only (__host__ __device__): kernels_side2.cuh and the CPU test harness (tests/hostsim.cpp)
// call the same code.  These rows carry string columns (device names, 'D' / 'C', cpu and pid tokens) and, for
// diskstat, pandas' default float text, so they have their own small row writers over plain byte sinks instead of
// the 13-column numeric plans of rows.cuh.  These files are small next to the traces: one thread per line, a length
// pass, a scan and a write pass.
#pragma once
#include "side.cuh"

namespace sofa {

enum { S2_DISK = 0, S2_BLK = 1, S2_STRACE = 2, S2_PYSTACKS = 3,
       S2_VECCSV = 4, S2_BLKCSV = 5 };      // re-ingest of diskstat_vector.csv / blktrace.csv for sofa_analyze
// outputs of sofa_b200_side2_emit: 13-column CSV rows of list 0 / 1, report.js records of list 0 / 1,
// diskstat_vector.csv rows, diskstat_vector_ui.csv rows
enum { S2_CSV0 = 0, S2_CSV1 = 1, S2_JSON0 = 2, S2_JSON1 = 3, S2_VEC = 4, S2_VEC_UI = 5 };

// One parsed line.
//   disk   : f = float(m[0]), i[0..5] = int(m[2..7]), sp[0] = m[1] (device name), h = FNV-1a of the name,
//            link = slot of the name in the device table
//   blk    : cls 1 = 'D', 2 = 'C', 3 = another accepted event; f = float(fields[3]), i[0] = start block,
//            sp[0] = fields[1] (cpu), sp[1] = fields[4] (pid), flag bit 0 = int(fields[7]) failed,
//            link = line of the 'D' a 'C' completes (-1: none)
//   strace : cls 1 = row, 2 = counted in the total only; f = t_begin, g = duration, flag bit 0 = duration is a float,
//            i[0] = tid, sp[0] = from the first to the last token of the name
//   py     : (pair of lines) cls 1 = kept; f = int(ts) / 10**6, sp[0] = the stack line, link = rank among the kept pairs
struct S2Span { uint32_t a, b; };
struct S2Line {
    double f, g;
    int64_t i[6];
    uint64_t h;
    S2Span sp[2];
    int32_t link;
    uint8_t cls, flag;
    uint16_t pad;
};

struct S2Count {
    static constexpr bool WORDS = false;
    long long len;
    SOFA_HD S2Count() : len(0) {}
    SOFA_HD void put(char) { len++; }
    SOFA_HD void put_word(uint64_t, int nb) { len += nb; }
};
struct S2Write {
    static constexpr bool WORDS = false;
    char* out; long long n;
    SOFA_HD void put(char ch) { out[n++] = ch; }
    SOFA_HD void put_word(uint64_t w, int nb) { for (int b = 0; b < nb; b++) { out[n++] = (char)(w & 0xff); w >>= 8; } }
};

// ---------------------------------------------------------------------------- number texts into a sink
template <class K> SOFA_HD bool s2_put_plan(K& k, const JPlan& p, bool repr) {
    if (p.kind == JP_BAD) return false;
    char b[40];
    int l = repr ? repr_emit(p, b) : json_emit(p, b);
    for (int q = 0; q < l; q++) k.put(b[q]);
    return true;
}
template <class K> SOFA_HD bool s2_put_repr(K& k, double x) { return s2_put_plan(k, repr_plan(x), true); }     // pandas' default float text
template <class K> SOFA_HD bool s2_put_jsonf(K& k, double x) { return s2_put_plan(k, json_plan(x), false); }   // to_json -> loads -> dump
template <class K> SOFA_HD bool s2_put_f6(K& k, double x) {                                                    // '%.6f'
    F6Plan p = fixed6_plan(x);
    if (!p.ok) return false;
    char b[40];
    int l = fixed6_emit(p, b);
    for (int q = 0; q < l; q++) k.put(b[q]);
    return true;
}

// round(x, 2) of a finite double (float.__round__: the correctly rounded two-decimal string of the exact binary
// value, half to even, read back by strtod).  Returns false when the codec does not cover the value.
SOFA_HD bool py_round2(double x, double* out) {
    uint64_t xb = d_bits(x);
    int be = (int)((xb >> 52) & 0x7ff);
    if (be == 0x7ff) return false;
    bool neg = (xb >> 63) != 0;
    uint64_t frac = xb & 0xfffffffffffffull;
    if (be == 0) { if (frac) return false; *out = x; return true; }        // +-0.0; subnormals not covered
    uint64_t m = frac | (1ull << 52);
    int e = be - 1075;
    if (e >= 0) { *out = x; return true; }                                  // an integer: "N.00" reads back as N
    int s = -e;
    if (s > 120) { *out = neg ? -0.0 : 0.0; return true; }
    u128 num = (u128)m * 100u;
    u128 q = num >> s;
    u128 rem = num - (q << s);
    u128 half = ((u128)1) << (s - 1);
    if (rem > half || (rem == half && ((uint64_t)q & 1))) q++;
    if (q >> 63) return false;
    if (q == 0) { *out = neg ? -0.0 : 0.0; return true; }
    return dec_to_double((uint64_t)q, -2, neg, out) == P_OK;
}

// ---------------------------------------------------------------------------- strings
struct S2Special {             // does a CSV field need quotes (QUOTE_MINIMAL)?
    static constexpr bool WORDS = false;
    bool special;
    SOFA_HD S2Special() : special(false) {}
    SOFA_HD void put(char ch) { if (ch == '"' || ch == ',' || ch == '\n' || ch == '\r') special = true; }
    SOFA_HD void put_word(uint64_t w, int nb) { for (int b = 0; b < nb; b++) { put((char)(w & 0xff)); w >>= 8; } }
};
template <class K> struct S2CsvQ {          // doubles the quotes of a quoted field
    static constexpr bool WORDS = false;
    K* k;
    SOFA_HD void put(char ch) { if (ch == '"') k->put('"'); k->put(ch); }
    SOFA_HD void put_word(uint64_t w, int nb) { for (int b = 0; b < nb; b++) { put((char)(w & 0xff)); w >>= 8; } }
};
template <class K> struct S2JsonQ {         // json.dump(ensure_ascii=True) escaping
    static constexpr bool WORDS = false;
    K* k; JsonEsc e;
    SOFA_HD void put(char ch) {
        if (e.need == 0 && json_plain((unsigned char)ch)) { k->put(ch); return; }
        char b[12]; int l = e.feed((unsigned char)ch, b);
        for (int q = 0; q < l; q++) k->put(b[q]);
    }
    SOFA_HD void put_word(uint64_t w, int nb) { for (int b = 0; b < nb; b++) { put((char)(w & 0xff)); w >>= 8; } }
};
// g(sink) feeds the string's bytes
template <class K, class G> SOFA_HD void s2_csv_str(K& k, const G& g) {
    S2Special sp; g(sp);
    if (!sp.special) { g(k); return; }
    k.put('"');
    S2CsvQ<K> q{&k}; g(q);
    k.put('"');
}
template <class K, class G> SOFA_HD bool s2_json_str(K& k, const G& g) {
    k.put('"');
    S2JsonQ<K> q{&k, JsonEsc()}; g(q);
    k.put('"');
    return !(q.e.bad || q.e.need);
}
struct S2ChrGen {
    char ch;
    template <class K> SOFA_HD void operator()(K& k) const { k.put(ch); }
};
struct S2SpanGen {
    const char* text; S2Span s;
    template <class K> SOFA_HD void operator()(K& k) const { for (uint32_t i = s.a; i < s.b; i++) k.put(text[i]); }
};

// ---------------------------------------------------------------------------- the view a row is written from
struct S2View {
    int table;
    const char* text; uint64_t nbytes;
    const uint32_t* lstart;          // n_lines + 1 entries: start of every line, then the end of the text (+1 when it ends without '\n')
    uint64_t n_lines; int last_has_nl;
    const S2Line* raw;               // one per line (pystacks: per pair of lines)
    uint64_t n;                      // entries of raw
    const uint32_t* kept;            // pystacks: kept pairs in file order
    uint64_t n_kept;
    int absolute; double time_base;
    // diskstat
    uint32_t n_dev;                  // distinct device names (the reference's n_dev)
    const int32_t* dev_secsize;      // per device-table slot
    const uint8_t* dev_flags;        // per slot: bit 0 = dropped (its total never moved), bit 1 = selected by --diskstat_filters
    uint32_t disk_fl;                // bit 0 / 1 / 2: the r_time / w_time / await_time column holds a float
    // strace / pystacks
    int64_t pid; int64_t cat; int cat_is_str; S2Span cat_span; int dur_all_int;
};

SOFA_HD int s2_line_len(const S2View& v, uint64_t i) {            // content length of line i, without its '\n'
    return (int)(v.lstart[i + 1] - v.lstart[i]) - 1;
}

// ---------------------------------------------------------------------------- diskstat.txt (:548-673)
SOFA_HD uint64_t s2_fnv(const char* s, int n) {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 1099511628211ull; }
    return h;
}
// line: "time,dev,sectors_read,sectors_written,reads,writes,ms_reading,ms_writing[,...]"; the reference drops the last
// character of every line (`[:-1]`, the '\n' - or a real character of a last line that has none: n is that length)
SOFA_HD int s2_parse_disk(const char* text, uint32_t base, int n, S2Line* r) {
    const char* s = text + base;
    r->cls = 0; r->flag = 0; r->pad = 0; r->link = -1; r->f = 0; r->g = 0; r->h = 0;
    for (int q = 0; q < 6; q++) r->i[q] = 0;
    int ca[8], nc = 0;
    for (int i = 0; i < n && nc < 8; i++) if (s[i] == ',') ca[nc++] = i;
    if (nc < 1) return E_FEW_TOKENS;                               // x[1] raises IndexError while the device list is built
    r->sp[0].a = base + ca[0] + 1; r->sp[0].b = base + (nc >= 2 ? ca[1] : n);
    r->h = s2_fnv(text + r->sp[0].a, (int)(r->sp[0].b - r->sp[0].a));
    // The reference converts a field only where a row uses it (float(m[0]) and int(m[2..5]) in every row the line takes
    // part in, int(m[6]) / int(m[7]) only when that row saw reads / writes): a bad field is an error of the rows that
    // need it.  flag bit q: field q + 2 is missing or not an int (bit 6: m[0] is not a float); pad bit q: Python takes
    // the spelling, this codec does not.
    int st = parse_double(s, ca[0], &r->f);
    if (st == P_UNSUPPORTED) r->pad |= 1u << 6; else if (st) r->flag |= 1u << 6;
    for (int q = 0; q < 6; q++) {
        if (nc < 2 + q) { r->flag |= (uint8_t)(1u << q); continue; }
        int a = ca[1 + q] + 1, b = (nc >= 3 + q) ? ca[2 + q] : n;
        st = parse_i64(s + a, b - a, &r->i[q]);
        if (st == P_UNSUPPORTED || (st == P_OK && (r->i[q] > (1ll << 52) || r->i[q] < -(1ll << 52)))) { r->pad |= (uint16_t)(1u << q); r->i[q] = 0; }
        else if (st) { r->flag |= (uint8_t)(1u << q); r->i[q] = 0; }
    }
    r->cls = 1;
    return E_OK;
}
SOFA_HD int s2_disk_need(const S2Line& x, unsigned bits) {           // the row needs the fields in `bits` (6 = m[0]) of line x
    if (x.flag & bits) return (x.flag & bits & 0x40u) ? E_BAD_FLOAT : E_BAD_INT;
    if (x.pad & bits) return E_UNSUPPORTED_NUM;
    return E_OK;
}
struct DiskRow {
    double t_begin, dur, thr, rbw, wbw, tbw, r_time, w_time, a_time;
    int64_t d_read, d_write, d_total, r_iops, w_iops, iops;
    uint8_t rt_f, wt_f, at_f;            // the value is a float (else the int 0)
};
SOFA_HD int s2_disk_row(const S2Line& m, const S2Line& l, int64_t secsize, int absolute, double time_base, DiskRow* o) {
    const int64_t lim = 1ll << 52;
    if (secsize <= 0 || secsize > (1 << 20)) return E_RANGE;
    int need = s2_disk_need(m, 0x4fu); if (need) return need;       // m[0], m[2..5] of both lines
    need = s2_disk_need(l, 0x4fu); if (need) return need;
    __int128 dr = ((__int128)m.i[0] - l.i[0]) * secsize, dw = ((__int128)m.i[1] - l.i[1]) * secsize;
    __int128 dt = dr + dw;
    if (dr >= lim || dr <= -lim || dw >= lim || dw <= -lim || dt >= lim || dt <= -lim) return E_RANGE;
    o->d_read = (int64_t)dr; o->d_write = (int64_t)dw; o->d_total = (int64_t)dt;
    o->t_begin = absolute ? m.f : d_sub(m.f, time_base);
    o->dur = d_sub(m.f, l.f);
    if (o->dur == 0.0) return E_FORMAT;                            // ZeroDivisionError in the reference
    o->tbw = d_div((double)o->d_total, o->dur);
    o->rbw = d_div((double)o->d_read, o->dur);
    o->wbw = d_div((double)o->d_write, o->dur);
    if (!d_isfinite(o->tbw) || !d_isfinite(o->rbw) || !d_isfinite(o->wbw)) return E_FORMAT;
    if (!py_round2(d_div(o->tbw, 1048576.0), &o->thr)) return E_FORMAT;
    int64_t r_io = m.i[2] - l.i[2], w_io = m.i[3] - l.i[3];
    const unsigned lazy = ((r_io > 0 || r_io + w_io > 0) ? 0x10u : 0u) | ((w_io > 0 || r_io + w_io > 0) ? 0x20u : 0u);
    need = s2_disk_need(m, lazy); if (need) return need;            // int(m[6]) / int(m[7]) only where the reference evaluates them
    need = s2_disk_need(l, lazy); if (need) return need;
    int64_t r_ms = m.i[4] - l.i[4], w_ms = m.i[5] - l.i[5];
    int e = trunc_i64(d_div((double)r_io, o->dur), &o->r_iops); if (e) return e;
    e = trunc_i64(d_div((double)w_io, o->dur), &o->w_iops); if (e) return e;
    e = trunc_i64(d_div((double)(r_io + w_io), o->dur), &o->iops); if (e) return e;
    o->rt_f = o->wt_f = o->at_f = 0; o->r_time = o->w_time = o->a_time = 0;
    if (r_io > 0) { o->r_time = d_div((double)r_ms, (double)r_io); o->rt_f = 1; }         // int / int: both exact doubles
    if (w_io > 0) { o->w_time = d_div((double)w_ms, (double)w_io); o->wt_f = 1; }
    if (r_io + w_io > 0) { o->a_time = d_div((double)(r_ms + w_ms), (double)(r_io + w_io)); o->at_f = 1; }
    return E_OK;
}

// ---------------------------------------------------------------------------- blktrace.txt (:688-781)
SOFA_HD bool s2_contains(const char* s, int n, const char* w) { return line_contains(s, n, w); }
// str.split(): up to maxt tokens; returns the TOTAL number of tokens of the line
SOFA_HD int s2_split(const char* s, int n, Tok* t, int maxt) {
    int k = 0, i = 0;
    while (i < n) {
        while (i < n && is_space((unsigned char)s[i])) i++;
        if (i >= n) break;
        int a = i;
        while (i < n && !is_space((unsigned char)s[i])) i++;
        if (k < maxt) { t[k].a = a; t[k].b = i; }
        k++;
    }
    return k;
}
// n: content length, has_nl: the line ends with '\n' (len(line) counts it)
SOFA_HD int s2_parse_blk(const char* text, uint32_t base, int n, bool has_nl, S2Line* r) {
    const char* s = text + base;
    r->cls = 0; r->flag = 0; r->link = -1; r->f = 0; r->g = 0; r->h = 0;
    if (n + (has_nl ? 1 : 0) <= 50) return E_OK;
    if (s2_contains(s, n, "Read") || s2_contains(s, n, "CPU") || s2_contains(s, n, "IO unplugs")) return E_OK;
    Tok t[8];
    int nt = s2_split(s, n, t, 8);
    if (nt < 7) return E_FEW_TOKENS;                               // fields[6] raises IndexError
    int st = parse_double(s + t[3].a, t[3].b - t[3].a, &r->f);
    if (st) return num_status_to_err(st, true);
    r->sp[0].a = base + t[1].a; r->sp[0].b = base + t[1].b;
    r->sp[1].a = base + t[4].a; r->sp[1].b = base + t[4].b;
    r->i[0] = 0;
    if (nt >= 8) {
        st = parse_i64(s + t[7].a, t[7].b - t[7].a, &r->i[0]);
        if (st == P_UNSUPPORTED) return E_UNSUPPORTED_NUM;
        if (st) { r->i[0] = 0; r->flag |= 1; }                     // `except:` -> 0, "blktrace maybe record failed!"
    } else r->flag |= 1;
    const int el = t[5].b - t[5].a;
    r->cls = (el == 1 && s[t[5].a] == 'D') ? 1 : ((el == 1 && s[t[5].a] == 'C') ? 2 : 3);
    return E_OK;
}

// ---------------------------------------------------------------------------- strace.txt (:1618-1704)
// plot_ratio is applied by the caller (line index).  min_time: cfg.strace_min_time.
SOFA_HD int s2_parse_strace(const char* text, uint32_t base, int n, int64_t pid, int absolute, double time_base, double min_time,
                            S2Line* r) {
    const char* s = text + base;
    r->cls = 0; r->flag = 0; r->link = -1; r->f = 0; r->g = 0; r->h = 0;
    const char* keys[10] = {"resume", "nanosleep", "clock_gettime", "brk", "stat", "close", "exited", "access", "lseek", "getrusage"};
    for (int q = 0; q < 10; q++) if (s2_contains(s, n, keys[q])) return E_OK;
    // tokens: the first three, and the last four (name = tokens [a, L-3), duration in token L-1)
    Tok first[3] = {}, last[4] = {};
    int L = 0, i = 0;
    while (i < n) {
        while (i < n && is_space((unsigned char)s[i])) i++;
        if (i >= n) break;
        int a = i;
        while (i < n && !is_space((unsigned char)s[i])) i++;
        if (L < 3) { first[L].a = a; first[L].b = i; }
        last[0] = last[1]; last[1] = last[2]; last[2] = last[3]; last[3].a = a; last[3].b = i;
        L++;
    }
    if (L < 1) return E_FEW_TOKENS;                                // fields[0] raises IndexError
    bool has_pid = false;
    for (int q = first[0].a; q + 3 <= first[0].b; q++) if (s[q] == 'p' && s[q + 1] == 'i' && s[q + 2] == 'd') has_pid = true;
    int na, ti;                                                    // first token of the name, token holding the time
    if (has_pid) {
        if (L < 3) return E_FEW_TOKENS;
        int e = first[1].a;
        while (e < first[1].b && s[e] != ']') e++;
        int st = parse_i64(s + first[1].a, e - first[1].a, &r->i[0]);
        if (st) return num_status_to_err(st, false);
        ti = 2; na = 3;
    } else {
        if (L < 2) return E_FEW_TOKENS;
        r->i[0] = pid;
        ti = 1; na = 1;
    }
    double t;
    int st = parse_double(s + first[ti].a, first[ti].b - first[ti].a, &t);
    if (st) return num_status_to_err(st, true);
    r->f = absolute ? t : d_sub(t, time_base);
    // ''.join(fields[na:-3]): tokens na .. L-4
    r->sp[0].a = r->sp[0].b = base;
    if (L - 3 > na) {
        // token na is one of first[] (na <= 3 needs first[3] when na == 3: recompute by walking)
        int k = 0, j = 0, a0 = 0;
        while (j < n) {
            while (j < n && is_space((unsigned char)s[j])) j++;
            if (j >= n) break;
            int a = j;
            while (j < n && !is_space((unsigned char)s[j])) j++;
            if (k == na) { a0 = a; break; }
            k++;
        }
        r->sp[0].a = base + a0;
        r->sp[0].b = base + last[0].b;                             // token L-4 = last[0] when L >= 4
    }
    // float(fields[-1].split('<')[1].split('>')[0]) or 0
    const Tok& lt = last[3];
    int p = lt.a;
    while (p < lt.b && s[p] != '<') p++;
    double dur = 0; bool is_f = false;
    if (p < lt.b) {
        int a = p + 1, b = a;
        while (b < lt.b && s[b] != '<' && s[b] != '>') b++;
        double d;
        int ps = parse_double(s + a, b - a, &d);
        if (ps == P_UNSUPPORTED) return E_UNSUPPORTED_NUM;
        if (ps == P_OK) { dur = d; is_f = true; }
    }
    r->g = dur; r->flag = is_f ? 1 : 0;
    r->cls = (dur < min_time) ? 2 : 1;
    return E_OK;
}
// the name: the span with its whitespace squeezed out
struct S2StraceName {
    const char* text; S2Span s;
    template <class K> SOFA_HD void operator()(K& k) const {
        for (uint32_t i = s.a; i < s.b; i++) if (!is_space((unsigned char)text[i])) k.put(text[i]);
    }
};

// ---------------------------------------------------------------------------- pystacks.txt (:1706-1761)
// pair j = lines 2j (microsecond stamp) and 2j + 1 (stack).  n0 / n1: content lengths.
SOFA_HD int s2_parse_py(const char* text, uint32_t b0, int n0, uint32_t b1, int n1, S2Line* r) {
    r->cls = 0; r->flag = 0; r->link = -1; r->f = 0; r->g = 0; r->h = 0;
    r->sp[0].a = b1; r->sp[0].b = b1 + (uint32_t)n1;
    if (s2_contains(text + b1, n1, "idle")) return E_OK;           // ignore_idle
    int64_t us;
    int st = parse_i64(text + b0, n0, &us);
    if (st) return num_status_to_err(st, false);
    if (us > (1ll << 53) || us < -(1ll << 53)) return E_UNSUPPORTED_NUM;
    r->f = d_div((double)us, 1e6);                                 // int / int, both exact doubles
    r->i[0] = us;
    r->cls = 1;
    return E_OK;
}
struct S2PyName {                 // fs.replace('\n', '').replace(';', '<br>')
    const char* text; S2Span s;
    template <class K> SOFA_HD void operator()(K& k) const {
        for (uint32_t i = s.a; i < s.b; i++) {
            char c = text[i];
            if (c == ';') { k.put('<'); k.put('b'); k.put('r'); k.put('>'); }
            else if (c != '\n') k.put(c);
        }
    }
};

// ---------------------------------------------------------------------------- CSVs read back by sofa_analyze
// diskstat_vector.csv (sofa_analyze.py:640-700 reads it with pd.read_csv): line 0 is the header.  The seven columns the
// profile uses are kept as doubles: f = d_read, g = d_write, i[0..4] = bits of d_disk_total, r_iops, w_iops, iops, await_time.
SOFA_HD double s2_vec_col(const S2Line& r, int col) {         // 0 d_read, 1 d_write, 2 d_disk_total, 3 r_iops, 4 w_iops, 5 iops, 6 await_time
    return col == 0 ? r.f : (col == 1 ? r.g : bits_d((uint64_t)r.i[col - 2]));
}
SOFA_HD int s2_parse_veccsv(const char* text, uint32_t base, int n, uint64_t line, S2Line* r) {
    const char* s = text + base;
    r->cls = 0; r->flag = 0; r->link = -1; r->f = 0; r->g = 0; r->h = 0;
    if (line == 0) return E_OK;
    int ca[11], nc = 0;
    for (int i = 0; i < n; i++) {
        if (s[i] == '"') return E_UNSUPPORTED_NUM;                // a quoted device name
        if (s[i] == ',') { if (nc >= 10) return E_FORMAT; ca[nc++] = i; }
    }
    if (nc != 10) return E_FEW_TOKENS;
    r->sp[0].a = base + ca[0] + 1; r->sp[0].b = base + ca[1];
    r->h = s2_fnv(text + r->sp[0].a, (int)(r->sp[0].b - r->sp[0].a));
    // columns: 0 time, 1 dev, 2 r_iops, 3 w_iops, 4 iops, 5 r_time, 6 w_time, 7 await_time, 8 d_read, 9 d_write, 10 d_disk_total
    const int src[7] = {8, 9, 10, 2, 3, 4, 7};
    for (int q = 0; q < 7; q++) {
        const int a = ca[src[q] - 1] + 1, b = src[q] == 10 ? n : ca[src[q]];
        double v;
        int st = parse_double(s + a, b - a, &v);
        if (st) return num_status_to_err(st, true);
        if (q == 0) r->f = v; else if (q == 1) r->g = v; else r->i[q - 2] = (int64_t)d_bits(v);
    }
    r->cls = 1;
    return E_OK;
}
// blktrace.csv: cls 2 = a row whose event is 'C', g = its duration (blktrace_latency_profile, sofa_analyze.py:627-631)
SOFA_HD int s2_parse_blkcsv(const char* text, uint32_t base, int n, uint64_t line, S2Line* r) {
    const char* s = text + base;
    r->cls = 0; r->flag = 0; r->link = -1; r->f = 0; r->g = 0; r->h = 0;
    if (line == 0) return E_OK;
    int ca[3], nc = 0;
    for (int i = 0; i < n && nc < 3; i++) {
        if (s[i] == '"') return E_UNSUPPORTED_NUM;
        if (s[i] == ',') ca[nc++] = i;
    }
    if (nc < 3) return E_FEW_TOKENS;
    r->cls = (ca[1] - ca[0] == 2 && s[ca[0] + 1] == 'C') ? 2 : 1;
    int st = parse_double(s + ca[1] + 1, ca[2] - ca[1] - 1, &r->g);
    if (st) return num_status_to_err(st, true);
    return E_OK;
}

// ---------------------------------------------------------------------------- 13-column rows
enum { C_INT = 0, C_FLT = 1, C_SPAN = 2, C_CHR = 3 };
struct S2Cell { int kind; int64_t i; double f; S2Span s; };
struct S2Row13 { S2Cell c[13]; bool repr; bool ybw; int name_kind; };   // c[11]: the name, written by s2_name
SOFA_HD void s2_cell_int(S2Cell* c, int64_t v) { c->kind = C_INT; c->i = v; c->f = 0; c->s.a = c->s.b = 0; }
SOFA_HD void s2_cell_flt(S2Cell* c, double v) { c->kind = C_FLT; c->i = 0; c->f = v; c->s.a = c->s.b = 0; }
SOFA_HD void s2_cell_span(S2Cell* c, S2Span s) { c->kind = C_SPAN; c->i = 0; c->f = 0; c->s = s; }
SOFA_HD void s2_cell_chr(S2Cell* c, char ch) { c->kind = C_CHR; c->i = (unsigned char)ch; c->f = 0; c->s.a = c->s.b = 0; }

struct S2NameArg { int kind; const char* text; S2Span s; int64_t a, b; double f; };
// kinds: 0 diskstat, 1 blktrace D, 2 blktrace C, 3 strace, 4 pystacks
struct S2NameGen {
    S2NameArg na;
    template <class K> SOFA_HD void operator()(K& k) const {
        if (na.kind == 0) {                                        // :640
            put_cstr(k, "diskstat_dev:"); S2SpanGen{na.text, na.s}(k); put_cstr(k, " (read|write): |");
            put_pad3(k, na.a); k.put('|'); put_pad3(k, na.b); put_cstr(k, "| bytes");
        } else if (na.kind == 1) { put_cstr(k, "starting_block="); put_i64(k, na.a); }
        else if (na.kind == 2) { put_cstr(k, "latency="); s2_put_f6(k, na.f); }
        else if (na.kind == 3) S2StraceName{na.text, na.s}(k);
        else S2PyName{na.text, na.s}(k);
    }
};

template <class K>
SOFA_HD int s2_csv13(K& k, const char* text, const S2Row13& r, const S2NameGen& ng) {
    bool ok = true;
    for (int c = 0; c < 13; c++) {
        if (c == 11) s2_csv_str(k, ng);
        else {
            const S2Cell& x = r.c[c];
            if (x.kind == C_INT) put_i64(k, x.i);
            else if (x.kind == C_FLT) ok &= r.repr ? s2_put_repr(k, x.f) : s2_put_f6(k, x.f);
            else if (x.kind == C_CHR) s2_csv_str(k, S2ChrGen{(char)x.i});
            else s2_csv_str(k, S2SpanGen{text, x.s});
        }
        k.put(c == 12 ? '\n' : ',');
    }
    return ok ? E_OK : E_FORMAT;
}
template <class K>
SOFA_HD int s2_json13(K& k, const char* text, const S2Row13& r, const S2NameGen& ng) {
    const char* keys[13] = {", {\"x\": ", ", \"event\": ", ", \"y\": ", ", \"deviceId\": ", ", \"copyKind\": ", ", \"payload\": ",
                            ", \"bandwidth\": ", ", \"pkt_src\": ", ", \"pkt_dst\": ", ", \"pid\": ", ", \"tid\": ", ", \"name\": ",
                            ", \"category\": "};
    bool ok = true, utf = true;
    for (int c = 0; c < 13; c++) {
        if (r.ybw && c == 2) put_cstr(k, ", \"duration\": ");
        else if (r.ybw && c == 6) put_cstr(k, ", \"y\": ");
        else put_cstr(k, keys[c]);
        if (c == 11) { utf &= s2_json_str(k, ng); continue; }
        const S2Cell& x = r.c[c];
        if (x.kind == C_INT) put_i64(k, x.i);
        else if (x.kind == C_FLT) ok &= s2_put_jsonf(k, x.f);
        else if (x.kind == C_CHR) utf &= s2_json_str(k, S2ChrGen{(char)x.i});
        else utf &= s2_json_str(k, S2SpanGen{text, x.s});
    }
    k.put('}');
    return !utf ? E_UTF8 : (ok ? E_OK : E_FORMAT);
}

// ---------------------------------------------------------------------------- row of entry `i` for output `what`
// returns E_OK (row written), -1 (this entry has no row in this output) or an error code
template <class K>
SOFA_HD int s2_emit_row(const S2View& v, uint64_t i, int what, K& k) {
    S2Row13 r; r.repr = false; r.ybw = false;
    for (int c = 0; c < 13; c++) s2_cell_int(&r.c[c], -1);
    S2NameGen ng; ng.na.text = v.text; ng.na.s.a = ng.na.s.b = 0; ng.na.a = ng.na.b = 0; ng.na.f = 0; ng.na.kind = 0;
    const bool json = (what == S2_JSON0 || what == S2_JSON1);
    const int list = (what == S2_CSV1 || what == S2_JSON1) ? 1 : 0;
    if (v.table == S2_DISK) {
        if (list) return -1;
        if (i < v.n_dev) return -1;
        const S2Line& m = v.raw[i];
        const uint8_t fl = v.dev_flags[m.link];
        if (fl & 1) return -1;                                     // a device whose total never moved (:650-653)
        if (what == S2_VEC_UI && !(fl & 2)) return -1;
        DiskRow d;
        int e = s2_disk_row(m, v.raw[i - v.n_dev], v.dev_secsize[m.link], v.absolute, v.time_base, &d);
        if (e) return e;
        if (what == S2_VEC || what == S2_VEC_UI) {
            bool ok = s2_put_repr(k, d.t_begin); k.put(',');
            s2_csv_str(k, S2SpanGen{v.text, m.sp[0]}); k.put(',');
            put_i64(k, d.r_iops); k.put(','); put_i64(k, d.w_iops); k.put(','); put_i64(k, d.iops); k.put(',');
            const double tv[3] = {d.r_time, d.w_time, d.a_time};
            const uint8_t tf[3] = {d.rt_f, d.wt_f, d.at_f};
            for (int q = 0; q < 3; q++) {
                if ((v.disk_fl >> q) & 1) ok &= s2_put_repr(k, tf[q] ? tv[q] : 0.0);
                else k.put('0');
                k.put(',');
            }
            ok &= s2_put_repr(k, d.rbw); k.put(',');
            ok &= s2_put_repr(k, d.wbw); k.put(',');
            ok &= s2_put_repr(k, d.tbw); k.put('\n');
            return ok ? E_OK : E_FORMAT;
        }
        r.repr = true; r.ybw = true;
        s2_cell_flt(&r.c[0], d.t_begin); s2_cell_flt(&r.c[2], d.dur); s2_cell_span(&r.c[3], m.sp[0]);
        s2_cell_int(&r.c[5], d.d_total); s2_cell_flt(&r.c[6], d.thr); s2_cell_int(&r.c[12], 0);
        ng.na.kind = 0; ng.na.s = m.sp[0]; ng.na.a = d.d_read; ng.na.b = d.d_write;
    } else if (v.table == S2_BLK) {
        const S2Line& m = v.raw[i];
        if (what == S2_VEC || what == S2_VEC_UI) return -1;
        if (list == 0) {
            if (m.cls != 1) return -1;
            s2_cell_flt(&r.c[0], m.f); s2_cell_chr(&r.c[1], 'D'); s2_cell_int(&r.c[2], m.i[0]);
            ng.na.kind = 1; ng.na.a = m.i[0];
        } else {
            if (m.cls != 2 || m.link < 0) return -1;
            const S2Line& d = v.raw[m.link];
            const double lat = d_sub(m.f, d.f);
            s2_cell_flt(&r.c[0], d.f); s2_cell_chr(&r.c[1], 'C'); s2_cell_flt(&r.c[2], lat);
            ng.na.kind = 2; ng.na.f = lat;
        }
        s2_cell_span(&r.c[3], m.sp[0]); s2_cell_span(&r.c[9], m.sp[1]); s2_cell_span(&r.c[10], m.sp[1]); s2_cell_span(&r.c[12], m.sp[0]);
    } else if (v.table == S2_STRACE) {
        const S2Line& m = v.raw[i];
        if (list || what == S2_VEC || what == S2_VEC_UI || m.cls != 1) return -1;
        s2_cell_flt(&r.c[0], m.f);
        if (v.dur_all_int) s2_cell_int(&r.c[2], 0); else s2_cell_flt(&r.c[2], m.g);
        s2_cell_int(&r.c[9], v.pid); s2_cell_int(&r.c[10], m.i[0]);
        if (v.cat_is_str) s2_cell_span(&r.c[12], v.cat_span); else s2_cell_int(&r.c[12], v.cat);
        ng.na.kind = 3; ng.na.s = m.sp[0];
    } else {
        const S2Line& m = v.raw[i];
        if (list || what == S2_VEC || what == S2_VEC_UI || m.cls != 1) return -1;
        // kept pairs 0 .. m-1: the last one is deleted (max key), the first one is lost by list_to_csv_and_traces
        if (m.link < 1 || (uint64_t)m.link + 1 >= v.n_kept) return -1;
        const S2Line& nx = v.raw[v.kept[m.link + 1]];
        s2_cell_flt(&r.c[0], v.absolute ? m.f : d_sub(m.f, v.time_base));
        s2_cell_flt(&r.c[2], d_sub(nx.f, m.f));
        if (v.cat_is_str) s2_cell_span(&r.c[12], v.cat_span); else s2_cell_int(&r.c[12], v.cat);
        ng.na.kind = 4; ng.na.s = m.sp[0];
    }
    return json ? s2_json13(k, v.text, r, ng) : s2_csv13(k, v.text, r, ng);
}

}  // namespace sofa
