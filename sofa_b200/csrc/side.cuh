// Recorder side files -> 13-column rows: mpstat.txt, nvsmi_query.txt,
// netstat.txt as sofa_preprocess.py:482-546, 1013-1084, 1235-1337 turn them into mpstat.csv /
// usr_sys.csv, nvsmi_trace.csv, netstat.csv / netbandwidth.csv and the mpstat_usr, nvsmi_mem,
// nvsmi_sm, tx_bandwidth, rx_bandwidth series of report.js.
//
// Row-level functions only (__host__ __device__): the kernels in kernels_side.cuh and the CPU
// test harness (tests/hostsim.cpp) call the same code.
#pragma once
#include "rows.cuh"

namespace sofa {

enum { RAW_MPSTAT = 0, RAW_NVSMI = 1, RAW_NETSTAT = 2, RAW_VMSTAT = 3, RAW_DMON = 4 };
constexpr int RAW_NF = 17;           // doubles kept per parsed line (vmstat has 17 columns)

// One parsed line of a side file.
//   mpstat : f[0..8] = the nine float fields, nf = number of ',' separated fields (0: blank line)
//   nvsmi  : i[0] = seconds since 1970-01-01 of the naive stamp read as UTC, i[1] = microseconds,
//            i[2] = gpu index, i[3] = sm %, f[0] = mem %
//   netstat: f[0] = time, i[0] = tx bytes, i[1] = rx bytes
//   vmstat : f[0..16] = the 17 columns of `vmstat -w 1`, nf = 17 (0: header or short line, skipped)
//   dmon   : i[0] = gpu index, i[1] = enc %, i[2] = dec % of `nvidia-smi dmon -s u`, nf = 5 (0: skipped)
struct RawRow { double f[RAW_NF]; int64_t i[4]; int32_t nf; int32_t flag; };

// ---------------------------------------------------------------------------- mpstat.txt
// np.genfromtxt(delimiter=',', invalid_raise=False) (:484): comments ('#') are cut, the line is
// stripped of " \r\n", blank lines vanish, rows whose field count differs from the first row's are
// dropped, each field goes through float() and becomes nan when that raises.
// parse_double said P_UNSUPPORTED: is the token something float() accepts (inf / nan spellings, digit
// groups with '_', more digits than the codec carries), or just text that float() rejects too?
template <typename S>
SOFA_HD bool float_token_is_valid_python(S s, int n) {
    int a = 0, b = n;
    while (a < b && is_space((unsigned char)s[a])) a++;
    while (b > a && is_space((unsigned char)s[b - 1])) b--;
    bool numeric = true;
    for (int i = a; i < b; i++) {
        char c = s[i];
        if (!((c >= '0' && c <= '9') || c == '.' || c == 'e' || c == 'E' || c == '+' || c == '-' || c == '_')) numeric = false;
    }
    if (numeric) return true;                      // conservatively "maybe valid": the caller refuses
    if (a < b && (s[a] == '+' || s[a] == '-')) a++;
    const char* words[3] = {"inf", "infinity", "nan"};
    for (int w = 0; w < 3; w++) {
        int L = 0; while (words[w][L]) L++;
        if (b - a != L) continue;
        bool eq = true;
        for (int i = 0; i < L; i++) { char c = s[a + i]; if (c >= 'A' && c <= 'Z') c = (char)(c + 32); if (c != words[w][i]) eq = false; }
        if (eq) return true;
    }
    return false;
}

template <typename S>
SOFA_HD int parse_raw_mpstat(S s, int n, RawRow* r) {
    for (int k = 0; k < 9; k++) r->f[k] = 0;
    r->flag = 0;
    int a = 0, b = n;
    for (int i = 0; i < n; i++) if (s[i] == '#') return E_UNSUPPORTED_NUM;      // genfromtxt comment syntax
    while (a < b && s[a] == ' ') a++;
    while (b > a && s[b - 1] == ' ') b--;
    if (a >= b) { r->nf = 0; return E_OK; }
    int f = 0, st = a;
    for (int i = a; i <= b; i++) {
        if (i == b || s[i] == ',') {
            if (f < 9) {
                double v;
                int ps = parse_double(s + st, i - st, &v);
                if (ps == P_UNSUPPORTED && float_token_is_valid_python(s + st, i - st)) return E_UNSUPPORTED_NUM;
                r->f[f] = ps == P_OK ? v : bits_d(0x7ff8000000000000ull);
            }
            f++; st = i + 1;
        }
    }
    r->nf = f;
    return E_OK;
}

struct SideOutRow {
    const double* ext;       // vmstat: the line's 17 parsed columns (durations and name ints are derived on the fly)
    double ts, f1, f2;       // mpstat: f1 = cpu_time (trace row) or usr (usr_sys row), f2 = sys; netstat: tx, rx bandwidth
    int64_t i[6];            // mpstat: core, usr, sys, idl, iow, irq (the %3d ints); nvsmi: id, sm, mem; netstat: int(tx), int(rx)
    int cls;                 // 0 dropped, 1 trace row, 2 usr_sys.csv row
};

SOFA_HD bool d_isnan(double x) { return (d_bits(x) & 0x7fffffffffffffffull) > 0x7ff0000000000000ull; }
SOFA_HD bool d_isfinite(double x) { return ((d_bits(x) >> 52) & 0x7ff) != 0x7ff; }
// int(x) of a finite float that fits int64 (the '%d' % float conversion)
SOFA_HD int trunc_i64(double x, int64_t* out) {
    if (!d_isfinite(x)) return E_FORMAT;
    if (x >= 9223372036854775808.0 || x <= -9223372036854775808.0) return E_RANGE;
    *out = (int64_t)x;
    return E_OK;
}

// Row i of mpstat[1:] against row i - stride (:498-543).  cur/prev: nine floats each.
SOFA_HD int side_mpstat_row(const double* cur, const double* prev, int absolute, double time_base, SideOutRow* o) {
    double d[9];
    for (int k = 0; k < 9; k++) d[k] = d_sub(cur[k], prev[k]);
    if (d_isnan(d[2]) || d_isnan(d[4]) || d_isnan(d[5]) || d_isnan(d[6]) || d_isnan(d[7])) return E_FORMAT;   // '%3d' % nan raises
    double total = d_add(d_add(d_add(d_add(d[2], d[4]), d[5]), d[6]), d[7]);
    o->cls = 0; o->ext = nullptr;
    if (total == 0.0) return E_OK;                                           // "No increases in mpstat values"
    double usr = d_div(d_mul(d[2], 100.0), total), sys = d_div(d_mul(d[4], 100.0), total);
    double idl = d_div(d_mul(d[5], 100.0), total), iow = d_div(d_mul(d[6], 100.0), total);
    double irq = d_div(d_mul(d[7], 100.0), total);
    o->ts = absolute ? cur[0] : d_sub(cur[0], time_base);
    if (cur[1] != -1.0) {
        o->cls = 1;
        o->f1 = d_mul(d_sub(total, d[5]), 0.01);                             // cpu_time
        o->f2 = 0;
        int e = trunc_i64(cur[1], &o->i[0]); if (e) return e;
        if ((double)o->i[0] != cur[1]) return E_UNSUPPORTED_NUM;             // deviceId column would not be integral
        if (o->i[0] < -2147483647 || o->i[0] > 2147483647) return E_RANGE;
        const double pct[5] = {usr, sys, idl, iow, irq};
        for (int k = 0; k < 5; k++) {
            e = trunc_i64(pct[k], &o->i[1 + k]); if (e) return e;
            if (o->i[1 + k] < -1000000 || o->i[1 + k] > 1000000) return E_RANGE;
        }
    } else {
        o->cls = 2;
        o->f1 = usr; o->f2 = sys;
        if (!d_isfinite(usr) || !d_isfinite(sys)) return E_FORMAT;
    }
    return E_OK;
}

// ---------------------------------------------------------------------------- nvsmi_query.txt
// "2020/09/13 12:26:40.513, Tesla V100-SXM2-16GB, 0, 23 %, 38 %" (:1030-1035).  The stamp must be
// the fixed-width form nvidia-smi writes (strptime would also take 1-digit fields: refused here).
SOFA_HD int64_t days_from_civil(int y, int m, int d) {
    y -= m <= 2;
    const int64_t era = (y >= 0 ? y : y - 399) / 400;
    const int yoe = (int)(y - era * 400);
    const int doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const int doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (int64_t)doe - 719468;
}
template <typename S>
SOFA_HD int two_digits(S s, int at) {
    unsigned a = (unsigned char)s[at] - '0', b = (unsigned char)s[at + 1] - '0';
    return (a > 9 || b > 9) ? -1 : (int)(a * 10 + b);
}
// has_nl: the line ended with '\n' (fields[-1] then carries it, and [:-2] removes "%\n")
template <typename S>
SOFA_HD int parse_raw_nvsmi(S s, int n, bool has_nl, RawRow* r) {
    r->nf = 0; r->flag = 0;
    for (int i = 0; i + 6 <= n; i++)
        if ((s[i] == 'f' || s[i] == 'F') && s[i + 1] == 'a' && s[i + 2] == 'i' && s[i + 3] == 'l' && s[i + 4] == 'e' && s[i + 5] == 'd')
            r->flag = 1;                                                      // :1017-1021 "No nvsmi query data."
    int ca[5], nc = 0;
    for (int i = 0; i < n && nc < 5; i++) if (s[i] == ',') ca[nc++] = i;
    if (nc < 4) return r->flag ? E_OK : E_FEW_TOKENS;
    if (r->flag) return E_OK;
    // fields[0]
    int L = ca[0];
    if (L < 21 || L > 26) return E_UNSUPPORTED_NUM;
    if (s[4] != '/' || s[7] != '/' || s[10] != ' ' || s[13] != ':' || s[16] != ':' || s[19] != '.') return E_UNSUPPORTED_NUM;
    int y1 = two_digits(s, 0), y2 = two_digits(s, 2), mo = two_digits(s, 5), dd = two_digits(s, 8);
    int hh = two_digits(s, 11), mi = two_digits(s, 14), ss = two_digits(s, 17);
    if (y1 < 0 || y2 < 0 || mo < 0 || dd < 0 || hh < 0 || mi < 0 || ss < 0) return E_UNSUPPORTED_NUM;
    int year = y1 * 100 + y2;
    if (year < 1971 || year > 2200) return E_UNSUPPORTED_NUM;
    if (mo < 1 || mo > 12 || hh > 23 || mi > 59 || ss > 59) return E_FORMAT;   // strptime / datetime() ValueError
    const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    bool leap = (year % 4 == 0 && year % 100 != 0) || year % 400 == 0;
    if (dd < 1 || dd > mdays[mo - 1] + ((mo == 2 && leap) ? 1 : 0)) return E_FORMAT;
    int64_t us = 0; int nd = 0;
    for (int i = 20; i < L; i++) { unsigned c = (unsigned char)s[i] - '0'; if (c > 9) return E_UNSUPPORTED_NUM; us = us * 10 + c; nd++; }
    for (; nd < 6; nd++) us *= 10;                                            // %f right-pads to microseconds
    r->i[0] = days_from_civil(year, mo, dd) * 86400 + hh * 3600 + mi * 60 + ss;
    r->i[1] = us;
    // fields[2], fields[3][:-2], fields[4][:-2]
    int64_t v;
    int st = parse_i64(s + ca[1] + 1, ca[2] - ca[1] - 1, &v);
    if (st) return num_status_to_err(st, false);
    if (v < -2147483647 || v > 2147483647) return E_RANGE;
    r->i[2] = v;
    int l3 = ca[3] - ca[2] - 1 - 2;
    st = parse_i64(s + ca[2] + 1, l3 < 0 ? 0 : l3, &v);
    if (st) return num_status_to_err(st, false);
    r->i[3] = v;
    int e4 = nc >= 5 ? ca[4] : n + (has_nl ? 1 : 0);
    int l4 = e4 - ca[3] - 1 - 2;
    if (l4 > n - ca[3] - 1) l4 = n - ca[3] - 1;
    st = parse_i64(s + ca[3] + 1, l4 < 0 ? 0 : l4, &v);
    if (st) return num_status_to_err(st, false);
    r->f[0] = (double)v;                 // mem (kept as an exact small integer)
    if (v < -1000000000000ll || v > 1000000000000ll || r->i[3] < -1000000000000ll || r->i[3] > 1000000000000ll) return E_RANGE;
    r->nf = 5;
    return E_OK;
}
// local_off: seconds the host's local zone is ahead of UTC (datetime.timestamp() of a naive
// datetime reads it as local time); tz_add = cfg.nvsmi_time_zone * 3600 (:1032)
SOFA_HD void side_nvsmi_row(const RawRow& r, int64_t local_off, double tz_add, int absolute, double time_base, SideOutRow* o) {
    double t = d_add((double)(r.i[0] - local_off), d_div((double)r.i[1], 1e6));
    t = d_add(t, tz_add);
    o->ts = absolute ? t : d_sub(t, time_base);
    o->ext = nullptr; o->f1 = 0; o->f2 = 0;
    o->i[0] = r.i[2]; o->i[1] = r.i[3]; o->i[2] = (int64_t)r.f[0];
    o->i[3] = o->i[4] = o->i[5] = 0;
    o->cls = 1;
}

// ---------------------------------------------------------------------------- netstat.txt
// "time,tx_bytes,rx_bytes" (:1250-1252)
template <typename S>
SOFA_HD int parse_raw_netstat(S s, int n, RawRow* r) {
    r->nf = 0; r->flag = 0;
    int ca[3], nc = 0;
    for (int i = 0; i < n && nc < 3; i++) if (s[i] == ',') ca[nc++] = i;
    if (nc < 2) return E_FEW_TOKENS;
    int st = parse_double(s, ca[0], &r->f[0]);
    if (st) return num_status_to_err(st, true);
    st = parse_i64(s + ca[0] + 1, ca[1] - ca[0] - 1, &r->i[0]);
    if (st) return num_status_to_err(st, false);
    int e2 = nc >= 3 ? ca[2] : n;
    st = parse_i64(s + ca[1] + 1, e2 - ca[1] - 1, &r->i[1]);
    if (st) return num_status_to_err(st, false);
    r->nf = 3;
    return E_OK;
}
SOFA_HD int side_netstat_row(const RawRow& cur, const RawRow& prev, int absolute, double time_base, SideOutRow* o) {
    double dt = d_sub(cur.f[0], prev.f[0]);
    if (dt == 0.0 || d_isnan(dt)) return E_FORMAT;                            // ZeroDivisionError in the reference
    __int128 dtx = (__int128)cur.i[0] - prev.i[0], drx = (__int128)cur.i[1] - prev.i[1];
    const __int128 lim = (__int128)1 << 62;
    if (dtx >= lim || dtx <= -lim || drx >= lim || drx <= -lim) return E_RANGE;
    o->f1 = d_div((double)(int64_t)dtx, dt);                                  // int / float: float(int) then IEEE division
    o->f2 = d_div((double)(int64_t)drx, dt);
    int e = trunc_i64(o->f1, &o->i[0]); if (e) return e;
    e = trunc_i64(o->f2, &o->i[1]); if (e) return e;
    o->i[2] = o->i[3] = o->i[4] = o->i[5] = 0;
    o->ts = absolute ? cur.f[0] : d_sub(cur.f[0], time_base);
    o->ext = nullptr; o->cls = 1;
    return E_OK;
}

// ---------------------------------------------------------------------------- vmstat.txt (:787-1008)
template <typename S>
SOFA_HD bool line_contains(S s, int n, const char* w) {
    int L = 0; while (w[L]) L++;
    for (int i = 0; i + L <= n; i++) {
        bool eq = true;
        for (int k = 0; k < L; k++) if (s[i + k] != w[k]) { eq = false; break; }
        if (eq) return true;
    }
    return false;
}
SOFA_HD int parse_raw_vmstat(const char* s, int n, RawRow* r) {
    r->nf = 0; r->flag = 0;
    if (line_contains(s, n, "procs") || line_contains(s, n, "swpd")) return E_OK;       // header lines (:810)
    Tok t[17];
    if (tokenize(s, n, t, 17) < 17) return E_OK;                                       // :813
    for (int q = 0; q < 17; q++) {
        int st = parse_double(s + t[q].a, t[q].b - t[q].a, &r->f[q]);
        if (st) return num_status_to_err(st, true);
    }
    r->nf = 17;
    return E_OK;
}
// accepted line number j (the reference's counter t); validates every value the name prints
SOFA_HD int side_vmstat_row(int64_t j, const double* ext, int absolute, double time_base, SideOutRow* o) {
    for (int q = 0; q < 17; q++) { int64_t v; int e = trunc_i64(d_add(ext[q], 1e-5), &v); if (e) return e; }
    o->ext = ext; o->f1 = 0; o->f2 = 0; o->cls = 1;
    o->i[0] = j; o->i[1] = o->i[2] = o->i[3] = o->i[4] = o->i[5] = 0;
    o->ts = absolute ? d_add((double)j, time_base) : (double)j;
    return E_OK;
}
// list order of vmstat.csv (bi, bo, in, cs, wa, st, usr, sys: :994-1008) -> column of `vmstat -w`
SOFA_HD int vmstat_list_col(int list) { const int c[8] = {8, 9, 10, 11, 15, 16, 12, 13}; return c[list & 7]; }

// ---------------------------------------------------------------------------- nvsmi.txt, `nvidia-smi dmon -s u` (:1092-1183)
SOFA_HD int parse_raw_dmon(const char* s, int n, RawRow* r) {
    r->nf = 0; r->flag = 0;
    if (line_contains(s, n, "failed") || line_contains(s, n, "Failed")) r->flag = 1;
    if (line_contains(s, n, "gpu") || line_contains(s, n, "Idx")) return E_OK;            // header lines (:1108)
    Tok t[5];
    if (tokenize(s, n, t, 5) < 5) return E_OK;
    int st = parse_i64(s + t[0].a, t[0].b - t[0].a, &r->i[0]);
    if (st) return num_status_to_err(st, false);
    for (int q = 0; q < 2; q++) {
        const Tok& k = t[3 + q];
        if (k.b - k.a == 1 && s[k.a] == '-') { r->i[1 + q] = 0; continue; }
        st = parse_i64(s + k.a, k.b - k.a, &r->i[1 + q]);
        if (st) return num_status_to_err(st, false);
    }
    if (r->i[0] < -2147483647 || r->i[0] > 2147483647) return E_RANGE;
    r->nf = 5;
    return E_OK;
}
// t = number of earlier accepted lines whose gpu index is 0 (the reference's counter, :1172-1173)
SOFA_HD void side_dmon_row(int64_t t, const RawRow& r, int absolute, double time_base, SideOutRow* o) {
    o->ext = nullptr; o->f1 = 0; o->f2 = 0; o->cls = t > 3 ? 1 : 0;                      // rows of the first seconds are dropped
    o->i[0] = r.i[0]; o->i[1] = r.i[1]; o->i[2] = r.i[2]; o->i[3] = t; o->i[4] = o->i[5] = 0;
    o->ts = absolute ? d_add((double)t, time_base) : (double)t;
}

// ---------------------------------------------------------------------------- 13-column view
// list: nvsmi 0 = sm, 1 = mem; netstat 0 = tx, 1 = rx; vmstat 0..7 = bi, bo, in, cs, wa, st, usr, sys; dmon 0 = enc, 1 = dec
SOFA_HD RowVals vals_side(int table, int list, int absolute, const SideOutRow& r) {
    RowVals v; v.valid = true; v.ts = r.ts; v.ts_is_f = true; v.ts_i = 0; v.dur = 0; v.dur_is_f = false; v.dur_i = -1; v.cat = 0;
    v.ev_is_f = false; v.ev_f = 0; v.ev_i = -1;
    v.dev = -1; v.kind = -1; v.payload = -1; v.bw_is_f = false; v.bw_f = 0; v.bw_i = -1;
    v.src = -1; v.dst = -1; v.pid = -1; v.tid = -1;
    if (table == RAW_MPSTAT) { v.dur_is_f = true; v.dur = r.f1; v.dev = r.i[0]; }
    else if (table == RAW_NVSMI) { v.ev_i = list; v.dur_i = list ? r.i[2] : r.i[1]; v.dev = r.i[0]; v.cat = r.i[0]; }
    else if (table == RAW_NETSTAT) { v.ev_i = list; v.bw_is_f = true; v.bw_f = list ? r.f2 : r.f1; }
    else if (table == RAW_VMSTAT) {
        v.ts_is_f = absolute != 0; v.ts_i = r.i[0]; v.dur_is_f = true; v.dur = d_add(r.ext[vmstat_list_col(list)], 1e-5); v.cat = -1;
    } else {
        v.ts_is_f = absolute != 0; v.ts_i = r.i[3]; v.ev_i = 2 + list; v.dur_i = list ? r.i[2] : r.i[1]; v.dev = r.i[0]; v.cat = r.i[0];
    }
    return v;
}
// float-column mask of the file (pandas dtype of each DataFrame column)
SOFA_HD uint32_t side_fmask(int table, int absolute) {
    if (table == RAW_MPSTAT) return 0xDu;
    if (table == RAW_NVSMI) return 0x1u;
    if (table == RAW_NETSTAT) return 0x41u;
    if (table == RAW_VMSTAT) return 0x4u | (absolute ? 1u : 0u);
    return absolute ? 1u : 0u;
}

template <class K> SOFA_HD void put_pad3(K& k, int64_t v) {                   // '%3d'
    int l = len_i64(v);
    for (; l < 3; l++) k.put(' ');
    put_i64(k, v);
}
struct SideName {
    int table, half; SideOutRow r;
    template <class K> SOFA_HD void gen(K& k) const {
        if (table == RAW_MPSTAT) {                                             // :524
            put_lit(k, "mpstat_core"); put_i64(k, r.i[0]); put_lit(k, " (usr|sys|idl|iow|irq): |");
            for (int q = 1; q <= 5; q++) { put_pad3(k, r.i[q]); k.put('|'); }
        } else if (table == RAW_NVSMI) {                                       // :1046-1047
            put_cstr(k, half ? "GPUID_mem=" : "GPUID_sm="); put_i64(k, r.i[0]); k.put('_'); put_i64(k, half ? r.i[2] : r.i[1]);
        } else if (table == RAW_NETSTAT) {                                     // :1287, 1304
            put_cstr(k, half ? "network_bandwidth_rx(bytes):" : "network_bandwidth_tx(bytes):"); put_i64(k, half ? r.i[1] : r.i[0]);
        } else if (table == RAW_VMSTAT) {                                      // :850-866
            const char* keys[17] = {"r=", "|b=", "|sw=", "|fr=", "|bu=", "|ca=", "|si=", "|so=", "|bi=", "|bo=", "|in=", "|cs=",
                                    "|usr=", "|sys=", "|idl=", "|wa=", "|st="};
            for (int q = 0; q < 17; q++) { put_cstr(k, keys[q]); put_i64(k, (int64_t)d_add(r.ext[q], 1e-5)); }
        } else {                                                               // :1132-1133
            put_cstr(k, half ? "GPUID_dec=" : "GPUID_enc="); put_i64(k, r.i[0]); k.put('_'); put_i64(k, half ? r.i[2] : r.i[1]);
        }
    }
};

template <> struct JsonYIsBw<SideName> { static SOFA_HD bool get(const SideName& ng) { return ng.table == RAW_NETSTAT; } };

// ---------------------------------------------------------------------------- usr_sys.csv / netbandwidth.csv rows
// "a,b,c\n" with pandas' default float text (repr) - :545, :1323
struct AuxRowPlan { int len; JPlan f[3]; };
SOFA_HD AuxRowPlan aux_row_plan(double a, double b, double c) {
    AuxRowPlan p;
    p.f[0] = repr_plan(a); p.f[1] = repr_plan(b); p.f[2] = repr_plan(c);
    bool bad = p.f[0].kind == JP_BAD || p.f[1].kind == JP_BAD || p.f[2].kind == JP_BAD;
    p.len = bad ? -1 : p.f[0].len + p.f[1].len + p.f[2].len + 3;
    return p;
}
template <typename O>
SOFA_HD int aux_row_emit(const AuxRowPlan& p, O out) {
    int n = 0;
    for (int k = 0; k < 3; k++) { n += repr_emit(p.f[k], out + n); out[n++] = k == 2 ? '\n' : ','; }
    return n;
}

}  // namespace sofa
