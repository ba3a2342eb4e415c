// CPU unit-test harness (TEST INFRASTRUCTURE): compiles the __host__ __device__ codecs of
// sofa_b200/csrc/*.cuh with g++ so their exact arithmetic can be checked against
// Python/numpy/pandas in the build container.  Never loaded by the product path.
#include <string.h>
#define SOFA_HOST_WIDE 1   // exercise the 8-byte-word sink paths (text is copied into a padded buffer)
#include "../sofa_b200/csrc/num.cuh"
using namespace sofa;
extern "C" {
int hs_parse_double(const char* s, int n, double* out) { return parse_double(s, n, out); }
int hs_parse_i64(const char* s, int n, long long* out) { int64_t v = 0; int r = parse_i64(s, n, &v); *out = v; return r; }
int hs_fmt_fixed6(double x, char* out) { return fmt_fixed6(x, out); }
int hs_len_fixed6(double x) { return len_fixed6(x); }
int hs_fmt_i64(long long v, char* out) { return fmt_i64((int64_t)v, out); }
int hs_json_double(double x, char* out) { return json_double(x, out); }
double hs_log10(double x) { return log10_cr(x); }
double hs_log2(double x) { return log2_cr(x); }
double hs_interp(double x, const double* xp, const double* fp, int n) { return interp1(x, xp, fp, n); }
double hs_ip_event_raw(unsigned long long ip, int h) { return ip_event_raw(ip, h); }
void hs_log10_many(const double* x, double* y, long n) { for (long i = 0; i < n; i++) y[i] = log10_cr(x[i]); }
void hs_log2_many(const double* x, double* y, long n) { for (long i = 0; i < n; i++) y[i] = log2_cr(x[i]); }
// batch helpers: strings are packed back to back, off[i]..off[i+1]
void hs_json_many(const double* x, long n, char* out, long* off) {
    long o = 0;
    for (long i = 0; i < n; i++) { off[i] = o; int l = json_double(x[i], out + o); JPlan p = json_plan(x[i]); if (l > 0 && p.len == l) o += l; else { out[o++] = '?'; } }
    off[n] = o;
}
void hs_fixed6_many(const double* x, long n, char* out, long* off) {
    long o = 0;
    for (long i = 0; i < n; i++) { off[i] = o; int l = fmt_fixed6(x[i], out + o); if (l > 0 && len_fixed6(x[i]) == l) o += l; else { out[o++] = '?'; } }
    off[n] = o;
}
}

// ---------------------------------------------------------------- row-level harness
#include "../sofa_b200/csrc/rows.cuh"
#include "../sofa_b200/csrc/side.cuh"
#include <vector>
#include <string>
struct HsCfg {
    double time_base, t_offset; int absolute_timestamp, perf_has_cycles, n_mhz;
    const double* mhz_xp; const double* mhz_fp;
    double ts_rescale, dt_rescale, payload_unit; int n_cudaproc;
    int idx[9];
    int api_mode;
};
static IngestCfg mk(const HsCfg* h) {
    IngestCfg c; c.time_base = h->time_base; c.t_offset = h->t_offset; c.absolute_timestamp = h->absolute_timestamp;
    c.perf_has_cycles = h->perf_has_cycles; c.n_mhz = h->n_mhz; c.mhz_xp = h->mhz_xp; c.mhz_fp = h->mhz_fp;
    c.ts_rescale = h->ts_rescale; c.dt_rescale = h->dt_rescale; c.payload_unit = h->payload_unit; c.n_cudaproc = h->n_cudaproc;
    c.idx_start = h->idx[0]; c.idx_dur = h->idx[1]; c.idx_size = h->idx[2]; c.idx_thr = h->idx[3]; c.idx_ctx = h->idx[4];
    c.idx_stream = h->idx[5]; c.idx_name = h->idx[6]; c.idx_srcctx = h->idx[7]; c.idx_dstctx = h->idx[8];
    c.api_mode = h->api_mode;
    return c;
}
// demangle table for viz names: sorted FNV hashes -> strings
struct HsDict { int n; const unsigned long long* hash; const long* off; const char* blob; };
struct DictVizName {
    PerfVizName<const char*> raw; const HsDict* d;
    template <class K> void gen(K& k) const {
        HashSink h; raw.gen(h);
        if (h.len >= 2 && h.c0 == '_' && h.c1 == 'Z' && d) {
            for (int i = 0; i < d->n; i++) if (d->hash[i] == h.h) { for (long j = d->off[i]; j < d->off[i + 1]; j++) k.put(d->blob[j]); return; }
        }
        raw.gen(k);
    }
};
extern "C" {
// mode 0: CSV of all rows (stride applies to net only, as in the reference); mode 1: JSON records of
// rows i%stride==0 (perf: viz names); mode 2: JSON records of rows whose name contains `kw`
// (perf: on the down-sampled viz frame; gpu: on the full frame).  Returns -1 or (1000*row + err).
long hs_process(int kind, const char* text, long nbytes, const HsCfg* h, int mode, int stride, unsigned fmask,
                const char* kw, const HsDict* dict, char* out, long* outlen, unsigned* flags, long* nrec) {
    IngestCfg c = mk(h);
    std::vector<char> padded((size_t)nbytes + 64, '\n');
    memcpy(padded.data() + 16, text, (size_t)nbytes);
    text = padded.data() + 16;
    long o = 0, row = 0, pos = 0, cnt = 0;
    unsigned fl = 0;
    FilterTable ft; memset(&ft, 0, sizeof(ft));
    if (kw) { size_t L = strlen(kw); const char* ks[1] = {kw}; filter_table_build(&ft, 1, ks, &L); }
    while (pos < nbytes) {
        long e = pos; while (e < nbytes && text[e] != '\n') e++;
        long linelen = e - pos;                    // without '\n'
        long with_nl = (e < nbytes) ? linelen + 1 : linelen;
        const char* line = text + pos;
        int err = 0, len = 0;
        bool sel = (row % stride) == 0;
        if (kind == SRC_PERF) {
            PerfRow r; err = parse_perf(line, (int)linelen, c, &r);
            if (!err) {
                RowVals v = vals_perf(r);
                PerfName<const char*> ng{line, r, true};
                if (mode == 0) { bool sp; int l = csv_row_len(v, fmask, ng, &sp); if (l < 0) err = E_FORMAT; else { len = csv_row_write(v, fmask, ng, sp, out + o); if (len != l) err = 99; } }
                else if (sel) {
                    DictVizName vz{PerfVizName<const char*>{ng}, dict};
                    bool take = true;
                    if (mode == 2) { MatchSink ms(&ft); vz.gen(ms); take = ms.hits() != 0; }
                    if (take) { int l = json_row_len(v, fmask, vz); if (l < 0) err = E_FORMAT; else { len = json_row_write(v, fmask, vz, out + o); if (len != l) err = 99; } }
                }
            }
        } else if (kind == SRC_GPU) {
            GpuRow r; err = parse_gpu(line, (int)with_nl, c, &r);
            if (!err) {
                if (sel) fl |= r.bw_is_float ? 1u : 0u;
                RowVals v = vals_gpu(r, c.n_cudaproc);
                GpuName<const char*> ng{line, r, true, c.api_mode != 0};
                if (mode == 0) { bool sp; int l = csv_row_len(v, fmask, ng, &sp); if (l < 0) err = E_FORMAT; else { len = csv_row_write(v, fmask, ng, sp, out + o); if (len != l) err = 99; } }
                else {
                    bool take = (mode == 1) ? sel : true;
                    if (mode == 2) { MatchSink ms(&ft); ng.gen(ms); take = ms.hits() != 0; }
                    if (take) { int l = json_row_len(v, fmask, ng); if (l < 0) err = E_FORMAT; else { len = json_row_write(v, fmask, ng, out + o); if (len != l) err = 99; } }
                }
            }
        } else {
            NetRow r; err = parse_net(line, (int)linelen, c, &r);
            if (!err) {
                if (sel) fl |= r.proto == 0 ? 2u : 4u;
                RowVals v = vals_net(r);
                NetName ng{r};
                if (sel) {
                    if (mode == 0) { bool sp; int l = csv_row_len(v, fmask, ng, &sp); if (l < 0) err = E_FORMAT; else { len = csv_row_write(v, fmask, ng, sp, out + o); if (len != l) err = 99; } }
                    else { int l = json_row_len(v, fmask, ng); if (l < 0) err = E_FORMAT; else { len = json_row_write(v, fmask, ng, out + o); if (len != l) err = 99; } }
                }
            }
        }
        if (err) return row * 1000 + err;
        o += len; row++; if (len > 0) cnt++;
        pos = e + 1;
    }
    *outlen = o; *flags = fl; *nrec = cnt;
    return -1;
}
unsigned long long hs_fnv(const char* s, long n) { HashSink h; for (long i = 0; i < n; i++) h.put(s[i]); return h.h; }
}

// ---------------------------------------------------------------- synthetic generator spec
#include "../sofa_b200/csrc/synth.cuh"
extern "C" long hs_synth(int src, unsigned long long seed, unsigned long long i0, unsigned long long n, int variant, char* out) {
    RawWrite<char*> w{out, 0};
    for (unsigned long long i = i0; i < i0 + n; i++) synth_line(w, src, seed, i, variant);
    return w.n;
}

// ---------------------------------------------------------------- analyze primitives
extern "C" {
double hs_pairwise(const double* a, unsigned long long n) { return pairwise_sum(a, n); }
double hs_pairwise_par(const double* a, unsigned long long n, int depth) {
    // emulates the two-kernel device evaluation: 2^depth sub-trees, then the combine sweep
    std::vector<double> part((size_t)1 << depth);
    for (uint32_t t = 0; t < (1u << depth); t++) {
        uint64_t s, l; pairwise_subtree(n, depth, t, &s, &l);
        double v = pairwise_sum(a + s, l);
        part[t] = v;
    }
    for (int lev = depth - 1; lev >= 0; lev--)
        for (uint32_t i = 0; i < (1u << lev); i++) part[i] = part[2 * i] + part[2 * i + 1];
    return 0.0 + part[0];
}
double hs_round6(double x) { return round6_value(x); }
}

// ---------------------------------------------------------------- tokenizer fuzz hook
extern "C" int hs_tokenize(const char* s, int n, int* ab, int maxt) {
    Tok t[16];
    int k = tokenize(s, n, t, maxt > 16 ? 16 : maxt);
    for (int i = 0; i < k; i++) { ab[2 * i] = t[i].a; ab[2 * i + 1] = t[i].b; }
    return k;
}

extern "C" int hs_repr6(double x, char* out) { JPlan p = repr6_plan(x); if (p.kind == JP_BAD) return -1; int l = json_emit(p, out); return l == p.len ? l : -2; }

// ---------------------------------------------------------------- recorder side files (side.cuh)
// CPU mirror of sofa_b200_side_preprocess + the side emitters: same row functions, the indexing
// (which line a row is differenced against, which rows survive) restated sequentially.
// mode = 16 * fmt + list; fmt 0 CSV, 1 JSON records, 2 aux rows (usr_sys.csv / netbandwidth.csv).
// Returns -1, or row*1000+err; *nodata = 1 when nvidia-smi reported a failure.
extern "C" long hs_side(int table, const char* text, long nbytes, int absolute, double time_base, long long local_off,
                        double tz_add, int mode, char* out, long* outlen, long* nrec, long* nlines, long* ncand_out, int* nodata) {
    std::vector<char> padded((size_t)nbytes + 64, '\n');
    memcpy(padded.data() + 16, text, (size_t)nbytes);
    text = padded.data() + 16;
    std::vector<RawRow> raw;
    long pos = 0, row = 0, first_err = -1;
    bool failed = false;
    while (pos < nbytes) {
        long e = pos; while (e < nbytes && text[e] != '\n') e++;
        bool has_nl = e < nbytes;
        RawRow r = RawRow();
        int err;
        if (table == RAW_MPSTAT) err = parse_raw_mpstat(text + pos, (int)(e - pos), &r);
        else if (table == RAW_NVSMI) err = parse_raw_nvsmi(text + pos, (int)(e - pos), has_nl, &r);
        else if (table == RAW_NETSTAT) err = parse_raw_netstat(text + pos, (int)(e - pos), &r);
        else if (table == RAW_VMSTAT) err = parse_raw_vmstat(text + pos, (int)(e - pos), &r);
        else err = parse_raw_dmon(text + pos, (int)(e - pos), &r);
        if (err && first_err < 0) first_err = row * 1000 + err;
        if (r.flag) failed = true;
        raw.push_back(r);
        row++; pos = e + 1;
    }
    *nlines = row; *nodata = 0; *outlen = 0; *nrec = 0; *ncand_out = 0;
    if ((table == RAW_NVSMI || table == RAW_DMON) && failed) { *nodata = 1; return -1; }
    if (first_err >= 0) return first_err;
    std::vector<SideOutRow> rows;
    if (table == RAW_MPSTAT) {
        std::vector<long> idx;
        long first = -1;
        for (long i = 0; i < (long)raw.size(); i++) if (raw[i].nf > 0) { first = i; break; }
        for (long i = 0; i < (long)raw.size(); i++) if (first >= 0 && raw[i].nf == raw[first].nf) idx.push_back(i);
        if (idx.size() < 2) return 0 * 1000 + E_FEW_TOKENS;
        double mx = -1e300; bool nan = false;
        for (size_t j = 1; j < idx.size(); j++) { double v = raw[idx[j]].f[1]; if (d_isnan(v)) nan = true; else if (v > mx) mx = v; }
        if (nan) return 1 * 1000 + E_FORMAT;
        double nc = mx + 1.0;
        if (!(nc >= 0.0 && nc < 1e9)) return 1 * 1000 + E_RANGE;
        long stride = (long)nc + 1;
        for (long i = 0; i + 1 < (long)idx.size(); i++) {
            SideOutRow q; q.cls = 0;
            if (i > stride) {
                int e = side_mpstat_row(raw[idx[i + 1]].f, raw[idx[i + 1 - stride]].f, absolute, time_base, &q);
                if (e) return idx[i + 1] * 1000 + e;
            }
            rows.push_back(q);
        }
    } else if (table == RAW_NVSMI) {
        for (auto& r : raw) { SideOutRow q; side_nvsmi_row(r, local_off, tz_add, absolute, time_base, &q); rows.push_back(q); }
    } else if (table == RAW_VMSTAT) {
        long j = 0;
        for (long i = 0; i < (long)raw.size(); i++) {
            if (raw[i].nf <= 0) continue;
            SideOutRow q;
            int e = side_vmstat_row(j++, raw[i].f, absolute, time_base, &q);
            if (e) return i * 1000 + e;
            rows.push_back(q);
        }
    } else if (table == RAW_DMON) {
        long t = 0;
        for (long i = 0; i < (long)raw.size(); i++) {
            if (raw[i].nf <= 0) continue;
            SideOutRow q;
            side_dmon_row(t, raw[i], absolute, time_base, &q);
            rows.push_back(q);
            if (raw[i].i[0] == 0) t++;
        }
    } else {
        for (long i = 0; i + 1 < (long)raw.size(); i++) {
            SideOutRow q;
            int e = side_netstat_row(raw[i + 1], raw[i], absolute, time_base, &q);
            if (e) return (i + 1) * 1000 + e;
            rows.push_back(q);
        }
    }
    *ncand_out = (long)rows.size();
    long o = 0, cnt = 0;
    const int fmt = mode >> 4, half = mode & 15;
    const uint32_t fmask = side_fmask(table, absolute);
    for (size_t i = 0; i < rows.size(); i++) {
        const SideOutRow& q = rows[i];
        if (fmt == 2) {
            if (table == RAW_MPSTAT && q.cls != 2) continue;
            AuxRowPlan p = aux_row_plan(q.ts, q.f1, q.f2);
            if (p.len < 0) return (long)i * 1000 + E_FORMAT;
            int l = aux_row_emit(p, out + o);
            if (l != p.len) return (long)i * 1000 + 99;
            o += l; cnt++;
            continue;
        }
        if (q.cls != 1) continue;
        SideName ng{table, half, q};
        RowVals v = vals_side(table, half, absolute, q);
        int l, len;
        if (fmt == 0) { bool sp; l = csv_row_len(v, fmask, ng, &sp); if (l < 0) return (long)i * 1000 + E_FORMAT; len = csv_row_write(v, fmask, ng, sp, out + o); }
        else { l = json_row_len(v, fmask, ng); if (l < 0) return (long)i * 1000 + E_FORMAT; len = json_row_write(v, fmask, ng, out + o); }
        if (len != l) return (long)i * 1000 + 99;
        o += len; cnt++;
    }
    *outlen = o; *nrec = cnt;
    return -1;
}
extern "C" int hs_repr(double x, char* out) { JPlan p = repr_plan(x); if (p.kind == JP_BAD) return -1; int l = repr_emit(p, out); return l == p.len ? l : -2; }

// ---------------------------------------------------------------- diskstat / blktrace / strace / pystacks (side2.cuh)
// CPU mirror of sofa_b200_side2_preprocess + _set_devs + _emit: the same row functions; the indexing (line starts,
// device-name table, 'D' -> 'C' links, ranks) is restated sequentially here.
#include "../sofa_b200/csrc/side2.cuh"
#include <map>
struct HsSide2 {
    double time_base, strace_min_time; long long pid, category; const char* category_text;
    int absolute, plot_ratio, category_is_str, category_len;
    const int* secsize; const unsigned char* ui_keep; int n_dev_given;       // per device in order of first appearance
};
struct HsSide2Info { long long n_lines, n_entries, n_kept, n_devs; double total; unsigned last_a, last_b; int flags, pad; };
// returns -1 or (1000 * entry + err)
extern "C" long hs_side2(int table, const char* text_in, long nbytes, const HsSide2* p, int what, char* out, long* outlen, long* nrows,
                         HsSide2Info* info) {
    std::vector<char> buf((size_t)nbytes + 256, '\n');
    memcpy(buf.data(), text_in, (size_t)nbytes);
    const unsigned cat_a = (unsigned)nbytes + 128;
    if (p->category_is_str) memcpy(buf.data() + cat_a, p->category_text, (size_t)p->category_len);
    const char* text = buf.data();
    memset(info, 0, sizeof *info);
    *outlen = 0; *nrows = 0;
    std::vector<uint32_t> lstart; lstart.push_back(0);
    for (long i = 0; i < nbytes; i++) if (text[i] == '\n') lstart.push_back((uint32_t)i + 1);
    const bool last_has_nl = nbytes == 0 || text[nbytes - 1] == '\n';
    if (!last_has_nl) lstart.push_back((uint32_t)nbytes + 1);
    const uint64_t n_lines = lstart.size() - 1;
    if (table == S2_PYSTACKS && (n_lines & 1)) return (long)(n_lines - 1) * 1000 + E_FEW_TOKENS;
    const uint64_t n = table == S2_PYSTACKS ? n_lines / 2 : n_lines;
    info->n_lines = (long long)n_lines; info->n_entries = (long long)n;
    std::vector<S2Line> raw(n + 1);
    for (uint64_t i = 0; i < n; i++) {
        S2Line r = {};
        int e = E_OK;
        if (table == S2_PYSTACKS) {
            e = s2_parse_py(text, lstart[2 * i], (int)(lstart[2 * i + 1] - lstart[2 * i]) - 1, lstart[2 * i + 1],
                            (int)(lstart[2 * i + 2] - lstart[2 * i + 1]) - 1, &r);
        } else {
            const bool has_nl = (i + 1 != n_lines) || last_has_nl;
            const int len = (int)(lstart[i + 1] - lstart[i]) - 1;
            if (table == S2_DISK) e = s2_parse_disk(text, lstart[i], has_nl ? len : (len > 0 ? len - 1 : 0), &r);
            else if (table == S2_BLK) e = s2_parse_blk(text, lstart[i], len, has_nl, &r);
            else if (i % (uint64_t)p->plot_ratio) r.link = -1;
            else e = s2_parse_strace(text, lstart[i], len, p->pid, p->absolute, p->time_base, p->strace_min_time, &r);
        }
        if (e) return (long)i * 1000 + e;
        raw[i] = r;
    }
    S2View v; memset(&v, 0, sizeof v);
    v.table = table; v.text = text; v.nbytes = (uint64_t)nbytes; v.lstart = lstart.data(); v.n_lines = n_lines; v.last_has_nl = last_has_nl;
    v.raw = raw.data(); v.n = n; v.absolute = p->absolute; v.time_base = p->time_base;
    v.pid = p->pid; v.cat = p->category; v.cat_is_str = p->category_is_str; v.cat_span.a = cat_a; v.cat_span.b = cat_a + (unsigned)p->category_len;
    std::vector<uint32_t> kept;
    std::vector<int32_t> secsize; std::vector<uint8_t> dflags;
    if (table == S2_DISK) {
        std::vector<std::string> names;
        for (uint64_t i = 0; i < n; i++) {
            std::string nm(text + raw[i].sp[0].a, text + raw[i].sp[0].b);
            size_t k = 0;
            while (k < names.size() && names[k] != nm) k++;
            if (k == names.size()) names.push_back(nm);
            raw[i].link = (int32_t)k;
        }
        v.n_dev = (uint32_t)names.size(); info->n_devs = (long long)names.size();
        secsize.assign(names.size() + 1, 512); dflags.assign(names.size() + 1, 0);
        for (size_t k = 0; k < names.size() && (int)k < p->n_dev_given; k++) { secsize[k] = p->secsize[k]; dflags[k] = p->ui_keep[k] ? 2 : 0; }
        v.dev_secsize = secsize.data(); v.dev_flags = dflags.data();
        std::vector<double> mx(names.size() + 1, 0.0); std::vector<char> has(names.size() + 1, 0);
        for (uint64_t i = v.n_dev; i < n; i++) {
            DiskRow d;
            int e = s2_disk_row(raw[i], raw[i - v.n_dev], secsize[raw[i].link], p->absolute, p->time_base, &d);
            if (e) return (long)i * 1000 + e;
            int k = raw[i].link;
            if (!has[k] || d.tbw > mx[k]) mx[k] = d.tbw;
            has[k] = 1;
            v.disk_fl |= (d.rt_f ? 1u : 0u) | (d.wt_f ? 2u : 0u) | (d.at_f ? 4u : 0u);
        }
        for (size_t k = 0; k < names.size(); k++) if (has[k] && mx[k] == 0.0) dflags[k] |= 1;
    } else if (table == S2_BLK) {
        std::map<int64_t, int32_t> last_d;
        long last = -1;
        for (uint64_t i = 0; i < n; i++) {
            if (raw[i].cls) { last = (long)i; if (raw[i].flag & 1) info->flags |= 1; }
            if (raw[i].cls == 1) last_d[raw[i].i[0]] = (int32_t)i;
            if (raw[i].cls == 2) { auto it = last_d.find(raw[i].i[0]); raw[i].link = it == last_d.end() ? -1 : it->second; }
        }
        if (last >= 0) { info->last_a = raw[last].sp[0].a; info->last_b = raw[last].sp[0].b; }
    } else if (table == S2_STRACE) {
        double t = 0; unsigned f = 0;
        for (uint64_t i = 0; i < n; i++) {
            if (raw[i].cls) t = d_add(t, raw[i].g);
            if (raw[i].cls == 1) f |= (raw[i].flag & 1) ? 1u : 2u;
        }
        info->total = t;
        v.dur_all_int = (f & 1) ? 0 : 1;
    } else {
        for (uint64_t i = 0; i < n; i++) if (raw[i].cls == 1) {
            raw[i].link = (int32_t)kept.size();
            if (!kept.empty() && !(raw[i].f > raw[kept.back()].f)) return (long)i * 1000 + E_UNSUPPORTED_NUM;
            kept.push_back((uint32_t)i);
        }
        v.kept = kept.data(); v.n_kept = kept.size(); info->n_kept = (long long)kept.size();
    }
    if (what < 0) return -1;
    long o = 0, cnt = 0;
    for (uint64_t i = 0; i < n; i++) {
        S2Count kc;
        int e = s2_emit_row(v, i, what, kc);
        if (e > 0) return (long)i * 1000 + e;
        if (e < 0) continue;
        S2Write kw{out + o, 0};
        s2_emit_row(v, i, what, kw);
        if (kw.n != kc.len) return (long)i * 1000 + 99;
        o += (long)kw.n; cnt++;
    }
    *outlen = o; *nrows = cnt;
    return -1;
}
