// diskstat.txt / blktrace.txt / strace.txt / pystacks.txt on the device.  Row logic is side2.cuh;
// these kernels index lines, build the small tables the rows need (device names, 'D' -> 'C' links, ranks) and run the
// length / write passes of the text outputs.
#pragma once
#include "kernels_emit.cuh"
#include "side2.cuh"

namespace sofa {

constexpr uint32_t S2_DEV_CAP = 1024;          // slots of the device-name table (at most half are used)

__global__ void __launch_bounds__(256) k_s2_nlmask(const char* __restrict__ text, uint64_t n, uint8_t* __restrict__ mask) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mask[i] = text[i] == '\n' ? 1 : 0;
}
// line starts from the newline positions; a text that does not end with '\n' has one more line, closed at nbytes + 1
__global__ void __launch_bounds__(256) k_s2_lstart(const uint32_t* __restrict__ nlpos, uint64_t n_nl, uint64_t nbytes, int ends_nl,
                                                  uint32_t* __restrict__ lstart) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) { lstart[0] = 0; if (!ends_nl) lstart[n_nl + 1] = (uint32_t)(nbytes + 1); }
    if (j < n_nl) lstart[j + 1] = nlpos[j] + 1;
}

struct S2ParseArgs {
    int table; const char* text; const uint32_t* lstart; uint64_t n_lines; int last_has_nl;
    S2Line* raw; uint64_t n;
    int absolute; double time_base; double min_time; int64_t pid; uint32_t ratio;
    unsigned long long* err;
};
__global__ void __launch_bounds__(128) k_s2_parse(S2ParseArgs a) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    S2Line r = {};
    int e = E_OK;
    if (a.table == S2_PYSTACKS) {
        const uint64_t l0 = 2 * i, l1 = 2 * i + 1;
        const int n0 = (int)(a.lstart[l0 + 1] - a.lstart[l0]) - 1, n1 = (int)(a.lstart[l1 + 1] - a.lstart[l1]) - 1;
        e = s2_parse_py(a.text, a.lstart[l0], n0, a.lstart[l1], n1, &r);
    } else {
        const bool last = i + 1 == a.n_lines;
        const bool has_nl = !last || a.last_has_nl;
        const int n = (int)(a.lstart[i + 1] - a.lstart[i]) - 1;
        if (a.table == S2_VECCSV) e = s2_parse_veccsv(a.text, a.lstart[i], n, i, &r);
        else if (a.table == S2_BLKCSV) e = s2_parse_blkcsv(a.text, a.lstart[i], n, i, &r);
        else if (a.table == S2_DISK) e = s2_parse_disk(a.text, a.lstart[i], has_nl ? n : (n > 0 ? n - 1 : 0), &r);
        else if (a.table == S2_BLK) e = s2_parse_blk(a.text, a.lstart[i], n, has_nl, &r);
        else {
            if (i % a.ratio) r.link = -1;
            else e = s2_parse_strace(a.text, a.lstart[i], n, a.pid, a.absolute, a.time_base, a.min_time, &r);
        }
    }
    if (e) { report_error(a.err, i, e); r.cls = 0; }
    a.raw[i] = r;
}

// ---------------------------------------------------------------------------- diskstat: device-name table
// keys[slot] = name hash | 1 (0 = empty), first[slot] = first line with that name
__global__ void __launch_bounds__(256) k_s2_disk_insert(S2Line* raw, uint64_t n, unsigned long long* keys, unsigned int* first,
                                                       unsigned int* used, unsigned long long* err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (raw[i].cls == 0) { raw[i].link = -1; return; }            // (the header line of a CSV read back by sofa_analyze)
    const unsigned long long key = raw[i].h | 1ull;
    uint32_t slot = (uint32_t)(key >> 20) & (S2_DEV_CAP - 1);
    for (uint32_t probe = 0; probe < S2_DEV_CAP; probe++, slot = (slot + 1) & (S2_DEV_CAP - 1)) {
        unsigned long long old = atomicCAS(&keys[slot], 0ull, key);
        if (old == 0ull) { if (atomicAdd(used, 1u) >= S2_DEV_CAP / 2) report_error(err, i, E_TOO_MANY_LINES); }
        if (old == 0ull || old == key) { atomicMin(&first[slot], (unsigned int)i); raw[i].link = (int32_t)slot; return; }
    }
    report_error(err, i, E_TOO_MANY_LINES);
}
// two different names with one hash would be merged silently: compare every name with its slot's first line
__global__ void __launch_bounds__(256) k_s2_disk_verify(const char* __restrict__ text, const S2Line* __restrict__ raw, uint64_t n,
                                                       const unsigned int* __restrict__ first, unsigned long long* err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || raw[i].link < 0) return;
    const S2Line& a = raw[i];
    const S2Line& b = raw[first[a.link]];
    bool same = (a.sp[0].b - a.sp[0].a) == (b.sp[0].b - b.sp[0].a);
    for (uint32_t q = 0; same && q < a.sp[0].b - a.sp[0].a; q++) same = text[a.sp[0].a + q] == text[b.sp[0].a + q];
    if (!same) report_error(err, i, E_UTF8 + 1);                 // "two names, one hash"
}
__global__ void __launch_bounds__(256) k_s2_disk_slots(const S2Line* __restrict__ raw, const unsigned long long* __restrict__ keys,
                                                      const unsigned int* __restrict__ first, uint32_t* __restrict__ span) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S2_DEV_CAP) return;
    if (keys[s] == 0ull) { span[2 * s] = 0xffffffffu; span[2 * s + 1] = 0; return; }
    span[2 * s] = raw[first[s]].sp[0].a; span[2 * s + 1] = raw[first[s]].sp[0].b;
}
// per device: ordered bits of max(total bandwidth) over its rows (0 = no row); fl |= which of r_time / w_time / await_time hold a float
__global__ void __launch_bounds__(128) k_s2_disk_stats(S2View v, unsigned long long* maxbits, unsigned int* fl, unsigned long long* err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + v.n_dev;
    if (i >= v.n) return;
    const S2Line& m = v.raw[i];
    DiskRow d;
    int e = s2_disk_row(m, v.raw[i - v.n_dev], v.dev_secsize[m.link], v.absolute, v.time_base, &d);
    if (e) { report_error(err, i, e); return; }
    uint64_t b = d_bits(d.tbw);
    b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    atomicMax(&maxbits[m.link], (unsigned long long)b);
    unsigned int f = (d.rt_f ? 1u : 0u) | (d.wt_f ? 2u : 0u) | (d.at_f ? 4u : 0u);
    if (f) atomicOr(fl, f);
}

// ---------------------------------------------------------------------------- blktrace: 'C' -> latest earlier 'D' of the same block
__global__ void __launch_bounds__(256) k_s2_blk_mask(const S2Line* __restrict__ raw, uint64_t n, uint8_t* __restrict__ mask,
                                                    unsigned long long* what) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t c = raw[i].cls;
    mask[i] = (c == 1 || c == 2) ? 1 : 0;
    if (c) atomicMax(&what[0], (unsigned long long)(i + 1));      // last accepted line + 1 (its cpu is what `cpuid` keeps)
    if (c && (raw[i].flag & 1)) atomicOr(&what[1], 1ull);         // "blktrace maybe record failed!"
}
__global__ void __launch_bounds__(256) k_s2_blk_keys(const S2Line* __restrict__ raw, const uint32_t* __restrict__ idx, uint64_t m,
                                                    uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    keys[j] = (uint64_t)raw[idx[j]].i[0];
    vals[j] = idx[j];
}
// keys sorted (stable: line order inside a key)
__global__ void __launch_bounds__(256) k_s2_blk_link(S2Line* raw, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t m) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t line = vals[j];
    if (raw[line].cls != 2) return;
    int32_t link = -1;
    for (uint64_t q = j; q-- > 0;) {
        if (keys[q] != keys[j]) break;
        if (raw[vals[q]].cls == 1) { link = (int32_t)vals[q]; break; }
    }
    raw[line].link = link;
}

// ---------------------------------------------------------------------------- strace
// total_strace_duration: a left-to-right Python float sum over the lines that reach it (:1667)
__global__ void k_s2_strace_total(const S2Line* __restrict__ raw, uint64_t n, double* out, unsigned int* fl) {
    if (blockIdx.x || threadIdx.x) return;
    double t = 0; unsigned int f = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t c = raw[i].cls;
        if (c) t = d_add(t, raw[i].g);
        if (c == 1) f |= (raw[i].flag & 1) ? 1u : 2u;             // bit 0: some kept duration is a float; bit 1: some is the int 0
    }
    *out = t; *fl = f;
}

// ---------------------------------------------------------------------------- pystacks
__global__ void __launch_bounds__(256) k_s2_cls_mask(const S2Line* __restrict__ raw, uint64_t n, uint8_t* __restrict__ mask) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mask[i] = raw[i].cls == 1 ? 1 : 0;
}
// rank of every kept pair; the keys must rise strictly (a dict with repeated or unordered keys would need the
// reference's insertion-order semantics: refused)
__global__ void __launch_bounds__(256) k_s2_py_rank(S2Line* raw, const uint32_t* __restrict__ kept, uint64_t m, unsigned long long* err) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    raw[kept[r]].link = (int32_t)r;
    if (r && !(raw[kept[r]].f > raw[kept[r - 1]].f)) report_error(err, kept[r], E_UNSUPPORTED_NUM);
}

// ---------------------------------------------------------------------------- analyze: groups of diskstat_vector.csv, 'C' rows of blktrace.csv
// mask of the entries of class `cls` (and, when slot >= 0, of device-table slot `slot`)
__global__ void __launch_bounds__(256) k_s2_pick(const S2Line* __restrict__ raw, uint64_t n, int cls, int slot, uint8_t* __restrict__ mask) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mask[i] = (raw[i].cls == cls && (slot < 0 || raw[i].link == slot)) ? 1 : 0;
}
// sort keys of one column of the picked rows (col < 0: the g field)
__global__ void __launch_bounds__(256) k_s2_col_keys(const S2Line* __restrict__ raw, const uint32_t* __restrict__ rows, uint64_t m, int col,
                                                    uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const S2Line& r = raw[rows[j]];
    keys[j] = f64_ordered(col < 0 ? r.g : s2_vec_col(r, col));
    vals[j] = rows[j];
}
// groupby().mean() of pandas: Kahan-compensated sum in row order (group_mean), one thread per column
__global__ void k_s2_group_means(const S2Line* __restrict__ raw, const uint32_t* __restrict__ rows, uint64_t m, double* __restrict__ out) {
    const int col = threadIdx.x;
    if (blockIdx.x || col >= 7) return;
    double sum = 0, comp = 0;
    for (uint64_t j = 0; j < m; j++) {
        const double y = d_sub(s2_vec_col(raw[rows[j]], col), comp);
        const double t = d_add(sum, y);
        comp = d_sub(d_sub(t, sum), y);
        if (comp != comp) comp = 0;
        sum = t;
    }
    out[col] = d_div(sum, (double)m);
}

// ---------------------------------------------------------------------------- text passes
__global__ void __launch_bounds__(128) k_s2_len(S2View v, int what, uint64_t* __restrict__ lens, unsigned long long* count,
                                               unsigned long long* err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    S2Count k;
    int e = s2_emit_row(v, i, what, k);
    if (e > 0) report_error(err, i, e);
    lens[i] = e == 0 ? (uint64_t)k.len : 0;
    if (e == 0) atomicAdd(count, 1ull);
}
__global__ void __launch_bounds__(128) k_s2_write(S2View v, int what, const uint64_t* __restrict__ offs, char* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    if (offs[i + 1] == offs[i]) return;
    S2Write k{out + offs[i], 0};
    s2_emit_row(v, i, what, k);
}

}  // namespace sofa
