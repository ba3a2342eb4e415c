// Recorder side files on the device: the parsed lines (RawCols, k_parse<SRC_RAW>)
// become the rows of mpstat.csv / nvsmi_trace.csv / netstat.csv.  Row arithmetic lives in side.cuh;
// these kernels only index: which earlier line a row is differenced against, which rows survive.
#pragma once
#include "kernels_ingest.cuh"

namespace sofa {

struct SideOutCols { double *ts, *f1, *f2; int64_t* i; uint8_t* cls; const double* ext; const uint32_t* ext_idx; };
// i: 6 per row; cls: 0 dropped, 1 trace row, 2 usr_sys row; ext: the raw RAW_NF doubles of line ext_idx[row] (vmstat)

#ifdef SOFA_K_SIDE
// what[0] = first line with nf > 0 (min), what[1] |= 1 when some line is flagged "failed"
__global__ void __launch_bounds__(256) k_side_first(RawCols r, uint64_t n, unsigned long long* what) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int nf = r.nf[i];
    if (nf > 0) atomicMin(&what[0], (unsigned long long)i);
    if (nf < 0) atomicOr(&what[1], 1ull);
}
// genfromtxt(invalid_raise=False): keep the lines whose field count equals the first line's
__global__ void __launch_bounds__(256) k_side_valid(RawCols r, uint64_t n, const unsigned long long* what, uint8_t* mask) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long first = what[0];
    mask[i] = (first < n && r.nf[i] == r.nf[first]) ? 1 : 0;
}
// max of column 1 over the kept lines j >= 1 (mpstat[1:, 1].max(), :489); what[2] = ordered bits, what[3] |= nan seen
__global__ void __launch_bounds__(256) k_side_maxcpu(RawCols r, const uint32_t* __restrict__ idx, uint64_t m, unsigned long long* what) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (j >= m) return;
    double v = r.f[(uint64_t)idx[j] * RAW_NF + 1];
    if (d_isnan(v)) { atomicOr(&what[3], 1ull); return; }
    uint64_t b = d_bits(v);
    b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    atomicMax(&what[2], (unsigned long long)b);
}
// rows i = j - 1 of mpstat[1:] (:492-543)
__global__ void __launch_bounds__(128) k_side_mpstat(RawCols r, const uint32_t* __restrict__ idx, uint64_t m, uint64_t stride,
                                                     int absolute, double time_base, SideOutCols o, unsigned long long* err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= m) return;
    o.cls[i] = 0;
    if (i <= stride) return;
    const double* cur = r.f + (uint64_t)idx[i + 1] * RAW_NF;
    const double* prev = r.f + (uint64_t)idx[i + 1 - stride] * RAW_NF;
    SideOutRow q;
    int e = side_mpstat_row(cur, prev, absolute, time_base, &q);
    if (e) { report_error(err, idx[i + 1], e); return; }
    o.cls[i] = (uint8_t)q.cls;
    o.ts[i] = q.ts; o.f1[i] = q.f1; o.f2[i] = q.f2;
    for (int k = 0; k < 6; k++) o.i[i * 6 + k] = q.i[k];
}
__global__ void __launch_bounds__(128) k_side_nvsmi(RawCols r, uint64_t n, long long local_off, double tz_add, int absolute,
                                                    double time_base, SideOutCols o) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RawRow rr = RawRow();
    for (int k = 0; k < 4; k++) rr.i[k] = r.i[i * 4 + k];
    rr.f[0] = r.f[i * RAW_NF];
    SideOutRow q;
    side_nvsmi_row(rr, local_off, tz_add, absolute, time_base, &q);
    o.cls[i] = 1; o.ts[i] = q.ts; o.f1[i] = 0; o.f2[i] = 0;
    for (int k = 0; k < 6; k++) o.i[i * 6 + k] = q.i[k];
}
// rows i = line i + 1 against line i (:1249-1321)
__global__ void __launch_bounds__(128) k_side_netstat(RawCols r, uint64_t n, int absolute, double time_base, SideOutCols o,
                                                      unsigned long long* err) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    RawRow cur = RawRow(), prev = RawRow();
    cur.f[0] = r.f[(i + 1) * RAW_NF]; cur.i[0] = r.i[(i + 1) * 4]; cur.i[1] = r.i[(i + 1) * 4 + 1];
    prev.f[0] = r.f[i * RAW_NF]; prev.i[0] = r.i[i * 4]; prev.i[1] = r.i[i * 4 + 1];
    SideOutRow q;
    int e = side_netstat_row(cur, prev, absolute, time_base, &q);
    if (e) { report_error(err, i + 1, e); o.cls[i] = 0; return; }
    o.cls[i] = 1; o.ts[i] = q.ts; o.f1[i] = q.f1; o.f2[i] = q.f2;
    for (int k = 0; k < 6; k++) o.i[i * 6 + k] = q.i[k];
}

// lines with nf > 0 (vmstat / dmon: the records that are not headers or short lines)
__global__ void __launch_bounds__(256) k_side_accept(RawCols r, uint64_t n, uint8_t* mask, unsigned long long* what) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int nf = r.nf[i];
    mask[i] = (nf > 0 ? 1 : 0) | ((nf > 0 && r.i[i * 4] == 0) ? 2 : 0);      // bit 1: dmon line of gpu 0 (advances the clock)
    if (nf < 0) atomicOr(&what[1], 1ull);
}
// exclusive count of mask bit `bit` before every element (blockoff: scanned per-block counts of k_mask_count)
__global__ void __launch_bounds__(256) k_excl_prefix(const uint8_t* __restrict__ mask, int bit, uint64_t n, const uint64_t* blockoff,
                                                     uint32_t* __restrict__ out) {
    __shared__ uint32_t sm[33];
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t total;
    uint32_t ex = block_excl_scan((i < n && ((mask[i] >> bit) & 1)) ? 1u : 0u, &total, sm);
    if (i < n) out[i] = (uint32_t)blockoff[blockIdx.x] + ex;
}
__global__ void __launch_bounds__(128) k_side_vmstat(RawCols r, const uint32_t* __restrict__ idx, uint64_t m, int absolute, double time_base,
                                                     SideOutCols o, unsigned long long* err) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    SideOutRow q;
    int e = side_vmstat_row((int64_t)j, r.f + (uint64_t)idx[j] * RAW_NF, absolute, time_base, &q);
    if (e) { report_error(err, idx[j], e); o.cls[j] = 0; return; }
    o.cls[j] = 1; o.ts[j] = q.ts; o.f1[j] = 0; o.f2[j] = 0;
    for (int k = 0; k < 6; k++) o.i[j * 6 + k] = q.i[k];
}
// tclock[line] = number of earlier accepted gpu-0 lines
__global__ void __launch_bounds__(128) k_side_dmon(RawCols r, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ tclock, uint64_t m,
                                                   int absolute, double time_base, SideOutCols o) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    uint64_t line = idx[j];
    RawRow rr = RawRow();
    for (int k = 0; k < 4; k++) rr.i[k] = r.i[line * 4 + k];
    SideOutRow q;
    side_dmon_row((int64_t)tclock[line], rr, absolute, time_base, &q);
    o.cls[j] = (uint8_t)q.cls; o.ts[j] = q.ts; o.f1[j] = 0; o.f2[j] = 0;
    for (int k = 0; k < 6; k++) o.i[j * 6 + k] = q.i[k];
}

#endif  // SOFA_K_SIDE

__device__ __forceinline__ SideOutRow load_sideout(const SideOutCols& o, uint64_t i) {
    SideOutRow q;
    q.ext = o.ext ? o.ext + (uint64_t)o.ext_idx[i] * RAW_NF : nullptr;
    q.ts = o.ts[i]; q.f1 = o.f1[i]; q.f2 = o.f2[i]; q.cls = o.cls[i];
#pragma unroll
    for (int k = 0; k < 6; k++) q.i[k] = o.i[i * 6 + k];
    return q;
}

}  // namespace sofa
