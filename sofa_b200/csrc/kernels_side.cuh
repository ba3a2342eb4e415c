// Recorder side files on the device: the parsed lines (RawCols, k_parse<SRC_RAW>)
// become the rows of mpstat.csv / nvsmi_trace.csv / netstat.csv.  Row arithmetic lives in side.cuh;
// these kernels only index: which earlier line a row is differenced against, which rows survive.
#pragma once
#include "kernels_ingest.cuh"

namespace sofa {

struct SideOutCols { double *ts, *f1, *f2; int64_t* i; uint8_t* cls; };       // i: 6 per row; cls: 0 dropped, 1 trace row, 2 usr_sys row

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
    double v = r.f[(uint64_t)idx[j] * 9 + 1];
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
    const double* cur = r.f + (uint64_t)idx[i + 1] * 9;
    const double* prev = r.f + (uint64_t)idx[i + 1 - stride] * 9;
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
    rr.f[0] = r.f[i * 9];
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
    cur.f[0] = r.f[(i + 1) * 9]; cur.i[0] = r.i[(i + 1) * 4]; cur.i[1] = r.i[(i + 1) * 4 + 1];
    prev.f[0] = r.f[i * 9]; prev.i[0] = r.i[i * 4]; prev.i[1] = r.i[i * 4 + 1];
    SideOutRow q;
    int e = side_netstat_row(cur, prev, absolute, time_base, &q);
    if (e) { report_error(err, i + 1, e); o.cls[i] = 0; return; }
    o.cls[i] = 1; o.ts[i] = q.ts; o.f1[i] = q.f1; o.f2[i] = q.f2;
    for (int k = 0; k < 6; k++) o.i[i * 6 + k] = q.i[k];
}

__device__ __forceinline__ SideOutRow load_sideout(const SideOutCols& o, uint64_t i) {
    SideOutRow q;
    q.ts = o.ts[i]; q.f1 = o.f1[i]; q.f2 = o.f2[i]; q.cls = o.cls[i];
#pragma unroll
    for (int k = 0; k < 6; k++) q.i[k] = o.i[i * 6 + k];
    return q;
}

}  // namespace sofa
