// Window reduction of the side tables (concurrency_breakdown, sofa_analyze.py:75-243).
// The reference masks every table once per window (O(W x N)); here every row finds its window
// with a binary search over the window end times (repeated t += 1/rate on the host, exactly as
// the reference accumulates them) and folds into per-window accumulators with integer atomics.
// Window k covers (end[k] - d, end[k]]; a row can fall into two windows only when rounding makes
// two windows overlap, which the membership test handles explicitly.
#pragma once
#include "kernels_sort.cuh"

namespace sofa {

struct WinAcc {                         // one per window
    unsigned int mp_count;
    int usr_max, sys_max, idl_max, iow_max, cpu_max, cpu_min;
    long long usr_sum, sys_sum, irq_sum;
    unsigned int nv_count;
    unsigned long long nv_max_bits, nv_min_bits;     // f64_ordered() of the duration
    long long nv_sum;                                // durations are integral (checked)
    unsigned long long nv_devmask;
    unsigned int tx_count, rx_count;
};

__device__ __forceinline__ int first_window(const double* __restrict__ ends, uint32_t W, double ts) {
    uint32_t lo = 0, hi = W;                 // smallest k with ts <= ends[k]
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (ts <= ends[mid]) hi = mid; else lo = mid + 1; }
    return (int)lo;
}
// calls f(k) for every window k the timestamp belongs to (at most two)
template <class F>
__device__ __forceinline__ void for_windows(const double* __restrict__ ends, uint32_t W, double delta, double ts, F&& f) {
    int k = first_window(ends, W, ts);
    for (int j = k; j < (int)W && j <= k + 1; j++)
        if (ts > d_sub(ends[j], delta) && ts <= ends[j]) f(j);
}

__global__ void __launch_bounds__(256) k_win_init(WinAcc* acc, uint32_t W) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= W) return;
    WinAcc a;
    a.mp_count = 0; a.usr_max = -2147483647; a.sys_max = -2147483647; a.idl_max = -2147483647; a.iow_max = -2147483647;
    a.cpu_max = 0; a.cpu_min = 100;                                   // :133-134
    a.usr_sum = 0; a.sys_sum = 0; a.irq_sum = 0;
    a.nv_count = 0; a.nv_max_bits = 0; a.nv_min_bits = ~0ull; a.nv_sum = 0; a.nv_devmask = 0; a.tx_count = 0; a.rx_count = 0;
    acc[k] = a;
}
__global__ void __launch_bounds__(256) k_win_mpstat(SideCols m, uint64_t n, const double* __restrict__ ends, uint32_t W, double delta,
                                                    WinAcc* acc) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t* r = m.ratios + i * 5;
    int usr = r[0], sys = r[1], idl = r[2], iow = r[3], irq = r[4];
    for_windows(ends, W, delta, m.ts[i], [&](int k) {
        WinAcc* a = acc + k;
        atomicAdd(&a->mp_count, 1u);
        atomicMax(&a->usr_max, usr); atomicMax(&a->sys_max, sys); atomicMax(&a->idl_max, idl); atomicMax(&a->iow_max, iow);
        atomicMax(&a->cpu_max, usr + sys + irq); atomicMin(&a->cpu_min, usr + sys + irq);
        atomicAdd((unsigned long long*)&a->usr_sum, (unsigned long long)(long long)usr);
        atomicAdd((unsigned long long*)&a->sys_sum, (unsigned long long)(long long)sys);
        atomicAdd((unsigned long long*)&a->irq_sum, (unsigned long long)(long long)irq);
    });
}
__global__ void __launch_bounds__(256) k_win_nvsmi(SideCols v, uint64_t n, const double* __restrict__ ends, uint32_t W, double delta,
                                                   WinAcc* acc, unsigned long long* devmask_all, unsigned int* bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double dev = v.dev[i];
    long long di = (long long)dev;
    if ((double)di != dev || di < 0 || di > 63) { atomicOr(bad, 1u); return; }
    atomicOr(devmask_all, 1ull << di);                         // len(set(df_nvsmi['deviceId'])) over the whole frame
    if (v.event[i] != 0.0) return;                             // SM rows only (:112)
    double d = v.dur[i];
    long long dl = (long long)d;
    if ((double)dl != d) { atomicOr(bad, 2u); return; }
    for_windows(ends, W, delta, v.ts[i], [&](int k) {
        WinAcc* a = acc + k;
        atomicAdd(&a->nv_count, 1u);
        atomicMax(&a->nv_max_bits, (unsigned long long)f64_ordered(d));
        atomicMin(&a->nv_min_bits, (unsigned long long)f64_ordered(d));
        atomicAdd((unsigned long long*)&a->nv_sum, (unsigned long long)dl);
        atomicOr(&a->nv_devmask, 1ull << di);
    });
}
// netstat rows: key = 2 * window + (event == 1) for the pairwise bandwidth sums (row order kept by the
// stable sort); rows outside every window or with another event get key ~0 (sorted last, ignored)
__global__ void __launch_bounds__(256) k_win_net_keys(SideCols b, uint64_t n, const double* __restrict__ ends, uint32_t W, double delta,
                                                      WinAcc* acc, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, unsigned int* bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = ~0ull;
    double ev = b.event[i];
    if (ev == 0.0 || ev == 1.0) {
        int hits = 0;
        for_windows(ends, W, delta, b.ts[i], [&](int k) {
            hits++;
            key = 2ull * (uint64_t)k + (ev == 1.0 ? 1ull : 0ull);
            atomicAdd(ev == 1.0 ? &acc[k].rx_count : &acc[k].tx_count, 1u);
        });
        if (hits > 1) atomicOr(bad, 4u);
    }
    keys[i] = key; vals[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256) k_win_net_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ seg_start,
                                                         uint32_t nseg, const double* __restrict__ sums, double* __restrict__ txrx) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    uint64_t key = keys[seg_start[s]];
    if (key != ~0ull) txrx[key] = sums[s];
}

// ---------------------------------------------------------------------------- nvsmi_profile (sofa_analyze.py:259-341)
// Sort key of an nvsmi_trace.csv row: (deviceId * 4 + event) << 32 | (duration + 2^31), so that every
// (deviceId, event) group is contiguous and ordered by duration; rows outside the region of interest
// get ~0.  tab[g] = {count, sum}; first[dev] = first row of that device (set() insertion order);
// span[0] / span[1] = first / last kept row.
__global__ void __launch_bounds__(256) k_nvsmi_keys(SideCols v, uint64_t n, int use_roi, double roi_begin, double roi_end,
                                                    uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                    unsigned long long* tab, unsigned long long* first, unsigned long long* span,
                                                    unsigned int* bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = ~0ull;
    double ts = v.ts[i];
    if (!use_roi || (ts > roi_begin && ts <= roi_end)) {
        double dev = v.dev[i], ev = v.event[i], d = v.dur[i];
        long long di = (long long)dev, ei = (long long)ev, dl = (long long)d;
        if ((double)di != dev || di < 0 || di > 63) { atomicOr(bad, 1u); }
        else if ((double)ei != ev || ei < 0 || ei > 3) { atomicOr(bad, 8u); }
        else if ((double)dl != d || dl <= -2147483648ll || dl >= 2147483648ll) { atomicOr(bad, 2u); }
        else {
            unsigned g = (unsigned)(di * 4 + ei);
            key = ((uint64_t)g << 32) | (uint64_t)(uint32_t)(dl + 2147483648ll);
            atomicAdd(&tab[2 * g], 1ull);
            atomicAdd(&tab[2 * g + 1], (unsigned long long)dl);
            atomicMin(&first[di], (unsigned long long)i);
            atomicMin(&span[0], (unsigned long long)i);
            atomicMax(&span[1], (unsigned long long)i);
        }
    }
    keys[i] = key; vals[i] = (uint32_t)i;
}
__global__ void k_gather_u64(const uint64_t* __restrict__ src, const uint64_t* __restrict__ pos, uint32_t n, uint64_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[pos[i]];
}

}  // namespace sofa
