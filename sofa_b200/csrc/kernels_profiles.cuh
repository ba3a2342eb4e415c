// Small-table profile functions of sofa_analyze (SURVEY.md §8 a15 verbose tables, a17, a8 network filters, the
// --spotlight_gpu region of interest).  Included by tu_analyze.cu only.
#pragma once
#include "kernels_sort.cuh"

namespace sofa {

// ---------------------------------------------------------------------------- net_profile --verbose (sofa_analyze.py:398)
// df.groupby(['pkt_src','pkt_dst','payload']).size(): open addressing on a tag of the triple; every slot keeps the
// min / max of its three components so that the host can prove that no two triples shared a slot.
struct NetTripleSlot { unsigned long long tag; long long src_min, src_max, dst_min, dst_max, pl_min, pl_max; unsigned long long count; };
__global__ void k_net_triples_init(NetTripleSlot* tab, uint64_t cap) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    NetTripleSlot s; s.tag = 0; s.count = 0;
    s.src_min = s.dst_min = s.pl_min = 0x7fffffffffffffffll; s.src_max = s.dst_max = s.pl_max = -0x7fffffffffffffffll - 1;
    tab[i] = s;
}
__global__ void __launch_bounds__(256) k_net_triples(SrcView src, uint64_t n, uint32_t stride, NetTripleSlot* tab, uint64_t cap_mask,
                                                     unsigned int* overflow) {
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n || row % stride) return;
    NetRow r = load_net(src, row);
    if (r.proto == 0) return;                                  // NaN keys are dropped by groupby
    uint64_t h = 1469598103934665603ull;
    h = (h ^ (uint64_t)r.src) * 1099511628211ull; h ^= h >> 29;
    h = (h ^ (uint64_t)r.dst) * 1099511628211ull; h ^= h >> 32;
    h = (h ^ (uint64_t)r.payload) * 1099511628211ull; h ^= h >> 31;
    unsigned long long tag = h | 1ull;
    uint64_t slot = (h >> 7) & cap_mask;
    for (uint64_t probe = 0; probe <= cap_mask && probe < 65536; probe++) {
        unsigned long long prev = atomicCAS(&tab[slot].tag, 0ull, tag);
        if (prev == 0ull || prev == tag) {
            atomicAdd(&tab[slot].count, 1ull);
            atomicMin(&tab[slot].src_min, (long long)r.src); atomicMax(&tab[slot].src_max, (long long)r.src);
            atomicMin(&tab[slot].dst_min, (long long)r.dst); atomicMax(&tab[slot].dst_max, (long long)r.dst);
            atomicMin(&tab[slot].pl_min, (long long)r.payload); atomicMax(&tab[slot].pl_max, (long long)r.payload);
            return;
        }
        slot = (slot + 1) & cap_mask;
    }
    atomicOr(overflow, 1u);
}

// ---------------------------------------------------------------------------- --network_filters (sofa_preprocess.py:1214-1231)
// bit k: payload > 0, pkt_src == base, pkt_dst == others[k] ("to_<k>");  bit 4 + k: payload > 0, pkt_dst == base,
// pkt_src == others[k] ("from_<k>")
struct NetFilterArgs { long long base, others[4]; int n; };
__global__ void __launch_bounds__(256) k_net_filter(SrcView src, uint64_t ncand, uint32_t stride, NetFilterArgs a,
                                                    uint8_t* __restrict__ mask, unsigned long long* counts) {
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t hit = 0;
    if (c < ncand) {
        NetRow r = load_net(src, c * stride);
        if (r.proto != 0 && r.payload > 0) {
            for (int k = 0; k < a.n; k++) {
                if (r.src == a.base && r.dst == a.others[k]) hit |= 1u << k;
                if (r.dst == a.base && r.src == a.others[k]) hit |= 1u << (4 + k);
            }
        }
        mask[c] = (uint8_t)hit;
    }
    for (int k = 0; k < 8; k++) {
        unsigned b = __ballot_sync(0xffffffffu, (hit >> k) & 1);
        if (lane_id() == 0 && b) atomicAdd(&counts[k], (unsigned long long)__popc(b));
    }
}

// ---------------------------------------------------------------------------- mpstat_profile (sofa_analyze.py:735-790)
// key = deviceId (core) for the stable sort; the five per-row times t_X = dt_all * X / 100 with
// dt_all = 0.1 where IDL == 100 else duration / ((100 - IDL) / 100), operation for operation (:742-747)
__global__ void __launch_bounds__(256) k_mp_keys(SideCols m, uint64_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                 unsigned long long* maxdev, unsigned int* bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double dev = m.dev[i];
    long long di = (long long)dev;
    if ((double)di != dev || di < 0 || di > 1000000) { atomicOr(bad, 1u); di = 0; }
    keys[i] = (uint64_t)di; vals[i] = (uint32_t)i;
    atomicMax(maxdev, (unsigned long long)di);
}
__global__ void __launch_bounds__(256) k_mp_values(SideCols m, uint64_t n, int which, double* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double idl = (double)m.ratios[i * 5 + 2];
    const double dt_all = idl == 100.0 ? 0.1 : d_div(m.dur[i], d_div(d_sub(100.0, idl), 100.0));
    out[i] = d_div(d_mul(dt_all, (double)m.ratios[i * 5 + which]), 100.0);
}

// ---------------------------------------------------------------------------- quantiles / mean of one column of a side table
// rows with event == ev (or all rows): ordered keys of the value for the sort, the values in row order for the sum
__global__ void __launch_bounds__(256) k_side_pick_count(SideCols t, uint64_t n, int use_event, double ev, uint64_t* blockoff) {
    __shared__ uint32_t sm[33];
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t total;
    block_excl_scan((i < n && (!use_event || t.event[i] == ev)) ? 1u : 0u, &total, sm);
    if (threadIdx.x == 0) blockoff[blockIdx.x] = total;
}
__global__ void __launch_bounds__(256) k_side_pick_write(SideCols t, uint64_t n, int use_event, double ev, int col, const uint64_t* blockoff,
                                                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, double* __restrict__ values) {
    __shared__ uint32_t sm[33];
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    bool f = i < n && (!use_event || t.event[i] == ev);
    uint32_t total;
    uint32_t ex = block_excl_scan(f ? 1u : 0u, &total, sm);
    if (f) {
        uint64_t o = blockoff[blockIdx.x] + ex;
        double v = col == 2 ? t.dur[i] : (col == 6 ? t.bw[i] : (col == 0 ? t.ts[i] : t.dev[i]));
        keys[o] = f64_ordered(v); vals[o] = (uint32_t)i; values[o] = v;
    }
}
__global__ void k_gather_ordered(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ pos, uint32_t n, double* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = keys[pos[i]];
    out[i] = bits_d((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k);
}

// sums of the five integers parsed out of the name column (mpstat: usr sys idl iow irq; vmstat: bi bo in cs)
__global__ void __launch_bounds__(256) k_side_int_sums(SideCols t, uint64_t n, unsigned long long* sums) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long v[5] = {0, 0, 0, 0, 0};
    if (i < n) for (int q = 0; q < 5; q++) v[q] = t.ratios[i * 5 + q];
    for (int q = 0; q < 5; q++) {
        long long s = v[q];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane_id() == 0 && s) atomicAdd(&sums[q], (unsigned long long)s);
    }
}

// ---------------------------------------------------------------------------- --spotlight_gpu region of interest (:872-895)
// the reference's state machine over the event == 0, deviceId == 0 rows of nvsmi_trace.csv, in row order: a
// saturating counter, sequential by nature - one thread walks the (10 Hz) table
__global__ void k_nvsmi_roi(SideCols v, uint64_t n, double* roi) {       // roi[0] begin, roi[1] end (in/out)
    if (blockIdx.x || threadIdx.x) return;
    int state = 0, sm_high = 0;
    const int trigger = 10;
    double rb = roi[0], re = roi[1];
    for (uint64_t i = 0; i < n; i++) {
        if (v.event[i] == 0.0 && v.dev[i] == 0.0) {
            const double d = v.dur[i];
            if (d >= 50.0) sm_high = sm_high + 1 < trigger ? sm_high + 1 : trigger;
            if (d < 10.0) sm_high = sm_high - 1 > 0 ? sm_high - 1 : 0;
            if (state == 0 && sm_high == trigger) { state = 1; rb = v.ts[i]; }
            else if (state == 1 && sm_high == 0) { state = 0; re = v.ts[i]; }
        }
    }
    roi[0] = rb; roi[1] = re;
}

}  // namespace sofa

namespace sofa {
// pairwise segment sums over a column that is NOT read through the "%.6f" round trip (computed values)
struct GatherRaw {
    const double* col; const uint32_t* idx;
    __device__ __forceinline__ double operator[](uint64_t i) const { return col[idx[i]]; }
    __device__ __forceinline__ GatherRaw operator+(uint64_t d) const { return GatherRaw{col, idx + d}; }
};
__global__ void __launch_bounds__(128) k_seg_pairwise_raw(const double* __restrict__ col, const uint32_t* __restrict__ vals,
                                                          const uint32_t* __restrict__ seg_start, uint32_t nseg, uint64_t n,
                                                          double* __restrict__ sums, uint32_t* __restrict__ big_list,
                                                          unsigned int* __restrict__ n_big) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    uint64_t a = seg_start[s], b = (s + 1 < nseg) ? seg_start[s + 1] : n;
    if (b - a > SEG_BIG) { big_list[atomicAdd(n_big, 1u)] = s; return; }
    GatherRaw g{col, vals + a};
    sums[s] = pairwise_sum(g, b - a);
}
__global__ void __launch_bounds__(256) k_seg_pairwise_big_raw(const double* __restrict__ col, const uint32_t* __restrict__ vals,
                                                              const uint32_t* __restrict__ seg_start, uint32_t nseg, uint64_t n,
                                                              double* __restrict__ sums, const uint32_t* __restrict__ big_list) {
    __shared__ double part[256];
    uint32_t s = big_list[blockIdx.x];
    uint64_t a = seg_start[s], b = (s + 1 < nseg) ? seg_start[s + 1] : n;
    uint64_t len = b - a;
    int depth = 0;
    while (depth < 8 && (len >> (depth + 1)) >= 256) depth++;
    if (threadIdx.x < (1u << depth)) {
        uint64_t st, l;
        pairwise_subtree(len, depth, threadIdx.x, &st, &l);
        GatherRaw g{col, vals + a + st};
        part[threadIdx.x] = pairwise_sum(g, l);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int lev = depth - 1; lev >= 0; lev--)
            for (uint32_t i = 0; i < (1u << lev); i++) part[i] = d_add(part[2 * i], part[2 * i + 1]);
        sums[s] = d_add(0.0, part[0]);
    }
}
}  // namespace sofa

namespace sofa {
// ---------------------------------------------------------------------------- AISI iter_profile (sofa_aisi.py:21-59, 429-453)
// Rows of one GPU (df_gpu[df_gpu.deviceId == dev]) falling into iteration i: times[i] <= timestamp < times[i+1] (CSV
// values).  Per row: iteration index + which of the five keywords ("_fw", "_bw", "gemm", "CUDA_COPY", "nccl") its name
// contains (one packed shift-and pass).  Per iteration, exact integer / index reductions with atomics: row count,
// first and last row (row order = frame order), payload sum.
struct IterAcc { unsigned long long count, first_row, last_row; long long payload; };
__global__ void __launch_bounds__(256) k_iter_rows(SrcView src, uint64_t n, int32_t dev, const double* __restrict__ times, uint32_t nt,
                                                   const PackedFilter* __restrict__ pf, uint32_t* __restrict__ iter_of, uint8_t* __restrict__ bits,
                                                   IterAcc* acc) {
    __shared__ PackedFilter spf;
    {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(pf);
        uint32_t* sp = reinterpret_cast<uint32_t*>(&spf);
        for (uint32_t i = threadIdx.x; i < sizeof(PackedFilter) / 4; i += blockDim.x) sp[i] = g[i];
    }
    __syncthreads();
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    uint32_t it = 0xffffffffu; uint32_t hit = 0;
    if (src.c.g.dev[row] == dev) {
        const double ts = round6_value(src.c.g.ts[row]);
        if (ts >= times[0] && ts < times[nt - 1]) {
            uint32_t lo = 0, hi = nt - 1;                       // largest lo with times[lo] <= ts
            while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (times[mid] <= ts) lo = mid; else hi = mid; }
            it = lo;
            GpuRow r = load_gpu(src, row);
            GpuName<const char*> ng{src.text + src.c.g.line_start[row], r, true, src.api_mode != 0};
            PackedMatchSink ms(&spf); ng.gen(ms); hit = ms.hits();
            atomicAdd(&acc[it].count, 1ull);
            atomicMin(&acc[it].first_row, (unsigned long long)row);
            atomicMax(&acc[it].last_row, (unsigned long long)row);
            atomicAdd((unsigned long long*)&acc[it].payload, (unsigned long long)r.payload);
        }
    }
    iter_of[row] = it; bits[row] = (uint8_t)hit;
}
// selection `sel` (0: every row of an iteration, 1 + k: rows whose name holds keyword k): key = iteration + 1, else 0;
// sel == 6: key = (iteration + 1) << 32 | tid (distinct streams per iteration)
__global__ void __launch_bounds__(256) k_iter_selkey(const uint32_t* __restrict__ iter_of, const uint8_t* __restrict__ bits,
                                                     const int32_t* __restrict__ tid, uint64_t n, int sel, uint64_t* __restrict__ key) {
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const uint32_t it = iter_of[row];
    uint64_t k = 0;
    if (it != 0xffffffffu) {
        if (sel == 0) k = (uint64_t)it + 1;
        else if (sel == 6) k = (((uint64_t)it + 1) << 32) | (uint32_t)tid[row];
        else if ((bits[row] >> (sel - 1)) & 1) k = (uint64_t)it + 1;
    }
    key[row] = k;
}
}  // namespace sofa
