// Device LSD radix sort (64-bit keys, 32-bit payload, stable) and segmented reductions.
//
// 8-bit digits; per pass: per-tile digit histograms -> one scan over [digit][tile] -> stable
// scatter.  Stability inside a tile comes from warp-synchronous ranking: every warp walks its
// contiguous 512-key chunk 32 keys at a time, __match_any_sync groups equal digits, the rank
// is (running per-warp digit count) + (equal-digit lanes below me).
//
// Used for: group-by-name (sort rows by 64-bit name hash, row order preserved inside a group
// -> numpy pairwise sums per group, sofa_analyze.py:34-44, 363-368) and sort_values(duration)
// (sofa_analyze.py:706).
#pragma once
#include "kernels_analyze.cuh"

namespace sofa {

constexpr int SORT_TPB = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_TPB * SORT_ITEMS;     // 4096 keys per block

__global__ void __launch_bounds__(SORT_TPB) k_radix_hist(const uint64_t* __restrict__ keys, uint64_t n, int shift,
                                                         uint64_t* __restrict__ hist, uint32_t nb) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint64_t base = (uint64_t)blockIdx.x * SORT_TILE;
#pragma unroll
    for (int s = 0; s < SORT_ITEMS; s++) {
        uint64_t i = base + (uint64_t)s * SORT_TPB + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(SORT_TPB) k_radix_scatter(const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                            uint64_t* __restrict__ kout, uint32_t* __restrict__ vout, uint64_t n,
                                                            int shift, const uint64_t* __restrict__ offs, uint32_t nb) {
    __shared__ unsigned int whist[SORT_TPB / 32][256];
    for (int i = threadIdx.x; i < (SORT_TPB / 32) * 256; i += SORT_TPB) (&whist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1;
    uint64_t base = (uint64_t)blockIdx.x * SORT_TILE + (uint64_t)w * (32 * SORT_ITEMS);
    uint64_t k[SORT_ITEMS]; uint32_t v[SORT_ITEMS]; uint32_t rank[SORT_ITEMS];
#pragma unroll
    for (int s = 0; s < SORT_ITEMS; s++) {
        uint64_t i = base + (uint64_t)s * 32 + lane;
        bool valid = i < n;
        k[s] = valid ? kin[i] : ~0ull;
        v[s] = valid ? vin[i] : 0u;
        uint32_t d = (uint32_t)(k[s] >> shift) & 255u;
        uint32_t m = __match_any_sync(0xffffffffu, d);
        rank[s] = whist[w][d] + __popc(m & lt);
        __syncwarp();
        if (lane == (uint32_t)(__ffs(m) - 1)) whist[w][d] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
    {   // digit = threadIdx.x: exclusive prefix over the warps, on top of the global offset
        uint32_t run = (uint32_t)offs[(uint64_t)threadIdx.x * nb + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < SORT_TPB / 32; ww++) { uint32_t c = whist[ww][threadIdx.x]; whist[ww][threadIdx.x] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < SORT_ITEMS; s++) {
        uint64_t i = base + (uint64_t)s * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)(k[s] >> shift) & 255u;
            uint32_t pos = whist[w][d] + rank[s];
            kout[pos] = k[s]; vout[pos] = v[s];
        }
    }
}

// order-preserving double -> u64 (ascending); descending = bitwise not
__device__ __forceinline__ uint64_t f64_ordered(double x) {
    uint64_t b = d_bits(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// ---------------------------------------------------------------------------- key builders
// (hash, row) of gpu rows whose name contains "nccl", row order (two-pass compaction)
__global__ void __launch_bounds__(256) k_pairs_count(const uint64_t* __restrict__ key, uint64_t n, uint64_t* blockoff) {
    __shared__ uint32_t sm[33];
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t total;
    block_excl_scan((i < n && key[i] != 0) ? 1u : 0u, &total, sm);
    if (threadIdx.x == 0) blockoff[blockIdx.x] = total;
}
__global__ void __launch_bounds__(256) k_pairs_write(const uint64_t* __restrict__ key, uint64_t n, const uint64_t* blockoff,
                                                     uint64_t* __restrict__ kout, uint32_t* __restrict__ vout) {
    __shared__ uint32_t sm[33];
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    bool f = i < n && key[i] != 0;
    uint32_t total;
    uint32_t ex = block_excl_scan(f ? 1u : 0u, &total, sm);
    if (f) { uint64_t o = blockoff[blockIdx.x] + ex; kout[o] = key[i]; vout[o] = (uint32_t)i; }
}
// FNV-1a of every row's full name (all kinds) + identity payload
template <int KIND>
__global__ void __launch_bounds__(256) k_name_hash_all(SrcView src, uint64_t n, uint64_t* __restrict__ kout, uint32_t* __restrict__ vout,
                                                        uint64_t hash_mask) {
    uint64_t row = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= n) return;
    with_row<KIND, false>(src, row, row, [&](const RowVals&, const auto& ng) {
        HashSink h; ng.gen(h);
        h.h &= hash_mask;                    // all ones, except in the collision test (sofa_b200_set_debug)
        kout[row] = h.h ? h.h : 1;
        vout[row] = (uint32_t)row;
    });
}
// Groups are keyed by a 64-bit hash of the name; this pass makes them exact: in the hash-sorted order every row whose
// key equals its predecessor's must have the same name bytes (equality is transitive along the segment).  A
// mismatch (two names with one hash) sets *collision: the host raises instead of merging two groups.
struct WinWrite {           // bytes [lo, lo + W) of the generated name -> buf
    static constexpr bool WORDS = false;
    char* buf; int lo, n;
    __device__ __forceinline__ void put(char ch) { int d = n - lo; if (d >= 0 && d < 128) buf[d] = ch; n++; }
    __device__ __forceinline__ void put_word(uint64_t w, int nb) { put_word_bytes(*this, w, nb); }
};
struct WinCompare {
    static constexpr bool WORDS = false;
    const char* buf; int lo, n; bool diff;
    __device__ __forceinline__ void put(char ch) { int d = n - lo; if (d >= 0 && d < 128 && buf[d] != ch) diff = true; n++; }
    __device__ __forceinline__ void put_word(uint64_t w, int nb) { put_word_bytes(*this, w, nb); }
};
template <int KIND>
__global__ void __launch_bounds__(128) k_verify_groups(SrcView src, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                       uint64_t m, unsigned int* collision) {
    uint64_t i = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    if (i == 0 || i >= m || keys[i] != keys[i - 1]) return;
    const uint64_t ra = vals[i - 1], rb = vals[i];
    char buf[128];
    bool diff = false;
    with_row<KIND, false>(src, ra, ra, [&](const RowVals&, const auto& na) {
        with_row<KIND, false>(src, rb, rb, [&](const RowVals&, const auto& nb) {
            for (int lo = 0;; lo += 128) {
                WinWrite w{buf, lo, 0}; na.gen(w);
                WinCompare cmp{buf, lo, 0, false}; nb.gen(cmp);
                if (cmp.diff || cmp.n != w.n) { diff = true; break; }
                if (lo + 128 >= w.n) break;
            }
        });
    });
    if (diff) atomicOr(collision, 1u);
}

// keys for sort_values(by='duration', ascending=False): descending order of the CSV value
__global__ void __launch_bounds__(256) k_dur_keys_desc(const double* __restrict__ dur, uint64_t n, uint64_t* __restrict__ kout,
                                                       uint32_t* __restrict__ vout) {
    uint64_t row = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= n) return;
    kout[row] = ~f64_ordered(round6_value(dur[row]));
    vout[row] = (uint32_t)row;
}

// ---------------------------------------------------------------------------- segments of equal keys
__global__ void __launch_bounds__(256) k_seg_count(const uint64_t* __restrict__ key, uint64_t n, uint64_t* blockoff) {
    __shared__ uint32_t sm[33];
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t total;
    block_excl_scan((i < n && (i == 0 || key[i] != key[i - 1])) ? 1u : 0u, &total, sm);
    if (threadIdx.x == 0) blockoff[blockIdx.x] = total;
}
__global__ void __launch_bounds__(256) k_seg_write(const uint64_t* __restrict__ key, uint64_t n, const uint64_t* blockoff,
                                                   uint32_t* __restrict__ seg_start) {
    __shared__ uint32_t sm[33];
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    bool f = i < n && (i == 0 || key[i] != key[i - 1]);
    uint32_t total;
    uint32_t ex = block_excl_scan(f ? 1u : 0u, &total, sm);
    if (f) seg_start[blockoff[blockIdx.x] + ex] = (uint32_t)i;
}

// values of a segment in sorted (= row) order, read through the "%.6f" round trip
struct GatherR6 {
    const double* col; const uint32_t* idx;
    __device__ __forceinline__ double operator[](uint64_t i) const { return round6_value(col[idx[i]]); }
    __device__ __forceinline__ GatherR6 operator+(uint64_t d) const { return GatherR6{col, idx + d}; }
};
// numpy pairwise sum of every segment's values.  Small segments: one thread each.  Segments longer
// than SEG_BIG are queued and summed by one block each (2^d sub-trees in parallel, tree-order combine).
constexpr uint32_t SEG_BIG = 2048;
__global__ void __launch_bounds__(128) k_seg_pairwise(const double* __restrict__ col, const uint32_t* __restrict__ vals,
                                                      const uint32_t* __restrict__ seg_start, uint32_t nseg, uint64_t n,
                                                      double* __restrict__ sums, uint32_t* __restrict__ big_list,
                                                      unsigned int* __restrict__ n_big) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    uint64_t a = seg_start[s], b = (s + 1 < nseg) ? seg_start[s + 1] : n;
    if (b - a > SEG_BIG) { big_list[atomicAdd(n_big, 1u)] = s; return; }
    GatherR6 g{col, vals + a};
    sums[s] = pairwise_sum(g, b - a);
}
__global__ void __launch_bounds__(256) k_seg_pairwise_big(const double* __restrict__ col, const uint32_t* __restrict__ vals,
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
        GatherR6 g{col, vals + a + st};
        part[threadIdx.x] = pairwise_sum(g, l);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int lev = depth - 1; lev >= 0; lev--)
            for (uint32_t i = 0; i < (1u << lev); i++) part[i] = d_add(part[2 * i], part[2 * i + 1]);
        sums[s] = d_add(0.0, part[0]);
    }
}
// keys for ordering segments by their sum, descending; payload = segment index
__global__ void __launch_bounds__(256) k_sum_keys_desc(const double* __restrict__ sums, uint32_t nseg, uint64_t* __restrict__ kout,
                                                       uint32_t* __restrict__ vout) {
    uint32_t s = blockIdx.x * 256 + threadIdx.x;
    if (s >= nseg) return;
    kout[s] = ~f64_ordered(sums[s]);
    vout[s] = s;
}

// names (full, '\0'-terminated) of a list of rows, in list order: lengths, then (after a prefix sum
// on the host - the list is short) the bytes; one thread per row.
template <int KIND>
__global__ void __launch_bounds__(128) k_rows_name_len(SrcView src, const uint32_t* __restrict__ rows, uint32_t nrows, uint32_t* lens) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    uint64_t row = rows[i];
    with_row<KIND, false>(src, row, row, [&](const RowVals&, const auto& ng) { LenSink l; ng.gen(l); lens[i] = (uint32_t)l.len; });
}
template <int KIND>
__global__ void __launch_bounds__(128) k_rows_name_write(SrcView src, const uint32_t* __restrict__ rows, uint32_t nrows,
                                                         const uint64_t* __restrict__ name_off, char* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    uint64_t row = rows[i];
    with_row<KIND, false>(src, row, row, [&](const RowVals&, const auto& ng) {
        RawWrite<char*> w{out + name_off[i], 0}; ng.gen(w);
        out[name_off[i] + w.n] = '\0';
    });
}
// ts / duration (CSV values) of a short list of perf rows
__global__ void k_rows_values(SrcView src, const uint32_t* __restrict__ rows, uint32_t nrows, double* ts, double* dur) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    ts[i] = round6_value(src.c.p.ts[rows[i]]);
    dur[i] = round6_value(src.c.p.dur[rows[i]]);
}

}  // namespace sofa
