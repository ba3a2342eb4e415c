// Analyze primitives on the resident columns (sofa_analyze.py:343-377 gpu_profile,
// sofa_common.py:23-177 comm_profile, sofa_analyze.py:34-44 top-k, :694-710 cpu_profile).
//
// The reference analyses pd.read_csv(<src>trace.csv), i.e. values that went through
// "%.6f": every float is read through round6_value().  Float sums reproduce numpy's
// pairwise tree exactly (stream-compact the group's rows in row order, evaluate 2^d
// sub-trees in parallel, combine in tree order); integer sums and counts use atomics.
#pragma once
#include "kernels_emit.cuh"

namespace sofa {

enum { P_ALL = 0, P_DEV_EQ, P_KIND_EQ, P_KIND_BIG, P_NAMEHASH, P_CELL_BIG };
enum { V_GPU_DUR = 0, V_GPU_BW = 1, V_PERF_DUR = 2 };

struct SelArgs {
    int pred, vsel;
    int64_t a0, a1;
    uint64_t h;
    const uint64_t* namehash;
    int bw_col_is_float;          // gputrace.csv bandwidth column printed as float (else the CSV holds int 0)
};

__device__ __forceinline__ bool sel_pred(const SrcView& s, const SelArgs& a, uint64_t row) {
    switch (a.pred) {
        case P_ALL: return true;
        case P_DEV_EQ: return s.c.g.dev[row] == (int32_t)a.a0;
        case P_KIND_EQ: return s.c.g.kind[row] == (uint8_t)a.a0;
        case P_KIND_BIG: return s.c.g.kind[row] == (uint8_t)a.a0 && s.c.g.payload[row] > 64000;
        case P_NAMEHASH: return a.namehash[row] == a.h;
        default: {
            uint8_t k = s.c.g.kind[row];
            return k != 0 && k != 8 && s.c.g.src[row] == (int32_t)a.a0 && s.c.g.dst[row] == (int32_t)a.a1 &&
                   s.c.g.payload[row] > 64000;
        }
    }
}
__device__ __forceinline__ double sel_value(const SrcView& s, const SelArgs& a, uint64_t row) {
    if (a.vsel == V_GPU_DUR) return round6_value(s.c.g.dur[row]);
    if (a.vsel == V_GPU_BW) return a.bw_col_is_float ? round6_value(s.c.g.bw[row]) : 0.0;
    return round6_value(s.c.p.dur[row]);
}

// Stream compaction of the selected rows, row order kept: a block covers SEL_ROWS consecutive rows,
// four per thread (one block scan per 1024 rows instead of per 256).
constexpr uint32_t SEL_ROWS = 1024;
__global__ void __launch_bounds__(256) k_select_count(SrcView src, uint64_t n, SelArgs a, uint64_t* blockoff) {
    __shared__ uint32_t sm[33];
    uint64_t row0 = (uint64_t)blockIdx.x * SEL_ROWS + (uint64_t)threadIdx.x * 4;
    uint32_t f = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) f += (row0 + k < n && sel_pred(src, a, row0 + k)) ? 1u : 0u;
    uint32_t total;
    block_excl_scan(f, &total, sm);
    if (threadIdx.x == 0) blockoff[blockIdx.x] = total;
}
__global__ void __launch_bounds__(256) k_select_write(SrcView src, uint64_t n, SelArgs a, const uint64_t* blockoff,
                                                      double* __restrict__ out) {
    __shared__ uint32_t sm[33];
    uint64_t row0 = (uint64_t)blockIdx.x * SEL_ROWS + (uint64_t)threadIdx.x * 4;
    bool f[4];
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { f[k] = row0 + k < n && sel_pred(src, a, row0 + k); cnt += f[k] ? 1u : 0u; }
    uint32_t total;
    uint32_t ex = block_excl_scan(cnt, &total, sm);
    uint64_t o = blockoff[blockIdx.x] + ex;
#pragma unroll
    for (int k = 0; k < 4; k++) if (f[k]) out[o++] = sel_value(src, a, row0 + k);
}

// 2^depth sub-trees of numpy's pairwise tree, one thread each; then the combine sweep.
__global__ void __launch_bounds__(128) k_pairwise_top(const double* __restrict__ a, uint64_t n, int depth, double* partial) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (1u << depth)) return;
    uint64_t s, l;
    pairwise_subtree(n, depth, t, &s, &l);
    // pairwise_sum adds the 0.0 identity; harmless on an inner node (x + 0.0 == x for x != -0.0,
    // and -0.0 can only come from an all -0.0 sub-tree, which durations never are)
    partial[t] = pairwise_sum(a + s, l);
}
__global__ void k_pairwise_combine(double* partial, int depth, double* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int lev = depth - 1; lev >= 0; lev--)
        for (uint32_t i = 0; i < (1u << lev); i++) partial[i] = d_add(partial[2 * i], partial[2 * i + 1]);
    *out = d_add(0.0, partial[0]);
}
// one level of the combine sweep, ping-pong buffers: out[i] = in[2i] + in[2i+1]
__global__ void __launch_bounds__(256) k_pairwise_level(const double* __restrict__ in, double* __restrict__ out, uint32_t n_out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_out) out[i] = d_add(in[2 * i], in[2 * i + 1]);
}
__global__ void k_pairwise_finish(const double* in, double* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = d_add(0.0, in[0]);
}
// python-loop accumulation in row order (sofa_common.py:108-119 accum_bw)
__global__ void k_seq_sum(const double* __restrict__ a, uint64_t n, double* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0;
    for (uint64_t i = 0; i < n; i++) s = d_add(s, a[i]);
    *out = s;
}

struct GpuStats {
    unsigned long long dev_hist[4096];     // deviceId + 1 in [0, 4096)
    long long max_dev;
    unsigned long long kind_count[16], big_count[16];
    unsigned long long payload_sum[16];    // two's complement int64 sums per copyKind
    unsigned long long cell_payload[256], cell_count[256], cell_big[256];   // (src, dst) in [0,16)^2
    unsigned int bad_dev, bad_cell;
};

__global__ void __launch_bounds__(256) k_gpu_stats(SrcView src, uint64_t n, GpuStats* st) {
    // block-private accumulators in shared memory, one flush per block
    __shared__ unsigned int hist[4096];
    __shared__ unsigned long long s_kind[16], s_big[16], s_pay[16];
    __shared__ unsigned long long s_cpay[256], s_ccnt[256], s_cbig[256];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) hist[i] = 0;
    if (threadIdx.x < 16) { s_kind[threadIdx.x] = 0; s_big[threadIdx.x] = 0; s_pay[threadIdx.x] = 0; }
    s_cpay[threadIdx.x] = 0; s_ccnt[threadIdx.x] = 0; s_cbig[threadIdx.x] = 0;
    __syncthreads();
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    long long mx = -0x7fffffffffffffffll;
    unsigned int bad_dev = 0, bad_cell = 0;
    for (uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride) {
        int32_t dev = src.c.g.dev[row];
        uint8_t k = src.c.g.kind[row];
        int64_t pl = src.c.g.payload[row];
        if (dev < -1 || dev > 4094) bad_dev = 1; else atomicAdd(&hist[dev + 1], 1u);
        if (dev > mx) mx = dev;
        atomicAdd(&s_kind[k & 15], 1ull);
        if (pl > 64000) atomicAdd(&s_big[k & 15], 1ull);
        if (k != 0) atomicAdd(&s_pay[k & 15], (unsigned long long)pl);
        if (k != 0 && k != 8) {
            int32_t s = src.c.g.src[row], d = src.c.g.dst[row];
            if (s < 0 || s > 15 || d < 0 || d > 15) bad_cell = 1;
            else {
                int c = s * 16 + d;
                atomicAdd(&s_cpay[c], (unsigned long long)pl);
                atomicAdd(&s_ccnt[c], 1ull);
                if (pl > 64000) atomicAdd(&s_cbig[c], 1ull);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) if (hist[i]) atomicAdd(&st->dev_hist[i], (unsigned long long)hist[i]);
    if (threadIdx.x < 16) {
        if (s_kind[threadIdx.x]) atomicAdd(&st->kind_count[threadIdx.x], s_kind[threadIdx.x]);
        if (s_big[threadIdx.x]) atomicAdd(&st->big_count[threadIdx.x], s_big[threadIdx.x]);
        if (s_pay[threadIdx.x]) atomicAdd(&st->payload_sum[threadIdx.x], s_pay[threadIdx.x]);
    }
    if (s_ccnt[threadIdx.x]) {
        atomicAdd(&st->cell_payload[threadIdx.x], s_cpay[threadIdx.x]);
        atomicAdd(&st->cell_count[threadIdx.x], s_ccnt[threadIdx.x]);
        if (s_cbig[threadIdx.x]) atomicAdd(&st->cell_big[threadIdx.x], s_cbig[threadIdx.x]);
    }
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 16)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 4)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    if (lane_id() == 0) atomicMax(&st->max_dev, mx);
    if (bad_dev) atomicOr(&st->bad_dev, 1u);
    if (bad_cell) atomicOr(&st->bad_cell, 1u);
}

// hash + "contains nccl" over the generated name in one pass
struct NcclHashSink {
    static constexpr bool WORDS = false;
    HashSink h; uint32_t w; bool hit;
    __device__ NcclHashSink() : w(0), hit(false) {}
    __device__ __forceinline__ void put(char ch) { h.put(ch); w = (w << 8) | (unsigned char)ch; hit |= (w == 0x6e63636cu); }
    __device__ __forceinline__ void put_word(uint64_t ww, int nb) { put_word_bytes(*this, ww, nb); }
};
// namehash[row] = FNV-1a of the full name when it contains "nccl" (else 0); distinct hashes
// go into an open-addressing table (value = smallest row).
__global__ void __launch_bounds__(256) k_nccl_hash(SrcView src, uint64_t n, uint64_t* __restrict__ namehash,
                                                   unsigned long long* keys, unsigned long long* rows, uint64_t cap_mask,
                                                   unsigned int* overflow, uint64_t hash_mask) {
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    uint64_t key = 0;
    if (src.c.g.kind[row] == 0) {
        GpuRow r = load_gpu(src, row);
        GpuName<const char*> ng{src.text + src.c.g.line_start[row], r, true, src.api_mode != 0};
        NcclHashSink s; ng.gen(s);
        s.h.h &= hash_mask;
        if (s.hit) key = s.h.h ? s.h.h : 1;
        if (s.hit && keys) {
            uint64_t slot = key & cap_mask;
            bool done = false;
            for (uint64_t probe = 0; probe <= cap_mask && probe < 8192; probe++) {
                unsigned long long prev = atomicCAS(&keys[slot], 0ull, (unsigned long long)key);
                if (prev == 0ull || prev == key) { atomicMin(&rows[slot], (unsigned long long)row); done = true; break; }
                slot = (slot + 1) & cap_mask;
            }
            if (!done) atomicOr(overflow, 1u);
        }
    }
    namehash[row] = key;
}
// ---------------------------------------------------------------------------- net_profile (sofa_analyze.py:385-493)
// Rows of nettrace.csv are the packets i % stride == 0.  namehash[row] = FNV-1a of the name for tcp packets
// (the names that contain "network:tcp:"), 0 otherwise; dur[row] = payload / 128e6 as the CSV holds it.
__global__ void __launch_bounds__(256) k_net_tcp_hash(SrcView src, uint64_t n, uint32_t stride, uint64_t* __restrict__ namehash,
                                                      double* __restrict__ dur) {
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    uint64_t key = 0;
    NetRow r = load_net(src, row);
    if (row % stride == 0 && r.proto == 2) {
        NetName ng{r};
        HashSink h; ng.gen(h);
        key = h.h ? h.h : 1;
    }
    namehash[row] = key;
    dur[row] = d_div((double)r.payload, 128.0e6);
}
// Order of the names "network:tcp:%d_to_%d_with_%d" as strings: compare src, dst, payload as decimal STRINGS.
// Inside the name a number is followed by '_' (greater than every digit), so a shorter number that is a prefix
// of a longer one sorts after it; the payload ends the string, so there the shorter one sorts first.
// strkey = the decimal digits left-aligned in 13 base-11 places, pad = 10 (inner numbers) or 0 with digits
// shifted to 1..10 (last number): integer order of the key = string order of the number.
__device__ __forceinline__ uint64_t net_strkey(int64_t v, bool last) {
    uint64_t u = v < 0 ? 0 : (uint64_t)v;                // (addresses and payloads of tcp rows are never negative)
    int nd = ndigits_u64(u);
    uint64_t key = 0;
    uint64_t p = 1;
    for (int i = 1; i < nd; i++) p *= 10;
    for (int i = 0; i < 13; i++) {
        uint64_t d;
        if (i < nd) { uint64_t q = u / p; d = (last ? 1 : 0) + q; u -= q * p; p /= 10; }
        else d = last ? 0 : 10;
        key = key * 11 + d;
    }
    return key;
}
// keys[i] = string-order key of one name component of packet rows[i]: which = 0 payload, 1 dst, 2 src
__global__ void __launch_bounds__(256) k_net_key(SrcView src, const uint32_t* __restrict__ rows, uint64_t m, int which,
                                                 uint64_t* __restrict__ keys) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint64_t row = rows[i];
    int64_t v = which == 0 ? src.c.n.payload[row] : (which == 1 ? src.c.n.dst[row] : src.c.n.src[row]);
    keys[i] = net_strkey(v, which == 0);
}
// segment heads of the packets sorted by name: (src, dst, payload) differs from the previous packet's
__device__ __forceinline__ bool net_new_group(const SrcView& s, const uint32_t* rows, uint64_t i) {
    if (i == 0) return true;
    uint64_t a = rows[i], b = rows[i - 1];
    return s.c.n.src[a] != s.c.n.src[b] || s.c.n.dst[a] != s.c.n.dst[b] || s.c.n.payload[a] != s.c.n.payload[b];
}
__global__ void __launch_bounds__(256) k_net_seg_count(SrcView src, const uint32_t* __restrict__ rows, uint64_t m, uint64_t* blockoff) {
    __shared__ uint32_t sm[33];
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t total;
    block_excl_scan((i < m && net_new_group(src, rows, i)) ? 1u : 0u, &total, sm);
    if (threadIdx.x == 0) blockoff[blockIdx.x] = total;
}
__global__ void __launch_bounds__(256) k_net_seg_write(SrcView src, const uint32_t* __restrict__ rows, uint64_t m, const uint64_t* blockoff,
                                                       uint32_t* __restrict__ seg_start) {
    __shared__ uint32_t sm[33];
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    bool f = i < m && net_new_group(src, rows, i);
    uint32_t total;
    uint32_t ex = block_excl_scan(f ? 1u : 0u, &total, sm);
    if (f) seg_start[blockoff[blockIdx.x] + ex] = (uint32_t)i;
}

// payload summed per (pkt_src, pkt_dst): open addressing on a 64-bit tag of the pair; every slot also keeps
// min / max of src and dst so that the host can prove that no two pairs shared a tag (exact, or a loud error).
struct NetPairSlot { unsigned long long tag; long long sum, src_min, src_max, dst_min, dst_max; unsigned long long count; };
__global__ void __launch_bounds__(256) k_net_pairs(SrcView src, uint64_t n, uint32_t stride, NetPairSlot* tab, uint64_t cap_mask,
                                                   unsigned long long* total, unsigned int* overflow) {
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n || row % stride) return;
    NetRow r = load_net(src, row);
    if (r.proto == 0) return;                                  // the all-NaN row of a non-IP line
    uint64_t h = 1469598103934665603ull;
    h = (h ^ (uint64_t)r.src) * 1099511628211ull; h ^= h >> 29;
    h = (h ^ (uint64_t)r.dst) * 1099511628211ull; h ^= h >> 32;
    unsigned long long tag = h | 1ull;
    uint64_t slot = (h >> 7) & cap_mask;
    for (uint64_t probe = 0; probe <= cap_mask && probe < 65536; probe++) {
        unsigned long long prev = atomicCAS(&tab[slot].tag, 0ull, tag);
        if (prev == 0ull || prev == tag) {
            atomicAdd((unsigned long long*)&tab[slot].sum, (unsigned long long)r.payload);
            atomicAdd(&tab[slot].count, 1ull);
            atomicMin(&tab[slot].src_min, (long long)r.src); atomicMax(&tab[slot].src_max, (long long)r.src);
            atomicMin(&tab[slot].dst_min, (long long)r.dst); atomicMax(&tab[slot].dst_max, (long long)r.dst);
            atomicAdd(total, (unsigned long long)r.payload);
            return;
        }
        slot = (slot + 1) & cap_mask;
    }
    atomicOr(overflow, 1u);
}
__global__ void k_net_pairs_init(NetPairSlot* tab, uint64_t cap) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    NetPairSlot s; s.tag = 0; s.sum = 0; s.count = 0;
    s.src_min = s.dst_min = 0x7fffffffffffffffll; s.src_max = s.dst_max = -0x7fffffffffffffffll - 1;
    tab[i] = s;
}

// names of the table's representative rows, each followed by '\0' (gpu names may hold '\n');
// order[] receives the slot of every exported name in output order.
template <int KIND>
__global__ void __launch_bounds__(256) k_table_names(SrcView src, const unsigned long long* keys, const unsigned long long* rows,
                                                     uint64_t cap, unsigned long long* cursor, unsigned int* n_out, char* out,
                                                     uint64_t out_cap, uint64_t* name_off, uint32_t* name_slot) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap || keys[s] == 0ull) return;
    uint64_t row = rows[s];
    with_row<KIND, false>(src, row, row, [&](const RowVals&, const auto& ng) {
        LenSink l; ng.gen(l);
        unsigned long long off = atomicAdd(cursor, (unsigned long long)(l.len + 1));
        unsigned int idx = atomicAdd(n_out, 1u);
        if (out && off + l.len + 1 <= out_cap) {
            RawWrite<char*> w{out + off, 0}; ng.gen(w);
            out[off + l.len] = '\0';
            name_off[idx] = off; name_slot[idx] = (uint32_t)s;
        }
    });
}

}  // namespace sofa
