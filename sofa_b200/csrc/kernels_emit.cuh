// Emit: columns (+ name spans in the resident text) -> CSV rows / report.js records,
// keyword filters, dtype flags, viz-name dictionary.
//
// Text emission is a size pass, a scan of per-block byte totals and a write pass.  The
// write pass formats each row into a shared-memory image of the block's output range
// (same 16-byte phase as the global destination) and flushes it with 128-bit coalesced
// stores; blocks whose range exceeds the staging buffer write straight to global memory.
#pragma once
#include "kernels_ingest.cuh"

namespace sofa {

struct DictView { int n; const uint64_t* hash; const uint64_t* off; const char* blob; };

struct SrcView {
    const char* text;
    uint64_t n_rows;
    AllCols c;
    const int32_t* viz_id;      // per down-sampled perf candidate: -1 raw, else dictionary slot
    DictView dict;
    int32_t n_cudaproc;
};

template <typename S>
struct DevVizName {
    PerfVizName<S> raw; int id; DictView d;
    template <class K> __device__ __forceinline__ void gen(K& k) const {
        if (id >= 0) { for (uint64_t j = d.off[id]; j < d.off[id + 1]; j++) k.put(d.blob[j]); }
        else raw.gen(k);
    }
};

__device__ __forceinline__ PerfRow load_perf(const SrcView& s, uint64_t row) {
    PerfRow r;
    r.ts = s.c.p.ts[row]; r.event = s.c.p.event[row]; r.dur = s.c.p.dur[row];
    r.pid = s.c.p.pid[row]; r.tid = s.c.p.tid[row];
    const uint16_t* pc = s.c.p.pieces + row * 6;
    r.evt_off = pc[0]; r.evt_len = pc[1]; r.sym_off = pc[2]; r.sym_len = pc[3]; r.dso_off = pc[4]; r.dso_len = pc[5];
    return r;
}
__device__ __forceinline__ GpuRow load_gpu(const SrcView& s, uint64_t row) {
    GpuRow r;
    r.ts = s.c.g.ts[row]; r.dur = s.c.g.dur[row]; r.bw = s.c.g.bw[row]; r.payload = s.c.g.payload[row];
    r.dev = s.c.g.dev[row]; r.tid = s.c.g.tid[row]; r.src = s.c.g.src[row]; r.dst = s.c.g.dst[row];
    r.kind = s.c.g.kind[row]; r.bw_is_float = s.c.g.bwf[row];
    r.name_off = s.c.g.noff[row]; r.name_len = s.c.g.nlen[row];
    return r;
}
__device__ __forceinline__ NetRow load_net(const SrcView& s, uint64_t row) {
    NetRow r;
    r.ts = s.c.n.ts[row]; r.payload = s.c.n.payload[row]; r.src = s.c.n.src[row]; r.dst = s.c.n.dst[row];
    r.proto = s.c.n.proto[row];
    return r;
}

// Calls f(RowVals, NameGen) for row `row` (candidate index `cand`); VIZ selects the
// down-sampled/demangled perf name used by report.js and the cpu filters.
template <int KIND, bool VIZ, class F>
__device__ __forceinline__ void with_row(const SrcView& s, uint64_t row, uint64_t cand, F&& f) {
    if constexpr (KIND == SRC_PERF) {
        PerfRow r = load_perf(s, row);
        const char* line = s.text + s.c.p.line_start[row];
        RowVals v = vals_perf(r);
        PerfName<const char*> full{line, r};
        if constexpr (VIZ) {
            DevVizName<const char*> ng{PerfVizName<const char*>{full}, s.viz_id ? s.viz_id[cand] : -1, s.dict};
            f(v, ng);
        } else {
            f(v, full);
        }
    } else if constexpr (KIND == SRC_GPU) {
        GpuRow r = load_gpu(s, row);
        const char* line = s.text + s.c.g.line_start[row];
        RowVals v = vals_gpu(r, s.n_cudaproc);
        GpuName<const char*> ng{line, r};
        f(v, ng);
    } else {
        NetRow r = load_net(s, row);
        RowVals v = vals_net(r);
        NetName ng{r};
        f(v, ng);
    }
}

enum { FMT_CSV = 0, FMT_JSON = 1 };

struct EmitArgs {
    uint64_t ncand; uint32_t stride; uint32_t fmask;
    const uint8_t* mask; int bit;          // optional row filter (mask[cand] >> bit) & 1
    uint32_t* rowlen;                      // len | special << 31
    uint64_t* blockoff;                    // per block byte total -> (after scan) byte offset
    char* out;
    unsigned long long* err;
    unsigned long long* count;             // selected rows
};

template <int KIND, int FMT>
__global__ void __launch_bounds__(EMIT_TPB) k_emit_size(SrcView src, EmitArgs a) {
    __shared__ uint32_t sm[33];
    uint64_t c = (uint64_t)blockIdx.x * EMIT_TPB + threadIdx.x;
    bool sel = c < a.ncand && (a.mask == nullptr || ((a.mask[c] >> a.bit) & 1));
    uint32_t len = 0, special = 0;
    if (sel) {
        uint64_t row = c * a.stride;
        int l = 0;
        with_row<KIND, FMT == FMT_JSON>(src, row, c, [&](const RowVals& v, const auto& ng) {
            if constexpr (FMT == FMT_CSV) { bool sp; l = csv_row_len(v, a.fmask, ng, &sp); special = sp ? 1u : 0u; }
            else l = json_row_len(v, a.fmask, ng);
        });
        if (l < 0) { report_error(a.err, row, l == -2 ? E_UTF8 : E_FORMAT); l = 0; }
        len = (uint32_t)l;
    }
    if (c < a.ncand) a.rowlen[c] = len | (special << 31);
    uint32_t total, cnt_total;
    block_excl_scan(len, &total, sm);
    block_excl_scan(sel ? 1u : 0u, &cnt_total, sm);
    if (threadIdx.x == 0) {
        a.blockoff[blockIdx.x] = total;
        if (cnt_total) atomicAdd(a.count, (unsigned long long)cnt_total);
    }
}

template <int KIND, int FMT>
__global__ void __launch_bounds__(EMIT_TPB) k_emit_write(SrcView src, EmitArgs a) {
    __shared__ __align__(16) unsigned char stage[EMIT_STAGE_BYTES + 32];
    __shared__ uint32_t sm[33];
    uint64_t c = (uint64_t)blockIdx.x * EMIT_TPB + threadIdx.x;
    uint32_t rl = c < a.ncand ? a.rowlen[c] : 0;
    uint32_t len = rl & 0x7fffffffu;
    bool special = (rl >> 31) != 0;
    uint32_t total;
    uint32_t ex = block_excl_scan(len, &total, sm);
    uint64_t tile_off = a.blockoff[blockIdx.x];
    bool staged = total <= EMIT_STAGE_BYTES;
    if (len) {
        char* dst = staged ? reinterpret_cast<char*>(stage) + (tile_off & 15) + ex : a.out + tile_off + ex;
        uint64_t row = c * a.stride;
        with_row<KIND, FMT == FMT_JSON>(src, row, c, [&](const RowVals& v, const auto& ng) {
            if constexpr (FMT == FMT_CSV) csv_row_write(v, a.fmask, ng, special, dst);
            else json_row_write(v, a.fmask, ng, dst);
        });
    }
    __syncthreads();
    if (staged && total) flush_staged(stage, a.out, tile_off, total);
}

// ---------------------------------------------------------------------------- dtype flags
template <int KIND>
__global__ void __launch_bounds__(256) k_row_flags(SrcView src, uint64_t ncand, uint32_t stride, unsigned int* flags) {
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int f = 0;
    if (c < ncand) {
        uint64_t row = c * stride;
        if (KIND == SRC_GPU) f = src.c.g.bwf[row] ? 1u : 0u;
        else if (KIND == SRC_NET) f = src.c.n.proto[row] == 0 ? 2u : 4u;
    }
    f = __reduce_or_sync(0xffffffffu, f);
    if (lane_id() == 0 && f) atomicOr(flags, f);
}

// ---------------------------------------------------------------------------- keyword filters
template <int KIND>
__global__ void __launch_bounds__(256) k_filter(SrcView src, uint64_t ncand, uint32_t stride, const FilterTable* __restrict__ ft,
                                                uint8_t* __restrict__ mask, unsigned long long* counts) {
    __shared__ FilterTable sft;
    {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(ft);
        uint32_t* s = reinterpret_cast<uint32_t*>(&sft);
        for (uint32_t i = threadIdx.x; i < sizeof(FilterTable) / 4; i += blockDim.x) s[i] = g[i];
    }
    __syncthreads();
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t hit = 0;
    if (c < ncand) {
        uint64_t row = c * stride;
        with_row<KIND, true>(src, row, c, [&](const RowVals& v, const auto& ng) {
            if (v.valid) { MatchSink ms(&sft); ng.gen(ms); hit = ms.hit; }
        });
        mask[c] = (uint8_t)hit;
    }
    for (int k = 0; k < sft.nk; k++) {
        unsigned b = __ballot_sync(0xffffffffu, (hit >> k) & 1);
        if (lane_id() == 0 && b) atomicAdd(&counts[k], (unsigned long long)__popc(b));
    }
}

// ---------------------------------------------------------------------------- viz-name dictionary
// Hash every down-sampled perf row's raw viz name; mangled ones (_Z...) go into an
// open-addressing table (key = FNV-1a 64, value = smallest row) for host demangling.
__global__ void __launch_bounds__(256) k_viz_hash(SrcView src, uint64_t ncand, uint32_t stride, uint64_t* __restrict__ vizhash,
                                                  unsigned long long* keys, unsigned long long* rows, uint64_t cap_mask,
                                                  unsigned int* overflow) {
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand) return;
    uint64_t row = c * stride;
    PerfRow r = load_perf(src, row);
    const char* line = src.text + src.c.p.line_start[row];
    PerfVizName<const char*> vz{PerfName<const char*>{line, r}};
    HashSink h; vz.gen(h);
    uint64_t key = 0;
    if (h.len >= 2 && h.c0 == '_' && h.c1 == 'Z') {
        key = h.h ? h.h : 1;
        uint64_t slot = key & cap_mask;
        bool done = false;
        for (uint64_t probe = 0; probe <= cap_mask; probe++) {
            unsigned long long prev = atomicCAS(&keys[slot], 0ull, (unsigned long long)key);
            if (prev == 0ull || prev == key) { atomicMin(&rows[slot], (unsigned long long)row); done = true; break; }
            slot = (slot + 1) & cap_mask;
            if (probe > 4096) break;
        }
        if (!done) atomicOr(overflow, 1u);
    }
    vizhash[c] = key;
}
// total bytes ("name\n" each) and count of occupied slots
__global__ void __launch_bounds__(256) k_dict_len(SrcView src, const unsigned long long* keys, const unsigned long long* rows,
                                                  uint64_t cap, unsigned long long* total, unsigned int* n) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap || keys[s] == 0ull) return;
    uint64_t row = rows[s];
    PerfRow r = load_perf(src, row);
    PerfVizName<const char*> vz{PerfName<const char*>{src.text + src.c.p.line_start[row], r}};
    LenSink l; vz.gen(l);
    atomicAdd(total, (unsigned long long)(l.len + 1));
    atomicAdd(n, 1u);
}
__global__ void __launch_bounds__(256) k_dict_write(SrcView src, const unsigned long long* keys, const unsigned long long* rows,
                                                    uint64_t cap, unsigned long long* cursor, char* out) {
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap || keys[s] == 0ull) return;
    uint64_t row = rows[s];
    PerfRow r = load_perf(src, row);
    PerfVizName<const char*> vz{PerfName<const char*>{src.text + src.c.p.line_start[row], r}};
    LenSink l; vz.gen(l);
    unsigned long long off = atomicAdd(cursor, (unsigned long long)(l.len + 1));
    RawWrite<char*> w{out + off, 0}; vz.gen(w);
    out[off + l.len] = '\n';
}
__global__ void __launch_bounds__(256) k_viz_resolve(const uint64_t* __restrict__ vizhash, uint64_t ncand, DictView d,
                                                     int32_t* __restrict__ viz_id, unsigned long long* err) {
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand) return;
    uint64_t key = vizhash[c];
    int32_t id = -1;
    if (key) {
        int lo = 0, hi = d.n - 1;
        while (lo <= hi) {
            int mid = (lo + hi) >> 1;
            uint64_t hm = d.hash[mid];
            if (hm == key) { id = mid; break; }
            if (hm < key) lo = mid + 1; else hi = mid - 1;
        }
        if (id < 0) report_error(err, c, E_FORMAT);
    }
    viz_id[c] = id;
}

}  // namespace sofa
