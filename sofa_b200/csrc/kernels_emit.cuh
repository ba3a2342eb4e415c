// Emit: columns (+ name spans in the resident text) -> CSV rows / report.js records,
// keyword filters, dtype flags, viz-name dictionary.
//
// Text emission is one kernel per output (k_emit_fused, below): digit plans + exact row length,
// decoupled look-back for the block's byte offset, rows printed into a packed shared-memory
// image (same 16-byte phase as the global destination), 128-bit coalesced flush; blocks whose
// range exceeds the image write straight to global memory.
#pragma once
#include "kernels_ingest.cuh"
#include "kernels_side.cuh"

namespace sofa {

struct DictView { int n; const uint64_t* hash; const uint64_t* off; const char* blob; };

// rows derived from a recorder side file: candidate c + base addresses row (c + base) % n_half of
// list (c + base) / n_half (nvsmi: sm, mem; netstat: tx, rx; mpstat: one list)
struct SideView { int table, absolute; uint64_t n_half, base; SideOutCols o; };

struct SrcView {
    SideView side;
    const char* text;
    uint64_t n_rows;
    AllCols c;
    const int32_t* viz_id;      // per down-sampled perf candidate: -1 raw, else dictionary slot
    DictView dict;
    int32_t n_cudaproc;
    int32_t api_mode;
};

template <typename S>
struct DevVizName {
    PerfVizName<S> raw; int id; DictView d;
    template <class K> __device__ __forceinline__ void gen(K& k) const {
        const PerfRow& r = raw.full.r;
        if (id < 0 && r.va_len == 0xffff) { raw.gen(k); return; }        // rare: the slice was not cached at parse
        // one code path for demangled (dictionary) and raw names - two spans, the second empty for a
        // dictionary entry - so that the lanes of a warp stay together
        const char* p0 = raw.full.line + r.sym_off;
        int l0 = r.va_len, l1 = r.vb_len;
        if (id >= 0) { p0 = d.blob + d.off[id]; l0 = (int)(d.off[id + 1] - d.off[id]); l1 = 0; }
        put_span(k, p0, l0, true);
        put_span(k, raw.full.line + r.dso_off, l1, true);
    }
};

__device__ __forceinline__ PerfRow load_perf(const SrcView& s, uint64_t row) {
    PerfRow r;
    r.ts = s.c.p.ts[row]; r.event = s.c.p.event[row]; r.dur = s.c.p.dur[row];
    r.pid = s.c.p.pid[row]; r.tid = s.c.p.tid[row];
    const uint4 pv = __ldg(reinterpret_cast<const uint4*>(s.c.p.pieces) + row);       // 8 x u16
    r.evt_off = (uint16_t)(pv.x & 0xffff); r.evt_len = (uint16_t)(pv.x >> 16);
    r.sym_off = (uint16_t)(pv.y & 0xffff); r.sym_len = (uint16_t)(pv.y >> 16);
    r.dso_off = (uint16_t)(pv.z & 0xffff); r.dso_len = (uint16_t)(pv.z >> 16);
    r.va_len = (uint16_t)(pv.w & 0xffff); r.vb_len = (uint16_t)(pv.w >> 16);
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
        PerfName<const char*> full{line, r, true};
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
        GpuName<const char*> ng{line, r, true, s.api_mode != 0};
        f(v, ng);
    } else if constexpr (KIND == SRC_SIDEOUT) {
        const uint64_t cc = cand + s.side.base;
        const int list = (int)(cc / s.side.n_half);
        SideName ng{s.side.table, list, load_sideout(s.side.o, cc - (uint64_t)list * s.side.n_half)};
        RowVals v = vals_side(s.side.table, list, s.side.absolute, ng.r);
        f(v, ng);
    } else {
        NetRow r = load_net(s, row);
        RowVals v = vals_net(r);
        NetName ng{r};
        f(v, ng);
    }
}

enum { FMT_CSV = 0, FMT_JSON = 1, FMT_COMM = 2, FMT_AUX = 3 };     // AUX: usr_sys.csv / netbandwidth.csv rows

// ---------------------------------------------------------------------------- dtype flags
template <int KIND>
__global__ void __launch_bounds__(256) k_row_flags(SrcView src, uint64_t ncand, uint32_t stride, unsigned int* flags) {
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int f = 0;
    if (c < ncand) {
        uint64_t row = c * stride;
        if (KIND == SRC_GPU) f = (src.c.g.bwf[row] & 1) ? 1u : 0u;
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
            if (v.valid) { MatchSink ms(&sft); ng.gen(ms); hit = ms.hits(); }
        });
        mask[c] = (uint8_t)hit;
    }
    for (int k = 0; k < sft.nk; k++) {
        unsigned b = __ballot_sync(0xffffffffu, (hit >> k) & 1);
        if (lane_id() == 0 && b) atomicAdd(&counts[k], (unsigned long long)__popc(b));
    }
}

#ifdef SOFA_K_VIZ
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
    PerfVizName<const char*> vz{PerfName<const char*>{line, r, true}};
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
    PerfVizName<const char*> vz{PerfName<const char*>{src.text + src.c.p.line_start[row], r, true}};
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
    PerfVizName<const char*> vz{PerfName<const char*>{src.text + src.c.p.line_start[row], r, true}};
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

#endif  // SOFA_K_VIZ

}  // namespace sofa

// ============================================================================ single-pass emit
// One kernel per output: every thread formats its row ONCE into a private slot of dynamic
// shared memory (st.shared byte stores), a block scan of the exact lengths packs the rows,
// a decoupled look-back over per-block totals (dynamic block tickets, so predecessors are
// always resident) yields the block's global byte offset, and each warp copies its rows out
// with coalesced byte stores.  No size pass, no per-row length array.
namespace sofa {

struct SmemRef {
    uint32_t a;
    __device__ __forceinline__ void operator=(char ch) const {
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"((int)(unsigned char)ch) : "memory");
    }
};
struct SmemPtr {
    uint32_t a;
    __device__ __forceinline__ SmemRef operator[](int i) const { return SmemRef{a + (uint32_t)i}; }
    __device__ __forceinline__ SmemPtr operator+(int i) const { return SmemPtr{a + (uint32_t)i}; }
};

struct FusedArgs {
    uint64_t ncand; uint32_t stride; uint32_t fmask;
    const uint8_t* mask; int bit;
    const uint32_t* index;         // optional compacted candidate list (ncand = its length)
    char* out; uint64_t cap;
    unsigned long long* state;     // per block: flag << 62 | bytes   (1 = aggregate, 2 = inclusive prefix)
    unsigned int* ticket;
    unsigned long long* err;
    unsigned long long* count;
    unsigned long long* total;
    unsigned int* overflow;
    uint32_t slot_bytes;           // dynamic shared memory available for slots
    // report.js main series only: the byte offset of every record, so that the filter series can be gathered from
    // this output instead of being formatted again (their records are the same bytes).  (Matching the filter keywords
    // inside this kernel as well was measured: 13.4 ms of extra time in here against 10.9 ms for the separate
    // k_filter pass - 80 instead of 72 registers - so the masks come from k_filter.)
    unsigned long long* rec_off;   // [ncand + 1] byte offset of each record in `out`; null: not recorded
};

constexpr unsigned long long LB_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// candidate indices whose filter bit is set, ascending (two-pass compaction)
static __global__ void __launch_bounds__(256) k_mask_count(const uint8_t* __restrict__ mask, int bit, uint64_t n, uint64_t* blockoff) {
    __shared__ uint32_t sm[33];
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t total;
    block_excl_scan((i < n && ((mask[i] >> bit) & 1)) ? 1u : 0u, &total, sm);
    if (threadIdx.x == 0) blockoff[blockIdx.x] = total;
}
static __global__ void __launch_bounds__(256) k_mask_write(const uint8_t* __restrict__ mask, int bit, uint64_t n, const uint64_t* blockoff,
                                                    uint32_t* __restrict__ index) {
    __shared__ uint32_t sm[33];
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    bool f = i < n && ((mask[i] >> bit) & 1);
    uint32_t total;
    uint32_t ex = block_excl_scan(f ? 1u : 0u, &total, sm);
    if (f) index[blockoff[blockIdx.x] + ex] = (uint32_t)i;
}

template <int FMT> struct RowPlanOf { typedef CsvRowPlan type; };
template <> struct RowPlanOf<FMT_JSON> { typedef JsonRowPlan type; };
template <> struct RowPlanOf<FMT_COMM> { typedef CommRowPlan type; };
template <> struct RowPlanOf<FMT_AUX> { typedef AuxRowPlan type; };

template <int KIND, int FMT>
__global__ void __launch_bounds__(EMIT_TPB) k_emit_fused(SrcView src, FusedArgs a) {
    extern __shared__ __align__(16) unsigned char stage[];      // packed image of this block's output range
    __shared__ uint32_t sm[33];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_base;
    if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t b = s_tile;
    const uint64_t ci = (uint64_t)b * EMIT_TPB + threadIdx.x;
    const bool inr = ci < a.ncand;
    const uint64_t c = a.index ? (inr ? a.index[ci] : 0) : ci;
    const bool sel = inr && (a.index != nullptr || a.mask == nullptr || ((a.mask[c] >> a.bit) & 1));
    const uint64_t row = c * a.stride;
    // ---- phase 1: digit plans + exact row length (every float is analysed once)
    typename RowPlanOf<FMT>::type plan;
    plan.len = 0;
    if (sel) {
        with_row<KIND, FMT == FMT_JSON>(src, row, c, [&](const RowVals& v, const auto& ng) {
            if constexpr (FMT == FMT_CSV) plan = csv_row_plan(v, a.fmask, ng);
            else if constexpr (FMT == FMT_COMM) plan = comm_row_plan(c, v, a.fmask);
            else if constexpr (FMT == FMT_AUX) plan = aux_row_plan(ng.r.ts, ng.r.f1, ng.r.f2);
            else plan = json_row_plan(v, a.fmask, ng);
        });
        if (plan.len < 0) { report_error(a.err, row, plan.len == -2 ? E_UTF8 : E_FORMAT); plan.len = 0; }
    }
    const uint32_t len = (uint32_t)plan.len;
    uint32_t total, cnt_total;
    const uint32_t ex = block_excl_scan(len, &total, sm);
    block_excl_scan(len ? 1u : 0u, &cnt_total, sm);
    // ---- phase 2: global byte offset of this block (decoupled look-back, one warp)
    if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        unsigned long long base = 0;
        if (b > 0) {
            if (lane == 0) st_release_u64(&a.state[b], (1ull << 62) | (unsigned long long)total);
            long long j = (long long)b - 1;
            while (true) {
                long long idx = j - (long long)lane;
                unsigned long long v = idx >= 0 ? ld_acquire_u64(&a.state[idx]) : (2ull << 62);
                unsigned int f = (unsigned int)(v >> 62);
                unsigned int pmask = __ballot_sync(0xffffffffu, f == 2);
                unsigned int emask = __ballot_sync(0xffffffffu, f == 0);
                int first = pmask ? (__ffs(pmask) - 1) : 32;
                unsigned int need = first >= 31 ? 0xffffffffu : ((1u << (first + 1)) - 1u);
                if (emask & need) continue;                      // a needed predecessor has not published yet
                unsigned long long contrib = ((int)lane <= first) ? (v & LB_MASK) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
                base += contrib;
                if (first < 32) break;
                j -= 32;
            }
        }
        if (lane == 0) {
            st_release_u64(&a.state[b], (2ull << 62) | (base + total));
            s_base = base;
            if (cnt_total) atomicAdd(a.count, (unsigned long long)cnt_total);
            if ((uint64_t)b + 1 == (a.ncand + EMIT_TPB - 1) / EMIT_TPB) *a.total = base + total;
        }
    }
    __syncthreads();
    const unsigned long long base = s_base;
    if constexpr (FMT == FMT_JSON) {
        if (a.rec_off) {
            if (inr) a.rec_off[ci] = base + ex;
            if (ci + 1 == a.ncand) a.rec_off[a.ncand] = base + ex + len;
        }
    }
    if (base + total > a.cap) { if (threadIdx.x == 0) atomicOr(a.overflow, 1u); return; }
    // ---- phase 3: print the row at its final position (shared-memory image, or global when the
    // block's range does not fit)
    const bool staged = total + 16 <= a.slot_bytes;
    if (len) {
        with_row<KIND, FMT == FMT_JSON>(src, row, c, [&](const RowVals& v, const auto& ng) {
            if (staged) {
                SmemPtr dst{(uint32_t)__cvta_generic_to_shared(stage) + (uint32_t)(base & 15) + ex};
                if constexpr (FMT == FMT_CSV) csv_row_emit(v, a.fmask, ng, plan, dst);
                else if constexpr (FMT == FMT_COMM) comm_row_emit(c, v, a.fmask, plan, dst);
                else if constexpr (FMT == FMT_AUX) aux_row_emit(plan, dst);
                else json_row_emit(v, a.fmask, ng, plan, dst);
            } else {
                char* dst = a.out + base + ex;
                if constexpr (FMT == FMT_CSV) csv_row_emit(v, a.fmask, ng, plan, dst);
                else if constexpr (FMT == FMT_COMM) comm_row_emit(c, v, a.fmask, plan, dst);
                else if constexpr (FMT == FMT_AUX) aux_row_emit(plan, dst);
                else json_row_emit(v, a.fmask, ng, plan, dst);
            }
        });
    }
    // ---- phase 4: the image goes out through one bulk-asynchronous store (block-uniform condition)
    if (staged && total) flush_staged_bulk(stage, a.out, base, total);
}


}  // namespace sofa
