// Host-side context of libsofa_b200.so shared by the translation units (one per kernel family so that
// nvcc compiles them in parallel): device buffers, streams, small-transfer staging, launch timing.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/sofa_b200.h"
#include "kernels_emit.cuh"

using namespace sofa;

extern thread_local std::string g_last_error;

enum {
    K_COUNT = 0, K_SCAN, K_PARSE, K_FLAGS, K_CSV_WRITE, K_JSON_WRITE, K_FILTER,
    K_VIZ, K_ANALYZE, K_SORT, K_SYNTH, K_GATHER, K_NKINDS
};

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return -1; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

struct Source {
    DevBuf text; uint64_t nbytes = 0, data_off = 0;
    bool uploaded = false, parsed = false;
    uint64_t n_rows = 0;
    DevBuf cols[16];
    IngestCfg cfg;
    DevBuf mhz;
    DevBuf mask; uint32_t mask_stride = 0; int nkw = 0;      // filter masks
    uint64_t filter_counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t flags_stride[2] = {0, 0}, flags_val[2] = {0, 0};      // cached sofa_b200_row_flags results
    DevBuf vizhash, viz_id; uint32_t viz_stride = 0; bool viz_ready = false;
    DevBuf dict_hash, dict_off, dict_blob; int dict_n = 0;
    // report.js main series kept resident for gathering the filter series (sofa_b200_emit_json_series)
    DevBuf json_main, rec_off; const char* gather_ptr = nullptr; uint64_t gather_gen = 0; bool gather_ok = false;
    uint32_t gather_stride = 0; double gather_total = 0;
    bool side_ready = false; int side_absolute = 0; uint64_t side_ncand = 0;          // raw side source: rows per list after sofa_b200_side_preprocess
};

struct TimedEvent { int kind; cudaEvent_t a, b; };

// diskstat / blktrace / strace / pystacks (capi_side2.inc): one table at a time, preprocess -> emits
struct Side2State {
    DevBuf text, lstart, raw, mask, idx, lens, tab;
    int table = -1; bool ready = false, devs_set = false;
    uint64_t nbytes = 0, n_lines = 0, n = 0, n_kept = 0; int last_has_nl = 1;
    int absolute = 0; double time_base = 0; int64_t pid = 0, cat = 0; int cat_is_str = 0; uint32_t cat_a = 0, cat_b = 0;
    uint32_t n_dev = 0, disk_fl = 0; int dur_all_int = 0;
    void release() { text.release(); lstart.release(); raw.release(); mask.release(); idx.release(); lens.release(); tab.release(); }
};

struct sofa_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    Source src[11];                      // 0 perf, 1 gpu, 2 net, 3..5 side tables (mpstat, nvsmi, netstat CSV), 6..10 raw recorder side files
    DevBuf out[3]; uint64_t out_n[3] = {0, 0, 0};
    // overlapped transfers: copy streams, the idle halves of the csv/json double buffers and the
    // D2H-complete event of each half
    cudaStream_t h2d = nullptr, d2h = nullptr;
    DevBuf out_alt[2];
    cudaEvent_t out_ev[2] = {nullptr, nullptr}, out_alt_ev[2] = {nullptr, nullptr};
    bool out_busy[2] = {false, false}, out_alt_busy[2] = {false, false};
    cudaEvent_t up_ev[11] = {};
    bool up_pending[11] = {};
    cudaEvent_t fetch_ev[16]; bool fetch_used[16]; int fetch_next = 0;
    DevBuf arena; uint64_t arena_cur = 0; bool arena_on = false;
    uint64_t arena_gen = 1;              // bumped whenever the arena wraps or is re-created (older pointers into it are dead)
    uint64_t async_fetches = 0;          // asynchronous fetches issued since the D2H stream was last drained
    // small host<->device traffic goes through mapped pinned memory + a copy kernel, never the
    // DMA engines (a tiny read-back queued behind a multi-GB transfer would stall the host)
    unsigned char* stage_h = nullptr; unsigned char* stage_d = nullptr; size_t stage_cap = 8u << 20, stage_cur = 0;
    struct SmallXfer { void* dst; size_t off, n; };
    std::vector<SmallXfer> pending;
    char* out_ptr[3] = {nullptr, nullptr, nullptr};
    DevBuf tiles, rowlen, blockoff, scratch, ftable, keys, rows;
    Side2State s2;
    std::vector<DevBuf> held; std::vector<uint64_t> held_n;      // outputs parked on the device until report.js is written (sofa_b200_hold_output)
    DevBuf gather[17];                   // rank 0 of a sharded trace: text + columns being gathered (sofa_b200_shard_reserve)
    DevBuf sk0, sv0, sk1, sv1, hist, segs, sums, small, scanpart;
    void* pin_h = nullptr; size_t pin_cap = 0;  // pinned host scratch for larger read-backs (grown on demand)
    unsigned long long* d_err = nullptr;        // [0] error slot, [1] counter, [2..9] filter counts, [10] misc
    unsigned int* d_flags = nullptr;            // [0] cr flag, [1] row flags, [2] overflow, [3] count
    uint64_t err_row = 0; int err_code = 0;
    int sm_count = 148;                  // multiprocessors of the device (persistent grids)
    uint64_t debug_hash_mask = ~0ull;    // name-group hashes are ANDed with this (tests force collisions with a narrow mask)
    bool profiling = false;
    cudaEvent_t mark[2] = {nullptr, nullptr};
    std::vector<TimedEvent> events;
    std::vector<cudaEvent_t> pool;
    sofa_stats_t stats;
};

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            g_last_error = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
            return SOFA_ERR_CUDA;                                                                    \
        }                                                                                            \
    } while (0)
#define FAIL(code, msg) do { g_last_error = (msg); return (code); } while (0)
#define ENSURE(buf, n) do { if ((buf).ensure(n)) FAIL(SOFA_ERR_CUDA, "cudaMalloc failed (" #buf ")"); } while (0)

#define CK_S(call) do { int r_s_ = (call); if (r_s_) return r_s_; } while (0)

static __global__ void k_copy_small(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
static int stream_sync(sofa_ctx* c) {
    CK(cudaStreamSynchronize(c->stream));
    for (auto& x : c->pending) memcpy(x.dst, c->stage_h + x.off, x.n);
    c->pending.clear();
    c->stage_cur = 0;
    return 0;
}
static int small_d2h(sofa_ctx* c, void* host_dst, const void* dev_src, size_t n) {
    if (n == 0) return 0;
    if (n > c->stage_cap / 2) { CK(cudaMemcpyAsync(host_dst, dev_src, n, cudaMemcpyDeviceToHost, c->stream)); return 0; }
    size_t off = (c->stage_cur + 15) & ~(size_t)15;
    if (off + n > c->stage_cap) { CK_S(stream_sync(c)); off = 0; }
    unsigned g = (unsigned)((n + 255) / 256); if (g > 64) g = 64;
    k_copy_small<<<g, 256, 0, c->stream>>>(c->stage_d + off, (const unsigned char*)dev_src, n);
    c->pending.push_back({host_dst, off, n});
    c->stage_cur = off + n;
    return 0;
}
static int small_h2d(sofa_ctx* c, void* dev_dst, const void* host_src, size_t n) {
    if (n == 0) return 0;
    if (n > c->stage_cap / 2) { CK(cudaMemcpyAsync(dev_dst, host_src, n, cudaMemcpyHostToDevice, c->stream)); CK(cudaStreamSynchronize(c->stream)); return 0; }
    size_t off = (c->stage_cur + 15) & ~(size_t)15;
    if (off + n > c->stage_cap) { CK_S(stream_sync(c)); off = 0; }
    memcpy(c->stage_h + off, host_src, n);
    unsigned g = (unsigned)((n + 255) / 256); if (g > 64) g = 64;
    k_copy_small<<<g, 256, 0, c->stream>>>((unsigned char*)dev_dst, c->stage_d + off, n);
    c->stage_cur = off + n;
    return 0;
}

static cudaEvent_t get_event(sofa_ctx* c) {
    if (!c->pool.empty()) { cudaEvent_t e = c->pool.back(); c->pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct LaunchTimer {
    sofa_ctx* c; int kind; cudaEvent_t a = nullptr, b = nullptr;
    LaunchTimer(sofa_ctx* ctx, int k) : c(ctx), kind(k) {
        c->stats.launches[kind]++;
        if (c->profiling) { a = get_event(c); b = get_event(c); cudaEventRecord(a, c->stream); }
    }
    ~LaunchTimer() { if (c->profiling) { cudaEventRecord(b, c->stream); c->events.push_back({kind, a, b}); } }
};
#define LAUNCH(ctx, kind, ...) do { LaunchTimer lt_(ctx, kind); __VA_ARGS__; } while (0)

static int drain_events(sofa_ctx* c) {
    if (c->events.empty()) return 0;
    CK_S(stream_sync(c));
    for (auto& ev : c->events) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev.a, ev.b);
        c->stats.ms[ev.kind] += ms;
        c->pool.push_back(ev.a); c->pool.push_back(ev.b);
    }
    c->events.clear();
    return 0;
}

static int check_kernel_error(sofa_ctx* c, const char* what) {
    unsigned long long e = ~0ull;
    CK_S(small_d2h(c, &e, c->d_err, 8));
    CK_S(stream_sync(c));
    CK(cudaGetLastError());
    if (e != ~0ull) {
        c->err_row = e >> 8; c->err_code = (int)(e & 0xff);
        char buf[160];
        snprintf(buf, sizeof buf, "%s: record %llu: error %d", what, (unsigned long long)c->err_row, c->err_code);
        g_last_error = buf;
        return SOFA_ERR_ROW;
    }
    return 0;
}
static int reset_err(sofa_ctx* c) {
    CK(cudaMemsetAsync(c->d_err, 0xff, 8, c->stream));
    CK(cudaMemsetAsync(c->d_err + 1, 0, 8 * 15, c->stream));
    CK(cudaMemsetAsync(c->d_flags, 0, 4 * 8, c->stream));
    return 0;
}

// column slots
enum { PC_TS = 0, PC_EVENT, PC_DUR, PC_PID, PC_TID, PC_LS, PC_PIECES };
enum { GC_TS = 0, GC_DUR, GC_BW, GC_PAYLOAD, GC_DEV, GC_TID, GC_SRC, GC_DST, GC_KIND, GC_BWF, GC_LS, GC_NOFF, GC_NLEN };
enum { NC_TS = 0, NC_PAYLOAD, NC_SRC, NC_DST, NC_PROTO, NC_DUR };
enum { SC_TS = 0, SC_EVENT, SC_DUR, SC_DEV, SC_BW, SC_RATIOS };

static SideCols side_cols(sofa_ctx* c, int slot) {
    Source& t = c->src[slot];
    SideCols sc;
    sc.ts = t.cols[SC_TS].as<double>(); sc.event = t.cols[SC_EVENT].as<double>(); sc.dur = t.cols[SC_DUR].as<double>();
    sc.dev = t.cols[SC_DEV].as<double>(); sc.bw = t.cols[SC_BW].as<double>(); sc.ratios = t.cols[SC_RATIOS].as<int32_t>();
    return sc;
}
static AllCols all_cols(sofa_ctx* c, int side_slot = 3, int raw_slot = 6) {
    AllCols a;
    a.s = side_cols(c, side_slot);
    a.r.f = c->src[raw_slot].cols[0].as<double>(); a.r.i = c->src[raw_slot].cols[1].as<int64_t>();
    a.r.nf = c->src[raw_slot].cols[2].as<int32_t>();
    Source& p = c->src[0]; Source& g = c->src[1]; Source& n = c->src[2];
    a.p.ts = p.cols[PC_TS].as<double>(); a.p.event = p.cols[PC_EVENT].as<double>(); a.p.dur = p.cols[PC_DUR].as<double>();
    a.p.pid = p.cols[PC_PID].as<int32_t>(); a.p.tid = p.cols[PC_TID].as<int32_t>();
    a.p.line_start = p.cols[PC_LS].as<uint64_t>(); a.p.pieces = p.cols[PC_PIECES].as<uint16_t>();
    a.g.ts = g.cols[GC_TS].as<double>(); a.g.dur = g.cols[GC_DUR].as<double>(); a.g.bw = g.cols[GC_BW].as<double>();
    a.g.payload = g.cols[GC_PAYLOAD].as<int64_t>(); a.g.dev = g.cols[GC_DEV].as<int32_t>(); a.g.tid = g.cols[GC_TID].as<int32_t>();
    a.g.src = g.cols[GC_SRC].as<int32_t>(); a.g.dst = g.cols[GC_DST].as<int32_t>(); a.g.kind = g.cols[GC_KIND].as<uint8_t>();
    a.g.bwf = g.cols[GC_BWF].as<uint8_t>(); a.g.line_start = g.cols[GC_LS].as<uint64_t>();
    a.g.noff = g.cols[GC_NOFF].as<uint16_t>(); a.g.nlen = g.cols[GC_NLEN].as<uint16_t>();
    a.n.ts = n.cols[NC_TS].as<double>(); a.n.payload = n.cols[NC_PAYLOAD].as<int64_t>(); a.n.src = n.cols[NC_SRC].as<int64_t>();
    a.n.dst = n.cols[NC_DST].as<int64_t>(); a.n.proto = n.cols[NC_PROTO].as<uint8_t>();
    return a;
}
static SrcView make_view(sofa_ctx* c, int k, bool with_viz) {
    SrcView v;
    Source& s = c->src[k];
    memset(&v.side, 0, sizeof v.side);
    v.text = s.text.as<char>();
    v.n_rows = s.n_rows;
    v.c = all_cols(c);
    v.viz_id = (with_viz && k == SRC_PERF && s.viz_ready) ? s.viz_id.as<int32_t>() : nullptr;
    v.dict.n = s.dict_n; v.dict.hash = s.dict_hash.as<uint64_t>(); v.dict.off = s.dict_off.as<uint64_t>();
    v.dict.blob = s.dict_blob.as<char>();
    v.n_cudaproc = s.cfg.n_cudaproc;
    v.api_mode = (k == SRC_GPU) ? s.cfg.api_mode : 0;
    return v;
}

static int alloc_cols(sofa_ctx* c, int k, uint64_t n) {
    Source& s = c->src[k];
    uint64_t m = n ? n : 1;
    if (k == SRC_PERF) {
        ENSURE(s.cols[PC_TS], m * 8); ENSURE(s.cols[PC_EVENT], m * 8); ENSURE(s.cols[PC_DUR], m * 8);
        ENSURE(s.cols[PC_PID], m * 4); ENSURE(s.cols[PC_TID], m * 4); ENSURE(s.cols[PC_LS], m * 8);
        ENSURE(s.cols[PC_PIECES], m * 16);
    } else if (k == SRC_GPU) {
        ENSURE(s.cols[GC_TS], m * 8); ENSURE(s.cols[GC_DUR], m * 8); ENSURE(s.cols[GC_BW], m * 8);
        ENSURE(s.cols[GC_PAYLOAD], m * 8); ENSURE(s.cols[GC_DEV], m * 4); ENSURE(s.cols[GC_TID], m * 4);
        ENSURE(s.cols[GC_SRC], m * 4); ENSURE(s.cols[GC_DST], m * 4); ENSURE(s.cols[GC_KIND], m);
        ENSURE(s.cols[GC_BWF], m); ENSURE(s.cols[GC_LS], m * 8); ENSURE(s.cols[GC_NOFF], m * 2); ENSURE(s.cols[GC_NLEN], m * 2);
    } else if (k == SRC_NET) {
        ENSURE(s.cols[NC_TS], m * 8); ENSURE(s.cols[NC_PAYLOAD], m * 8); ENSURE(s.cols[NC_SRC], m * 8);
        ENSURE(s.cols[NC_DST], m * 8); ENSURE(s.cols[NC_PROTO], m);
    } else if (k >= 6) {
        ENSURE(s.cols[0], m * 8 * RAW_NF); ENSURE(s.cols[1], m * 32); ENSURE(s.cols[2], m * 4);
    } else {
        ENSURE(s.cols[SC_TS], m * 8); ENSURE(s.cols[SC_EVENT], m * 8); ENSURE(s.cols[SC_DUR], m * 8);
        ENSURE(s.cols[SC_DEV], m * 8); ENSURE(s.cols[SC_BW], m * 8); ENSURE(s.cols[SC_RATIOS], m * 20);
    }
    return 0;
}

static inline uint64_t cdiv(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

static int ensure_pinned(sofa_ctx* c, size_t n) {
    if (n <= c->pin_cap) return 0;
    if (c->pin_h) { cudaFreeHost(c->pin_h); c->pin_h = nullptr; c->pin_cap = 0; }
    size_t want = n + n / 4 + 4096;
    CK(cudaHostAlloc(&c->pin_h, want, cudaHostAllocDefault));
    c->pin_cap = want;
    return 0;
}

// In-place exclusive scan of a[0..n), a[n] = total: one block for short arrays, three launches otherwise.
static int launch_scan(sofa_ctx* c, uint64_t* a, uint64_t n) {
    if (n <= 4 * (uint64_t)SCAN_BLOCK_ELEMS) {
        LAUNCH(c, K_SCAN, (k_scan_u64<<<1, 1024, 0, c->stream>>>(a, n)));
        return 0;
    }
    uint64_t nparts = cdiv(n, SCAN_BLOCK_ELEMS);
    if (nparts > 0x7fffffffull) FAIL(SOFA_ERR_ARG, "scan: array too long");
    ENSURE(c->scanpart, (nparts + 2) * 8);
    uint64_t* part = c->scanpart.as<uint64_t>();
    LAUNCH(c, K_SCAN, (k_scan_part<<<(unsigned)nparts, 1024, 0, c->stream>>>(a, n, part)));
    LAUNCH(c, K_SCAN, (k_scan_u64<<<1, 1024, 0, c->stream>>>(part, nparts)));
    LAUNCH(c, K_SCAN, (k_scan_add<<<(unsigned)nparts, 1024, 0, c->stream>>>(a, n, part, (uint32_t)nparts)));
    return 0;
}

// Flip the double buffer of output slot `s` (0 csv, 1 json) before an emit overwrites it: the
// emit goes into the half whose last asynchronous fetch is oldest, after that fetch completed.
static int flip_output(sofa_ctx* c, int s) {
    if (s > 1 || c->arena_on) return 0;
    if (!c->out_busy[s]) return 0;                 // current half was never fetched asynchronously
    std::swap(c->out[s], c->out_alt[s]);
    std::swap(c->out_ev[s], c->out_alt_ev[s]);
    std::swap(c->out_busy[s], c->out_alt_busy[s]);
    if (c->out_busy[s]) { CK(cudaEventSynchronize(c->out_ev[s])); c->out_busy[s] = false; }
    return 0;
}
// Buffer for an emit into output slot s with capacity `cap`: from the arena (bump allocation;
// wraps after draining the D2H stream) when one is set, else the slot's own (double) buffer.
static int out_alloc(sofa_ctx* c, int s, uint64_t cap, char** p) {
    if (c->arena_on && s < 2) {
        uint64_t need = (cap + 64 + 255) & ~255ull;
        if (c->arena_cur + need > c->arena.cap) {
            CK(cudaStreamSynchronize(c->d2h));
            c->arena_cur = 0; c->arena_gen++; c->async_fetches = 0;
            if (need > c->arena.cap) { CK_S(stream_sync(c)); ENSURE(c->arena, need); }
        }
        *p = c->arena.as<char>() + c->arena_cur;
        c->out_ptr[s] = *p;
        return 0;
    }
    ENSURE(c->out[s], cap + 64);
    *p = c->out[s].as<char>();
    c->out_ptr[s] = *p;
    return 0;
}
static void out_commit(sofa_ctx* c, int s, uint64_t total) {
    c->out_n[s] = total;
    if (c->arena_on && s < 2) c->arena_cur += (total + 64 + 255) & ~255ull;
}
static int wait_upload(sofa_ctx* c, int k) {
    if (c->up_pending[k]) { CK(cudaStreamWaitEvent(c->stream, c->up_ev[k], 0)); c->up_pending[k] = false; }
    return 0;
}

// ---- shared across translation units
enum { RC_F = 0, RC_I, RC_NF, RC_IDX, RC_TS, RC_F1, RC_F2, RC_I6, RC_CLS, RC_MASK, RC_TCLOCK };      // column slots of a raw side source
struct SideEmit { sofa::SideView v; uint64_t ncand; };
int need_parsed(sofa_ctx* c, int k);
int fmask_for(sofa_ctx* c, int k, uint32_t stride, uint32_t* fmask);
// emit_fused<FMT> of the old single translation unit: fmt is FMT_CSV / FMT_JSON / FMT_COMM / FMT_AUX
int sofa_emit_fused(sofa_ctx* c, int fmt, int k, uint32_t stride, const uint8_t* mask, int bit, uint32_t fmask,
                    int outslot, uint64_t sel_hint, uint64_t* nbytes, uint64_t* nrec, const SideEmit* se = nullptr);
// per-instantiation launchers (inst/*.cu)
void launch_parse(int tpl_kind, unsigned grid, size_t smem, cudaStream_t st, const char* text, uint64_t begin, uint64_t end,
                  uint32_t tile_bytes, uint32_t max_lines, const uint64_t* tile_row_base, const sofa::IngestCfg& cfg,
                  const sofa::AllCols& cols, unsigned long long* err);
int csv_mask_newlines(sofa_ctx* c, char* text, uint64_t begin, uint64_t end, uint64_t* counts, uint64_t nchunks, int (*scan)(sofa_ctx*, uint64_t*, uint64_t));
void launch_parse_csv(int kind, unsigned grid, size_t smem, cudaStream_t st, const char* text, uint64_t begin, uint64_t end,
                      uint32_t tile_bytes, uint32_t max_lines, const uint64_t* tile_row_base, const sofa::IngestCfg& cfg,
                      const sofa::AllCols& cols, unsigned long long* err);
int parse_tma_occupancy(int kind, size_t smem);
void launch_parse_tma(int kind, unsigned grid, size_t smem, cudaStream_t st, const char* text, uint64_t begin, uint64_t end,
                      uint32_t tile_bytes, uint32_t max_lines, uint64_t ntiles, const uint64_t* tile_row_base, const sofa::IngestCfg& cfg,
                      const sofa::AllCols& cols, unsigned long long* err);
void launch_count_lines(unsigned grid, cudaStream_t st, const char* text, uint64_t begin, uint64_t end, uint32_t tile_bytes,
                        uint64_t* tile_counts, unsigned int* cr_flag);
void launch_emit(int kind, int fmt, unsigned grid, size_t smem, cudaStream_t st, const sofa::SrcView& v, const sofa::FusedArgs& a);
int set_gather_attr(void);
int set_kernel_attrs(void);       // dynamic shared memory opt-in of every instantiation; 0 or a cudaError_t
