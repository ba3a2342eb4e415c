// libsofa_b200.so: C ABI (include/sofa_b200.h) over the sm_100a kernels.
// Single translation unit; CUDA runtime only (no torch, no thrust/cub).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/sofa_b200.h"
#include "kernels_emit.cuh"
#include "kernels_analyze.cuh"
#include "kernels_sort.cuh"
#include "kernels_window.cuh"
#include "synth.cuh"

using namespace sofa;

static thread_local std::string g_last_error;

enum {
    K_COUNT = 0, K_SCAN, K_PARSE, K_FLAGS, K_CSV_WRITE, K_JSON_WRITE, K_FILTER,
    K_VIZ, K_ANALYZE, K_SORT, K_SYNTH, K_NKINDS
};
static const char* KERNEL_NAMES[K_NKINDS] = {
    "count_lines", "scan", "parse", "row_flags", "csv_emit", "json_emit", "filter",
    "viz_dict", "analyze", "sort", "synth"};

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
    bool side_ready = false; int side_absolute = 0; uint64_t side_ncand = 0;          // raw side source: rows per list after sofa_b200_side_preprocess
};

struct TimedEvent { int kind; cudaEvent_t a, b; };

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
    // small host<->device traffic goes through mapped pinned memory + a copy kernel, never the
    // DMA engines (a tiny read-back queued behind a multi-GB transfer would stall the host)
    unsigned char* stage_h = nullptr; unsigned char* stage_d = nullptr; size_t stage_cap = 8u << 20, stage_cur = 0;
    struct SmallXfer { void* dst; size_t off, n; };
    std::vector<SmallXfer> pending;
    char* out_ptr[3] = {nullptr, nullptr, nullptr};
    DevBuf tiles, rowlen, blockoff, scratch, ftable, keys, rows;
    DevBuf sk0, sv0, sk1, sv1, hist, segs, sums, small, scanpart;
    void* pin_h = nullptr; size_t pin_cap = 0;  // pinned host scratch for larger read-backs (grown on demand)
    unsigned long long* d_err = nullptr;        // [0] error slot, [1] counter, [2..9] filter counts, [10] misc
    unsigned int* d_flags = nullptr;            // [0] cr flag, [1] row flags, [2] overflow, [3] count
    uint64_t err_row = 0; int err_code = 0;
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

__global__ void k_copy_small(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, size_t n) {
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
            c->arena_cur = 0;
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

extern "C" {

int sofa_b200_abi_version(void) { return SOFA_B200_ABI_VERSION; }
const char* sofa_b200_last_error(void) { return g_last_error.c_str(); }
const char* sofa_b200_kernel_name(int kind) { return (kind >= 0 && kind < K_NKINDS) ? KERNEL_NAMES[kind] : ""; }

int sofa_b200_error_row(sofa_ctx* c, uint64_t* row, int32_t* code) {
    if (!c) FAIL(SOFA_ERR_ARG, "null ctx");
    if (row) *row = c->err_row;
    if (code) *code = c->err_code;
    return 0;
}

int sofa_b200_create(sofa_ctx** out, int device) {
    if (!out) FAIL(SOFA_ERR_ARG, "null out");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev == 0) FAIL(SOFA_ERR_CUDA, "no CUDA device: libsofa_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) FAIL(SOFA_ERR_ARG, "bad device index");
    CK(cudaSetDevice(device));
    sofa_ctx* c = new sofa_ctx();
    c->device = device;
    memset(&c->stats, 0, sizeof c->stats);
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) { CK(cudaEventCreateWithFlags(&c->out_ev[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->out_alt_ev[i], cudaEventDisableTiming)); }
    for (int i = 0; i < 11; i++) CK(cudaEventCreateWithFlags(&c->up_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < 16; i++) { CK(cudaEventCreateWithFlags(&c->fetch_ev[i], cudaEventDisableTiming)); c->fetch_used[i] = false; }
    CK(cudaHostAlloc((void**)&c->stage_h, c->stage_cap, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&c->stage_d, c->stage_h, 0));
    CK(cudaMalloc(&c->d_err, 8 * 16));
    CK(cudaMalloc(&c->d_flags, 4 * 8));
    // dynamic shared memory opt-in for the parse kernels
    int maxsm = 0;
    CK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CK(cudaFuncSetAttribute(k_parse<SRC_PERF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CK(cudaFuncSetAttribute(k_parse<SRC_GPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CK(cudaFuncSetAttribute(k_parse<SRC_NET>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CK(cudaFuncSetAttribute(k_parse<SRC_SIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CK(cudaFuncSetAttribute(k_parse<SRC_RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CK(cudaFuncSetAttribute(k_emit_fused<SRC_PERF, FMT_CSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k_emit_fused<SRC_PERF, FMT_JSON>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k_emit_fused<SRC_GPU, FMT_CSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k_emit_fused<SRC_GPU, FMT_JSON>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k_emit_fused<SRC_NET, FMT_CSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k_emit_fused<SRC_NET, FMT_JSON>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k_emit_fused<SRC_GPU, FMT_COMM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k_emit_fused<SRC_SIDEOUT, FMT_CSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k_emit_fused<SRC_SIDEOUT, FMT_JSON>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k_emit_fused<SRC_SIDEOUT, FMT_AUX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    *out = c;
    return 0;
}

int sofa_b200_destroy(sofa_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->h2d) { cudaStreamSynchronize(c->h2d); cudaStreamDestroy(c->h2d); }
    if (c->d2h) { cudaStreamSynchronize(c->d2h); cudaStreamDestroy(c->d2h); }
    for (int i = 0; i < 2; i++) { if (c->out_ev[i]) cudaEventDestroy(c->out_ev[i]); if (c->out_alt_ev[i]) cudaEventDestroy(c->out_alt_ev[i]); c->out_alt[i].release(); }
    for (int i = 0; i < 11; i++) if (c->up_ev[i]) cudaEventDestroy(c->up_ev[i]);
    for (int i = 0; i < 16; i++) cudaEventDestroy(c->fetch_ev[i]);
    for (auto& ev : c->events) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    for (auto& e : c->pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) if (c->mark[i]) cudaEventDestroy(c->mark[i]);
    for (int k = 0; k < 11; k++) {
        Source& s = c->src[k];
        s.text.release(); s.mhz.release(); s.mask.release(); s.vizhash.release(); s.viz_id.release();
        s.dict_hash.release(); s.dict_off.release(); s.dict_blob.release();
        for (auto& b : s.cols) b.release();
        if (k < 3) c->out[k].release();
    }
    c->tiles.release(); c->rowlen.release(); c->blockoff.release(); c->scratch.release(); c->ftable.release();
    c->keys.release(); c->rows.release();
    c->sk0.release(); c->sv0.release(); c->sk1.release(); c->sv1.release(); c->hist.release(); c->segs.release();
    c->sums.release(); c->small.release(); c->arena.release(); c->scanpart.release();
    cudaFree(c->d_err); cudaFree(c->d_flags);
    if (c->stage_h) cudaFreeHost(c->stage_h);
    if (c->pin_h) cudaFreeHost(c->pin_h);
    cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

void* sofa_b200_host_alloc(uint64_t nbytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, nbytes ? nbytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void sofa_b200_host_free(void* p) { if (p) cudaFreeHost(p); }

int sofa_b200_upload(sofa_ctx* c, int k, const void* text, uint64_t nbytes, uint64_t data_off) {
    if (!c || k < 0 || k > 10 || (!text && nbytes)) FAIL(SOFA_ERR_ARG, "upload: bad argument");
    if (data_off > nbytes) FAIL(SOFA_ERR_ARG, "upload: data_off beyond the text");
    CK(cudaSetDevice(c->device));
    Source& s = c->src[k];
    ENSURE(s.text, nbytes + 64);
    if (nbytes) CK(cudaMemcpyAsync(s.text.p, text, nbytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(s.text.as<char>() + nbytes, '\n', 64, c->stream));
    CK_S(stream_sync(c));
    c->stats.h2d_bytes += nbytes;
    s.nbytes = nbytes; s.data_off = data_off; s.uploaded = true; s.parsed = false; s.viz_ready = false; s.n_rows = 0;
    s.mask_stride = 0; s.nkw = 0;
    return 0;
}

static uint32_t pick_tile_bytes(int k) { return (k == SRC_NET || k >= 3) ? 16384u : 24576u; }
enum { RC_F = 0, RC_I, RC_NF, RC_IDX, RC_TS, RC_F1, RC_F2, RC_I6, RC_CLS, RC_MASK, RC_TCLOCK };      // column slots of a raw side source

int sofa_b200_parse(sofa_ctx* c, int k, const sofa_ingest_params_t* p, uint64_t* n_rows) {
    if (!c || k < 0 || k > 10 || !p) FAIL(SOFA_ERR_ARG, "parse: bad argument");
    Source& s = c->src[k];
    if (!s.uploaded) FAIL(SOFA_ERR_STATE, "parse: nothing uploaded for this source");
    CK(cudaSetDevice(c->device));
    c->pending.clear(); c->stage_cur = 0;
    IngestCfg cfg;
    cfg.time_base = p->time_base; cfg.t_offset = p->t_offset; cfg.absolute_timestamp = p->absolute_timestamp;
    cfg.perf_has_cycles = p->perf_has_cycles; cfg.n_mhz = p->n_mhz > 0 ? p->n_mhz : 1;
    cfg.ts_rescale = p->ts_rescale; cfg.dt_rescale = p->dt_rescale; cfg.payload_unit = p->payload_unit;
    cfg.n_cudaproc = p->n_cudaproc;
    cfg.idx_start = p->idx[0]; cfg.idx_dur = p->idx[1]; cfg.idx_size = p->idx[2]; cfg.idx_thr = p->idx[3];
    cfg.idx_ctx = p->idx[4]; cfg.idx_stream = p->idx[5]; cfg.idx_name = p->idx[6]; cfg.idx_srcctx = p->idx[7];
    cfg.idx_dstctx = p->idx[8];
    cfg.api_mode = p->api_mode;
    cfg.side_table = k >= 6 ? k - 6 : (k >= 3 ? k - 3 : 0);
    if (k == SRC_GPU && (cfg.idx_start < 0 || cfg.idx_dur < 0 || cfg.idx_name < 0))
        FAIL(SOFA_ERR_ARG, "parse: gpu trace needs Start, Duration and Name columns");
    ENSURE(s.mhz, 16 * (size_t)cfg.n_mhz + 16);
    if (p->n_mhz > 0 && p->mhz_xp && p->mhz_fp) {
        CK_S(small_h2d(c, s.mhz.p, p->mhz_xp, 8 * (size_t)cfg.n_mhz));
        CK_S(small_h2d(c, s.mhz.as<double>() + cfg.n_mhz, p->mhz_fp, 8 * (size_t)cfg.n_mhz));
    } else {
        double d[2] = {0.0, 3000.0};
        CK_S(small_h2d(c, s.mhz.p, &d[0], 8));
        CK_S(small_h2d(c, s.mhz.as<double>() + 1, &d[1], 8));
        CK_S(stream_sync(c));
    }
    cfg.mhz_xp = s.mhz.as<double>(); cfg.mhz_fp = s.mhz.as<double>() + cfg.n_mhz;
    s.cfg = cfg;
    s.parsed = false; s.viz_ready = false; s.mask_stride = 0;
    s.flags_stride[0] = s.flags_stride[1] = 0;
    if (wait_upload(c, k)) return SOFA_ERR_CUDA;
    uint64_t begin = s.data_off, end = s.nbytes;
    if (end <= begin) { s.n_rows = 0; s.parsed = true; if (n_rows) *n_rows = 0; return 0; }
    uint32_t tile = pick_tile_bytes(k);
    uint64_t ntiles = cdiv(end - begin, tile);
    if (ntiles > 0x7fffffffull) FAIL(SOFA_ERR_ARG, "parse: source too large for one call");
    ENSURE(c->tiles, (ntiles + 2) * 8);
    if (reset_err(c)) return SOFA_ERR_CUDA;
    LAUNCH(c, K_COUNT, (k_count_lines<<<(unsigned)ntiles, PARSE_TPB, 0, c->stream>>>(
        s.text.as<char>(), begin, end, tile, c->tiles.as<uint64_t>(), c->d_flags)));
    CK_S(launch_scan(c, c->tiles.as<uint64_t>(), ntiles));
    uint64_t total = 0; unsigned int cr = 0;
    CK_S(small_d2h(c, &total, c->tiles.as<uint64_t>() + ntiles, 8));
    CK_S(small_d2h(c, &cr, c->d_flags, 4));
    CK_S(stream_sync(c));
    CK(cudaGetLastError());
    if (cr) { c->err_row = 0; c->err_code = E_CR; FAIL(SOFA_ERR_ROW, "parse: carriage return in trace text (not supported)"); }
    if (alloc_cols(c, k, total)) return SOFA_ERR_CUDA;
    if (k >= 6 && total) CK(cudaMemsetAsync(s.cols[RC_NF].p, 0, total * 4, c->stream));
    s.n_rows = total;
    uint32_t max_lines = tile / 8;
    size_t smem = (((size_t)tile + OVERLAP_BYTES + 48) & ~(size_t)15) + (size_t)max_lines * 4;
    AllCols cols = all_cols(c, (k >= 3 && k < 6) ? k : 3, k >= 6 ? k : 6);
    if (k >= 6)
        LAUNCH(c, K_PARSE, (k_parse<SRC_RAW><<<(unsigned)ntiles, PARSE_TPB, smem, c->stream>>>(
            s.text.as<char>(), begin, end, tile, max_lines, c->tiles.as<uint64_t>(), cfg, cols, c->d_err)));
    else if (k >= 3)
        LAUNCH(c, K_PARSE, (k_parse<SRC_SIDE><<<(unsigned)ntiles, PARSE_TPB, smem, c->stream>>>(
            s.text.as<char>(), begin, end, tile, max_lines, c->tiles.as<uint64_t>(), cfg, cols, c->d_err)));
    else if (k == SRC_PERF)
        LAUNCH(c, K_PARSE, (k_parse<SRC_PERF><<<(unsigned)ntiles, PARSE_TPB, smem, c->stream>>>(
            s.text.as<char>(), begin, end, tile, max_lines, c->tiles.as<uint64_t>(), cfg, cols, c->d_err)));
    else if (k == SRC_GPU)
        LAUNCH(c, K_PARSE, (k_parse<SRC_GPU><<<(unsigned)ntiles, PARSE_TPB, smem, c->stream>>>(
            s.text.as<char>(), begin, end, tile, max_lines, c->tiles.as<uint64_t>(), cfg, cols, c->d_err)));
    else
        LAUNCH(c, K_PARSE, (k_parse<SRC_NET><<<(unsigned)ntiles, PARSE_TPB, smem, c->stream>>>(
            s.text.as<char>(), begin, end, tile, max_lines, c->tiles.as<uint64_t>(), cfg, cols, c->d_err)));
    int r = check_kernel_error(c, "parse");
    if (r) return r;
    s.parsed = true;
    if (n_rows) *n_rows = total;
    return 0;
}

static int need_parsed(sofa_ctx* c, int k) {
    if (!c || k < 0 || k > 2) FAIL(SOFA_ERR_ARG, "bad source");
    if (!c->src[k].parsed) FAIL(SOFA_ERR_STATE, "source not parsed");
    CK(cudaSetDevice(c->device));
    c->pending.clear(); c->stage_cur = 0;          // drop read-backs orphaned by an earlier failed call
    return 0;
}

int sofa_b200_row_flags(sofa_ctx* c, int k, uint32_t stride, uint32_t* flags) {
    int r = need_parsed(c, k); if (r) return r;
    if (stride == 0 || !flags) FAIL(SOFA_ERR_ARG, "row_flags: bad argument");
    Source& s = c->src[k];
    *flags = 0;
    if (k == SRC_PERF || s.n_rows == 0) return 0;
    for (int i = 0; i < 2; i++) if (s.flags_stride[i] == stride) { *flags = s.flags_val[i]; return 0; }
    uint64_t ncand = cdiv(s.n_rows, stride);
    CK(cudaMemsetAsync(c->d_flags + 1, 0, 4, c->stream));
    SrcView v = make_view(c, k, false);
    if (k == SRC_GPU) LAUNCH(c, K_FLAGS, (k_row_flags<SRC_GPU><<<(unsigned)cdiv(ncand, 256), 256, 0, c->stream>>>(v, ncand, stride, c->d_flags + 1)));
    else LAUNCH(c, K_FLAGS, (k_row_flags<SRC_NET><<<(unsigned)cdiv(ncand, 256), 256, 0, c->stream>>>(v, ncand, stride, c->d_flags + 1)));
    CK_S(small_d2h(c, flags, c->d_flags + 1, 4));
    CK_S(stream_sync(c));
    CK(cudaGetLastError());
    { int slot = s.flags_stride[0] == 0 ? 0 : 1; s.flags_stride[slot] = stride; s.flags_val[slot] = *flags; }
    return 0;
}

static int fmask_for(sofa_ctx* c, int k, uint32_t stride, uint32_t* fmask) {
    if (k == SRC_PERF) { *fmask = 0x7; return 0; }
    uint32_t fl = 0;
    int r = sofa_b200_row_flags(c, k, stride, &fl);
    if (r) return r;
    if (k == SRC_GPU) { *fmask = 0x5u | ((fl & 1) ? 0x40u : 0u); return 0; }
    if (fl & 2) {
        if (!(fl & 4)) FAIL(SOFA_ERR_ROW, "every selected packet row is empty: the reference raises (0 columns)");
        *fmask = 0x17ff;
    } else *fmask = 0x45;
    return 0;
}

}  // extern "C"


struct SideEmit { SideView v; uint64_t ncand; };

template <int FMT>
static int emit_fused(sofa_ctx* c, int k, uint32_t stride, const uint8_t* mask, int bit, uint32_t fmask,
                      int outslot, uint64_t sel_hint, uint64_t* nbytes, uint64_t* nrec, const SideEmit* se = nullptr) {
    Source& s = c->src[k];
    uint64_t ncand = se ? se->ncand : cdiv(s.n_rows, stride);
    c->out_n[outslot] = 0;
    if (nbytes) *nbytes = 0;
    if (nrec) *nrec = 0;
    if (flip_output(c, outslot)) return SOFA_ERR_CUDA;
    if (ncand == 0) return 0;
    if (ncand >= 0xffffffffull) FAIL(SOFA_ERR_ARG, "emit: too many rows for one call");
    uint64_t text_share = s.nbytes > s.data_off ? s.nbytes - s.data_off : 0;
    if (se) text_share = 0;                      // names are generated, not copied from the text
    const uint32_t* index = nullptr;
    uint64_t nitems = ncand;
    if (mask) {
        // filtered series: compact the selected candidates first so the emit grid is dense
        uint64_t nb = cdiv(ncand, 256);
        ENSURE(c->blockoff, (nb + 2) * 8);
        LAUNCH(c, K_FILTER, (k_mask_count<<<(unsigned)nb, 256, 0, c->stream>>>(mask, bit, ncand, c->blockoff.as<uint64_t>())));
        CK_S(launch_scan(c, c->blockoff.as<uint64_t>(), nb));
        uint64_t cnt = 0;
        CK_S(small_d2h(c, &cnt, c->blockoff.as<uint64_t>() + nb, 8));
        CK_S(stream_sync(c));
        if (cnt == 0) return 0;
        ENSURE(c->rowlen, cnt * 4 + 16);
        LAUNCH(c, K_FILTER, (k_mask_write<<<(unsigned)nb, 256, 0, c->stream>>>(mask, bit, ncand, c->blockoff.as<uint64_t>(), c->rowlen.as<uint32_t>())));
        index = c->rowlen.as<uint32_t>();
        nitems = cnt;
        text_share = (uint64_t)((double)text_share * ((double)cnt / (double)ncand) * 2.0) + (1 << 20);
        sel_hint = cnt;
    }
    uint64_t nblocks = cdiv(nitems, EMIT_TPB);
    ENSURE(c->tiles, (nblocks + 2) * 8);
    SrcView v = make_view(c, k, true);
    if (se) v.side = se->v;
    uint64_t per_row = FMT == FMT_CSV ? 96 : ((FMT == FMT_COMM || FMT == FMT_AUX) ? 64 : 230);
    if (FMT == FMT_COMM) text_share = 0;
    uint64_t cap = sel_hint * per_row + text_share + (1 << 16);
    // shared-memory image per block: sized for 128 typical rows of the format (larger ranges fall back
    // to direct global writes); smaller images = more resident blocks per SM
    // bytes of the per-block image; measured at 100 M events: CSV 24K 37.5 ms, 28K 26.5, 32K 27.8, 40K 30.2; JSON 36K 73.9,
    // 40K 57.5, 44K 57.5 (a block whose rows do not fit falls back to direct global stores - the cliff).  Env overrides for tuning.
    static const uint32_t slot_csv = getenv("SOFA_B200_SLOT_CSV") ? (uint32_t)atoi(getenv("SOFA_B200_SLOT_CSV")) : 32u * 1024u;
    static const uint32_t slot_json = getenv("SOFA_B200_SLOT_JSON") ? (uint32_t)atoi(getenv("SOFA_B200_SLOT_JSON")) : 40u * 1024u;
    uint32_t slot_bytes = (FMT == FMT_COMM || FMT == FMT_AUX) ? 16 * 1024 : (FMT == FMT_JSON ? slot_json : slot_csv);
    if (!se && nitems) {
        // traces with long names (C++ symbols of several hundred bytes): grow the image with the average line
        // length so that a block's 128 rows still fit - the direct-to-global fallback is several times slower
        double avg_line = (double)(s.nbytes > s.data_off ? s.nbytes - s.data_off : 0) / (double)(s.n_rows ? s.n_rows : 1);
        double est_row = (FMT == FMT_JSON ? 200.0 : 70.0) + avg_line;
        uint32_t want = (uint32_t)(est_row * EMIT_TPB * 1.1);
        if (FMT != FMT_COMM && want > slot_bytes + slot_bytes / 8) slot_bytes = (want + 1023u) & ~1023u;     // only when clearly needed
    }
    if (slot_bytes < 4096u) slot_bytes = 4096u;
    if (slot_bytes > 63u * 1024u) slot_bytes = 63u * 1024u;
    for (int attempt = 0; attempt < 6; attempt++, cap *= 2) {
        char* outp = nullptr;
        { int ra = out_alloc(c, outslot, cap, &outp); if (ra) return ra; }
        if (reset_err(c)) return SOFA_ERR_CUDA;
        CK(cudaMemsetAsync(c->tiles.p, 0, (nblocks + 1) * 8, c->stream));
        FusedArgs a;
        a.ncand = nitems; a.stride = stride; a.fmask = fmask; a.mask = nullptr; a.bit = 0; a.index = index;
        a.out = outp; a.cap = cap;
        a.state = c->tiles.as<unsigned long long>(); a.ticket = c->d_flags + 4;
        a.err = c->d_err; a.count = c->d_err + 1; a.total = c->d_err + 11; a.overflow = c->d_flags + 2;
        a.slot_bytes = slot_bytes;
        int kk = FMT == FMT_JSON ? K_JSON_WRITE : K_CSV_WRITE;
        unsigned g = (unsigned)nblocks;
        if constexpr (FMT == FMT_COMM) {
            if (k != SRC_GPU) FAIL(SOFA_ERR_ARG, "comm.csv is a gpu-trace output");
            LAUNCH(c, kk, (k_emit_fused<SRC_GPU, FMT><<<g, EMIT_TPB, slot_bytes + 32, c->stream>>>(v, a)));
        } else if constexpr (FMT == FMT_AUX) {
            if (!se) FAIL(SOFA_ERR_ARG, "aux rows are a side-source output");
            LAUNCH(c, kk, (k_emit_fused<SRC_SIDEOUT, FMT><<<g, EMIT_TPB, slot_bytes + 32, c->stream>>>(v, a)));
        } else {
            if (se) LAUNCH(c, kk, (k_emit_fused<SRC_SIDEOUT, FMT><<<g, EMIT_TPB, slot_bytes + 32, c->stream>>>(v, a)));
            else if (k == SRC_PERF) LAUNCH(c, kk, (k_emit_fused<SRC_PERF, FMT><<<g, EMIT_TPB, slot_bytes + 32, c->stream>>>(v, a)));
            else if (k == SRC_GPU) LAUNCH(c, kk, (k_emit_fused<SRC_GPU, FMT><<<g, EMIT_TPB, slot_bytes + 32, c->stream>>>(v, a)));
            else LAUNCH(c, kk, (k_emit_fused<SRC_NET, FMT><<<g, EMIT_TPB, slot_bytes + 32, c->stream>>>(v, a)));
        }
        unsigned long long total = 0, cnt = 0; unsigned int ovf = 0;
        CK_S(small_d2h(c, &total, c->d_err + 11, 8));
        CK_S(small_d2h(c, &cnt, c->d_err + 1, 8));
        CK_S(small_d2h(c, &ovf, c->d_flags + 2, 4));
        int r = check_kernel_error(c, FMT == FMT_CSV ? "emit_csv" : "emit_json");
        if (r) return r;
        if (ovf) { if (total + 64 > cap * 2) cap = total / 2 + 64; continue; }
        out_commit(c, outslot, total);
        if (nbytes) *nbytes = total;
        if (nrec) *nrec = cnt;
        return 0;
    }
    FAIL(SOFA_ERR_CAP, "emit: output capacity estimate failed repeatedly");
}

extern "C" {

int sofa_b200_emit_csv(sofa_ctx* c, int k, uint32_t stride, uint64_t* nbytes) {
    int r = need_parsed(c, k); if (r) return r;
    if (stride == 0) FAIL(SOFA_ERR_ARG, "emit_csv: stride must be >= 1");
    uint32_t fmask = 0;
    r = fmask_for(c, k, stride, &fmask); if (r) return r;
    return emit_fused<FMT_CSV>(c, k, stride, nullptr, 0, fmask, 0, cdiv(c->src[k].n_rows, stride), nbytes, nullptr);
}

int sofa_b200_emit_comm_csv(sofa_ctx* c, uint64_t* nbytes) {
    int r = need_parsed(c, SRC_GPU); if (r) return r;
    uint32_t fmask = 0;
    r = fmask_for(c, SRC_GPU, 1, &fmask); if (r) return r;
    return emit_fused<FMT_COMM>(c, SRC_GPU, 1, nullptr, 0, fmask, 0, c->src[SRC_GPU].n_rows, nbytes, nullptr);
}

int sofa_b200_emit_json(sofa_ctx* c, int k, uint32_t stride, int32_t filter_idx, uint64_t* nbytes, uint64_t* nrec) {
    int r = need_parsed(c, k); if (r) return r;
    if (stride == 0) FAIL(SOFA_ERR_ARG, "emit_json: stride must be >= 1");
    Source& s = c->src[k];
    if (k == SRC_PERF && !(s.viz_ready && s.viz_stride == stride))
        FAIL(SOFA_ERR_STATE, "emit_json(perf): call viz_mangled + set_demangled for this stride first");
    const uint8_t* mask = nullptr;
    if (filter_idx >= 0) {
        if (filter_idx >= s.nkw || s.mask_stride != stride) FAIL(SOFA_ERR_STATE, "emit_json: filter not set for this stride");
        mask = s.mask.as<uint8_t>();
    }
    uint32_t fmask = 0;
    r = fmask_for(c, k, stride, &fmask); if (r) return r;
    uint64_t nb = 0, nr = 0;
    uint64_t hint = filter_idx >= 0 ? s.filter_counts[filter_idx] : cdiv(s.n_rows, stride);
    r = emit_fused<FMT_JSON>(c, k, stride, mask, filter_idx >= 0 ? filter_idx : 0, fmask, 1, hint, &nb, &nr);
    if (r) return r;
    if (nbytes) *nbytes = nb >= 2 ? nb - 2 : 0;       // the leading ", " is dropped by fetch
    if (nrec) *nrec = nr;
    return 0;
}

int sofa_b200_fetch(sofa_ctx* c, int which, void* host_out, uint64_t cap) {
    if (!c || which < 0 || which > 2 || !host_out) FAIL(SOFA_ERR_ARG, "fetch: bad argument");
    CK(cudaSetDevice(c->device));
    uint64_t n = c->out_n[which], skip = 0;
    if (which == 1) { skip = n >= 2 ? 2 : n; n -= skip; }
    if (cap < n) FAIL(SOFA_ERR_CAP, "fetch: buffer too small");
    if (n) {
        const char* srcp = which < 2 ? c->out_ptr[which] : c->out[which].as<char>();
        CK(cudaMemcpyAsync(host_out, srcp + skip, n, cudaMemcpyDeviceToHost, c->stream));
        CK_S(stream_sync(c));
        c->stats.d2h_bytes += n;
    }
    return 0;
}

int sofa_b200_upload_async(sofa_ctx* c, int k, const void* text, uint64_t nbytes, uint64_t data_off) {
    if (!c || k < 0 || k > 10 || (!text && nbytes)) FAIL(SOFA_ERR_ARG, "upload_async: bad argument");
    if (data_off > nbytes) FAIL(SOFA_ERR_ARG, "upload_async: data_off beyond the text");
    CK(cudaSetDevice(c->device));
    Source& s = c->src[k];
    if (nbytes + 64 > s.text.cap) { CK_S(stream_sync(c)); CK(cudaStreamSynchronize(c->h2d)); ENSURE(s.text, nbytes + 64); }
    // the copy must not overtake kernels of an earlier pass that still read this buffer
    CK(cudaEventRecord(c->up_ev[k], c->stream));
    CK(cudaStreamWaitEvent(c->h2d, c->up_ev[k], 0));
    if (nbytes) CK(cudaMemcpyAsync(s.text.p, text, nbytes, cudaMemcpyHostToDevice, c->h2d));
    CK(cudaMemsetAsync(s.text.as<char>() + nbytes, '\n', 64, c->h2d));
    CK(cudaEventRecord(c->up_ev[k], c->h2d));
    c->up_pending[k] = true;
    c->stats.h2d_bytes += nbytes;
    s.nbytes = nbytes; s.data_off = data_off; s.uploaded = true; s.parsed = false; s.viz_ready = false; s.n_rows = 0;
    s.mask_stride = 0; s.nkw = 0;
    return 0;
}

int sofa_b200_fetch_async(sofa_ctx* c, int which, void* host_out, uint64_t cap, int32_t* ticket) {
    if (!c || which < 0 || which > 1 || !host_out) FAIL(SOFA_ERR_ARG, "fetch_async: output buffer 0 or 1");
    CK(cudaSetDevice(c->device));
    uint64_t n = c->out_n[which], skip = 0;
    if (which == 1) { skip = n >= 2 ? 2 : n; n -= skip; }
    if (cap < n) FAIL(SOFA_ERR_CAP, "fetch_async: buffer too small");
    int t = c->fetch_next;
    c->fetch_next = (c->fetch_next + 1) & 15;
    if (c->fetch_used[t]) CK(cudaEventSynchronize(c->fetch_ev[t]));
    if (n) CK(cudaMemcpyAsync(host_out, c->out_ptr[which] + skip, n, cudaMemcpyDeviceToHost, c->d2h));
    CK(cudaEventRecord(c->out_ev[which], c->d2h));
    CK(cudaEventRecord(c->fetch_ev[t], c->d2h));
    c->out_busy[which] = true;
    c->fetch_used[t] = true;
    c->stats.d2h_bytes += n;
    if (ticket) *ticket = t;
    return 0;
}
int sofa_b200_wait_fetch(sofa_ctx* c, int32_t ticket) {
    if (!c || ticket < 0 || ticket > 15) FAIL(SOFA_ERR_ARG, "wait_fetch: bad ticket");
    CK(cudaSetDevice(c->device));
    if (c->fetch_used[ticket]) CK(cudaEventSynchronize(c->fetch_ev[ticket]));
    return 0;
}
int sofa_b200_sync(sofa_ctx* c) {
    if (!c) FAIL(SOFA_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->h2d));
    CK_S(stream_sync(c));
    CK(cudaStreamSynchronize(c->d2h));
    c->arena_cur = 0;
    return 0;
}
int sofa_b200_set_output_arena(sofa_ctx* c, uint64_t nbytes) {
    if (!c) FAIL(SOFA_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    CK_S(stream_sync(c));
    CK(cudaStreamSynchronize(c->d2h));
    if (nbytes == 0) { c->arena.release(); c->arena_on = false; c->arena_cur = 0; return 0; }
    ENSURE(c->arena, nbytes);
    c->arena_on = true; c->arena_cur = 0;
    return 0;
}

int sofa_b200_set_filters(sofa_ctx* c, int k, uint32_t stride, int32_t nkw, const char* const* kws, uint64_t* counts) {
    int r = need_parsed(c, k); if (r) return r;
    if (stride == 0 || nkw < 0 || nkw > 8 || (nkw && !kws)) FAIL(SOFA_ERR_ARG, "set_filters: 0..8 keywords");
    Source& s = c->src[k];
    if (k == SRC_PERF && !(s.viz_ready && s.viz_stride == stride))
        FAIL(SOFA_ERR_STATE, "set_filters(perf): call viz_mangled + set_demangled for this stride first");
    FilterTable ft;
    memset(&ft, 0, sizeof ft);
    size_t lens[8] = {0};
    for (int i = 0; i < nkw; i++) {
        lens[i] = strlen(kws[i]);
        if (lens[i] == 0 || lens[i] > 32) FAIL(SOFA_ERR_ARG, "set_filters: keyword length must be 1..32 bytes");
    }
    filter_table_build(&ft, nkw, kws, lens);
    uint64_t ncand = cdiv(s.n_rows, stride);
    ENSURE(c->ftable, sizeof(FilterTable));
    ENSURE(s.mask, ncand + 16);
    if (reset_err(c)) return SOFA_ERR_CUDA;
    CK_S(small_h2d(c, c->ftable.p, &ft, sizeof ft));
    CK_S(stream_sync(c));
    if (ncand) {
        SrcView v = make_view(c, k, true);
        unsigned g = (unsigned)cdiv(ncand, 256);
        const FilterTable* dft = c->ftable.as<FilterTable>();
        if (k == SRC_PERF) LAUNCH(c, K_FILTER, (k_filter<SRC_PERF><<<g, 256, 0, c->stream>>>(v, ncand, stride, dft, s.mask.as<uint8_t>(), c->d_err + 2)));
        else if (k == SRC_GPU) LAUNCH(c, K_FILTER, (k_filter<SRC_GPU><<<g, 256, 0, c->stream>>>(v, ncand, stride, dft, s.mask.as<uint8_t>(), c->d_err + 2)));
        else LAUNCH(c, K_FILTER, (k_filter<SRC_NET><<<g, 256, 0, c->stream>>>(v, ncand, stride, dft, s.mask.as<uint8_t>(), c->d_err + 2)));
    }
    unsigned long long cn[8] = {0};
    CK_S(small_d2h(c, cn, c->d_err + 2, 64));
    CK_S(stream_sync(c));
    CK(cudaGetLastError());
    if (counts) for (int i = 0; i < nkw; i++) counts[i] = cn[i];
    for (int i = 0; i < 8; i++) s.filter_counts[i] = i < nkw ? cn[i] : 0;
    s.mask_stride = stride; s.nkw = nkw;
    return 0;
}

int sofa_b200_viz_mangled(sofa_ctx* c, uint32_t stride, uint64_t* nbytes, uint32_t* n_names) {
    int r = need_parsed(c, SRC_PERF); if (r) return r;
    if (stride == 0) FAIL(SOFA_ERR_ARG, "viz_mangled: stride must be >= 1");
    Source& s = c->src[SRC_PERF];
    uint64_t ncand = cdiv(s.n_rows, stride);
    s.viz_ready = false; s.viz_stride = stride;
    c->out_n[2] = 0;
    if (nbytes) *nbytes = 0;
    if (n_names) *n_names = 0;
    ENSURE(s.vizhash, (ncand + 1) * 8);
    ENSURE(s.viz_id, (ncand + 1) * 4);
    if (ncand == 0) return 0;
    SrcView v = make_view(c, SRC_PERF, false);
    for (uint64_t cap = 1ull << 16;; cap <<= 2) {
        if (cap > (1ull << 30)) FAIL(SOFA_ERR_CUDA, "viz dictionary overflow");
        ENSURE(c->keys, cap * 8); ENSURE(c->rows, cap * 8);
        if (reset_err(c)) return SOFA_ERR_CUDA;
        CK(cudaMemsetAsync(c->keys.p, 0, cap * 8, c->stream));
        CK(cudaMemsetAsync(c->rows.p, 0xff, cap * 8, c->stream));
        LAUNCH(c, K_VIZ, (k_viz_hash<<<(unsigned)cdiv(ncand, 256), 256, 0, c->stream>>>(
            v, ncand, stride, s.vizhash.as<uint64_t>(), c->keys.as<unsigned long long>(), c->rows.as<unsigned long long>(),
            cap - 1, c->d_flags + 2)));
        unsigned int ovf = 0;
        CK_S(small_d2h(c, &ovf, c->d_flags + 2, 4));
        CK_S(stream_sync(c));
        CK(cudaGetLastError());
        if (!ovf) {
            LAUNCH(c, K_VIZ, (k_dict_len<<<(unsigned)cdiv(cap, 256), 256, 0, c->stream>>>(
                v, c->keys.as<unsigned long long>(), c->rows.as<unsigned long long>(), cap, c->d_err + 1, c->d_flags + 3)));
            unsigned long long total = 0; unsigned int n = 0;
            CK_S(small_d2h(c, &total, c->d_err + 1, 8));
            CK_S(small_d2h(c, &n, c->d_flags + 3, 4));
            CK_S(stream_sync(c));
            // occupancy above 50% makes probing slow and overflow likely next time: grow
            if ((uint64_t)n * 2 > cap) continue;
            ENSURE(c->out[2], total + 64);
            CK(cudaMemsetAsync(c->d_err + 1, 0, 8, c->stream));
            if (total)
                LAUNCH(c, K_VIZ, (k_dict_write<<<(unsigned)cdiv(cap, 256), 256, 0, c->stream>>>(
                    v, c->keys.as<unsigned long long>(), c->rows.as<unsigned long long>(), cap, c->d_err + 1, c->out[2].as<char>())));
            CK_S(stream_sync(c));
            CK(cudaGetLastError());
            c->out_n[2] = total;
            if (nbytes) *nbytes = total;
            if (n_names) *n_names = n;
            return 0;
        }
    }
}

int sofa_b200_set_demangled(sofa_ctx* c, uint32_t n, const char* mangled, const uint64_t* moff,
                            const char* demangled, const uint64_t* doff) {
    int r = need_parsed(c, SRC_PERF); if (r) return r;
    Source& s = c->src[SRC_PERF];
    if (s.viz_stride == 0) FAIL(SOFA_ERR_STATE, "set_demangled: call viz_mangled first");
    if (n && (!mangled || !moff || !demangled || !doff)) FAIL(SOFA_ERR_ARG, "set_demangled: null table");
    std::vector<std::pair<uint64_t, uint32_t>> order(n);
    for (uint32_t i = 0; i < n; i++) {
        HashSink h;
        for (uint64_t j = moff[i]; j < moff[i + 1]; j++) h.put(mangled[j]);
        order[i] = {h.h ? h.h : 1, i};
    }
    std::sort(order.begin(), order.end());
    std::vector<uint64_t> hashes(n + 1), offs(n + 2);
    std::string blob;
    for (uint32_t i = 0; i < n; i++) {
        if (i && order[i].first == order[i - 1].first) FAIL(SOFA_ERR_ARG, "set_demangled: duplicate or colliding names");
        hashes[i] = order[i].first;
        offs[i] = blob.size();
        uint32_t src = order[i].second;
        blob.append(demangled + doff[src], demangled + doff[src + 1]);
    }
    offs[n] = blob.size();
    ENSURE(s.dict_hash, (n + 1) * 8); ENSURE(s.dict_off, (n + 2) * 8); ENSURE(s.dict_blob, blob.size() + 16);
    if (n) {
        CK_S(small_h2d(c, s.dict_hash.p, hashes.data(), n * 8));
        CK_S(small_h2d(c, s.dict_blob.p, blob.data(), blob.size()));
    }
    CK_S(small_h2d(c, s.dict_off.p, offs.data(), (n + 1) * 8));
    CK_S(stream_sync(c));
    s.dict_n = (int)n;
    uint64_t ncand = cdiv(s.n_rows, s.viz_stride);
    if (ncand) {
        if (reset_err(c)) return SOFA_ERR_CUDA;
        DictView d{(int)n, s.dict_hash.as<uint64_t>(), s.dict_off.as<uint64_t>(), s.dict_blob.as<char>()};
        LAUNCH(c, K_VIZ, (k_viz_resolve<<<(unsigned)cdiv(ncand, 256), 256, 0, c->stream>>>(
            s.vizhash.as<uint64_t>(), ncand, d, s.viz_id.as<int32_t>(), c->d_err)));
        r = check_kernel_error(c, "set_demangled (a mangled viz name has no demangled entry)");
        if (r) return r;
    }
    s.viz_ready = true;
    return 0;
}

int sofa_b200_get_column(sofa_ctx* c, int k, int col, void* host_out, uint64_t cap_bytes) {
    int r = need_parsed(c, k); if (r) return r;
    Source& s = c->src[k];
    int slot = -1; size_t w = 0;
    if (k == SRC_PERF) {
        switch (col) { case 0: slot = PC_TS; w = 8; break; case 1: slot = PC_EVENT; w = 8; break; case 2: slot = PC_DUR; w = 8; break;
                       case 9: slot = PC_PID; w = 4; break; case 10: slot = PC_TID; w = 4; break; }
    } else if (k == SRC_GPU) {
        switch (col) { case 0: slot = GC_TS; w = 8; break; case 2: slot = GC_DUR; w = 8; break; case 6: slot = GC_BW; w = 8; break;
                       case 5: slot = GC_PAYLOAD; w = 8; break; case 3: slot = GC_DEV; w = 4; break; case 10: slot = GC_TID; w = 4; break;
                       case 7: slot = GC_SRC; w = 4; break; case 8: slot = GC_DST; w = 4; break; case 4: slot = GC_KIND; w = 1; break; }
    } else {
        switch (col) { case 0: slot = NC_TS; w = 8; break; case 5: slot = NC_PAYLOAD; w = 8; break; case 7: slot = NC_SRC; w = 8; break;
                       case 8: slot = NC_DST; w = 8; break; case 4: slot = NC_PROTO; w = 1; break; }
    }
    if (slot < 0) FAIL(SOFA_ERR_ARG, "get_column: no such stored column for this source");
    uint64_t nb = s.n_rows * w;
    if (cap_bytes < nb) FAIL(SOFA_ERR_CAP, "get_column: buffer too small");
    if (nb) { CK(cudaMemcpyAsync(host_out, s.cols[slot].p, nb, cudaMemcpyDeviceToHost, c->stream)); CK_S(stream_sync(c)); }
    return 0;
}

int sofa_b200_mark(sofa_ctx* c, int which) {
    if (!c || which < 0 || which > 1) FAIL(SOFA_ERR_ARG, "mark: bad argument");
    CK(cudaSetDevice(c->device));
    if (!c->mark[which]) CK(cudaEventCreate(&c->mark[which]));
    CK(cudaEventRecord(c->mark[which], c->stream));
    return 0;
}
int sofa_b200_elapsed_ms(sofa_ctx* c, double* ms) {
    if (!c || !ms || !c->mark[0] || !c->mark[1]) FAIL(SOFA_ERR_ARG, "elapsed_ms: marks not set");
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->mark[1]));
    float f = 0;
    CK(cudaEventElapsedTime(&f, c->mark[0], c->mark[1]));
    *ms = f;
    return 0;
}

int sofa_b200_set_profiling(sofa_ctx* c, int on) {
    if (!c) FAIL(SOFA_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    if (drain_events(c)) return SOFA_ERR_CUDA;
    c->profiling = on != 0;
    return 0;
}
int sofa_b200_get_stats(sofa_ctx* c, sofa_stats_t* out) {
    if (!c || !out) FAIL(SOFA_ERR_ARG, "null argument");
    CK(cudaSetDevice(c->device));
    if (drain_events(c)) return SOFA_ERR_CUDA;
    *out = c->stats;
    return 0;
}
int sofa_b200_reset_stats(sofa_ctx* c) {
    if (!c) FAIL(SOFA_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    if (drain_events(c)) return SOFA_ERR_CUDA;
    memset(&c->stats, 0, sizeof c->stats);
    return 0;
}

int sofa_b200_text_fetch(sofa_ctx* c, int k, void* host_out, uint64_t cap) {
    if (!c || k < 0 || k > 2 || !host_out) FAIL(SOFA_ERR_ARG, "text_fetch: bad argument");
    Source& s = c->src[k];
    if (!s.uploaded) FAIL(SOFA_ERR_STATE, "text_fetch: no text");
    if (cap < s.nbytes) FAIL(SOFA_ERR_CAP, "text_fetch: buffer too small");
    CK(cudaSetDevice(c->device));
    if (s.nbytes) { CK(cudaMemcpyAsync(host_out, s.text.p, s.nbytes, cudaMemcpyDeviceToHost, c->stream)); CK_S(stream_sync(c)); }
    return 0;
}

#include "capi_analyze.inc"
#include "capi_window.inc"
#include "capi_side.inc"
#include "capi_synth.inc"

}  // extern "C"
